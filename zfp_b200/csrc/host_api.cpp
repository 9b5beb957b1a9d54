// host_api.cpp - host side of libzfp_b200: bit stream object, field and stream records,
// parameter/mode arithmetic, header I/O and the execution dispatch of zfp_compress /
// zfp_decompress.  Written from the behaviour documented in the reference (citations per
// function); the ABI (struct layouts, enum values, signatures) is the reference's.
//
// Execution dispatch ("changed subsystem (1)" of the north star): the reference indexes a
// function table ftable[exec][strided][dims-1][type-1] (src/zfp.c:1055-1093, 1126-1153) whose
// CUDA row is fixed-rate-only and has no 4-D entries.  Here the CUDA policy accepts every
// (type, dims, mode) and forwards to one backend call; the serial and OpenMP policies are not
// implemented in this library and return 0.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

#include "../../include/zfp_b200_backend.h"
#include "bitstream_impl.h"

extern "C" {

// ------------------------------------------------------------------------------------------------
// bit stream (include/zfp/bitstream.inl).  Word I/O goes through two helpers so that a stream
// opened on a device buffer stays usable from the host for the few words a header needs.
// ------------------------------------------------------------------------------------------------
const size_t stream_word_bits = 64;

static bool is_device_memory(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice;
}

static uint64 load_word(const uint64* p)
{
  if (is_device_memory(p)) {
    uint64 v = 0;
    cudaMemcpy(&v, p, sizeof(v), cudaMemcpyDeviceToHost);
    return v;
  }
  return *p;
}

static void store_word(uint64* p, uint64 v)
{
  if (is_device_memory(p))
    cudaMemcpy(p, &v, sizeof(v), cudaMemcpyHostToDevice);
  else
    *p = v;
}

bitstream* stream_open(void* buffer, size_t bytes)
{
  bitstream* s = static_cast<bitstream*>(malloc(sizeof(bitstream)));
  if (s) {
    s->begin = static_cast<uint64*>(buffer);
    s->end = s->begin + bytes / sizeof(uint64);
    stream_rewind(s);
  }
  return s;
}

void stream_close(bitstream* s) { free(s); }
bitstream_count stream_alignment(void) { return 64; }
void* stream_data(const bitstream* s) { return s->begin; }
size_t stream_size(const bitstream* s) { return (size_t)(s->ptr - s->begin) * sizeof(uint64); }
size_t stream_capacity(const bitstream* s) { return (size_t)(s->end - s->begin) * sizeof(uint64); }
bitstream_offset stream_rtell(const bitstream* s) { return (bitstream_offset)(s->ptr - s->begin) * 64 - s->bits; }
bitstream_offset stream_wtell(const bitstream* s) { return (bitstream_offset)(s->ptr - s->begin) * 64 + s->bits; }

void stream_rewind(bitstream* s)
{
  s->ptr = s->begin;
  s->buffer = 0;
  s->bits = 0;
}

// reading keeps the not-yet-consumed high part of the last fetched word in `buffer`
void stream_rseek(bitstream* s, bitstream_offset offset)
{
  const unsigned r = (unsigned)(offset & 63);
  s->ptr = s->begin + (offset >> 6);
  s->buffer = 0;
  s->bits = 0;
  if (r) {
    s->buffer = load_word(s->ptr++) >> r;
    s->bits = 64 - r;
  }
}

// writing keeps the already-written low part of the current word in `buffer`
void stream_wseek(bitstream* s, bitstream_offset offset)
{
  const unsigned r = (unsigned)(offset & 63);
  s->ptr = s->begin + (offset >> 6);
  s->buffer = 0;
  s->bits = 0;
  if (r) {
    s->buffer = load_word(s->ptr) & ((uint64(1) << r) - 1);
    s->bits = r;
  }
}

uint64 stream_read_bits(bitstream* s, bitstream_count n)
{
  uint64 value = s->buffer;
  if (n <= s->bits) {
    s->bits -= n;
    s->buffer = n < 64 ? s->buffer >> n : 0;
    return n < 64 ? value & ((uint64(1) << n) - 1) : value;
  }
  // need another word: low part from the buffer, high part from the new word
  const uint64 w = load_word(s->ptr++);
  const size_t have = s->bits, take = n - have;  // 1 <= take <= 64
  value |= have < 64 ? w << have : 0;
  s->bits = 64 - take;
  s->buffer = take < 64 ? w >> take : 0;
  return n < 64 ? value & ((uint64(1) << n) - 1) : value;
}

uint stream_read_bit(bitstream* s) { return (uint)stream_read_bits(s, 1); }

uint64 stream_write_bits(bitstream* s, uint64 value, bitstream_count n)
{
  const uint64 rest = n < 64 ? value >> n : 0;
  if (!n) return rest;
  if (n < 64) value &= (uint64(1) << n) - 1;
  const size_t have = s->bits;
  s->buffer |= value << have;
  if (have + n >= 64) {
    store_word(s->ptr++, s->buffer);
    s->buffer = have ? value >> (64 - have) : 0;
    s->bits = have + n - 64;
  }
  else
    s->bits = have + n;
  return rest;
}

uint stream_write_bit(bitstream* s, uint bit)
{
  stream_write_bits(s, bit, 1);
  return bit;
}

void stream_skip(bitstream* s, bitstream_size n) { stream_rseek(s, stream_rtell(s) + n); }

void stream_pad(bitstream* s, bitstream_size n)
{
  while (n >= 64) { stream_write_bits(s, 0, 64); n -= 64; }
  stream_write_bits(s, 0, (bitstream_count)n);
}

bitstream_count stream_align(bitstream* s)
{
  const bitstream_count r = s->bits;
  if (r) stream_skip(s, r);
  return r;
}

bitstream_count stream_flush(bitstream* s)
{
  const bitstream_count r = (64 - s->bits) % 64;
  if (r) stream_pad(s, r);
  return r;
}

// ------------------------------------------------------------------------------------------------
// fields (src/zfp.c:107-470)
// ------------------------------------------------------------------------------------------------
const uint zfp_codec_version = ZFP_CODEC;
const uint zfp_library_version = 0x1010;
const char* const zfp_version_string = "zfp-b200 backend (zfp codec 5, API of zfp 1.0.1)";

size_t zfp_type_size(zfp_type type)
{
  switch (type) {
    case zfp_type_int32: case zfp_type_float: return 4;
    case zfp_type_int64: case zfp_type_double: return 8;
    default: return 0;
  }
}

zfp_field* zfp_field_alloc(void) { return static_cast<zfp_field*>(calloc(1, sizeof(zfp_field))); }

static zfp_field* new_field(void* data, zfp_type type, size_t nx, size_t ny, size_t nz, size_t nw)
{
  zfp_field* f = zfp_field_alloc();
  if (f) {
    f->type = type;
    f->nx = nx; f->ny = ny; f->nz = nz; f->nw = nw;
    f->data = data;
  }
  return f;
}

zfp_field* zfp_field_1d(void* data, zfp_type type, size_t nx) { return new_field(data, type, nx, 0, 0, 0); }
zfp_field* zfp_field_2d(void* data, zfp_type type, size_t nx, size_t ny) { return new_field(data, type, nx, ny, 0, 0); }
zfp_field* zfp_field_3d(void* data, zfp_type type, size_t nx, size_t ny, size_t nz) { return new_field(data, type, nx, ny, nz, 0); }
zfp_field* zfp_field_4d(void* data, zfp_type type, size_t nx, size_t ny, size_t nz, size_t nw) { return new_field(data, type, nx, ny, nz, nw); }
void zfp_field_free(zfp_field* field) { free(field); }
void* zfp_field_pointer(const zfp_field* field) { return field->data; }
zfp_type zfp_field_type(const zfp_field* field) { return field->type; }
uint zfp_field_precision(const zfp_field* field) { return (uint)(8 * zfp_type_size(field->type)); }

uint zfp_field_dimensionality(const zfp_field* f)
{
  if (!f->nx) return 0;
  if (!f->ny) return 1;
  if (!f->nz) return 2;
  return f->nw ? 4 : 3;
}

static void field_arrays(const zfp_field* f, size_t n[4], ptrdiff_t s[4])
{
  const size_t dim[4] = { f->nx, f->ny, f->nz, f->nw };
  const ptrdiff_t str[4] = { f->sx, f->sy, f->sz, f->sw };
  ptrdiff_t contiguous = 1;
  for (int i = 0; i < 4; i++) {
    n[i] = dim[i];
    s[i] = str[i] ? str[i] : contiguous;
    contiguous *= (ptrdiff_t)(dim[i] ? dim[i] : 1);
  }
}

// lowest / highest element offsets reached (src/zfp.c field_index_span)
static void field_span(const zfp_field* f, ptrdiff_t* lo, ptrdiff_t* hi)
{
  size_t n[4];
  ptrdiff_t s[4];
  field_arrays(f, n, s);
  *lo = *hi = 0;
  for (uint i = 0; i < zfp_field_dimensionality(f); i++) {
    ptrdiff_t reach = s[i] * (ptrdiff_t)(n[i] - 1);
    if (reach < 0) *lo += reach; else *hi += reach;
  }
}

void* zfp_field_begin(const zfp_field* field)
{
  if (!field->data) return NULL;
  ptrdiff_t lo, hi;
  field_span(field, &lo, &hi);
  return static_cast<uchar*>(field->data) + lo * (ptrdiff_t)zfp_type_size(field->type);
}

size_t zfp_field_size(const zfp_field* f, size_t* size)
{
  const uint dims = zfp_field_dimensionality(f);
  const size_t n[4] = { f->nx, f->ny, f->nz, f->nw };
  size_t total = 1;
  for (uint i = 0; i < 4; i++) {
    if (size && i < dims) size[i] = n[i];
    total *= n[i] ? n[i] : 1;
  }
  return total;
}

size_t zfp_field_size_bytes(const zfp_field* field)
{
  ptrdiff_t lo, hi;
  field_span(field, &lo, &hi);
  return (size_t)(hi - lo + 1) * zfp_type_size(field->type);
}

size_t zfp_field_blocks(const zfp_field* f)
{
  const uint dims = zfp_field_dimensionality(f);
  const size_t n[4] = { f->nx, f->ny, f->nz, f->nw };
  if (!dims) return 0;
  size_t blocks = 1;
  for (uint i = 0; i < dims; i++)
    blocks *= (n[i] + 3) / 4;
  return blocks;
}

zfp_bool zfp_field_stride(const zfp_field* f, ptrdiff_t* stride)
{
  if (stride) {
    size_t n[4];
    ptrdiff_t s[4];
    field_arrays(f, n, s);
    for (uint i = 0; i < zfp_field_dimensionality(f); i++)
      stride[i] = s[i];
  }
  return f->sx || f->sy || f->sz || f->sw;
}

zfp_bool zfp_field_is_contiguous(const zfp_field* field)
{
  ptrdiff_t lo, hi;
  field_span(field, &lo, &hi);
  return (size_t)(hi - lo + 1) == zfp_field_size(field, NULL);
}

// 52-bit metadata: per-dimension (n-1) fields of 48/dims bits, then 2 bits dims-1, 2 bits type-1
// (src/zfp.c zfp_field_metadata / zfp_field_set_metadata)
uint64 zfp_field_metadata(const zfp_field* f)
{
  const uint dims = zfp_field_dimensionality(f);
  const size_t n[4] = { f->nx, f->ny, f->nz, f->nw };
  if (!dims) return ZFP_META_NULL;
  const uint width = 48 / dims;
  uint64 meta = 0;
  for (int i = (int)dims - 1; i >= 0; i--) {
    const uint64 v = (uint64)(n[i] - 1);
    if (v >> width) return ZFP_META_NULL;
    meta = (meta << width) + v;
  }
  meta = (meta << 2) + (dims - 1);
  meta = (meta << 2) + (uint64)(f->type - 1);
  return meta;
}

zfp_bool zfp_field_set_metadata(zfp_field* f, uint64 meta)
{
  if (meta >> ZFP_META_BITS) return zfp_false;
  f->type = (zfp_type)((meta & 3u) + 1); meta >>= 2;
  const uint dims = (uint)(meta & 3u) + 1; meta >>= 2;
  const uint width = 48 / dims;
  size_t n[4] = { 0, 0, 0, 0 };
  for (uint i = 0; i < dims; i++) {
    // 1-D sizes are limited to 32 bits even though 48 are stored
    const uint64 mask = dims == 1 ? 0xffffffffull : ((uint64(1) << width) - 1);
    n[i] = (size_t)(meta & mask) + 1;
    meta >>= width;
  }
  f->nx = n[0]; f->ny = n[1]; f->nz = n[2]; f->nw = n[3];
  f->sx = f->sy = f->sz = f->sw = 0;
  return zfp_true;
}

void zfp_field_set_pointer(zfp_field* field, void* data) { field->data = data; }

zfp_type zfp_field_set_type(zfp_field* field, zfp_type type)
{
  if (!zfp_type_size(type)) return zfp_type_none;
  field->type = type;
  return type;
}

void zfp_field_set_size_1d(zfp_field* f, size_t nx) { f->nx = nx; f->ny = f->nz = f->nw = 0; }
void zfp_field_set_size_2d(zfp_field* f, size_t nx, size_t ny) { f->nx = nx; f->ny = ny; f->nz = f->nw = 0; }
void zfp_field_set_size_3d(zfp_field* f, size_t nx, size_t ny, size_t nz) { f->nx = nx; f->ny = ny; f->nz = nz; f->nw = 0; }
void zfp_field_set_size_4d(zfp_field* f, size_t nx, size_t ny, size_t nz, size_t nw) { f->nx = nx; f->ny = ny; f->nz = nz; f->nw = nw; }
void zfp_field_set_stride_1d(zfp_field* f, ptrdiff_t sx) { f->sx = sx; f->sy = f->sz = f->sw = 0; }
void zfp_field_set_stride_2d(zfp_field* f, ptrdiff_t sx, ptrdiff_t sy) { f->sx = sx; f->sy = sy; f->sz = f->sw = 0; }
void zfp_field_set_stride_3d(zfp_field* f, ptrdiff_t sx, ptrdiff_t sy, ptrdiff_t sz) { f->sx = sx; f->sy = sy; f->sz = sz; f->sw = 0; }
void zfp_field_set_stride_4d(zfp_field* f, ptrdiff_t sx, ptrdiff_t sy, ptrdiff_t sz, ptrdiff_t sw) { f->sx = sx; f->sy = sy; f->sz = sz; f->sw = sw; }

// ------------------------------------------------------------------------------------------------
// compressed-stream object and its four parameters (src/zfp.c:536-915)
// ------------------------------------------------------------------------------------------------
static void drop_exec_params(zfp_stream* zfp)
{
  if (zfp->exec.params) {
    if (zfp->exec.policy == zfp_exec_cuda) {
      zfp_exec_params_cuda* p = static_cast<zfp_exec_params_cuda*>(zfp->exec.params);
      if (p->magic == ZFP_B200_PARAMS_MAGIC && p->index) zfp_b200_index_destroy(p->index);
    }
    free(zfp->exec.params);
    zfp->exec.params = NULL;
  }
}

zfp_stream* zfp_stream_open(bitstream* stream)
{
  zfp_stream* zfp = static_cast<zfp_stream*>(malloc(sizeof(zfp_stream)));
  if (zfp) {
    zfp->minbits = ZFP_MIN_BITS;
    zfp->maxbits = ZFP_MAX_BITS;
    zfp->maxprec = ZFP_MAX_PREC;
    zfp->minexp = ZFP_MIN_EXP;
    zfp->stream = stream;
    zfp->exec.policy = zfp_exec_serial;
    zfp->exec.params = NULL;
  }
  return zfp;
}

void zfp_stream_close(zfp_stream* zfp)
{
  drop_exec_params(zfp);
  free(zfp);
}

bitstream* zfp_stream_bit_stream(const zfp_stream* zfp) { return zfp->stream; }
void zfp_stream_set_bit_stream(zfp_stream* zfp, bitstream* stream) { zfp->stream = stream; }
void zfp_stream_rewind(zfp_stream* zfp) { stream_rewind(zfp->stream); }
size_t zfp_stream_flush(zfp_stream* zfp) { return stream_flush(zfp->stream); }
size_t zfp_stream_align(zfp_stream* zfp) { return stream_align(zfp->stream); }
size_t zfp_stream_compressed_size(const zfp_stream* zfp) { return stream_size(zfp->stream); }

// which of the named modes do the four parameters spell? (src/zfp.c:566-608; tests are ordered)
zfp_mode zfp_stream_compression_mode(const zfp_stream* zfp)
{
  const uint lo = zfp->minbits, hi = zfp->maxbits, prec = zfp->maxprec;
  const int emin = zfp->minexp;
  if (lo > hi || prec < 1 || prec > 64) return zfp_mode_null;
  if (lo == ZFP_MIN_BITS && hi == ZFP_MAX_BITS && prec == ZFP_MAX_PREC && emin == ZFP_MIN_EXP) return zfp_mode_expert;
  if (lo == hi && hi >= 1 && hi <= ZFP_MAX_BITS && prec >= ZFP_MAX_PREC && emin == ZFP_MIN_EXP) return zfp_mode_fixed_rate;
  const bool unbounded = lo <= ZFP_MIN_BITS && hi >= ZFP_MAX_BITS;
  if (unbounded && emin == ZFP_MIN_EXP) return zfp_mode_fixed_precision;  // prec >= 1 holds here
  if (unbounded && prec >= ZFP_MAX_PREC && emin >= ZFP_MIN_EXP) return zfp_mode_fixed_accuracy;
  if (unbounded && prec >= ZFP_MAX_PREC && emin < ZFP_MIN_EXP) return zfp_mode_reversible;
  return zfp_mode_expert;
}

double zfp_stream_rate(const zfp_stream* zfp, uint dims)
{
  return zfp_stream_compression_mode(zfp) == zfp_mode_fixed_rate ? (double)zfp->maxbits / (double)(1u << (2 * dims)) : 0.0;
}

uint zfp_stream_precision(const zfp_stream* zfp)
{
  return zfp_stream_compression_mode(zfp) == zfp_mode_fixed_precision ? zfp->maxprec : 0;
}

double zfp_stream_accuracy(const zfp_stream* zfp)
{
  return zfp_stream_compression_mode(zfp) == zfp_mode_fixed_accuracy ? ldexp(1.0, zfp->minexp) : 0.0;
}

void zfp_stream_params(const zfp_stream* zfp, uint* minbits, uint* maxbits, uint* maxprec, int* minexp)
{
  if (minbits) *minbits = zfp->minbits;
  if (maxbits) *maxbits = zfp->maxbits;
  if (maxprec) *maxprec = zfp->maxprec;
  if (minexp) *minexp = zfp->minexp;
}

zfp_bool zfp_stream_set_params(zfp_stream* zfp, uint minbits, uint maxbits, uint maxprec, int minexp)
{
  if (minbits > maxbits || maxprec < 1 || maxprec > 64) return zfp_false;
  zfp->minbits = minbits;
  zfp->maxbits = maxbits;
  zfp->maxprec = maxprec;
  zfp->minexp = minexp;
  return zfp_true;
}

void zfp_stream_set_reversible(zfp_stream* zfp)
{
  zfp_stream_set_params(zfp, ZFP_MIN_BITS, ZFP_MAX_BITS, ZFP_MAX_PREC, ZFP_MIN_EXP - 1);
}

// bits per block = round(4^d * rate), at least the float header, optionally word aligned
// (src/zfp.c:759-784)
double zfp_stream_set_rate(zfp_stream* zfp, double rate, zfp_type type, uint dims, zfp_bool align)
{
  const uint values = 1u << (2 * dims);
  uint bits = (uint)floor(values * rate + 0.5);
  const uint header = type == zfp_type_float ? 9u : type == zfp_type_double ? 12u : 0u;
  if (bits < header) bits = header;
  if (align) bits = (bits + 63u) & ~63u;
  zfp_stream_set_params(zfp, bits, bits, ZFP_MAX_PREC, ZFP_MIN_EXP);
  return (double)bits / values;
}

uint zfp_stream_set_precision(zfp_stream* zfp, uint precision)
{
  const uint p = (precision == 0 || precision > ZFP_MAX_PREC) ? ZFP_MAX_PREC : precision;
  zfp_stream_set_params(zfp, ZFP_MIN_BITS, ZFP_MAX_BITS, p, ZFP_MIN_EXP);
  return p;
}

// minexp = floor(log2(tolerance)) (src/zfp.c:797-811)
double zfp_stream_set_accuracy(zfp_stream* zfp, double tolerance)
{
  int emin = ZFP_MIN_EXP;
  if (tolerance > 0) {
    frexp(tolerance, &emin);  // tolerance = m * 2^emin, 0.5 <= m < 1
    emin -= 1;
  }
  zfp_stream_set_params(zfp, ZFP_MIN_BITS, ZFP_MAX_BITS, ZFP_MAX_PREC, emin);
  return tolerance > 0 ? ldexp(1.0, emin) : 0.0;
}

// compact parameter encoding (src/zfp.c:634-690): 12-bit codes
//   [0, 2047] fixed rate (maxbits-1) | [2048, 2175] fixed precision | 2176 reversible |
//   [2177, 4094] fixed accuracy (minexp + 1074) ; otherwise 64 bits: 0xfff marker + 4 raw fields
uint64 zfp_stream_mode(const zfp_stream* zfp)
{
  switch (zfp_stream_compression_mode(zfp)) {
    case zfp_mode_fixed_rate:
      if (zfp->maxbits <= 2048) return zfp->maxbits - 1;
      break;
    case zfp_mode_fixed_precision:
      if (zfp->maxprec <= 128) return 2048 + (zfp->maxprec - 1);
      break;
    case zfp_mode_fixed_accuracy:
      if (zfp->minexp <= 843) return 2177 + (uint64)(zfp->minexp - ZFP_MIN_EXP);
      break;
    case zfp_mode_reversible:
      return 2176;
    default:
      break;
  }
  auto clampu = [](uint v, uint hi) { return (v < 1 ? 1u : v > hi ? hi : v) - 1; };
  int e = zfp->minexp + 16495;
  e = e < 0 ? 0 : e > 0x7fff ? 0x7fff : e;
  uint64 mode = (uint64)e;
  mode = (mode << 7) + clampu(zfp->maxprec, 0x80u);
  mode = (mode << 15) + clampu(zfp->maxbits, 0x8000u);
  mode = (mode << 15) + clampu(zfp->minbits, 0x8000u);
  mode = (mode << 12) + 0xfffu;
  return mode;
}

zfp_mode zfp_stream_set_mode(zfp_stream* zfp, uint64 mode)
{
  uint minbits = ZFP_MIN_BITS, maxbits = ZFP_MAX_BITS, maxprec = ZFP_MAX_PREC;
  int minexp = ZFP_MIN_EXP;
  if (mode <= ZFP_MODE_SHORT_MAX) {
    if (mode < 2048) minbits = maxbits = (uint)mode + 1;
    else if (mode < 2176) maxprec = (uint)mode - 2048 + 1;
    else if (mode == 2176) minexp = ZFP_MIN_EXP - 1;
    else minexp = (int)mode - 2177 + ZFP_MIN_EXP;
  }
  else {
    mode >>= 12;
    minbits = (uint)(mode & 0x7fffu) + 1; mode >>= 15;
    maxbits = (uint)(mode & 0x7fffu) + 1; mode >>= 15;
    maxprec = (uint)(mode & 0x7fu) + 1; mode >>= 7;
    minexp = (int)(mode & 0x7fffu) - 16495;
  }
  if (!zfp_stream_set_params(zfp, minbits, maxbits, maxprec, minexp)) return zfp_mode_null;
  return zfp_stream_compression_mode(zfp);
}

// conservative buffer size (src/zfp.c:711-742); callers size device buffers with it
size_t zfp_stream_maximum_size(const zfp_stream* zfp, const zfp_field* field)
{
  zfp_b200_desc d;
  memset(&d, 0, sizeof(d));
  d.type = (int)field->type;
  d.dims = zfp_field_dimensionality(field);
  if (!d.dims || !zfp_type_size(field->type)) return 0;
  d.n[0] = field->nx; d.n[1] = field->ny; d.n[2] = field->nz; d.n[3] = field->nw;
  d.minbits = zfp->minbits; d.maxbits = zfp->maxbits; d.maxprec = zfp->maxprec; d.minexp = zfp->minexp;
  return zfp_b200_capacity(&d, 0);
}

// ------------------------------------------------------------------------------------------------
// execution policy (src/zfp.c:893-990)
// ------------------------------------------------------------------------------------------------
zfp_exec_policy zfp_stream_execution(const zfp_stream* zfp) { return zfp->exec.policy; }
uint zfp_stream_omp_threads(const zfp_stream*) { return 0; }
uint zfp_stream_omp_chunk_size(const zfp_stream*) { return 0; }

zfp_bool zfp_stream_set_execution(zfp_stream* zfp, zfp_exec_policy policy)
{
  if (policy != zfp_exec_serial && policy != zfp_exec_cuda) return zfp_false;  // no OpenMP here
  if (zfp->exec.policy != policy) drop_exec_params(zfp);
  zfp->exec.policy = policy;
  return zfp_true;
}

// OpenMP execution is the reference library's; as in a reference build without OpenMP the request
// fails and the current policy stays (src/zfp.c:956-974 via zfp_stream_set_execution)
zfp_bool zfp_stream_set_omp_threads(zfp_stream*, uint) { return zfp_false; }
zfp_bool zfp_stream_set_omp_chunk_size(zfp_stream*, uint) { return zfp_false; }

// mode configurations (src/zfp.c:466-533)
zfp_config zfp_config_none(void)
{
  zfp_config c;
  memset(&c, 0, sizeof(c));
  c.mode = zfp_mode_null;
  return c;
}
zfp_config zfp_config_rate(double rate, zfp_bool align)
{
  zfp_config c = zfp_config_none();
  c.mode = zfp_mode_fixed_rate;
  c.arg.rate = align ? -rate : +rate;
  return c;
}
zfp_config zfp_config_precision(uint precision)
{
  zfp_config c = zfp_config_none();
  c.mode = zfp_mode_fixed_precision;
  c.arg.precision = precision;
  return c;
}
zfp_config zfp_config_accuracy(double tolerance)
{
  zfp_config c = zfp_config_none();
  c.mode = zfp_mode_fixed_accuracy;
  c.arg.tolerance = tolerance;
  return c;
}
zfp_config zfp_config_reversible(void)
{
  zfp_config c = zfp_config_none();
  c.mode = zfp_mode_reversible;
  return c;
}
zfp_config zfp_config_expert(uint minbits, uint maxbits, uint maxprec, int minexp)
{
  zfp_config c = zfp_config_none();
  c.mode = zfp_mode_expert;
  c.arg.expert.minbits = minbits;
  c.arg.expert.maxbits = maxbits;
  c.arg.expert.maxprec = maxprec;
  c.arg.expert.minexp = minexp;
  return c;
}

zfp_exec_params_cuda* zfp_stream_cuda_params(zfp_stream* zfp)
{
  if (zfp->exec.policy != zfp_exec_cuda) return NULL;
  if (!zfp->exec.params) {
    zfp_exec_params_cuda* p = static_cast<zfp_exec_params_cuda*>(calloc(1, sizeof(zfp_exec_params_cuda)));
    if (!p) return NULL;
    p->magic = ZFP_B200_PARAMS_MAGIC;
    zfp->exec.params = p;
  }
  zfp_exec_params_cuda* p = static_cast<zfp_exec_params_cuda*>(zfp->exec.params);
  return p->magic == ZFP_B200_PARAMS_MAGIC ? p : NULL;
}

// ------------------------------------------------------------------------------------------------
// the hot path (src/zfp.c:1051-1180)
// ------------------------------------------------------------------------------------------------
size_t zfp_compress(zfp_stream* zfp, const zfp_field* field)
{
  if (!zfp_type_size(field->type) || !zfp_field_dimensionality(field)) return 0;
  if (zfp->exec.policy != zfp_exec_cuda) return 0;  // serial / OpenMP live in the reference library
  zfp_stream_cuda_params(zfp);                      // variable-rate compress parks its block index here
  const size_t bytes = zfp_b200_compress_stream(zfp, field);
  if (!bytes) return 0;
  stream_flush(zfp->stream);  // no-op: the backend leaves the stream word aligned
  return stream_size(zfp->stream);
}

size_t zfp_decompress(zfp_stream* zfp, zfp_field* field)
{
  if (!zfp_type_size(field->type) || !zfp_field_dimensionality(field)) return 0;
  if (zfp->exec.policy != zfp_exec_cuda) return 0;
  if (!zfp_b200_decompress_stream(zfp, field)) return 0;
  stream_align(zfp->stream);  // no-op: the backend leaves the stream word aligned
  return stream_size(zfp->stream);
}

// header = 'z','f','p',codec (32 bits) | field metadata (52 bits) | mode (12 or 64 bits)
// (src/zfp.c:1182-1249)
size_t zfp_write_header(zfp_stream* zfp, const zfp_field* field, uint mask)
{
  size_t bits = 0;
  uint64 meta = 0;
  if (mask & ZFP_HEADER_META) {
    meta = zfp_field_metadata(field);
    if (meta == ZFP_META_NULL) return 0;
  }
  if (mask & ZFP_HEADER_MAGIC) {
    const uint64 magic = (uint64)'z' | ((uint64)'f' << 8) | ((uint64)'p' << 16) | ((uint64)zfp_codec_version << 24);
    stream_write_bits(zfp->stream, magic, ZFP_MAGIC_BITS);
    bits += ZFP_MAGIC_BITS;
  }
  if (mask & ZFP_HEADER_META) {
    stream_write_bits(zfp->stream, meta, ZFP_META_BITS);
    bits += ZFP_META_BITS;
  }
  if (mask & ZFP_HEADER_MODE) {
    const uint64 mode = zfp_stream_mode(zfp);
    const uint size = mode > ZFP_MODE_SHORT_MAX ? ZFP_MODE_LONG_BITS : ZFP_MODE_SHORT_BITS;
    stream_write_bits(zfp->stream, mode, size);
    bits += size;
  }
  return bits;
}

size_t zfp_read_header(zfp_stream* zfp, zfp_field* field, uint mask)
{
  size_t bits = 0;
  if (mask & ZFP_HEADER_MAGIC) {
    const uint64 expect = (uint64)'z' | ((uint64)'f' << 8) | ((uint64)'p' << 16) | ((uint64)zfp_codec_version << 24);
    // the reference stops reading at the first mismatching byte (short-circuit ||)
    for (int i = 0; i < 4; i++)
      if (stream_read_bits(zfp->stream, 8) != ((expect >> (8 * i)) & 0xff)) return 0;
    bits += ZFP_MAGIC_BITS;
  }
  if (mask & ZFP_HEADER_META) {
    if (!zfp_field_set_metadata(field, stream_read_bits(zfp->stream, ZFP_META_BITS))) return 0;
    bits += ZFP_META_BITS;
  }
  if (mask & ZFP_HEADER_MODE) {
    uint64 mode = stream_read_bits(zfp->stream, ZFP_MODE_SHORT_BITS);
    bits += ZFP_MODE_SHORT_BITS;
    if (mode > ZFP_MODE_SHORT_MAX) {
      mode += stream_read_bits(zfp->stream, ZFP_MODE_LONG_BITS - ZFP_MODE_SHORT_BITS) << ZFP_MODE_SHORT_BITS;
      bits += ZFP_MODE_LONG_BITS - ZFP_MODE_SHORT_BITS;
    }
    if (zfp_stream_set_mode(zfp, mode) == zfp_mode_null) return 0;
  }
  return bits;
}

}  // extern "C"
