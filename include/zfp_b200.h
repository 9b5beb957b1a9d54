/*
 * zfp_b200.h - host-side C interface of the B200 (sm_100a) zfp execution backend.
 *
 * This header restates, from scratch, the part of zfp's public C API that lies on the
 * whole-array compress/decompress path, with IDENTICAL type layouts, enum values and function
 * signatures, so that code written against the reference's <zfp.h> compiles and links against
 * libzfp_b200 unchanged and a zfp_stream / zfp_field / bitstream created by either library can
 * be handed to the other.  Each group cites the reference declaration it mirrors
 * (paths relative to the reference tree).
 *
 * What is NOT here: the low-level per-block API, promote/demote helpers, the C++ array classes
 * and cfp - none of them is on the accelerated path.
 *
 * The backend executes zfp_exec_cuda only.  zfp_compress / zfp_decompress under
 * zfp_exec_serial or zfp_exec_omp return 0 ("unsupported", the reference's own convention,
 * src/zfp.c:1110-1113): there is deliberately no CPU fallback in this library.
 */
#ifndef ZFP_B200_H
#define ZFP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar typedefs (include/zfp/internal/zfp/types.h:25-56) ------------------------------ */
typedef unsigned char uchar;
typedef unsigned int uint;
typedef int32_t int32;
typedef uint32_t uint32;
typedef int64_t int64;
typedef uint64_t uint64;

/* ---- limits and header layout (include/zfp.h:18-50) ---------------------------------------- */
#define ZFP_MIN_BITS 1
#define ZFP_MAX_BITS 16658
#define ZFP_MAX_PREC 64
#define ZFP_MIN_EXP (-1074)

#define ZFP_HEADER_NONE 0x0u
#define ZFP_HEADER_MAGIC 0x1u
#define ZFP_HEADER_META 0x2u
#define ZFP_HEADER_MODE 0x4u
#define ZFP_HEADER_FULL 0x7u

#define ZFP_MAGIC_BITS 32
#define ZFP_META_BITS 52
#define ZFP_MODE_SHORT_BITS 12
#define ZFP_MODE_LONG_BITS 64
#define ZFP_HEADER_MAX_BITS 148
#define ZFP_MODE_SHORT_MAX ((1u << ZFP_MODE_SHORT_BITS) - 2)
#define ZFP_META_NULL ((uint64)-1)

#define ZFP_CODEC 5 /* wire-format version (include/zfp/version.h:14) */

/* ---- enums (include/zfp.h:59-72, 95-128) ---------------------------------------------------- */
typedef int zfp_bool;
enum { zfp_false = 0, zfp_true = 1 };

typedef enum { zfp_exec_serial = 0, zfp_exec_omp = 1, zfp_exec_cuda = 2 } zfp_exec_policy;

typedef enum {
  zfp_mode_null = 0,
  zfp_mode_expert = 1,
  zfp_mode_fixed_rate = 2,
  zfp_mode_fixed_precision = 3,
  zfp_mode_fixed_accuracy = 4,
  zfp_mode_reversible = 5
} zfp_mode;

/* compression mode and parameter settings (include/zfp.h:106-119) */
typedef struct {
  zfp_mode mode;
  union {
    double rate;      /* compressed bits/value (negative for word alignment) */
    uint precision;   /* uncompressed bits/value */
    double tolerance; /* absolute error tolerance */
    struct {
      uint minbits, maxbits, maxprec;
      int minexp;
    } expert;
  } arg;
} zfp_config;

typedef enum {
  zfp_type_none = 0,
  zfp_type_int32 = 1,
  zfp_type_int64 = 2,
  zfp_type_float = 3,
  zfp_type_double = 4
} zfp_type;

/* ---- bit stream (include/zfp/bitstream.h:8-94; layout include/zfp/bitstream.inl:133-143) ---- */
typedef struct bitstream bitstream; /* opaque to callers; 64-bit words, LSB first */
typedef uint64 bitstream_offset;
typedef bitstream_offset bitstream_size;
typedef size_t bitstream_count;

extern const size_t stream_word_bits; /* always 64 */

bitstream* stream_open(void* buffer, size_t bytes); /* buffer may be a HOST or a DEVICE pointer */
void stream_close(bitstream* s);
bitstream_count stream_alignment(void);
void* stream_data(const bitstream* s);
size_t stream_size(const bitstream* s);
size_t stream_capacity(const bitstream* s);
bitstream_offset stream_rtell(const bitstream* s);
bitstream_offset stream_wtell(const bitstream* s);
void stream_rewind(bitstream* s);
void stream_rseek(bitstream* s, bitstream_offset offset);
void stream_wseek(bitstream* s, bitstream_offset offset);
/* the bit-level accessors below work on host buffers and, through small synchronous copies,
 * on device buffers too (the reference would fault on a device buffer) */
uint stream_read_bit(bitstream* s);
uint stream_write_bit(bitstream* s, uint bit);
uint64 stream_read_bits(bitstream* s, bitstream_count n);
uint64 stream_write_bits(bitstream* s, uint64 value, bitstream_count n);
void stream_skip(bitstream* s, bitstream_size n);
void stream_pad(bitstream* s, bitstream_size n);
bitstream_count stream_align(bitstream* s);
bitstream_count stream_flush(bitstream* s);

/* ---- execution, stream and field records (include/zfp.h:75-93, 131-136) ------------------- */
typedef struct {
  uint threads;
  uint chunk_size;
} zfp_exec_params_omp;

typedef struct {
  zfp_exec_policy policy;
  void* params; /* owned by the library; for zfp_exec_cuda a zfp_exec_params_cuda (zfp_b200_backend.h) */
} zfp_execution;

typedef struct {
  uint minbits;
  uint maxbits;
  uint maxprec;
  int minexp;
  bitstream* stream;
  zfp_execution exec;
} zfp_stream;

typedef struct {
  zfp_type type;
  size_t nx, ny, nz, nw;    /* 0 marks an unused dimension */
  ptrdiff_t sx, sy, sz, sw; /* element strides; 0 = contiguous a[nw][nz][ny][nx] */
  void* data;               /* HOST or DEVICE pointer */
} zfp_field;

extern const uint zfp_codec_version;
extern const uint zfp_library_version;
extern const char* const zfp_version_string;

size_t zfp_type_size(zfp_type type);

/* ---- fields (include/zfp.h:437-582; src/zfp.c:107-470) -------------------------------------- */
zfp_field* zfp_field_alloc(void);
zfp_field* zfp_field_1d(void* data, zfp_type type, size_t nx);
zfp_field* zfp_field_2d(void* data, zfp_type type, size_t nx, size_t ny);
zfp_field* zfp_field_3d(void* data, zfp_type type, size_t nx, size_t ny, size_t nz);
zfp_field* zfp_field_4d(void* data, zfp_type type, size_t nx, size_t ny, size_t nz, size_t nw);
void zfp_field_free(zfp_field* field);
void* zfp_field_pointer(const zfp_field* field);
void* zfp_field_begin(const zfp_field* field);
zfp_type zfp_field_type(const zfp_field* field);
uint zfp_field_precision(const zfp_field* field);
uint zfp_field_dimensionality(const zfp_field* field);
size_t zfp_field_size(const zfp_field* field, size_t* size);
size_t zfp_field_size_bytes(const zfp_field* field);
size_t zfp_field_blocks(const zfp_field* field);
zfp_bool zfp_field_stride(const zfp_field* field, ptrdiff_t* stride);
zfp_bool zfp_field_is_contiguous(const zfp_field* field);
uint64 zfp_field_metadata(const zfp_field* field);
void zfp_field_set_pointer(zfp_field* field, void* data);
zfp_type zfp_field_set_type(zfp_field* field, zfp_type type);
void zfp_field_set_size_1d(zfp_field* field, size_t nx);
void zfp_field_set_size_2d(zfp_field* field, size_t nx, size_t ny);
void zfp_field_set_size_3d(zfp_field* field, size_t nx, size_t ny, size_t nz);
void zfp_field_set_size_4d(zfp_field* field, size_t nx, size_t ny, size_t nz, size_t nw);
void zfp_field_set_stride_1d(zfp_field* field, ptrdiff_t sx);
void zfp_field_set_stride_2d(zfp_field* field, ptrdiff_t sx, ptrdiff_t sy);
void zfp_field_set_stride_3d(zfp_field* field, ptrdiff_t sx, ptrdiff_t sy, ptrdiff_t sz);
void zfp_field_set_stride_4d(zfp_field* field, ptrdiff_t sx, ptrdiff_t sy, ptrdiff_t sz, ptrdiff_t sw);
zfp_bool zfp_field_set_metadata(zfp_field* field, uint64 meta);

/* ---- compressed-stream object and parameters (include/zfp.h:160-330; src/zfp.c:536-915) ---- */
zfp_stream* zfp_stream_open(bitstream* stream);
void zfp_stream_close(zfp_stream* zfp);
bitstream* zfp_stream_bit_stream(const zfp_stream* zfp);
void zfp_stream_set_bit_stream(zfp_stream* zfp, bitstream* stream);
void zfp_stream_rewind(zfp_stream* zfp);
size_t zfp_stream_flush(zfp_stream* zfp);
size_t zfp_stream_align(zfp_stream* zfp);

zfp_mode zfp_stream_compression_mode(const zfp_stream* zfp);
double zfp_stream_rate(const zfp_stream* zfp, uint dims);
uint zfp_stream_precision(const zfp_stream* zfp);
double zfp_stream_accuracy(const zfp_stream* zfp);
uint64 zfp_stream_mode(const zfp_stream* zfp);
void zfp_stream_params(const zfp_stream* zfp, uint* minbits, uint* maxbits, uint* maxprec, int* minexp);
size_t zfp_stream_compressed_size(const zfp_stream* zfp);
size_t zfp_stream_maximum_size(const zfp_stream* zfp, const zfp_field* field);

void zfp_stream_set_reversible(zfp_stream* zfp);
double zfp_stream_set_rate(zfp_stream* zfp, double rate, zfp_type type, uint dims, zfp_bool align);
uint zfp_stream_set_precision(zfp_stream* zfp, uint precision);
double zfp_stream_set_accuracy(zfp_stream* zfp, double tolerance);
zfp_mode zfp_stream_set_mode(zfp_stream* zfp, uint64 mode);
zfp_bool zfp_stream_set_params(zfp_stream* zfp, uint minbits, uint maxbits, uint maxprec, int minexp);

/* ---- execution policy (include/zfp.h:332-373; src/zfp.c:893-990) --------------------------- */
zfp_exec_policy zfp_stream_execution(const zfp_stream* zfp);
zfp_bool zfp_stream_set_execution(zfp_stream* zfp, zfp_exec_policy policy); /* zfp_exec_omp -> zfp_false */
uint zfp_stream_omp_threads(const zfp_stream* zfp);
uint zfp_stream_omp_chunk_size(const zfp_stream* zfp);
/* OpenMP execution lives in the reference library: both return zfp_false and leave the policy alone
 * (include/zfp.h:322-333; the reference returns zfp_false the same way when built without OpenMP) */
zfp_bool zfp_stream_set_omp_threads(zfp_stream* zfp, uint threads);
zfp_bool zfp_stream_set_omp_chunk_size(zfp_stream* zfp, uint chunk_size);

/* ---- mode configurations (include/zfp.h:335-371; src/zfp.c:466-533) -------------------------- */
zfp_config zfp_config_none(void);
zfp_config zfp_config_rate(double rate, zfp_bool align);
zfp_config zfp_config_precision(uint precision);
zfp_config zfp_config_accuracy(double tolerance);
zfp_config zfp_config_reversible(void);
zfp_config zfp_config_expert(uint minbits, uint maxbits, uint maxprec, int minexp);

/* ---- the hot path (include/zfp.h:585-627; src/zfp.c:1051-1249) ----------------------------- */
/* Both return the cumulative stream size in bytes, 0 on failure / unsupported.  field->data and
 * the bit stream's buffer may each live on the host or on the device (src/cuda_zfp/cuZFP.cu
 * pointer semantics); the timed path is device-resident on both sides. */
size_t zfp_compress(zfp_stream* zfp, const zfp_field* field);
size_t zfp_decompress(zfp_stream* zfp, zfp_field* field);
size_t zfp_write_header(zfp_stream* zfp, const zfp_field* field, uint mask);
size_t zfp_read_header(zfp_stream* zfp, zfp_field* field, uint mask);

#ifdef __cplusplus
}
#endif

#endif /* ZFP_B200_H */
