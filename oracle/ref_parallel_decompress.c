/* ref_parallel_decompress.c - chunk-parallel FIXED-RATE decompression on the host, for the CPU side of the
 * baseline table.  TEST / BASELINE INFRASTRUCTURE, never linked into the product.
 *
 * The reference has OpenMP compression but no OpenMP decompression (src/zfp.c:1137-1138 holds NULL rows;
 * docs/source/execution.rst:302-304 announces it for a later release).  In fixed-rate mode block b sits at bit
 * b * maxbits, so the strategy of src/template/ompcompress.c:168-198 carries over unchanged: cut the block
 * range into chunks, give every thread its own bitstream over the same buffer, seek, and decode blocks with the
 * reference's OWN low-level block API (include/zfp.h:700-770, zfp_decode_[partial_]block_strided_*).  Nothing of
 * the codec is restated here; this file only drives libzfp_ref.so.
 */
#include <stddef.h>
#include <stdint.h>
#include "zfp.h"

typedef size_t (*dec_full)(zfp_stream*, void*, ptrdiff_t, ptrdiff_t, ptrdiff_t, ptrdiff_t);

static size_t decode_one(zfp_stream* z, zfp_type type, unsigned dims, char* p, const size_t* ext, const ptrdiff_t* s, int full)
{
#define CALL(T, SUF)                                                                                              \
  switch (dims) {                                                                                                 \
    case 1: return full ? zfp_decode_block_strided_##SUF##_1(z, (T*)p, s[0])                                      \
                        : zfp_decode_partial_block_strided_##SUF##_1(z, (T*)p, ext[0], s[0]);                     \
    case 2: return full ? zfp_decode_block_strided_##SUF##_2(z, (T*)p, s[0], s[1])                                \
                        : zfp_decode_partial_block_strided_##SUF##_2(z, (T*)p, ext[0], ext[1], s[0], s[1]);       \
    case 3: return full ? zfp_decode_block_strided_##SUF##_3(z, (T*)p, s[0], s[1], s[2])                          \
                        : zfp_decode_partial_block_strided_##SUF##_3(z, (T*)p, ext[0], ext[1], ext[2], s[0], s[1], s[2]); \
    default: return full ? zfp_decode_block_strided_##SUF##_4(z, (T*)p, s[0], s[1], s[2], s[3])                   \
                         : zfp_decode_partial_block_strided_##SUF##_4(z, (T*)p, ext[0], ext[1], ext[2], ext[3], s[0], s[1], s[2], s[3]); \
  }
  switch (type) {
    case zfp_type_int32: CALL(int32, int32)
    case zfp_type_int64: CALL(int64, int64)
    case zfp_type_float: CALL(float, float)
    default: CALL(double, double)
  }
#undef CALL
}

/* Decompress a fixed-rate stream of a contiguous nx x ny x nz x nw array (unused dimensions 0) with `threads`
 * OpenMP threads.  Returns the number of bits consumed (blocks * maxbits), 0 on a parameter error. */
uint64_t zfp_ref_parallel_decompress(void* words, size_t bytes, uint64_t start_bit, int type, const size_t n[4], unsigned maxbits,
                                     void* out, int threads)
{
  unsigned dims = n[3] ? 4 : n[2] ? 3 : n[1] ? 2 : n[0] ? 1 : 0;
  size_t nb[4] = { 1, 1, 1, 1 }, esize = (type == zfp_type_int32 || type == zfp_type_float) ? 4 : 8;
  ptrdiff_t s[4] = { 1, 0, 0, 0 };
  uint64_t blocks = 1;
  unsigned d;
  int64_t b;
  if (!dims || !maxbits) return 0;
  for (d = 0; d < dims; d++) {
    nb[d] = (n[d] + 3) / 4;
    blocks *= nb[d];
    if (d) s[d] = s[d - 1] * (ptrdiff_t)n[d - 1];
  }
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
  {
    bitstream* bs = stream_open(words, bytes);
    zfp_stream* z = zfp_stream_open(bs);
    zfp_stream_set_params(z, maxbits, maxbits, ZFP_MAX_PREC, ZFP_MIN_EXP);
#pragma omp for schedule(static)
    for (b = 0; b < (int64_t)blocks; b++) {
      size_t c[4], ext[4] = { 1, 1, 1, 1 }, r = (size_t)b;
      ptrdiff_t off = 0;
      int full = 1;
      for (d = 0; d < dims; d++) {
        c[d] = r % nb[d];
        r /= nb[d];
        ext[d] = n[d] - 4 * c[d] < 4 ? n[d] - 4 * c[d] : 4;
        full = full && ext[d] == 4;
        off += s[d] * (ptrdiff_t)(4 * c[d]);
      }
      stream_rseek(bs, start_bit + (uint64_t)b * maxbits);
      decode_one(z, (zfp_type)type, dims, (char*)out + off * (ptrdiff_t)esize, ext, s, full);
    }
    zfp_stream_close(z);
    stream_close(bs);
  }
  return blocks * (uint64_t)maxbits;
}
