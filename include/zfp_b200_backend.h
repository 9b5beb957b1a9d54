/*
 * zfp_b200_backend.h - the C-ABI boundary between zfp's host dispatch and the sm_100a kernels.
 *
 * Plain C, no CUDA or torch types in any signature (a cudaStream_t travels as void*).
 *
 * (1) Drop-in symbols.  The reference's CUDA shims call exactly two functions
 *         size_t cuda_compress(zfp_stream*, const zfp_field*);     src/cuda_zfp/cuZFP.h:9
 *         void   cuda_decompress(zfp_stream*, zfp_field*);          src/cuda_zfp/cuZFP.h:10
 *     from src/template/cudacompress.c:5-34 and cudadecompress.c:5-34.  libzfp_b200 exports both
 *     with the same signatures and the same post-conditions on the host `bitstream`
 *     (src/cuda_zfp/cuZFP.cu:406-411, 486-490), so a reference libzfp built with
 *     -DZFP_WITH_CUDA links against this library instead of src/cuda_zfp.
 *     Unlike the code they replace they accept every mode, 1-4 D, a non-zero stream offset.
 *
 * (2) Raw entry points on plain pointers and sizes, used by bindings (ctypes/cgo/JNI style) and
 *     by the multi-GPU slab driver: one call = one slab.
 *
 * (3) The block-offset index that makes variable-rate streams decodable in parallel
 *     (the zfp format itself stores none: docs/source/execution.rst:292-300).
 */
#ifndef ZFP_B200_BACKEND_H
#define ZFP_B200_BACKEND_H

#include "zfp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) drop-in replacements for src/cuda_zfp/cuZFP.h:9-10 ------------------------------- */
size_t cuda_compress(zfp_stream* stream, const zfp_field* field);
void cuda_decompress(zfp_stream* stream, zfp_field* field);
/* the same two operations with an explicit status: bytes from the start of the stream, 0 = failed */
size_t zfp_b200_compress_stream(zfp_stream* stream, const zfp_field* field);
size_t zfp_b200_decompress_stream(zfp_stream* stream, zfp_field* field);

/* ---- execution parameters hung off zfp_stream.exec.params for zfp_exec_cuda ----------------- */
typedef struct zfp_b200_index zfp_b200_index; /* opaque; device-resident */

typedef struct {
  uint64 magic;          /* ZFP_B200_PARAMS_MAGIC; anything else is ignored */
  void* cuda_stream;     /* cudaStream_t to launch on; NULL = legacy default stream */
  int device_only_sync;  /* 0: calls return after the work is complete (reference semantics);
                            1: fixed-rate calls return as soon as the work is enqueued */
  zfp_b200_index* index; /* produced by the last variable-rate compress on this zfp_stream, consumed
                            by decompress; owned by the zfp_stream */
} zfp_exec_params_cuda;

#define ZFP_B200_PARAMS_MAGIC 0x7a66704232303021ull

/* get (creating on demand) the CUDA execution parameters of a stream whose policy is
 * zfp_exec_cuda; NULL otherwise */
zfp_exec_params_cuda* zfp_stream_cuda_params(zfp_stream* zfp);

/* ---- (2) raw slab entry points ----------------------------------------------------------------- */
typedef struct {
  int type;        /* zfp_type */
  uint dims;       /* 1..4 */
  size_t n[4];     /* nx, ny, nz, nw */
  ptrdiff_t s[4];  /* element strides, 0 = contiguous default */
  uint minbits, maxbits, maxprec;
  int minexp;
} zfp_b200_desc;

enum {
  ZFP_B200_OK = 0,
  ZFP_B200_EINVAL = 1,     /* bad type / dims / parameters */
  ZFP_B200_ECUDA = 2,      /* a CUDA call failed; see zfp_b200_last_error() */
  ZFP_B200_ENOINDEX = 3    /* variable-rate decode of a stream without a usable index */
};

/* Encode the field at device pointer d_data into device words d_words starting at bit
 * start_bit (bits below start_bit in the first word are preserved).  *end_bit receives the bit
 * offset one past the last block; words beyond it up to the next word boundary are zero, as after
 * stream_flush.  For variable-rate parameters, if index != NULL it is filled with the block
 * lengths.  All work is enqueued on cuda_stream; the call synchronises that stream only when it
 * must read a size back (variable rate). */
int zfp_b200_encode(const zfp_b200_desc* desc, const void* d_data, void* d_words, uint64 start_bit,
                    uint64* end_bit, zfp_b200_index* index, void* cuda_stream);

/* Stream-ordered variant: the end position goes to DEVICE memory (*d_end_bit) and the call never
 * synchronises cuda_stream, also for variable-rate parameters.  This is what the multi-GPU slab path
 * uses: encode, exchange the slab lengths (ncclAllGather on the same stream) and place the slab
 * (zfp_b200_bitcopy_ranked) without a host round trip. */
int zfp_b200_encode_async(const zfp_b200_desc* desc, const void* d_data, void* d_words, uint64 start_bit,
                          uint64* d_end_bit, zfp_b200_index* index, void* cuda_stream);

/* Decode; mirror of the above.  index may be NULL for fixed-rate parameters.  For variable-rate
 * parameters a NULL index makes the backend rebuild one by scanning the stream (slow, sequential
 * in the stream order by the nature of the format). */
int zfp_b200_decode(const zfp_b200_desc* desc, void* d_data, const void* d_words, uint64 start_bit,
                    uint64* end_bit, const zfp_b200_index* index, void* cuda_stream);

/* Stream-ordered decode: never synchronises cuda_stream.  Variable-rate parameters need the index of
 * this stream; *d_status (device memory, zeroed by the caller) becomes non-zero if some block does not parse
 * to the length the index records for it - the decoded array is then not to be trusted. */
int zfp_b200_decode_async(const zfp_b200_desc* desc, void* d_data, const void* d_words, uint64 start_bit,
                          const zfp_b200_index* index, unsigned int* d_status, void* cuda_stream);

/* Random access: decode only blocks [block0, block1) of the stream (stream order, x fastest:
 * b = bx + BX*(by + BY*(bz + BZ*bw)), src/template/ompcompress.c:168-198) into their places in the
 * field at d_data; values of other blocks are left untouched.  Fixed rate needs no index (block b
 * starts at start_bit + b*maxbits); variable rate uses the block-offset index (NULL = rebuild it by
 * scanning the stream).  This is the parallel form of what the reference offers only through its
 * C++ compressed-array classes (include/zfp/index.hpp:160-530). */
int zfp_b200_decode_blocks(const zfp_b200_desc* desc, void* d_data, const void* d_words, uint64 start_bit,
                           uint64 block0, uint64 block1, const zfp_b200_index* index, void* cuda_stream);

/* Random access by coordinates: decode, in ONE launch, every block that intersects the box
 * lo[i] <= index_i < hi[i] (i = 0 is x, like desc->n; entries beyond desc->dims are ignored) into its place in
 * the field at d_data.  Blocks are decoded whole, so up to 3 positions outside the box along each dimension are
 * written too.  (The reference offers this only one block at a time through its C++ array classes,
 * include/zfp/index.hpp:160-315.) */
int zfp_b200_decode_box(const zfp_b200_desc* desc, void* d_data, const void* d_words, uint64 start_bit,
                        const size_t* lo, const size_t* hi, const zfp_b200_index* index, void* cuda_stream);

/* Bit-granular device copy dst[dst_bit, dst_bit+nbits) = src[src_bit, ...): places a slab stream
 * produced at another bit phase / on another GPU into a global stream.  Destination words fully
 * inside the range are overwritten, partially covered ones OR-merged (their target bits must be 0). */
int zfp_b200_bitcopy(void* d_dst_words, uint64 dst_bit, const void* d_src_words, uint64 src_bit, uint64 nbits,
                     void* cuda_stream);

/* Placement of slab `rank` when the slab lengths are still on the device: d_lengths[r] = bits of slab r
 * (all-gathered on cuda_stream).  The slab stream at bit 0 of d_src_words goes to
 * start_bit + d_lengths[0] + ... + d_lengths[rank-1] of d_dst_words (same merge rules as
 * zfp_b200_bitcopy); that position is also stored to *d_base_out when given.  d_dst_words == NULL only
 * computes the position. */
int zfp_b200_bitcopy_ranked(void* d_dst_words, uint64 start_bit, const uint64* d_lengths, uint rank,
                            const void* d_src_words, uint64* d_base_out, void* cuda_stream);

/* 1 if the parameters make every block the same size (minbits == maxbits) */
int zfp_b200_is_fixed_rate(const zfp_b200_desc* desc);
/* number of 4^d blocks in the field */
size_t zfp_b200_blocks(const zfp_b200_desc* desc);
/* capacity in bytes a caller must provide for d_words (zfp_stream_maximum_size formula,
 * src/zfp.c:711-742, plus the words covering start_bit) */
size_t zfp_b200_capacity(const zfp_b200_desc* desc, uint64 start_bit);

/* ---- (2b) several GPUs, one process ------------------------------------------------------------------
 * Slab i of the array (block-aligned along the slowest dimension, desc[i] = its extents + the common
 * parameters) lives on device devices[i].  zfp_b200_multi_compress encodes all slabs concurrently, each at
 * bit 0 of d_words[i] on its own device and stream, exchanges the slab bit lengths with ncclAllGather
 * enqueued on those streams and derives each slab's base bit in the global stream on the device; the host
 * synchronises once, at the end, to read the lengths and bases back (either array may be NULL).  Fixed-rate
 * slabs need no exchange but go through the same call.  NCCL (ncclCommInitAll) is loaded at run time. */
typedef struct zfp_b200_multi zfp_b200_multi;
zfp_b200_multi* zfp_b200_multi_create(int ndev, const int* devices);
void zfp_b200_multi_destroy(zfp_b200_multi* m);
int zfp_b200_multi_devices(const zfp_b200_multi* m);
void* zfp_b200_multi_stream(const zfp_b200_multi* m, int i); /* the cudaStream_t used on device i */
int zfp_b200_multi_compress(zfp_b200_multi* m, const zfp_b200_desc* descs, const void* const* d_slabs,
                            void* const* d_words, zfp_b200_index* const* indexes, uint64* slab_bits, uint64* slab_base);
int zfp_b200_multi_decompress(zfp_b200_multi* m, const zfp_b200_desc* descs, void* const* d_slabs,
                              const void* const* d_words, zfp_b200_index* const* indexes);

/* ---- (3) block-offset index -------------------------------------------------------------------- */
zfp_b200_index* zfp_b200_index_create(void);
void zfp_b200_index_destroy(zfp_b200_index* index);
size_t zfp_b200_index_blocks(const zfp_b200_index* index);
/* total coded bits of the blocks (sum of the lengths) as recorded by the encode that filled the index; 0 after an import */
uint64 zfp_b200_index_bits(const zfp_b200_index* index);
/* serialised form: 16-bit coded length per block, block order = stream order */
size_t zfp_b200_index_export(const zfp_b200_index* index, uint16_t* host_lengths, size_t capacity);
/* Rebuild the index of a variable-rate stream that arrived without one (a file written by another zfp), in parallel:
 * the stream in [d_words, d_words + words_bytes) is cut into segments that are parsed speculatively and stitched
 * (a parse that starts off a block boundary falls onto the true chain of blocks after a few hundred blocks).  The
 * result is treated like an imported index: the decode checks every length and falls back to walking the stream
 * sequentially (~6 us per block) if it does not hold.  ZFP_B200_EINVAL when not applicable (fixed rate, 4-D, fewer
 * than 16384 blocks or less than 4 Mbit of stream): decode with index == NULL then.  zfp_decompress does this by itself for host-API callers, whose
 * bitstream knows where its buffer ends. */
int zfp_b200_index_rebuild(const zfp_b200_desc* desc, const void* d_words, uint64 start_bit, size_t words_bytes,
                           zfp_b200_index* index, void* cuda_stream);
int zfp_b200_index_import(zfp_b200_index* index, const uint16_t* host_lengths, size_t blocks);

/* ---- diagnostics --------------------------------------------------------------------------------- */
const char* zfp_b200_last_error(void);
/* number of kernels this library launched since load (bench.py's gpu_launches counter) */
uint64 zfp_b200_launch_count(void);
/* release cached device scratch */
void zfp_b200_release_scratch(void);

#ifdef __cplusplus
}
#endif

#endif /* ZFP_B200_BACKEND_H */
