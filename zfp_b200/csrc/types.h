// types.h - plain structs shared by the host dispatch and the kernel translation units.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace zb {

constexpr int kMinExp = -1074;  // ZFP_MIN_EXP

enum : int { T_INT32 = 1, T_INT64 = 2, T_FLOAT = 3, T_DOUBLE = 4 };  // zfp_type numbering

struct Geom {
  uint64_t n[4];    // extent per dimension (1 for unused)
  int64_t s[4];     // element strides (resolved, never 0)
  uint64_t nb[4];   // blocks per dimension
  uint64_t nblocks;
  int vec_rows;     // 1: sx == 1 and every 4-value row starts 16/32-byte aligned
  // decode of a coordinate box in one launch (zfp_b200_decode_box): when box != 0 the decode kernels walk the
  // list of blocks that intersect the box - be[0]*be[1]*be[2]*be[3] of them, x fastest - and map list position
  // i to block number sum_d (bl[d] + i_d) * prod_{e<d} nb[e]
  int box;
  uint32_t bl[4], be[4];
};

// list position inside the box -> block number of the array
__host__ __device__ inline uint64_t box_block(const Geom& g, uint64_t i)
{
  uint64_t b = 0, mul = 1;
  for (int d = 0; d < 4; d++) {
    const uint64_t q = i / g.be[d], c = i - q * g.be[d];
    b += (g.bl[d] + c) * mul;
    mul *= g.nb[d];
    i = q;
  }
  return b;
}

struct Params {
  uint32_t minbits, maxbits, maxprec;
  int32_t minexp;
};

// OUT: 0 fixed rate, word-aligned blocks (plain stores); 1 fixed rate, blocks share words
// (OR-merge into a zeroed destination); 2 variable rate (scratch slot per block + length)
struct EncodeArgs {
  const void* data;
  Geom g;
  Params prm;
  void* out;
  uint64_t start_bit;
  uint32_t slot_words;
  uint16_t* lengths;
  uint64_t b0, b1;  // block range [b0, b1)
  cudaStream_t st;
  int staged;       // OUT 0 only: use the shared-memory staged fast path
};

// OFFS: 0 fixed rate (offset = start + b*maxbits), 1 per-block offsets from the index scan
struct DecodeArgs {
  void* data;
  Geom g;
  Params prm;
  const void* in;
  uint64_t start_bit;
  const uint64_t* offsets;
  const uint16_t* lengths;  // OFFS 1: coded length of every block (bounds the staged reads)
  cudaStream_t st;
  int staged;       // OFFS 0 only: use the shared-memory staged fast path when the stream is word aligned
  uint64_t b0, b1;  // block range [b0, b1) to decode (random access); the whole field is [0, nblocks)
  uint32_t* check;  // OFFS 1: set to non-zero by any block whose parsed length differs from lengths[b] (nullptr: no check)
};

// single-pass variable-rate encode (kernels_var1.cuh): device buffers of one launch
struct Var1Bufs {
  void* status;                   // Var1Status[tiles], zeroed
  unsigned int* ticket;           // zeroed
  unsigned long long* carry;      // {end position, last 64 bits} before this launch; updated by it
  void* overflow;                 // Var1Overflow[overflow_capacity]
  unsigned int* overflow_count;   // accumulates over the launches of one array
  unsigned int overflow_capacity;
  int sms;                        // multiprocessors (grid of the clean-up kernel)
  int cleanup;                    // 0: encode blocks [b0, b1) only; 1: only run the clean-up kernel
};

// one translation unit per scalar type and direction (inst_*.cu) defines these
template <int TYPE> cudaError_t launch_encode_t(int dims, int out_mode, const EncodeArgs& a);
template <int TYPE> cudaError_t launch_decode_t(int dims, int offs_mode, const DecodeArgs& a);
// 3-D, variable rate: encode + place blocks [a.b0, a.b1) in one pass, then re-encode the overflowing ones; returns
// the tile size through *tile_blocks when a == nullptr-like query is wanted (see backend.cu)
template <int TYPE> cudaError_t launch_encode_var1_t(const EncodeArgs& a, const Var1Bufs& v);
template <int TYPE> int var1_tile_blocks();
// sequential rebuild of the block-length index of a variable-rate stream: blocks a.b0 .. nblocks-1, the first of
// them at bit a.start_bit
template <int TYPE> cudaError_t launch_index_t(int dims, const DecodeArgs& a, uint16_t* lengths);

// speculative segment-parallel rebuild (kernels.cuh spec_index_kernel), 1-3 D
struct SpecIndexArgs {
  const void* in;
  uint64_t start_bit;    // first block of the stream
  uint64_t avail_bits;   // end of the buffer the stream lies in
  uint64_t seg_bits;     // segment t covers bits [start_bit + t * seg_bits, + seg_bits)
  uint32_t nseg;
  uint32_t margin_bits;  // a walk stops this far before avail_bits (worst-case block + a word)
  Params prm;
  uint64_t* exit;        // [nseg] where the walk of segment t left the segment = where segment t+1 is entered
  uint32_t* cnt;         // [nseg] blocks between the entry and the exit of segment t
  uint32_t* changed;     // pass 1: set when some exit moved
  const uint64_t* off;   // [nseg] pass 2: number of the first block of segment t
  uint16_t* lengths;     // pass 2
  uint64_t nblocks;
};
template <int TYPE> cudaError_t launch_spec_index_t(int dims, int pass, const SpecIndexArgs& a, cudaStream_t st);

}  // namespace zb
