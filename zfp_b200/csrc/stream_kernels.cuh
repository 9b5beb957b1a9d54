// stream_kernels.cuh - block-length scan and bit-granular stream assembly (variable-rate modes).
#pragma once

#include "codec.cuh"
#include "scan_util.cuh"

namespace zb {

// ------------------------------------------------------------------------------------------------
// block-length scan (exclusive prefix of 16-bit lengths into 64-bit bit offsets)
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanTile = kScanThreads * kScanPerThread;  // 4096 blocks per tile

// 64-bit flavour for the tile-level scan: 1024 tiles of up to 4096 * 16658 bits overflow 32 bits
__device__ __forceinline__ uint64_t cta_excl_scan64(uint64_t v, uint64_t& total)
{
  __shared__ uint64_t ws[33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  uint64_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  __syncthreads();
  if (lane == 31) ws[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint64_t s = lane < nw ? ws[lane] : 0, si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) si += t;
    }
    ws[lane] = si - s;
    if (lane == 31) ws[32] = si;
  }
  __syncthreads();
  total = ws[32];
  return ws[wid] + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums(const uint16_t* __restrict__ lengths, uint64_t n, uint64_t* __restrict__ tile_sum, uint32_t cap)
{
  // (cap: no block is longer than the worst case the buffer was sized for - an index that came from outside,
  // zfp_b200_index_import, cannot steer the decoder past the stream; the decode then reports the mismatch)
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; i++)
    s += base + i < n ? min((uint32_t)lengths[base + i], cap) : 0u;
  uint32_t total;
  cta_excl_scan(s, total);
  if (threadIdx.x == 0)
    tile_sum[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile sums, starting at cursor[1]; cursor <- {begin, end}
__global__ void __launch_bounds__(1024)
scan_tile_offsets(uint64_t* __restrict__ tile_sum, uint64_t ntiles, uint64_t* __restrict__ cursor)
{
  __shared__ uint64_t carry;
  if (threadIdx.x == 0) { carry = cursor[1]; cursor[0] = cursor[1]; }
  __syncthreads();
  for (uint64_t t0 = 0; t0 < ntiles; t0 += blockDim.x) {
    uint64_t i = t0 + threadIdx.x;
    uint64_t v = i < ntiles ? tile_sum[i] : 0;
    uint64_t total;
    uint64_t excl = cta_excl_scan64(v, total);
    if (i < ntiles)
      tile_sum[i] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) cursor[1] = carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply(const uint16_t* __restrict__ lengths, uint64_t n, const uint64_t* __restrict__ tile_off,
           uint64_t* __restrict__ offsets, uint32_t cap)
{
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
  uint32_t len[kScanPerThread], s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; i++) {
    len[i] = base + i < n ? min((uint32_t)lengths[base + i], cap) : 0u;
    s += len[i];
  }
  uint32_t total;
  uint64_t o = tile_off[blockIdx.x] + cta_excl_scan(s, total);
#pragma unroll
  for (int i = 0; i < kScanPerThread; i++) {
    if (base + i < n) offsets[base + i] = o;
    o += len[i];
  }
}

// ------------------------------------------------------------------------------------------------
// stream assembly helpers
// ------------------------------------------------------------------------------------------------

// clear bits >= (bit % 64) of the word holding `bit` (keeps a header written before the payload)
__global__ void clear_word_tail(uint64_t* words, uint64_t bit)
{
  if (bit & 63)
    words[bit >> 6] &= (1ull << (bit & 63)) - 1;
}

// zero the words that start inside [cursor[0], cursor[1]) (device-side bounds)
__global__ void zero_new_words(uint64_t* __restrict__ words, const uint64_t* __restrict__ cursor)
{
  const uint64_t w0 = (cursor[0] + 63) >> 6, w1 = (cursor[1] + 63) >> 6;
  for (uint64_t i = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < w1; i += (uint64_t)gridDim.x * blockDim.x)
    words[i] = 0;
}

// concatenate the coded blocks of a chunk: block b's bits move from its scratch slot to its bit
// offset in the stream (stream_copy semantics, include/zfp/bitstream.inl:412-424).  Eight lanes per
// block, one destination word per lane and step: the slot is read with neighbouring lanes on
// neighbouring words, words inside the block go out with plain stores and only the (at most two)
// words shared with the neighbouring blocks are OR-merged into the pre-zeroed destination.
constexpr int kCompactLanes = 8;
__global__ void __launch_bounds__(256)
compact_blocks(const uint64_t* __restrict__ scratch, uint32_t slot_words, const uint16_t* __restrict__ lengths,
               const uint64_t* __restrict__ offsets, uint64_t nblocks, void* __restrict__ out)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b = t / kCompactLanes;
  if (b >= nblocks)
    return;
  const uint32_t lane = (uint32_t)(t % kCompactLanes);
  const uint64_t* src = scratch + b * slot_words;
  const uint64_t len = lengths[b], off = offsets[b];
  uint64_t* dst = static_cast<uint64_t*>(out);
  const uint64_t w0 = off >> 6, w1 = (off + len + 63) >> 6;  // destination words [w0, w1)
  for (uint64_t w = w0 + lane; w < w1; w += kCompactLanes) {
    const uint64_t lo = w * 64 > off ? w * 64 : off;
    const uint64_t hi = (w + 1) * 64 < off + len ? (w + 1) * 64 : off + len;
    const uint32_t n = (uint32_t)(hi - lo);       // bits of this block in word w
    const uint64_t sbit = lo - off;                // first source bit
    const uint32_t sh = (uint32_t)(sbit & 63);
    uint64_t v = src[sbit >> 6] >> sh;
    if (sh && sh + n > 64)
      v |= src[(sbit >> 6) + 1] << (64 - sh);
    v &= lowmask64(n);
    v <<= (uint32_t)(lo & 63);
    if (n == 64)
      dst[w] = v;
    else if (v)
      atomicOr(reinterpret_cast<unsigned long long*>(dst + w), (unsigned long long)v);
  }
}

// Bit-granular device copy: dst[dst_bit, dst_bit+nbits) = src[src_bit, src_bit+nbits).  Words of dst
// fully inside the range are overwritten; the (at most two) partially covered words are OR-merged,
// so their target bits must be zero beforehand.  Used to place slab streams produced on different
// GPUs (or at a different bit phase) into one stream.
__global__ void __launch_bounds__(256)
bitcopy_kernel(uint64_t* __restrict__ dst, uint64_t dst_bit, const uint64_t* __restrict__ src, uint64_t src_bit, uint64_t nbits)
{
  const uint64_t w0 = dst_bit >> 6, w1 = (dst_bit + nbits + 63) >> 6;  // destination words [w0, w1)
  for (uint64_t w = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < w1; w += (uint64_t)gridDim.x * blockDim.x) {
    // destination bit range covered by this word, clipped to the copy range
    const uint64_t lo = w * 64 > dst_bit ? w * 64 : dst_bit;
    const uint64_t hi = (w + 1) * 64 < dst_bit + nbits ? (w + 1) * 64 : dst_bit + nbits;
    const uint32_t len = (uint32_t)(hi - lo);
    const uint64_t s = src_bit + (lo - dst_bit);  // first source bit
    const uint32_t sh = (uint32_t)(s & 63);
    uint64_t v = src[s >> 6] >> sh;
    if (sh && sh + len > 64)
      v |= src[(s >> 6) + 1] << (64 - sh);
    v &= lowmask64(len);
    v <<= (uint32_t)(lo & 63);
    if (len == 64)
      dst[w] = v;
    else
      atomicOr(reinterpret_cast<unsigned long long*>(dst + w), (unsigned long long)v);
  }
}

}  // namespace zb
