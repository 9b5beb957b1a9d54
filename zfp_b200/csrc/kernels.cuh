// kernels.cuh - array-level kernels: block gather/scatter, encode/decode launch bodies,
// block-length scan, bit-granular stream compaction.
//
// Stream order is block order with x fastest (src/template/compress.c:58-109,
// ompcompress.c:168-198): block b = bx + BX*(by + BY*(bz + BZ*bw)).
#pragma once

#include <type_traits>

#include "codec.cuh"

namespace zb {

constexpr int kThreads = 128;  // threads per CTA for the codec kernels (4 warps)

// ------------------------------------------------------------------------------------------------
// block <-> array
// ------------------------------------------------------------------------------------------------
template <int DIMS>
struct BlockPos {
  int64_t offset;    // element offset of the block's first value
  uint32_t ext[3];   // valid extent (1..4) per dimension
  bool full;
};

template <int DIMS>
__device__ __forceinline__ BlockPos<DIMS> locate(const Geom& g, uint64_t b)
{
  BlockPos<DIMS> p;
  p.offset = 0;
  p.full = true;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (d < DIMS) {
      uint64_t q = d + 1 < DIMS ? b / g.nb[d] : 0;
      uint64_t c = d + 1 < DIMS ? b - q * g.nb[d] : b;
      b = q;
      uint64_t org = 4 * c, left = g.n[d] - org;
      p.ext[d] = left < 4 ? (uint32_t)left : 4u;
      p.full &= left >= 4;
      p.offset += g.s[d] * (int64_t)org;
    }
    else
      p.ext[d] = 1;
  }
  return p;
}

template <class T> struct Vec4;  // four consecutive scalars, moved with 128-bit accesses
template <> struct Vec4<float> {
  __device__ static __forceinline__ void load(const float* p, float& a, float& b, float& c, float& d)
  { float4 t = __ldg(reinterpret_cast<const float4*>(p)); a = t.x; b = t.y; c = t.z; d = t.w; }
  __device__ static __forceinline__ void store(float* p, float a, float b, float c, float d)
  { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
};
template <> struct Vec4<int32_t> {
  __device__ static __forceinline__ void load(const int32_t* p, int32_t& a, int32_t& b, int32_t& c, int32_t& d)
  { int4 t = __ldg(reinterpret_cast<const int4*>(p)); a = t.x; b = t.y; c = t.z; d = t.w; }
  __device__ static __forceinline__ void store(int32_t* p, int32_t a, int32_t b, int32_t c, int32_t d)
  { *reinterpret_cast<int4*>(p) = make_int4(a, b, c, d); }
};
template <> struct Vec4<double> {
  __device__ static __forceinline__ void load(const double* p, double& a, double& b, double& c, double& d)
  {
    double2 t = __ldg(reinterpret_cast<const double2*>(p)), u = __ldg(reinterpret_cast<const double2*>(p) + 1);
    a = t.x; b = t.y; c = u.x; d = u.y;
  }
  __device__ static __forceinline__ void store(double* p, double a, double b, double c, double d)
  {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
  }
};
template <> struct Vec4<int64_t> {
  __device__ static __forceinline__ void load(const int64_t* p, int64_t& a, int64_t& b, int64_t& c, int64_t& d)
  {
    longlong2 t = __ldg(reinterpret_cast<const longlong2*>(p)), u = __ldg(reinterpret_cast<const longlong2*>(p) + 1);
    a = t.x; b = t.y; c = u.x; d = u.y;
  }
  __device__ static __forceinline__ void store(int64_t* p, int64_t a, int64_t b, int64_t c, int64_t d)
  {
    reinterpret_cast<longlong2*>(p)[0] = make_longlong2(a, b);
    reinterpret_cast<longlong2*>(p)[1] = make_longlong2(c, d);
  }
};

// partial-block padding along one axis (src/template/encode.c:8-27):
// 1 valid -> (a,a,a,a), 2 -> (a,b,b,a), 3 -> (a,b,c,a)
template <class T>
__device__ __forceinline__ void pad4(T& a, T& b, T& c, T& d, uint32_t m)
{
  b = m < 2 ? a : b;
  c = m < 3 ? b : c;
  d = m < 4 ? a : d;
}

template <int DIMS, class Scalar>
__device__ __forceinline__ void gather(Scalar (&v)[1 << (2 * DIMS)], const Scalar* data, const Geom& g,
                                       const BlockPos<DIMS>& pos)
{
  constexpr int N = 1 << (2 * DIMS);
  const Scalar* p = data + pos.offset;
  if (pos.full && g.vec_rows) {
#pragma unroll
    for (int r = 0; r < N / 4; r++) {
      int64_t o = (DIMS > 1 ? g.s[1] * (r & 3) : 0) + (DIMS > 2 ? g.s[2] * (r >> 2) : 0);
      Vec4<Scalar>::load(p + o, v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t x = i & 3, y = (i >> 2) & 3, z = (i >> 4) & 3;
    bool ok = x < pos.ext[0] && (DIMS < 2 || y < pos.ext[1]) && (DIMS < 3 || z < pos.ext[2]);
    int64_t o = g.s[0] * x + (DIMS > 1 ? g.s[1] * y : 0) + (DIMS > 2 ? g.s[2] * z : 0);
    v[i] = ok ? __ldg(p + o) : Scalar(0);
  }
  if (!pos.full) {
    // x lines, then y lines, then z lines; lines lying outside the valid region are rewritten by
    // the later passes, which reproduces the nesting of gather_partial (encode3.c:17-31)
#pragma unroll
    for (int l = 0; l < N / 4; l++)
      pad4(v[4 * l], v[4 * l + 1], v[4 * l + 2], v[4 * l + 3], pos.ext[0]);
    if (DIMS > 1) {
#pragma unroll
      for (int i = 0; i < N; i++)
        if (((i >> 2) & 3) == 0)
          pad4(v[i], v[(i + 4) % N], v[(i + 8) % N], v[(i + 12) % N], pos.ext[1]);
    }
    if (DIMS > 2) {
#pragma unroll
      for (int i = 0; i < N; i++)
        if (((i >> 4) & 3) == 0)
          pad4(v[i], v[(i + 16) % N], v[(i + 32) % N], v[(i + 48) % N], pos.ext[2]);
    }
  }
}

template <int DIMS, class Scalar>
__device__ __forceinline__ void scatter(const Scalar (&v)[1 << (2 * DIMS)], Scalar* data, const Geom& g,
                                        const BlockPos<DIMS>& pos)
{
  constexpr int N = 1 << (2 * DIMS);
  Scalar* p = data + pos.offset;
  if (pos.full && g.vec_rows) {
#pragma unroll
    for (int r = 0; r < N / 4; r++) {
      int64_t o = (DIMS > 1 ? g.s[1] * (r & 3) : 0) + (DIMS > 2 ? g.s[2] * (r >> 2) : 0);
      Vec4<Scalar>::store(p + o, v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t x = i & 3, y = (i >> 2) & 3, z = (i >> 4) & 3;
    bool ok = x < pos.ext[0] && (DIMS < 2 || y < pos.ext[1]) && (DIMS < 3 || z < pos.ext[2]);
    int64_t o = g.s[0] * x + (DIMS > 1 ? g.s[1] * y : 0) + (DIMS > 2 ? g.s[2] * z : 0);
    if (ok)
      p[o] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// codec kernels.  OUT: 0 fixed rate, word-aligned blocks (plain stores)
//                      1 fixed rate, blocks share words (OR-merge into a zeroed destination)
//                      2 variable rate: each block goes to its own scratch slot, length recorded
// ------------------------------------------------------------------------------------------------
template <int TYPE, int DIMS, int OUT>
__global__ void __launch_bounds__(kThreads)
encode_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm,
              void* __restrict__ out, uint64_t start_bit, uint32_t slot_words,
              uint16_t* __restrict__ lengths, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  PW* sp = reinterpret_cast<PW*>(smem_raw) + (threadIdx.x >> 5) * (TR::P * 32) + (threadIdx.x & 31);

  const uint64_t b = block0 + (uint64_t)blockIdx.x * kThreads + threadIdx.x;
  if (b >= block1)
    return;
  const BlockPos<DIMS> pos = locate<DIMS>(g, b);
  typename TR::Scalar v[N];
  gather<DIMS>(v, data, g, pos);

  BitWriter<OUT == 1 ? 1 : 0> bw;
  if (OUT == 2)
    bw.init(out, (b - block0) * (uint64_t)slot_words * 64);
  else
    bw.init(out, start_bit + b * (uint64_t)prm.maxbits);
  uint32_t bits = encode_block<TYPE, DIMS>(v, prm, bw, sp);
  bw.flush();
  if (OUT == 2)
    lengths[b] = (uint16_t)bits;
}

// OFFS: 0 fixed rate (offset = start + b*maxbits), 1 per-block offsets from the index scan
template <int TYPE, int DIMS, int OFFS>
__global__ void __launch_bounds__(kThreads)
decode_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm,
              const void* __restrict__ in, uint64_t start_bit, const uint64_t* __restrict__ offsets)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  PW* sp = reinterpret_cast<PW*>(smem_raw) + (threadIdx.x >> 5) * (TR::P * 32) + (threadIdx.x & 31);

  const uint64_t b = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
  if (b >= g.nblocks)
    return;
  BitReader br;
  br.init(in, OFFS ? offsets[b] : start_bit + b * (uint64_t)prm.maxbits);
  typename TR::Scalar v[N];
  decode_block<TYPE, DIMS>(v, prm, br, sp);
  const BlockPos<DIMS> pos = locate<DIMS>(g, b);
  scatter<DIMS>(v, data, g, pos);
}

// ------------------------------------------------------------------------------------------------
// block-length scan (exclusive prefix of 16-bit lengths into 64-bit bit offsets)
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanTile = kScanThreads * kScanPerThread;  // 4096 blocks per tile

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// exclusive scan across the CTA of one uint32 per thread; returns the exclusive prefix and the total
__device__ __forceinline__ uint32_t cta_excl_scan(uint32_t v, uint32_t& total)
{
  __shared__ uint32_t ws[33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const uint32_t incl = warp_incl_scan(v);
  __syncthreads();  // readers of the previous call are done with ws
  if (lane == 31) ws[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = lane < nw ? ws[lane] : 0;
    uint32_t si = warp_incl_scan(s);
    ws[lane] = si - s;
    if (lane == 31) ws[32] = si;
  }
  __syncthreads();
  total = ws[32];
  return ws[wid] + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums(const uint16_t* __restrict__ lengths, uint64_t n, uint64_t* __restrict__ tile_sum)
{
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; i++)
    s += base + i < n ? lengths[base + i] : 0u;
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
    // tile sums fit 32 bits (4096 * 16658), their running total does not: scan 32-bit, carry 64-bit
    uint32_t total;
    uint32_t excl = cta_excl_scan((uint32_t)v, total);
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
           uint64_t* __restrict__ offsets)
{
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPerThread;
  uint32_t len[kScanPerThread], s = 0;
#pragma unroll
  for (int i = 0; i < kScanPerThread; i++) {
    len[i] = base + i < n ? lengths[base + i] : 0u;
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
// offset in the stream (stream_copy semantics, include/zfp/bitstream.inl:412-424)
__global__ void __launch_bounds__(256)
compact_blocks(const uint64_t* __restrict__ scratch, uint32_t slot_words, const uint16_t* __restrict__ lengths,
               const uint64_t* __restrict__ offsets, uint64_t nblocks, void* __restrict__ out)
{
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks)
    return;
  const uint64_t* src = scratch + b * slot_words;
  uint32_t len = lengths[b];
  BitWriter<1> bw;
  bw.init(out, offsets[b]);
  for (uint32_t i = 0; len; i++) {
    uint32_t c = len < 64 ? len : 64;
    bw.put(src[i] & lowmask64(c), c);
    len -= c;
  }
  bw.flush();
}

}  // namespace zb
