// kernels.cuh - array-level kernels: block gather/scatter, encode/decode launch bodies,
// block-length scan, bit-granular stream compaction.
//
// Stream order is block order with x fastest (src/template/compress.c:58-109,
// ompcompress.c:168-198): block b = bx + BX*(by + BY*(bz + BZ*bw)).
#pragma once

#include <type_traits>

#include "codec.cuh"

namespace zb {

#ifndef ZB_THREADS
#define ZB_THREADS 64
#endif
constexpr int kThreads = ZB_THREADS;  // threads per CTA for the codec kernels (2 warps: more CTAs fit the shared-memory budget)

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
  const bool small = (g.nblocks >> 32) == 0;  // 32-bit division is several times cheaper
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (d < DIMS) {
      uint64_t q = 0, c = b;
      if (d + 1 < DIMS) {
        if (small) {
          const uint32_t q32 = (uint32_t)b / (uint32_t)g.nb[d];
          c = (uint32_t)b - q32 * (uint32_t)g.nb[d];
          q = q32;
        }
        else {
          q = b / g.nb[d];
          c = b - q * g.nb[d];
        }
      }
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

// L2 prefetch of the rows of a block another CTA will gather about one wave of CTAs later (full blocks with
// contiguous rows only): the gather then waits for an L2 hit instead of HBM
template <int DIMS, class Scalar>
__device__ __forceinline__ void prefetch_block(const Scalar* data, const Geom& g, uint64_t b)
{
  constexpr int N = 1 << (2 * DIMS);
  if (b >= g.nblocks || !g.vec_rows)
    return;
  const BlockPos<DIMS> pos = locate<DIMS>(g, b);
  if (!pos.full)
    return;
  const Scalar* p = data + pos.offset;
#pragma unroll
  for (int r = 0; r < N / 4; r++) {
    const int64_t o = (DIMS > 1 ? g.s[1] * (r & 3) : 0) + (DIMS > 2 ? g.s[2] * (r >> 2) : 0);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
  }
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
template <int TYPE, int DIMS, int OUT, bool REV>
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
  uint32_t bits = encode_block<TYPE, DIMS, REV>(v, prm, bw, sp);
  bw.flush();
  if (OUT == 2)
    lengths[b] = (uint16_t)bits;
}

// Fixed rate with word-aligned blocks, fast path: bits are staged per lane in shared memory
// (ColWriter, codec.cuh) and leave with 16-byte stores.  Dynamic shared memory per warp:
// planes (P*32 plane words) followed by (maxbits/32 + kStageSlack) * 32 staging words.
#ifndef ZB_MINBLOCKS64
#define ZB_MINBLOCKS64 5  // launch-bounds hint for the 64-bit staged kernels; 5 measured best of {4,5,6,8} on B200
#endif
// encode_staged_kernel / encode_var_kernel: CTA size and CTAs per SM by plane width.  The 64-bit kernels
// measured 3-4 % faster with 4 warps per CTA (512^3 fp64 rate 8: 0.605 -> 0.581 ms; 192 threads and
// CTA-wide rendezvous points gave nothing more); the 32-bit ones keep 2 warps.
#ifndef ZB_ENC64_THREADS
#define ZB_ENC64_THREADS 128
#endif
// Two experiments on the staged encode of 64-bit 3-D blocks, both bit-exact, neither faster (1024^3 fp64 rate 8,
// 4.31 ms without): an L2 prefetch of the rows another CTA gathers one wave later (4.57 ms), and table-driven
// plane steps while only coefficients 0..7 are significant (encode_planes_small8: 11 % fewer instructions, 4.39 ms).
#ifndef ZB_PREFETCH
#define ZB_PREFETCH 0
#endif
#ifndef ZB_SMALL8
#define ZB_SMALL8 0
#endif
#ifndef ZB_SMALL8_F32
#define ZB_SMALL8_F32 0  // the same steps for 3-D blocks of 32-bit values (experiment)
#endif
#ifndef ZB_ENC64_CTAS
#define ZB_ENC64_CTAS 3  // CTAs of 128 threads per SM the lossy 64-bit encode kernels are compiled for (register cap 168)
#endif
#ifndef ZB_REV32_CTAS
#define ZB_REV32_CTAS 6  // CTAs of 64 threads per SM the reversible 32-bit kernels are compiled for (register cap 170)
#endif
#ifndef ZB_ENC32_THREADS
#define ZB_ENC32_THREADS kThreads
#endif
#ifndef ZB_DEC32_THREADS
#define ZB_DEC32_THREADS kThreads
#endif
#ifndef ZB_ENC_SYNC
#define ZB_ENC_SYNC 0  // 1 with ZB_ENC64_THREADS = 256 / 384: warps on one scheduler rendezvous before the long stages
#endif
// kernels that carry the encoder's 8-coefficient table (encode_planes_small8): 2-D blocks; 3-D blocks by experiment switch
template <int TYPE, int N, bool REV>
constexpr bool kEncSmall8 = !REV && (N == 16 || (N == 64 && (Traits<TYPE>::P == 64 ? ZB_SMALL8 : ZB_SMALL8_F32)));
template <int TYPE> struct EncCfg {
  static constexpr int threads = Traits<TYPE>::P == 64 ? ZB_ENC64_THREADS : ZB_ENC32_THREADS;
  __host__ __device__ static constexpr int min_ctas(bool rev)
  {
    // 64-bit: 384 threads per SM (<= 168 registers per thread), as with 6 CTAs of 64 threads
    return Traits<TYPE>::P == 64 ? (rev ? 2 : ZB_ENC64_CTAS) * 128 / threads : (rev ? ZB_REV32_CTAS : 9) * 64 / threads;
  }
};
// the staged kernels' CTA shape: 2-D blocks of 32-bit values carry a 9-10 KB table per CTA (encode_planes_small8 /
// decode_pair_small8), so their CTAs are larger (8 warps share one copy; 4 CTAs per SM)
#ifndef ZB_STAGED32_2D_THREADS
#define ZB_STAGED32_2D_THREADS 256
#endif
template <int TYPE, int DIMS> struct SEncCfg {
  static constexpr bool wide = Traits<TYPE>::P == 32 && DIMS == 2;
  static constexpr int threads = wide ? ZB_STAGED32_2D_THREADS : EncCfg<TYPE>::threads;
  __host__ __device__ static constexpr int min_ctas(bool rev) { return wide ? 1024 / ZB_STAGED32_2D_THREADS : EncCfg<TYPE>::min_ctas(rev); }
};
constexpr int kStageSlack = 10;   // words of overshoot room: a plane may exceed the budget by < 200 bits and an append stores two words ahead
constexpr int kStagedPlanes = 32;  // plane words resident at a time in the lockstep kernels (a 32-plane half or a 16-plane window)

template <int TYPE, int DIMS, bool REV>
__global__ void __launch_bounds__(SEncCfg<TYPE, DIMS>::threads, SEncCfg<TYPE, DIMS>::min_ctas(REV))
encode_staged_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm,
                     uint64_t* __restrict__ out, uint64_t start_bit)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  const uint32_t words = prm.maxbits >> 5;  // 32-bit words per block (maxbits % 32 == 0)
  const uint32_t warp_bytes = kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (words + kStageSlack) * 32 * 4;
  // (byte offsets kept opaque: left alone, the compiler re-derives them from the thread index inside the
  // coder loops - a dozen instructions per iteration - rather than hold two registers across the transform)
  uint32_t sp_off = (threadIdx.x >> 5) * warp_bytes + (threadIdx.x & 31) * (uint32_t)sizeof(PW);
  uint32_t stage_off = (threadIdx.x >> 5) * warp_bytes + kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (threadIdx.x & 31) * 4u;
  asm volatile("" : "+r"(sp_off), "+r"(stage_off));
  PW* sp = reinterpret_cast<PW*>(reinterpret_cast<char*>(smem_raw) + sp_off);
  uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + stage_off);

  // no early exit: lanes past the end redo the last block and discard it, so that warp-wide votes
  // inside encode_block always see 32 lanes
  const uint64_t b_raw = (uint64_t)blockIdx.x * SEncCfg<TYPE, DIMS>::threads + threadIdx.x;
  const bool valid = b_raw < g.nblocks;
  const uint64_t b = valid ? b_raw : g.nblocks - 1;
  const BlockPos<DIMS> pos = locate<DIMS>(g, b);
  typename TR::Scalar v[N];
  gather<DIMS>(v, data, g, pos);
#if ZB_PREFETCH
  if constexpr (N == 64 && TR::P == 64) {
    uint32_t nsm;
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    constexpr int kWaveThreads = SEncCfg<TYPE, DIMS>::min_ctas(REV) * SEncCfg<TYPE, DIMS>::threads;  // per multiprocessor
    prefetch_block<DIMS>(data, g, b_raw + (uint64_t)nsm * kWaveThreads);
  }
#endif

  ColWriter bw;
  bw.init(stage);
  if constexpr (N == 4) {
    // the plane-string table of the 4-value blocks (kEncLut4), one copy per CTA behind the warps' buffers
    uint32_t* lut = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + (SEncCfg<TYPE, DIMS>::threads / 32) * warp_bytes);
    for (int i = threadIdx.x; i < kEncLut4Words; i += SEncCfg<TYPE, DIMS>::threads)
      lut[i] = kEncLut4[i];
    __syncthreads();
    bw.lut = (uint32_t)__cvta_generic_to_shared(lut);
  }
  if constexpr (kEncSmall8<TYPE, N, REV>) {
    // the table of the small-universe plane steps (encode_planes_small8), one copy per CTA
    uint32_t* lut = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + (SEncCfg<TYPE, DIMS>::threads / 32) * warp_bytes);  // (behind the warps' buffers)
    for (int i = threadIdx.x; i < kEncLut8Words / 4; i += SEncCfg<TYPE, DIMS>::threads)
      reinterpret_cast<uint4*>(lut)[i] = __ldg(reinterpret_cast<const uint4*>(kEncLut8) + i);
    __syncthreads();
    bw.lut = (uint32_t)__cvta_generic_to_shared(lut);
  }
#if ZB_ENC_SYNC
  encode_block<TYPE, DIMS, REV, ColWriter, (Traits<TYPE>::P == 64 && DIMS == 3 && !REV) ? SEncCfg<TYPE, DIMS>::threads / 4 : 0>(v, prm, bw, sp);
#else
  encode_block<TYPE, DIMS, REV>(v, prm, bw, sp);
#endif
  bw.finish(words);

  if (valid) {
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(out + (start_bit >> 6)) + b * (uint64_t)words;
    if ((words & 3) == 0 && (reinterpret_cast<uintptr_t>(dst32) & 15) == 0) {
      // 16-byte stores: a block's words are contiguous in the stream
      uint4* dst4 = reinterpret_cast<uint4*>(dst32);
#pragma unroll 4
      for (uint32_t w = 0; w < words; w += 4)
        dst4[w >> 2] = make_uint4(stage[w * 32], stage[(w + 1) * 32], stage[(w + 2) * 32], stage[(w + 3) * 32]);
    }
    else if ((words & 1) == 0) {
      uint64_t* dst = reinterpret_cast<uint64_t*>(dst32);
#pragma unroll 4
      for (uint32_t w = 0; w < words; w += 2)
        dst[w >> 1] = (uint64_t)stage[w * 32] | ((uint64_t)stage[(w + 1) * 32] << 32);
    }
    else {
      // an odd number of 32-bit words per block (e.g. 1-D at 8 bits/value): word stores, coalesced across lanes
      for (uint32_t w = 0; w < words; w++)
        dst32[w] = stage[w * 32];
    }
  }
}

// Fixed rate, word-aligned blocks, fast path: each lane first copies its block's words to a
// shared-memory column (all loads in flight at once instead of one dependent global load per
// word inside the serial decoder), then decodes from there (ColReader, codec.cuh).
// decode_staged_kernel: CTA size and CTAs per SM by plane width.  The 64-bit kernels run 6 warps
// per CTA and rendezvous once after the stream parse: their tail (inverse transposes, lifting, cast)
// is ~50 KB of straight-line code, far more than the 32 KB instruction cache level, and warps that
// walk it together share the fetches (measured 512^3 fp64 rate 8: 0.90 -> 0.81 ms).
#ifndef ZB_DEC64_THREADS
#define ZB_DEC64_THREADS 192
#endif
template <int TYPE> struct DecCfg {
  static constexpr int threads = Traits<TYPE>::P == 64 ? ZB_DEC64_THREADS : ZB_DEC32_THREADS;
  static constexpr int min_ctas(bool rev) { return Traits<TYPE>::P == 64 ? 384 / ZB_DEC64_THREADS : (rev ? ZB_REV32_CTAS : 9) * 64 / ZB_DEC32_THREADS; }
};
template <int TYPE, int DIMS> struct SDecCfg {
  static constexpr bool wide = Traits<TYPE>::P == 32 && DIMS == 2;
  static constexpr int threads = wide ? ZB_STAGED32_2D_THREADS : DecCfg<TYPE>::threads;
  __host__ __device__ static constexpr int min_ctas(bool rev) { return wide ? 1024 / ZB_STAGED32_2D_THREADS : DecCfg<TYPE>::min_ctas(rev); }
};
// table-driven plane steps while only coefficients 0..7 are significant (decode_pair_small8): 2-D blocks;
// 3-D blocks with -DZB_DSMALL8_3D=1 (experiment)
#ifndef ZB_DSMALL8_3D
#define ZB_DSMALL8_3D 0
#endif
template <int N, bool REV> constexpr bool kDecSmall8 = !REV && (N == 16 || (ZB_DSMALL8_3D && N == 64));
constexpr int kReadSlack = 5;  // zero words after the block: a plane's reads reach 64 + 32 bits past the position, rounded up to words

template <int TYPE, int DIMS, bool REV>
__global__ void __launch_bounds__(SDecCfg<TYPE, DIMS>::threads, SDecCfg<TYPE, DIMS>::min_ctas(REV))
decode_staged_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm,
                     const uint64_t* __restrict__ in, uint64_t start_bit, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  __shared__ uint32_t run_table[32];  // test-bit positions of the run decoder (ColReader::run_mask)
  if (threadIdx.x < 32)
    ColReader::fill_run_table(run_table, threadIdx.x);
  __syncthreads();
  const uint32_t words = prm.maxbits >> 5;
  const uint32_t warp_bytes = kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (words + kReadSlack) * 32 * 4;
  // (byte offsets kept opaque: left alone, the compiler re-derives them from the thread index inside the
  // coder loops - a dozen instructions per iteration - rather than hold two registers across the transform)
  uint32_t sp_off = (threadIdx.x >> 5) * warp_bytes + (threadIdx.x & 31) * (uint32_t)sizeof(PW);
  uint32_t stage_off = (threadIdx.x >> 5) * warp_bytes + kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (threadIdx.x & 31) * 4u;
  asm volatile("" : "+r"(sp_off), "+r"(stage_off));
  PW* sp = reinterpret_cast<PW*>(reinterpret_cast<char*>(smem_raw) + sp_off);
  uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + stage_off);

  const uint64_t b_raw = block0 + (uint64_t)blockIdx.x * SDecCfg<TYPE, DIMS>::threads + threadIdx.x;
  const bool valid = b_raw < block1;  // no early exit (warp-wide votes in decode_block)
  const uint64_t b_list = valid ? b_raw : block1 - 1;
  const uint64_t b = g.box ? box_block(g, b_list) : b_list;
  const uint32_t* src32 = reinterpret_cast<const uint32_t*>(in + (start_bit >> 6)) + b * (uint64_t)words;
  const uint64_t* src = reinterpret_cast<const uint64_t*>(src32);
  if ((words & 3) == 0 && (reinterpret_cast<uintptr_t>(src32) & 15) == 0) {
    const uint4* src4 = reinterpret_cast<const uint4*>(src32);
#pragma unroll 4
    for (uint32_t w = 0; w < words; w += 4) {
      const uint4 v = __ldg(src4 + (w >> 2));
      stage[w * 32] = v.x;
      stage[(w + 1) * 32] = v.y;
      stage[(w + 2) * 32] = v.z;
      stage[(w + 3) * 32] = v.w;
    }
  }
  else if ((words & 1) == 0) {
#pragma unroll 4
    for (uint32_t w = 0; w < words; w += 2) {
      const uint64_t v = __ldg(src + (w >> 1));
      stage[w * 32] = (uint32_t)v;
      stage[(w + 1) * 32] = (uint32_t)(v >> 32);
    }
  }
  else {
    for (uint32_t w = 0; w < words; w++)
      stage[w * 32] = __ldg(src32 + w);
  }
#pragma unroll
  for (int j = 0; j < kReadSlack; j++)
    stage[(words + j) * 32] = 0;

  ColReader br;
  br.init(stage);
  br.set_run_table(run_table);
  if constexpr (kDecSmall8<N, REV>) {
    // the table of the small-universe plane steps (decode_pair_small8), one copy per CTA behind the warps' buffers
    uint32_t* lut = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + (SDecCfg<TYPE, DIMS>::threads / 32) * warp_bytes);
    for (int i = threadIdx.x; i < kDecLut8hBytes / 16; i += SDecCfg<TYPE, DIMS>::threads)
      reinterpret_cast<uint4*>(lut)[i] = __ldg(reinterpret_cast<const uint4*>(kDecLut8h) + i);
    __syncthreads();
    br.lut8 = (uint32_t)__cvta_generic_to_shared(lut);
  }
  typename TR::Scalar v[N];
  decode_block<TYPE, DIMS, REV>(v, prm, br, sp);
  if (valid) {
    const BlockPos<DIMS> pos = locate<DIMS>(g, b);
    scatter<DIMS>(v, data, g, pos);
  }
}

// Variable rate, fast path.  Encode: same lockstep coder, but the block's words leave for its
// scratch slot (the column is a kVarStageWords-word window that is drained at plane boundaries when
// it runs low) and the coded length is recorded; a scan + compaction pass then places the blocks.
// Decode: the block starts at an arbitrary bit offset (from the index scan); a window of its words
// is staged in the column and slid forward at plane boundaries for long blocks.
constexpr int kVarStageWords = 64;  // 2048 bits: more than almost every variable-rate block

template <int TYPE, int DIMS, bool REV>
__global__ void __launch_bounds__(EncCfg<TYPE>::threads, EncCfg<TYPE>::min_ctas(REV))
encode_var_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm,
                  uint32_t* __restrict__ slots, uint32_t slot_words32, uint16_t* __restrict__ lengths,
                  uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  constexpr uint32_t warp_bytes = kStagedPlanes * 32 * (uint32_t)sizeof(PW) + kVarStageWords * 32 * 4;
  // (byte offsets kept opaque: left alone, the compiler re-derives them from the thread index inside the
  // coder loops - a dozen instructions per iteration - rather than hold two registers across the transform)
  uint32_t sp_off = (threadIdx.x >> 5) * warp_bytes + (threadIdx.x & 31) * (uint32_t)sizeof(PW);
  uint32_t stage_off = (threadIdx.x >> 5) * warp_bytes + kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (threadIdx.x & 31) * 4u;
  asm volatile("" : "+r"(sp_off), "+r"(stage_off));
  PW* sp = reinterpret_cast<PW*>(reinterpret_cast<char*>(smem_raw) + sp_off);
  uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + stage_off);

  const uint64_t b_raw = block0 + (uint64_t)blockIdx.x * EncCfg<TYPE>::threads + threadIdx.x;
  const bool valid = b_raw < block1;
  const uint64_t b = valid ? b_raw : block1 - 1;
  const BlockPos<DIMS> pos = locate<DIMS>(g, b);
  typename TR::Scalar v[N];
  gather<DIMS>(v, data, g, pos);

  ColWriter bw;
  // lanes past the end write to the last block's slot too; they carry identical bits
  bw.init(stage, slots + (b - block0) * (uint64_t)slot_words32, kVarStageWords);
  const uint32_t bits = encode_block<TYPE, DIMS, REV>(v, prm, bw, sp);
  bw.finish_slot();
  if (valid)
    lengths[b] = (uint16_t)bits;
}

template <int TYPE, int DIMS, bool REV>
__global__ void __launch_bounds__(DecCfg<TYPE>::threads, DecCfg<TYPE>::min_ctas(REV))
decode_var_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, const uint32_t* __restrict__ in,
                  const uint64_t* __restrict__ offsets, const uint16_t* __restrict__ lengths, uint64_t block0, uint64_t block1,
                  uint32_t* __restrict__ check)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  __shared__ uint32_t run_table[32];  // test-bit positions of the run decoder (ColReader::run_mask)
  if (threadIdx.x < 32)
    ColReader::fill_run_table(run_table, threadIdx.x);
  __syncthreads();
  constexpr uint32_t warp_bytes = kStagedPlanes * 32 * (uint32_t)sizeof(PW) + kVarStageWords * 32 * 4;
  // (byte offsets kept opaque: left alone, the compiler re-derives them from the thread index inside the
  // coder loops - a dozen instructions per iteration - rather than hold two registers across the transform)
  uint32_t sp_off = (threadIdx.x >> 5) * warp_bytes + (threadIdx.x & 31) * (uint32_t)sizeof(PW);
  uint32_t stage_off = (threadIdx.x >> 5) * warp_bytes + kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (threadIdx.x & 31) * 4u;
  asm volatile("" : "+r"(sp_off), "+r"(stage_off));
  PW* sp = reinterpret_cast<PW*>(reinterpret_cast<char*>(smem_raw) + sp_off);
  uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + stage_off);

  const uint64_t b_raw = block0 + (uint64_t)blockIdx.x * DecCfg<TYPE>::threads + threadIdx.x;
  const bool valid = b_raw < block1;
  const uint64_t b_list = valid ? b_raw : block1 - 1;
  const uint64_t b = g.box ? box_block(g, b_list) : b_list;
  const uint64_t off = offsets[b];
  // (an indexed length beyond the worst case of a block - the per-block term of zfp_stream_maximum_size,
  // src/zfp.c:711-742, as in the scan - is not believed: the staging below never reads past that)
  constexpr uint32_t header = TR::is_fp ? (REV ? 2 + TR::EBITS + TR::PBITS : 1 + TR::EBITS) : (REV ? TR::PBITS : 0);
  uint32_t cap = header + N - 1 + N * (prm.maxprec < (uint32_t)TR::P ? prm.maxprec : (uint32_t)TR::P);
  cap = cap > prm.maxbits ? prm.maxbits : cap;
  cap = cap < prm.minbits ? prm.minbits : cap;
  const uint32_t phase = (uint32_t)(off & 31), len = min((uint32_t)lengths[b], cap);
  ColReader br;
  br.init_var(stage, kVarStageWords, in + (off >> 5), (phase + len + 31) >> 5, phase);
  br.set_run_table(run_table);
  typename TR::Scalar v[N];
  const uint32_t bits = decode_block<TYPE, DIMS, REV>(v, prm, br, sp);
  // the index the offsets came from must describe THIS stream: its length against the parsed one
  if (check && bits != len)
    atomicOr(check, 1u);
  if (valid) {
    const BlockPos<DIMS> pos = locate<DIMS>(g, b);
    scatter<DIMS>(v, data, g, pos);
  }
}

// OFFS: 0 fixed rate (offset = start + b*maxbits), 1 per-block offsets from the index scan
template <int TYPE, int DIMS, int OFFS, bool REV>
__global__ void __launch_bounds__(kThreads)
decode_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm,
              const void* __restrict__ in, uint64_t start_bit, const uint64_t* __restrict__ offsets, uint64_t block0,
              uint64_t block1, const uint16_t* __restrict__ lengths, uint32_t* __restrict__ check)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  PW* sp = reinterpret_cast<PW*>(smem_raw) + (threadIdx.x >> 5) * (TR::P * 32) + (threadIdx.x & 31);
  uint32_t lut4 = 0;
  if constexpr (N == 4) {
    // the decoder table of the 4-value blocks (kDecLut4), one copy per CTA; these blocks deposit their planes straight
    // into the coefficients (decode_planes4_direct), so there is no plane buffer and sp is never dereferenced
    uint4* lut = reinterpret_cast<uint4*>(smem_raw);
    for (int i = threadIdx.x; i < kDecLut4Bytes / 16; i += kThreads)
      lut[i] = __ldg(reinterpret_cast<const uint4*>(kDecLut4) + i);
    __syncthreads();
    lut4 = (uint32_t)__cvta_generic_to_shared(lut);
  }

  const uint64_t b_list = block0 + (uint64_t)blockIdx.x * kThreads + threadIdx.x;
  if (b_list >= block1)
    return;
  const uint64_t b = g.box ? box_block(g, b_list) : b_list;
  const uint64_t bitpos = OFFS ? offsets[b] : start_bit + b * (uint64_t)prm.maxbits;
  typename TR::Scalar v[N];
  uint32_t bits = 0;
  bool small = false;
  if constexpr (N == 4 && OFFS == 0) {
    if (prm.maxbits <= 64) {  // fixed rate, the whole block in a register pair (its maxbits bits are all its own)
      SmallReader br;
      br.init(in, bitpos, prm.maxbits);
      br.lut4 = lut4;
      bits = decode_block<TYPE, DIMS, REV>(v, prm, br, sp);
      small = true;
    }
  }
  if (!small) {
    BitReader br;
    br.init(in, bitpos);
    br.lut4 = lut4;
    bits = decode_block<TYPE, DIMS, REV>(v, prm, br, sp);
  }
  if (OFFS == 1 && check && lengths && bits != lengths[b])
    atomicOr(check, 1u);
  const BlockPos<DIMS> pos = locate<DIMS>(g, b);
  scatter<DIMS>(v, data, g, pos);
}

// Index rebuild for a variable-rate stream that arrives without block lengths: block b's position
// is only known once blocks 0..b-1 have been parsed (the zfp format stores no offsets,
// docs/source/execution.rst:292-300).  index_scan_kernel is the plain answer: ONE thread walks the stream
// (length-only parse, 5.5-7 us per block) from block b0 at bit start_bit.  Correct, and slow; streams produced
// by this backend carry their index and never come here.
template <int TYPE, int DIMS, bool REV>
__global__ void index_scan_kernel(const void* __restrict__ in, uint64_t start_bit, uint64_t b0, uint64_t nblocks, Params prm,
                                  uint16_t* __restrict__ lengths)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  uint64_t pos = start_bit;
  for (uint64_t b = b0; b < nblocks; b++) {
    BitReader br;
    br.init(in, pos);
    typename TR::Scalar v[N];
    const uint32_t bits = decode_block<TYPE, DIMS, REV, BitReader, 0, true>(v, prm, br, nullptr);  // (length only)
    lengths[b] = (uint16_t)bits;
    pos += bits;
  }
}

// Speculative segment-parallel rebuild.  A parse that starts at a wrong bit position produces garbage blocks, but
// each of them ends somewhere, and as soon as one ends on a true block boundary the walk is on the true chain for
// good; with blocks of a few hundred bits that takes a few hundred blocks.  So: one thread per segment of the
// stream (segments are much longer than that).
//   pass 0: walk from the segment's first bit to its end; exit[t] = where the walk left the segment.  For all but
//           pathological streams the walk has met the true chain by then, so exit[t] is where the true chain
//           enters segment t + 1.
//   pass 1: walk segment t from that entry (segment 0: the stream's first block), count the blocks, and check that
//           the walk leaves through exit[t]; an exit that moves is corrected and the pass repeated.
//   pass 2: (block numbers now known from a prefix sum of the counts) walk once more and store the lengths.
// The result is a CANDIDATE: the decode checks every indexed length against what the block parses to and falls
// back to the sequential rebuild on any mismatch, so a walk that never converged costs time, not correctness.
// A walk stops margin_bits before the end of the buffer (garbage blocks must not read past it); whatever is left
// there is finished by index_scan_kernel.
template <int TYPE, int DIMS, bool REV, int PASS>
__global__ void __launch_bounds__(64) spec_index_kernel(SpecIndexArgs a)
{
  using TR = Traits<TYPE>;
  constexpr int N = 1 << (2 * DIMS);
  // one walker per WARP (lane 0): walkers in one warp would diverge at every loop of the parse and run one after the other
  const uint32_t t = (blockIdx.x * 64 + threadIdx.x) >> 5;
  if ((threadIdx.x & 31) != 0 || t >= a.nseg)
    return;
  const uint64_t seg_begin = a.start_bit + (uint64_t)t * a.seg_bits;
  uint64_t seg_end = seg_begin + a.seg_bits;
  const uint64_t stop = a.avail_bits > a.margin_bits ? a.avail_bits - a.margin_bits : 0;  // no walk starts a block beyond this
  if (t + 1 == a.nseg || seg_end > stop)
    seg_end = stop;
  uint64_t pos = (PASS == 0 || t == 0) ? seg_begin : a.exit[t - 1];
  uint64_t b = PASS == 2 ? a.off[t] : 0;
  uint32_t n = 0;
  while (pos < seg_end) {
    BitReader br;
    br.init(a.in, pos);
    typename TR::Scalar v[N];
    const uint32_t bits = decode_block<TYPE, DIMS, REV, BitReader, 0, true>(v, a.prm, br, nullptr);
    if (PASS == 2) {
      if (b < a.nblocks)
        a.lengths[b] = (uint16_t)bits;
      b++;
    }
    n++;
    pos += bits ? bits : 1;  // (a block is never empty; guard against a stuck walk on garbage parameters)
  }
  if (PASS == 0)
    a.exit[t] = pos;
  if (PASS == 1) {
    a.cnt[t] = n;
    if (a.exit[t] != pos) {
      a.exit[t] = pos;
      *a.changed = 1;
    }
  }
}

}  // namespace zb
