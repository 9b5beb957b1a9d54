// kernels_q4.cuh - fixed-rate kernels for 3-D blocks of 64-bit values with FOUR LANES PER BLOCK in the
// transform stages.
//
// Why: with one thread per block (kernels.cuh) everything that is data parallel inside a block is
// straight-line code on 64 values in registers - 48 unrolled 64-bit lifts, 64x64 bit-matrix transposes:
// 55 KB of instructions that every warp streams through once per 32 blocks, and 128 registers of data
// per thread (12 warps per multiprocessor).  ncu on the B200: issue slots half empty with
// "no instruction" the top stall as soon as more warps are resident (the warp-specialised experiment,
// DESIGN.md): the kernels are bound by instruction FETCH - the hot code is twice the 32 KB instruction
// cache level and no two warps are at the same place in it.
//
// Here a warp still owns 32 consecutive blocks of the stream and the embedded coder still runs one
// block per lane (codec.cuh, plane-lockstep), but the transform side works on 8 blocks at a time, four
// lanes per block, in a loop of four passes:
//   lane (q, s), q = block of the pass, s = 0..3:
//     slice layout    values (x, y, z = s)     : gather / scatter rows, cast, lifts along x and y
//     column layout   values (x, y = s, z)     : lift along z            (exchange through shared memory)
//     sequency layout coefficients 16 s .. 16 s + 15 of the zfp order (codec3.c perm_3)
//                                              : negabinary, 16 x 16 bit-plane transposes
//   Plane words meet the coder in the same [plane][block] shared-memory matrix as before; each lane
//   contributes / takes its 16-bit slice of a block's 64-bit plane word.
// The pass body is ~700 instructions (11 KB) executed four times per warp, the data per lane is 16
// values, and the 32-bit halves of the coefficients that are not being coded (decode: already decoded
// high halves, encode: low halves kept for the lower planes) wait in registers: ~120 registers,
// 16 warps per multiprocessor, hot code under 20 KB.
//
// Semantics are those of decode_staged_kernel / encode_staged_kernel (reference src/template/
// decode3.c, encode3.c, decode.c, encode.c, decodef.c, encodef.c); lossy modes only.
#pragma once

#include "kernels.cuh"

namespace zb {

constexpr int kQ4Threads = 128;                      // 4 warps, each independent
constexpr uint32_t kQ4PlaneBytes = 32 * 32 * 8;      // [32 planes][32 blocks] 64-bit plane words
constexpr uint32_t kQ4ColStride = 656;               // column / slice exchange: bytes per block (41 x 16: conflict-free 16-byte accesses)
constexpr uint32_t kQ4SliceStride = 160;             //   bytes per z slice inside it (4 rows of 32 bytes + 32)
constexpr uint32_t kQ4SeqStride = 528;               // sequency exchange: bytes per block (64 values + 16)
constexpr uint32_t kQ4ExchBytes = 8 * kQ4ColStride;  // 8 blocks per pass

__host__ __device__ constexpr uint32_t q4_warp_bytes(uint32_t words)
{
  // plane matrix + per-block metadata + max(stream column of the warp, exchange buffer)
  const uint32_t col = (words + (uint32_t)kReadSlack + (uint32_t)kStageSlack) * 32 * 4;
  return kQ4PlaneBytes + 256 + (col > kQ4ExchBytes ? col : kQ4ExchBytes);
}

// sequency order of the 16 coefficients a lane holds in the sequency layout, as spatial indices
// x + 4 y + 16 z packed one per byte (codec3.c perm_3 through zfp_perm_tables.h)
__host__ __device__ constexpr uint32_t q4_perm_word(int t, int w)
{
  return (uint32_t)perm_at<3>(16 * t + 4 * w) | ((uint32_t)perm_at<3>(16 * t + 4 * w + 1) << 8) |
         ((uint32_t)perm_at<3>(16 * t + 4 * w + 2) << 16) | ((uint32_t)perm_at<3>(16 * t + 4 * w + 3) << 24);
}
struct Q4Perm {
  uint32_t w[4];
  __device__ __forceinline__ void init(uint32_t t)
  {
#pragma unroll
    for (int i = 0; i < 4; i++)
      w[i] = t == 0 ? q4_perm_word(0, i) : t == 1 ? q4_perm_word(1, i) : t == 2 ? q4_perm_word(2, i) : q4_perm_word(3, i);
  }
  // spatial index of the lane's l-th coefficient (l compile time after unrolling)
  __device__ __forceinline__ uint32_t at(int l) const { return (w[l >> 2] >> (8 * (l & 3))) & 0xffu; }
};

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds_v2(uint32_t addr, uint64_t& a, uint64_t& b)
{
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint64_t a, uint64_t b)
{
  asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void sts_u64(uint32_t addr, uint64_t a)
{
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(a) : "memory");
}

// the 16-bit slices [16 t, 16 t + 16) of planes base .. base + 31 of block `blk` -> the 32-bit half
// (bit p = plane base + p) of the lane's 16 coefficients.  `planes` = shared address of the matrix.
template <int NEG>
__device__ __forceinline__ void q4_planes_to_half(uint32_t (&a)[16], uint32_t planes, uint32_t blk, uint32_t t)
{
  const uint32_t src = planes + blk * 8 + t * 2;
#pragma unroll
  for (int i = 0; i < 16; i++)
    a[i] = __byte_perm(lds_u16(src + i * 256), lds_u16(src + (16 + i) * 256), 0x5410);
  transpose16x2<NEG>(a);
}

// ---------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------
template <int TYPE>
__global__ void __launch_bounds__(kQ4Threads, 4)
decode_q4_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, const uint64_t* __restrict__ in,
                 uint64_t start_bit, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  static_assert(TR::P == 64, "64-bit types only");
  constexpr int N = 64, P = 64, NEG = 2;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ uint64_t smem_raw[];
  const uint32_t words = prm.maxbits >> 5;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t warp_off = warp * q4_warp_bytes(words);
  asm volatile("" : "+r"(warp_off));  // (kept opaque: see encode_staged_kernel)
  char* base = reinterpret_cast<char*>(smem_raw) + warp_off;
  uint64_t* planes64 = reinterpret_cast<uint64_t*>(base);
  const uint32_t planes = (uint32_t)__cvta_generic_to_shared(base);
  int16_t* m_emax = reinterpret_cast<int16_t*>(base + kQ4PlaneBytes + 64);   // [32] block exponent
  uint32_t* column = reinterpret_cast<uint32_t*>(base + kQ4PlaneBytes + 256);
  const uint32_t exch = planes + kQ4PlaneBytes + 256;                        // aliases the column (dead once parsing is over)

  // ---- parse, one block per lane ------------------------------------------------------------------
  const uint64_t b_first = block0 + ((uint64_t)blockIdx.x * (kQ4Threads / 32) + warp) * 32;
  if (b_first >= block1)
    return;  // whole warp (the warps of a CTA do not synchronise with each other)
  const uint64_t b_raw = b_first + lane;
  const uint64_t b = b_raw < block1 ? b_raw : block1 - 1;  // lanes past the end redo the last block (warp votes need 32 lanes)
  uint32_t* stage = column + lane;
  {
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(in + (start_bit >> 6)) + b * (uint64_t)words;
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
    else {
      const uint64_t* src = reinterpret_cast<const uint64_t*>(src32);
#pragma unroll 4
      for (uint32_t w = 0; w < words; w += 2) {
        const uint64_t v = __ldg(src + (w >> 1));
        stage[w * 32] = (uint32_t)v;
        stage[(w + 1) * 32] = (uint32_t)(v >> 32);
      }
    }
#pragma unroll
    for (int j = 0; j < kReadSlack; j++)
      stage[(words + j) * 32] = 0;
  }
  ColReader br;
  br.init(stage);
  uint32_t hbits = 0, maxprec = prm.maxprec;
  int emax = 0;
  bool zero = false;
  if constexpr (TR::is_fp) {  // block header (decodef.c:10-24): '0' = all-zero block, else '1' + biased exponent
    hbits = 1;
    zero = !br.get(1);
    if (!zero) {
      hbits += TR::EBITS;
      emax = (int)br.get(TR::EBITS) - TR::EBIAS;
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 3);
    }
  }
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  LockDecodeState st = { prm.maxbits - hbits, 0, P, P, zero };
  uint64_t* sp = planes64 + lane;
  // planes the parse does not reach read as zero
#pragma unroll
  for (int k = 0; k < 32; k++)
    sp[k * 32] = 0;
  decode_planes_lockstep<N>(br, kmin, 32, 32, st, sp);
  m_emax[lane] = (int16_t)emax;
  __syncwarp();

  // ---- planes 63..32 -> high halves of the coefficients, four lanes per block ---------------------
  const uint32_t q = lane >> 2, t = lane & 3;
  uint32_t hi[4][16];
#pragma unroll 1
  for (int j = 0; j < 4; j++) {
    uint32_t a[16];
    q4_planes_to_half<NEG>(a, planes, 8 * j + q, t);
    switch (j) {  // (register arrays want compile-time indices; the pass body is not replicated)
      case 0:
#pragma unroll
        for (int i = 0; i < 16; i++) hi[0][i] = a[i];
        break;
      case 1:
#pragma unroll
        for (int i = 0; i < 16; i++) hi[1][i] = a[i];
        break;
      case 2:
#pragma unroll
        for (int i = 0; i < 16; i++) hi[2][i] = a[i];
        break;
      default:
#pragma unroll
        for (int i = 0; i < 16; i++) hi[3][i] = a[i];
        break;
    }
  }
  __syncwarp();

  // ---- planes 31..0 if some block of the warp still has budget and precision for them ------------
  const bool low = __any_sync(FULL, !st.done && st.k > kmin && st.bits != 0);
  if (low) {
#pragma unroll
    for (int k = 0; k < 32; k++)
      sp[k * 32] = 0;
    decode_planes_lockstep<N>(br, kmin, 0, 0, st, sp);
    __syncwarp();
  }

  // ---- per pass: low halves, inverse negabinary, order, lifting, cast, scatter --------------------
  Q4Perm perm;
  perm.init(t);
#pragma unroll 1
  for (int j = 0; j < 4; j++) {
    uint32_t lo[16], h[16];
    if (low)
      q4_planes_to_half<NEG>(lo, planes, 8 * j + q, t);
    else {
#pragma unroll
      for (int i = 0; i < 16; i++)
        lo[i] = NegaWord<NEG>::w32;
    }
    switch (j) {
      case 0:
#pragma unroll
        for (int i = 0; i < 16; i++) h[i] = hi[0][i];
        break;
      case 1:
#pragma unroll
        for (int i = 0; i < 16; i++) h[i] = hi[1][i];
        break;
      case 2:
#pragma unroll
        for (int i = 0; i < 16; i++) h[i] = hi[2][i];
        break;
      default:
#pragma unroll
        for (int i = 0; i < 16; i++) h[i] = hi[3][i];
        break;
    }
    __syncwarp();  // the exchange buffer is free (previous pass read, or the parse done with the column)
    // sequency layout -> column layout: coefficient 16 t + l goes to its place x + 4 y + 16 z
    {
      const uint32_t dst = exch + q * kQ4SeqStride;
#pragma unroll
      for (int l = 0; l < 16; l++) {
        const uint64_t c = (((uint64_t)h[l] << 32) | lo[l]) - 0xaaaaaaaaaaaaaaaaull;  // uint2int, XOR half done by the transposes
        sts_u64(dst + perm.at(l) * 8, c);
      }
    }
    __syncwarp();
    Int r[16];  // column layout: r[4 z + x] = value (x, y = t, z)
    {
      const uint32_t src = exch + q * kQ4SeqStride + t * 32;
#pragma unroll
      for (int z = 0; z < 4; z++) {
        uint64_t v0, v1, v2, v3;
        lds_v2(src + z * 128, v0, v1);
        lds_v2(src + z * 128 + 16, v2, v3);
        r[4 * z] = (Int)v0; r[4 * z + 1] = (Int)v1; r[4 * z + 2] = (Int)v2; r[4 * z + 3] = (Int)v3;
      }
    }
    // inverse lift along z (decode3.c inv_xform: z first)
#pragma unroll
    for (int x = 0; x < 4; x++)
      inv_lift(r[x], r[4 + x], r[8 + x], r[12 + x]);
    __syncwarp();
    // column layout -> slice layout
    {
      const uint32_t dst = exch + q * kQ4ColStride + t * 32;
#pragma unroll
      for (int z = 0; z < 4; z++) {
        sts_v2(dst + z * kQ4SliceStride, (uint64_t)r[4 * z], (uint64_t)r[4 * z + 1]);
        sts_v2(dst + z * kQ4SliceStride + 16, (uint64_t)r[4 * z + 2], (uint64_t)r[4 * z + 3]);
      }
    }
    __syncwarp();
    Int s[16];  // slice layout: s[4 y + x] = value (x, y, z = t)
    {
      const uint32_t src = exch + q * kQ4ColStride + t * kQ4SliceStride;
#pragma unroll
      for (int y = 0; y < 4; y++) {
        uint64_t v0, v1, v2, v3;
        lds_v2(src + y * 32, v0, v1);
        lds_v2(src + y * 32 + 16, v2, v3);
        s[4 * y] = (Int)v0; s[4 * y + 1] = (Int)v1; s[4 * y + 2] = (Int)v2; s[4 * y + 3] = (Int)v3;
      }
    }
#pragma unroll
    for (int x = 0; x < 4; x++)
      inv_lift(s[x], s[4 + x], s[8 + x], s[12 + x]);          // along y
#pragma unroll
    for (int y = 0; y < 4; y++)
      inv_lift(s[4 * y], s[4 * y + 1], s[4 * y + 2], s[4 * y + 3]);  // along x
    // cast and scatter the slice z = t of block 8 j + q
    const uint32_t blk = 8 * j + q;
    Scalar v[16];
    if constexpr (TR::is_fp)
      cast_inv<TR>(v, s, (int)m_emax[blk]);
    else {
#pragma unroll
      for (int i = 0; i < 16; i++)
        v[i] = (Scalar)s[i];
    }
    const uint64_t bb = b_first + blk;
    if (bb < block1) {
      const BlockPos<3> pos = locate<3>(g, bb);
      Scalar* p = data + pos.offset + g.s[2] * (int64_t)t;
      if (pos.full && g.vec_rows) {
#pragma unroll
        for (int y = 0; y < 4; y++)
          Vec4<Scalar>::store(p + g.s[1] * y, v[4 * y], v[4 * y + 1], v[4 * y + 2], v[4 * y + 3]);
      }
      else if (t < pos.ext[2]) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint32_t x = i & 3, y = i >> 2;
          if (x < pos.ext[0] && y < pos.ext[1])
            p[g.s[0] * x + g.s[1] * y] = v[i];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------

// exponent of the block maximum (block_emax, codec.cuh) with the 64 values spread over four lanes
template <class TR>
__device__ __forceinline__ int q4_block_emax(const typename TR::Scalar (&v)[16])
{
  static_assert(sizeof(typename TR::Scalar) == 8, "double");
  int32_t smax = (int32_t)0x80000000;
  uint32_t umax = 0, any = 0;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const uint64_t w0 = FpBits<typename TR::Scalar>::bits(v[i]), w1 = FpBits<typename TR::Scalar>::bits(v[i + 1]);
    const uint32_t h0 = (uint32_t)(w0 >> 32), h1 = (uint32_t)(w1 >> 32);
    smax = __vimax3_s32(smax, (int32_t)h0, (int32_t)h1);
    umax = __vimax3_u32(umax, h0, h1);
    any |= (h0 & 0x7fffffffu) | (uint32_t)w0 | (h1 & 0x7fffffffu) | (uint32_t)w1;
  }
#pragma unroll
  for (int d = 1; d <= 2; d <<= 1) {
    smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, d));
    umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, d));
    any |= __shfl_xor_sync(0xffffffffu, any, d);
  }
  const uint32_t a = (uint32_t)smax & 0x7fffffffu, b = umax & 0x7fffffffu;
  const uint32_t top = a > b ? a : b;
  const int E = (int)(top >> 20);
  if (E) return E - TR::EBIAS + 1;
  return any ? 1 - TR::EBIAS : -TR::EBIAS;  // subnormal maximum / all-zero block
}

// the lane's 16 x 32-bit halves (bit p = plane base + p of coefficient 16 t + l) -> its 16-bit slices of
// planes base .. base + 31 of block `blk` in the plane matrix
template <int NEG>
__device__ __forceinline__ void q4_half_to_planes(uint32_t (&a)[16], uint32_t planes, uint32_t blk, uint32_t t)
{
  transpose16x2<NEG>(a);
  const uint32_t dst = planes + blk * 8 + t * 2;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst + i * 256), "h"((unsigned short)a[i]) : "memory");
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst + (16 + i) * 256), "h"((unsigned short)(a[i] >> 16)) : "memory");
  }
}

template <int TYPE>
__global__ void __launch_bounds__(kQ4Threads, 4)
encode_q4_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, uint64_t* __restrict__ out,
                 uint64_t start_bit)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  static_assert(TR::P == 64, "64-bit types only");
  constexpr int N = 64, P = 64, NEG = 1;
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ uint64_t smem_raw[];
  const uint32_t words = prm.maxbits >> 5;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t warp_off = warp * q4_warp_bytes(words);
  asm volatile("" : "+r"(warp_off));
  char* base = reinterpret_cast<char*>(smem_raw) + warp_off;
  uint64_t* planes64 = reinterpret_cast<uint64_t*>(base);
  const uint32_t planes = (uint32_t)__cvta_generic_to_shared(base);
  int16_t* m_emax = reinterpret_cast<int16_t*>(base + kQ4PlaneBytes + 64);
  uint32_t* column = reinterpret_cast<uint32_t*>(base + kQ4PlaneBytes + 256);
  const uint32_t exch = planes + kQ4PlaneBytes + 256;  // aliases the stream column, which the coder starts only after the passes

  const uint64_t b_first = ((uint64_t)blockIdx.x * (kQ4Threads / 32) + warp) * 32;
  if (b_first >= g.nblocks)
    return;  // whole warp
  const uint32_t q = lane >> 2, t = lane & 3;
  Q4Perm perm;
  perm.init(t);

  // ---- passes: gather, cast, lifting, order, negabinary; planes 63..32 to the matrix ----------------
  uint32_t lo[4][16];
#pragma unroll 1
  for (int j = 0; j < 4; j++) {
    const uint32_t blk = 8 * j + q;
    const uint64_t bb_raw = b_first + blk;
    const uint64_t bb = bb_raw < g.nblocks ? bb_raw : g.nblocks - 1;  // past the end: the last block again, discarded later
    const BlockPos<3> pos = locate<3>(g, bb);
    Scalar v[16];  // slice layout: v[4 y + x] = value (x, y, z = t)
    {
      const Scalar* p = data + pos.offset;
      if (pos.full && g.vec_rows) {
        p += g.s[2] * (int64_t)t;
#pragma unroll
        for (int y = 0; y < 4; y++)
          Vec4<Scalar>::load(p + g.s[1] * y, v[4 * y], v[4 * y + 1], v[4 * y + 2], v[4 * y + 3]);
      }
      else {
        // partial block / strided array: the slice a padded z line would copy (encode.c:8-27 pad_block:
        // 1 valid -> a a a a, 2 -> a b b a, 3 -> a b c a), then the same rule along x and y
        const uint32_t ez = pos.ext[2];
        const uint32_t zsrc = t < ez ? t : (ez == 2 && t == 2) ? 1u : 0u;
        p += g.s[2] * (int64_t)zsrc;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint32_t x = i & 3, y = i >> 2;
          const bool ok = x < pos.ext[0] && y < pos.ext[1];
          v[i] = ok ? __ldg(p + g.s[0] * x + g.s[1] * y) : Scalar(0);
        }
#pragma unroll
        for (int y = 0; y < 4; y++)
          pad4(v[4 * y], v[4 * y + 1], v[4 * y + 2], v[4 * y + 3], pos.ext[0]);
#pragma unroll
        for (int x = 0; x < 4; x++)
          pad4(v[x], v[4 + x], v[8 + x], v[12 + x], pos.ext[1]);
      }
    }
    Int s[16];
    if constexpr (TR::is_fp) {
      const int emax = q4_block_emax<TR>(v);
      cast_fwd<TR>(s, v, emax);
      if (t == 0)
        m_emax[blk] = (int16_t)emax;
    }
    else {
#pragma unroll
      for (int i = 0; i < 16; i++)
        s[i] = (Int)v[i];
    }
    // forward transform (encode3.c fwd_xform): along x, then y ...
#pragma unroll
    for (int y = 0; y < 4; y++)
      fwd_lift(s[4 * y], s[4 * y + 1], s[4 * y + 2], s[4 * y + 3]);
#pragma unroll
    for (int x = 0; x < 4; x++)
      fwd_lift(s[x], s[4 + x], s[8 + x], s[12 + x]);
    __syncwarp();  // exchange buffer free
    // slice layout -> column layout
    {
      const uint32_t dst = exch + q * kQ4ColStride + t * kQ4SliceStride;
#pragma unroll
      for (int y = 0; y < 4; y++) {
        sts_v2(dst + y * 32, (uint64_t)s[4 * y], (uint64_t)s[4 * y + 1]);
        sts_v2(dst + y * 32 + 16, (uint64_t)s[4 * y + 2], (uint64_t)s[4 * y + 3]);
      }
    }
    __syncwarp();
    Int r[16];  // column layout: r[4 z + x] = value (x, y = t, z)
    {
      const uint32_t src = exch + q * kQ4ColStride + t * 32;
#pragma unroll
      for (int z = 0; z < 4; z++) {
        uint64_t v0, v1, v2, v3;
        lds_v2(src + z * kQ4SliceStride, v0, v1);
        lds_v2(src + z * kQ4SliceStride + 16, v2, v3);
        r[4 * z] = (Int)v0; r[4 * z + 1] = (Int)v1; r[4 * z + 2] = (Int)v2; r[4 * z + 3] = (Int)v3;
      }
    }
    // ... then along z
#pragma unroll
    for (int x = 0; x < 4; x++)
      fwd_lift(r[x], r[4 + x], r[8 + x], r[12 + x]);
    __syncwarp();
    // column layout -> sequency layout; pre-negabinary words (the XOR half of int2uint happens in the transposes)
    {
      const uint32_t dst = exch + q * kQ4SeqStride + t * 32;
#pragma unroll
      for (int z = 0; z < 4; z++) {
        sts_v2(dst + z * 128, (uint64_t)r[4 * z] + 0xaaaaaaaaaaaaaaaaull, (uint64_t)r[4 * z + 1] + 0xaaaaaaaaaaaaaaaaull);
        sts_v2(dst + z * 128 + 16, (uint64_t)r[4 * z + 2] + 0xaaaaaaaaaaaaaaaaull, (uint64_t)r[4 * z + 3] + 0xaaaaaaaaaaaaaaaaull);
      }
    }
    __syncwarp();
    uint32_t hi[16], l32[16];
    {
      const uint32_t src = exch + q * kQ4SeqStride;
#pragma unroll
      for (int l = 0; l < 16; l++) {
        uint64_t c;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(c) : "r"(src + perm.at(l) * 8));
        hi[l] = (uint32_t)(c >> 32);
        l32[l] = (uint32_t)c;
      }
    }
    q4_half_to_planes<NEG>(hi, planes, blk, t);
    switch (j) {  // low halves wait in registers for the lower planes
      case 0:
#pragma unroll
        for (int i = 0; i < 16; i++) lo[0][i] = l32[i];
        break;
      case 1:
#pragma unroll
        for (int i = 0; i < 16; i++) lo[1][i] = l32[i];
        break;
      case 2:
#pragma unroll
        for (int i = 0; i < 16; i++) lo[2][i] = l32[i];
        break;
      default:
#pragma unroll
        for (int i = 0; i < 16; i++) lo[3][i] = l32[i];
        break;
    }
  }
  __syncwarp();

  // ---- embedded coder, one block per lane -------------------------------------------------------------
  uint32_t* stage = column + lane;
  ColWriter bw;
  bw.init(stage);
  uint32_t hbits = 0, maxprec = prm.maxprec;
  bool coded = true;
  if constexpr (TR::is_fp) {  // block header (encodef.c:61-75): '1' + biased exponent, or a lone '0'
    const int emax = (int)m_emax[lane];
    maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 3);
    const uint32_t e = maxprec ? (uint32_t)(emax + TR::EBIAS) : 0;
    coded = e != 0;
    hbits = coded ? 1 + TR::EBITS : 1;
    bw.put(coded ? 2 * (uint64_t)e + 1 : 0, hbits);
  }
  const uint32_t budget = prm.maxbits - hbits, start = bw.tell();
  const uint32_t limit = start + budget < start ? 0xffffffffu : start + budget;
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  LockState st = { 0, P, !coded };
  uint64_t* sp = planes64 + lane;
  encode_planes_lockstep<N>(bw, limit, kmin, 32, 32, st, sp);
  if (__any_sync(FULL, !st.done && st.k > kmin && bw.tell() < limit)) {
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      uint32_t a[16];
      switch (j) {
        case 0:
#pragma unroll
          for (int i = 0; i < 16; i++) a[i] = lo[0][i];
          break;
        case 1:
#pragma unroll
          for (int i = 0; i < 16; i++) a[i] = lo[1][i];
          break;
        case 2:
#pragma unroll
          for (int i = 0; i < 16; i++) a[i] = lo[2][i];
          break;
        default:
#pragma unroll
          for (int i = 0; i < 16; i++) a[i] = lo[3][i];
          break;
      }
      q4_half_to_planes<NEG>(a, planes, 8 * j + q, t);
    }
    __syncwarp();
    encode_planes_lockstep<N>(bw, limit, kmin, 0, 0, st, sp);
  }
  bw.finish(words);
  __syncwarp();

  const uint64_t b = b_first + lane;
  if (b < g.nblocks) {
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(out + (start_bit >> 6)) + b * (uint64_t)words;
    if ((words & 3) == 0 && (reinterpret_cast<uintptr_t>(dst32) & 15) == 0) {
      uint4* dst4 = reinterpret_cast<uint4*>(dst32);
#pragma unroll 4
      for (uint32_t w = 0; w < words; w += 4)
        dst4[w >> 2] = make_uint4(stage[w * 32], stage[(w + 1) * 32], stage[(w + 2) * 32], stage[(w + 3) * 32]);
    }
    else {
      uint64_t* dst = reinterpret_cast<uint64_t*>(dst32);
#pragma unroll 4
      for (uint32_t w = 0; w < words; w += 2)
        dst[w >> 1] = (uint64_t)stage[w * 32] | ((uint64_t)stage[(w + 1) * 32] << 32);
    }
  }
}

}  // namespace zb
