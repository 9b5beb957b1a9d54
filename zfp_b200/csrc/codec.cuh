// codec.cuh - per-block zfp codec for sm_100a, one 4^d block per thread (d <= 3).
//
// Work split inside a warp (32 blocks at a time):
//   * everything that is data parallel inside a block - strided gather, block-floating-point
//     cast, lifting transform, sequency reorder, negabinary - runs on the block's 4^d values held
//     in registers with compile-time indices (no local memory);
//   * bit planes are formed by an in-register 32x32 bit-matrix transpose (5 butterfly stages),
//     not by per-bit loops, and parked in shared memory as [plane][lane] so that every lane's
//     accesses are conflict free;
//   * the inherently sequential embedded coder then walks the planes of its block with
//     popc/ctz arithmetic ("plane strings": n verbatim bits, then group-tested unary runs).
//
// Reference semantics reproduced (paths in the reference tree):
//   gather/pad      src/template/encode{1,2,3}.c, encode.c:8-27
//   exponent/cast   src/template/encodef.c:10-59, codecf.c:5-32
//   lifting         src/template/encode.c:30-56, decode.c:8-45, revencode.c:6-38, revdecode.c:6-38
//   order/negabin   src/template/encode.c:75-88, decode.c:63-76, codec{1,2,3}.c
//   embedded coder  src/template/encode.c:91-256, decode.c:79-278
//   reversible      src/template/revencode.c:41-79, revencodef.c:6-80, revdecode*.c, revcodecf.c
#pragma once

#include <cstdint>
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>

#include "zfp_perm_tables.h"

namespace zb {

constexpr int kMinExp = -1074;  // ZFP_MIN_EXP

enum : int { T_INT32 = 1, T_INT64 = 2, T_FLOAT = 3, T_DOUBLE = 4 };

template <int TYPE> struct Traits;
template <> struct Traits<T_INT32> {
  using Scalar = int32_t; using Int = int32_t; using UInt = uint32_t;
  static constexpr int P = 32, EBITS = 0, EBIAS = 0, MANT = 0, PBITS = 5;
  static constexpr bool is_fp = false;
};
template <> struct Traits<T_INT64> {
  using Scalar = int64_t; using Int = int64_t; using UInt = uint64_t;
  static constexpr int P = 64, EBITS = 0, EBIAS = 0, MANT = 0, PBITS = 6;
  static constexpr bool is_fp = false;
};
template <> struct Traits<T_FLOAT> {
  using Scalar = float; using Int = int32_t; using UInt = uint32_t;
  static constexpr int P = 32, EBITS = 8, EBIAS = 127, MANT = 23, PBITS = 5;
  static constexpr bool is_fp = true;
};
template <> struct Traits<T_DOUBLE> {
  using Scalar = double; using Int = int64_t; using UInt = uint64_t;
  static constexpr int P = 64, EBITS = 11, EBIAS = 1023, MANT = 52, PBITS = 6;
  static constexpr bool is_fp = true;
};

struct Geom {
  uint64_t n[4];    // extent per dimension (1 for unused)
  int64_t s[4];     // element strides (resolved, never 0)
  uint64_t nb[4];   // blocks per dimension
  uint64_t nblocks;
  int vec_rows;     // 1: sx == 1 and every 4-value row starts 16/32-byte aligned
};

struct Params {
  uint32_t minbits, maxbits, maxprec;
  int32_t minexp;
};

__constant__ uint8_t c_perm1[4] = ZFP_B200_PERM1_INIT;
__constant__ uint8_t c_perm2[16] = ZFP_B200_PERM2_INIT;
__constant__ uint8_t c_perm3[64] = ZFP_B200_PERM3_INIT;
__constant__ uint8_t c_perm4[256] = ZFP_B200_PERM4_INIT;

// compile-time lookup (switch, no table in memory) so that reordering is pure register renaming
// once the loops are unrolled
template <int DIMS>
__host__ __device__ constexpr int perm_at(int i)
{
  if (DIMS == 1) { switch (i) { ZFP_B200_PERM1_CASES } }
  else if (DIMS == 2) { switch (i) { ZFP_B200_PERM2_CASES } }
  else { switch (i) { ZFP_B200_PERM3_CASES } }
}

// ------------------------------------------------------------------------------------------------
// small bit helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t lowmask64(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1); }
__device__ __forceinline__ uint32_t ctz64(uint64_t x) { return (uint32_t)__ffsll((long long)x) - 1; }

// ------------------------------------------------------------------------------------------------
// bit writer: LSB-first into 64-bit words (include/zfp/bitstream.inl:288-313 semantics).
// MODE 0: the destination words are private to this thread from the first bit on (word-aligned
//         start): plain stores.
// MODE 1: neighbouring blocks share words: destination pre-zeroed, every word is OR-merged with a
//         fire-and-forget reduction.
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct BitWriter {
  unsigned long long* w;
  uint64_t acc;
  uint32_t fill;

  __device__ __forceinline__ void init(void* words, uint64_t bitpos)
  {
    w = reinterpret_cast<unsigned long long*>(words) + (bitpos >> 6);
    fill = (uint32_t)(bitpos & 63);
    acc = 0;
  }
  __device__ __forceinline__ void emit(uint64_t word)
  {
    if (MODE == 0)
      *w = word;
    else if (word)
      atomicOr(w, (unsigned long long)word);
    w++;
  }
  // append the low len bits of v (v < 2^len, 0 <= len <= 64)
  __device__ __forceinline__ void put(uint64_t v, uint32_t len)
  {
    acc |= v << fill;
    uint32_t total = fill + len;
    if (total >= 64) {
      emit(acc);
      acc = fill ? v >> (64 - fill) : 0;
      total -= 64;
    }
    fill = total;
  }
  __device__ __forceinline__ void pad(uint32_t nbits)
  {
    while (nbits >= 64) { put(0, 64); nbits -= 64; }
    put(0, nbits);
  }
  __device__ __forceinline__ void flush()
  {
    if (fill) { emit(acc); acc = 0; fill = 0; }
  }
};

// bit reader with a 64-bit look-ahead window
struct BitReader {
  const uint64_t* w;  // next word to fetch
  uint64_t buf;       // unread bits, LSB first
  uint32_t avail;     // number of valid bits in buf (0..64)

  __device__ __forceinline__ void init(const void* words, uint64_t bitpos)
  {
    w = reinterpret_cast<const uint64_t*>(words) + (bitpos >> 6);
    uint32_t sh = (uint32_t)(bitpos & 63);
    buf = __ldg(w++) >> sh;
    avail = 64 - sh;
  }
  // next len bits without consuming, 0 <= len <= 64
  __device__ __forceinline__ uint64_t peek(uint32_t len)
  {
    uint64_t v = buf;
    if (len > avail) {
      uint64_t nx = __ldg(w);
      v |= avail < 64 ? nx << avail : 0;
    }
    return v & lowmask64(len);
  }
  __device__ __forceinline__ void skip(uint32_t len)
  {
    if (len < avail) {
      buf >>= len;
      avail -= len;
    }
    else if (len == avail) {
      buf = 0;  // fetch lazily: never touch a word the stream may not own
      avail = 0;
    }
    else {
      uint32_t need = len - avail;  // 1..64
      uint64_t nx = __ldg(w++);
      buf = need < 64 ? nx >> need : 0;
      avail = 64 - need;
    }
  }
  __device__ __forceinline__ uint64_t get(uint32_t len)
  {
    uint64_t v = peek(len);
    skip(len);
    return v;
  }
};

// ------------------------------------------------------------------------------------------------
// lifting transforms on register-resident values (wrapping arithmetic via unsigned types)
// ------------------------------------------------------------------------------------------------
template <class Int>
__device__ __forceinline__ void fwd_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  x = (Int)((U)x + (U)w); x >>= 1; w = (Int)((U)w - (U)x);
  z = (Int)((U)z + (U)y); z >>= 1; y = (Int)((U)y - (U)z);
  x = (Int)((U)x + (U)z); x >>= 1; z = (Int)((U)z - (U)x);
  w = (Int)((U)w + (U)y); w >>= 1; y = (Int)((U)y - (U)w);
  w = (Int)((U)w + (U)(y >> 1)); y = (Int)((U)y - (U)(w >> 1));
}

template <class Int>
__device__ __forceinline__ void inv_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  y = (Int)((U)y + (U)(w >> 1)); w = (Int)((U)w - (U)(y >> 1));
  y = (Int)((U)y + (U)w); w = (Int)((U)w * 2u - (U)y);
  z = (Int)((U)z + (U)x); x = (Int)((U)x * 2u - (U)z);
  y = (Int)((U)y + (U)z); z = (Int)((U)z * 2u - (U)y);
  w = (Int)((U)w + (U)x); x = (Int)((U)x * 2u - (U)w);
}

template <class Int>
__device__ __forceinline__ void rev_fwd_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  U a = (U)x, b = (U)y, c = (U)z, d = (U)w;
  d -= c; c -= b; b -= a;
  d -= c; c -= b;
  d -= c;
  y = (Int)b; z = (Int)c; w = (Int)d;
}

template <class Int>
__device__ __forceinline__ void rev_inv_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  U a = (U)x, b = (U)y, c = (U)z, d = (U)w;
  d += c;
  c += b; d += c;
  b += a; c += b; d += c;
  y = (Int)b; z = (Int)c; w = (Int)d;
}

// KIND 0: fwd_lift, 1: inv_lift, 2: rev_fwd_lift, 3: rev_inv_lift
template <int KIND, class Int>
__device__ __forceinline__ void lift4(Int& x, Int& y, Int& z, Int& w)
{
  if (KIND == 0) fwd_lift(x, y, z, w);
  else if (KIND == 1) inv_lift(x, y, z, w);
  else if (KIND == 2) rev_fwd_lift(x, y, z, w);
  else rev_inv_lift(x, y, z, w);
}

template <int KIND, int DIMS, int AXIS, class Int>
__device__ __forceinline__ void lift_axis(Int (&p)[1 << (2 * DIMS)])
{
  constexpr int N = 1 << (2 * DIMS), st = 1 << (2 * AXIS);
#pragma unroll
  for (int i = 0; i < N; i++)
    if (((i >> (2 * AXIS)) & 3) == 0)
      lift4<KIND>(p[i], p[i + st], p[i + 2 * st], p[i + 3 * st]);
}

// forward: x, y, z (encode{1,2,3}.c fwd_xform); inverse: z, y, x (decode{1,2,3}.c inv_xform)
template <int KIND, int DIMS, class Int>
__device__ __forceinline__ void xform_fwd(Int (&p)[1 << (2 * DIMS)])
{
  lift_axis<KIND, DIMS, 0>(p);
  if (DIMS > 1) lift_axis<KIND, DIMS, (DIMS > 1 ? 1 : 0)>(p);
  if (DIMS > 2) lift_axis<KIND, DIMS, (DIMS > 2 ? 2 : 0)>(p);
}
template <int KIND, int DIMS, class Int>
__device__ __forceinline__ void xform_inv(Int (&p)[1 << (2 * DIMS)])
{
  if (DIMS > 2) lift_axis<KIND, DIMS, (DIMS > 2 ? 2 : 0)>(p);
  if (DIMS > 1) lift_axis<KIND, DIMS, (DIMS > 1 ? 1 : 0)>(p);
  lift_axis<KIND, DIMS, 0>(p);
}

// ------------------------------------------------------------------------------------------------
// 32x32 bit-matrix transpose in registers: on return bit i of a[j] is the former bit j of a[i]
// ------------------------------------------------------------------------------------------------
template <int J>
__device__ __forceinline__ void transpose32_stage(uint32_t (&a)[32])
{
  constexpr uint32_t m = J == 16 ? 0x0000ffffu : J == 8 ? 0x00ff00ffu : J == 4 ? 0x0f0f0f0fu : J == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
  for (int k = 0; k < 32; k++)
    if (!(k & J)) {
      // swap (row k, columns c+J) with (row k+J, columns c) for the columns c selected by m
      uint32_t t = ((a[k] >> J) ^ a[k + J]) & m;
      a[k + J] ^= t;
      a[k] ^= t << J;
    }
}

__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
  transpose32_stage<16>(a);
  transpose32_stage<8>(a);
  transpose32_stage<4>(a);
  transpose32_stage<2>(a);
  transpose32_stage<1>(a);
}

// Plane storage in shared memory: plane k of the calling lane is sp[k * 32] where sp already
// points at the lane's column (conflict free for 4- and 8-byte words).
template <int N> struct PlaneWord { using type = uint32_t; };
template <> struct PlaneWord<64> { using type = uint64_t; };

// coefficients (sequency order) -> bit planes
template <class UInt, int N>
__device__ __forceinline__ void to_planes(const UInt (&u)[N], typename PlaneWord<N>::type* sp)
{
  constexpr int P = 8 * (int)sizeof(UInt);
  constexpr int G = (N + 31) / 32;  // groups of 32 coefficients (1 or 2)
#pragma unroll
  for (int h = 0; h < P / 32; h++) {
    uint32_t a[G][32];
#pragma unroll
    for (int g = 0; g < G; g++) {
#pragma unroll
      for (int i = 0; i < 32; i++)
        a[g][i] = (32 * g + i < N) ? (uint32_t)(u[(32 * g + i) % N] >> (32 * h)) : 0u;
      transpose32(a[g]);
    }
#pragma unroll
    for (int k = 0; k < 32; k++) {
      if (G == 2)
        sp[(32 * h + k) * 32] = (typename PlaneWord<N>::type)((uint64_t)a[0][k] | ((uint64_t)a[G - 1][k] << 32));
      else
        sp[(32 * h + k) * 32] = (typename PlaneWord<N>::type)a[0][k];
    }
  }
}

// bit planes -> coefficients; planes below kstop were never written and read as zero
template <class UInt, int N>
__device__ __forceinline__ void from_planes(UInt (&u)[N], const typename PlaneWord<N>::type* sp, int kstop)
{
  constexpr int P = 8 * (int)sizeof(UInt);
  constexpr int G = (N + 31) / 32;
#pragma unroll
  for (int i = 0; i < N; i++)
    u[i] = 0;
#pragma unroll
  for (int h = 0; h < P / 32; h++) {
    uint32_t a[G][32];
#pragma unroll
    for (int k = 0; k < 32; k++) {
      typename PlaneWord<N>::type x = (32 * h + k >= kstop) ? sp[(32 * h + k) * 32] : 0;
      a[0][k] = (uint32_t)x;
      if (G == 2)
        a[G - 1][k] = (uint32_t)((uint64_t)x >> 32);
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      transpose32(a[g]);
#pragma unroll
      for (int i = 0; i < 32; i++)
        if (32 * g + i < N)
          u[(32 * g + i) % N] |= (UInt)((UInt)a[g][i] << (32 * h));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// embedded coder on plane words (N <= 64)
// ------------------------------------------------------------------------------------------------

// Emit planes P-1 .. P-maxprec truncated at `budget` bits; returns the bits used.
template <int N, int P, class Writer>
__device__ __forceinline__ uint32_t encode_planes(Writer& bw, uint32_t budget, uint32_t maxprec,
                                                  const typename PlaneWord<N>::type* sp)
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  for (int k = P - 1; bits && k >= kmin; k--) {
    uint64_t x = sp[k * 32];
    // bits of the n coefficients already significant, verbatim
    uint32_t m = n < bits ? n : bits;
    bits -= m;
    bw.put(x & lowmask64(m), m);
    // remaining coefficients: group test, then the distance to the next one-bit in unary
    uint64_t r = n < 64 ? x >> n : 0;
    while (bits && n < N) {
      if (!r) {
        bw.put(0, 1);
        bits--;
        break;
      }
      uint32_t c = ctz64(r);
      bool implied = n + c == N - 1;  // a one-bit in the last slot is not written
      uint32_t len = implied ? c + 1 : c + 2;
      uint64_t val = implied ? 1ull : (1ull | (2ull << c));
      if (len > bits) {
        len = bits;
        val &= lowmask64(len);
      }
      bw.put(val, len);
      bits -= len;
      n += c + 1;
      r >>= c;
      r >>= 1;
    }
  }
  return budget - bits;
}

// Mirror image.  Decoded planes are stored to sp[k*32]; returns bits consumed and, through
// kstop, the lowest plane index that was written.
template <int N, int P>
__device__ __forceinline__ uint32_t decode_planes(BitReader& br, uint32_t budget, uint32_t maxprec,
                                                  typename PlaneWord<N>::type* sp, int& kstop)
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  int k = P - 1;
  for (; bits && k >= kmin; k--) {
    uint32_t m = n < bits ? n : bits;
    bits -= m;
    uint64_t x = br.get(m);
    while (bits && n < N) {
      bits--;
      if (!br.get(1))
        break;
      // scan for the next one-bit: at most L more bits may be read
      uint32_t L = N - 1 - n;
      if (bits < L) L = bits;
      uint64_t t = br.peek(L);
      uint32_t c = t ? ctz64(t) : L;
      uint32_t used = t ? c + 1 : L;
      br.skip(used);
      bits -= used;
      n += c;
      x |= 1ull << n;  // deposited even if the scan ran dry (decode.c:103-111)
      n++;
    }
    sp[k * 32] = (typename PlaneWord<N>::type)x;
  }
  kstop = k + 1;
  return budget - bits;
}

// ------------------------------------------------------------------------------------------------
// floating-point helpers
// ------------------------------------------------------------------------------------------------
template <class T> struct FpBits;
template <> struct FpBits<float> {
  using U = uint32_t;
  __device__ static __forceinline__ U bits(float f) { return __float_as_uint(f); }
  __device__ static __forceinline__ float make(U b) { return __uint_as_float(b); }
};
template <> struct FpBits<double> {
  using U = uint64_t;
  __device__ static __forceinline__ U bits(double f) { return (U)__double_as_longlong(f); }
  __device__ static __forceinline__ double make(U b) { return __longlong_as_double((long long)b); }
};

// frexp exponent of the largest finite magnitude in the block, clamped as encodef.c:10-27;
// NaNs never win the comparison (encodef.c:35-37 uses `max < f`)
template <class TR, int N>
__device__ __forceinline__ int block_emax(const typename TR::Scalar (&v)[N])
{
  using U = typename TR::UInt;
  const U absmask = ~(U)0 >> 1, infbits = (U)((1u << TR::EBITS) - 1) << TR::MANT;
  U m = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    U a = FpBits<typename TR::Scalar>::bits(v[i]) & absmask;
    a = a > infbits ? 0 : a;
    m = a > m ? a : m;
  }
  int E = (int)(m >> TR::MANT);
  // inf: frexp leaves the exponent 0 in glibc; the value is irrelevant but keep it
  if (m == infbits) return 0 > 1 - TR::EBIAS ? 0 : 1 - TR::EBIAS;
  if (E) return E - TR::EBIAS + 1;
  return m ? 1 - TR::EBIAS : -TR::EBIAS;
}

// 2^e as Scalar for e in the normal range, +inf above it, 0 / subnormal below it
template <class Scalar> __device__ __forceinline__ Scalar pow2(int e);
template <> __device__ __forceinline__ float pow2<float>(int e)
{
  if (e > 127) return __uint_as_float(0x7f800000u);
  if (e >= -126) return __uint_as_float((uint32_t)(e + 127) << 23);
  return e >= -149 ? __uint_as_float(1u << (e + 149)) : 0.0f;
}
template <> __device__ __forceinline__ double pow2<double>(int e)
{
  if (e > 1023) return __longlong_as_double(0x7ff0000000000000ll);
  if (e >= -1022) return __longlong_as_double((long long)(e + 1023) << 52);
  return e >= -1074 ? __longlong_as_double(1ll << (e + 1074)) : 0.0;
}

__device__ __forceinline__ int32_t cvt_rz(float p) { return __float2int_rz(p); }
__device__ __forceinline__ int64_t cvt_rz(double p) { return __double2ll_rz(p); }
__device__ __forceinline__ float cvt_rn(int32_t i, float) { return __int2float_rn(i); }
__device__ __forceinline__ double cvt_rn(int64_t i, double) { return __ll2double_rn(i); }

// (Int)(2^(P-2-emax) * f), truncating; an infinite scale reproduces the x86 "integer
// indefinite" result of the reference build (oracle/zfp_oracle.c cast_fwd)
template <class TR, int N>
__device__ __forceinline__ void cast_fwd(typename TR::Int (&q)[N], const typename TR::Scalar (&v)[N], int emax)
{
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  const int se = TR::P - 2 - emax;
  const Scalar s = pow2<Scalar>(se);
  const bool overflow = se > (TR::P == 32 ? 127 : 1023);
#pragma unroll
  for (int i = 0; i < N; i++) {
    Int r = cvt_rz(s * v[i]);
    q[i] = overflow ? (Int)((typename TR::UInt)1 << (TR::P - 1)) : r;
  }
}

// (Scalar)i * 2^(emax-(P-2))
template <class TR, int N>
__device__ __forceinline__ void cast_inv(typename TR::Scalar (&v)[N], const typename TR::Int (&q)[N], int emax)
{
  using Scalar = typename TR::Scalar;
  const Scalar s = pow2<Scalar>(emax - (TR::P - 2));
#pragma unroll
  for (int i = 0; i < N; i++)
    v[i] = s * cvt_rn(q[i], Scalar());
}

template <class TR>
__device__ __forceinline__ uint32_t block_precision(int emax, uint32_t maxprec, int minexp, int dims)
{
  int p = emax - minexp + 2 * dims + 2;
  p = p < 0 ? 0 : p;
  return (uint32_t)p < maxprec ? (uint32_t)p : maxprec;
}

// negabinary
template <class Int>
__device__ __forceinline__ typename std::make_unsigned<Int>::type int2uint(Int x)
{
  using U = typename std::make_unsigned<Int>::type;
  const U mask = (U)0xaaaaaaaaaaaaaaaaull;
  return ((U)x + mask) ^ mask;
}
template <class UInt>
__device__ __forceinline__ typename std::make_signed<UInt>::type uint2int(UInt u)
{
  const UInt mask = (UInt)0xaaaaaaaaaaaaaaaaull;
  return (typename std::make_signed<UInt>::type)((u ^ mask) - mask);
}

// ------------------------------------------------------------------------------------------------
// whole-block encode / decode for one thread.  `sp` is the lane's plane column in shared memory.
// Returns the number of bits the block occupies in the stream.
// ------------------------------------------------------------------------------------------------
template <int TYPE, int DIMS, class Writer>
__device__ __forceinline__ uint32_t encode_block(const typename Traits<TYPE>::Scalar (&v)[1 << (2 * DIMS)],
                                                 const Params& prm, Writer& bw,
                                                 typename PlaneWord<(1 << (2 * DIMS))>::type* sp)
{
  using TR = Traits<TYPE>;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int N = 1 << (2 * DIMS), P = TR::P;
  const bool reversible = prm.minexp < kMinExp;
  uint32_t bits = 0, maxprec = prm.maxprec;
  Int q[N];

  if constexpr (TR::is_fp) {
    const int emax = block_emax<TR>(v);
    if (!reversible) {
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, DIMS);
      const uint32_t e = maxprec ? (uint32_t)(emax + TR::EBIAS) : 0;
      if (!e) {
        bw.put(0, 1);
        bits = 1;
        if (bits < prm.minbits) { bw.pad(prm.minbits - bits); bits = prm.minbits; }
        return bits;
      }
      bits = 1 + TR::EBITS;
      bw.put(2 * (uint64_t)e + 1, bits);
      cast_fwd<TR>(q, v, emax);
    }
    else {
      // try the block-floating-point route and verify it bit for bit (revencodef.c:6-27)
      bool same = true;
      if (emax != -TR::EBIAS) {
        typename TR::Scalar back[N];
        cast_fwd<TR>(q, v, emax);
        cast_inv<TR>(back, q, emax);
#pragma unroll
        for (int i = 0; i < N; i++)
          same &= FpBits<typename TR::Scalar>::bits(back[i]) == FpBits<typename TR::Scalar>::bits(v[i]);
      }
      else {
#pragma unroll
        for (int i = 0; i < N; i++) {
          q[i] = 0;
          same &= FpBits<typename TR::Scalar>::bits(v[i]) == 0;
        }
      }
      if (same) {
        const uint32_t e = (uint32_t)(emax + TR::EBIAS);
        if (!e) {
          bw.put(0, 1);
          return 1;  // no minbits padding on this path (revencodef.c:64-69)
        }
        bw.put(1, 2);
        bw.put(e, TR::EBITS);
        bits = 2 + TR::EBITS;
      }
      else {
        // sign-magnitude bit patterns -> two's complement (revencodef.c:29-41)
#pragma unroll
        for (int i = 0; i < N; i++) {
          Int x = (Int)FpBits<typename TR::Scalar>::bits(v[i]);
          q[i] = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        }
        bw.put(3, 2);
        bits = 2;
      }
    }
  }
  else {
#pragma unroll
    for (int i = 0; i < N; i++)
      q[i] = (Int)v[i];
  }

  UInt u[N];
  if (!reversible) {
    xform_fwd<0, DIMS>(q);
#pragma unroll
    for (int i = 0; i < N; i++)
      u[i] = int2uint(q[perm_at<DIMS>(i)]);
  }
  else {
    xform_fwd<2, DIMS>(q);
    UInt any = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      u[i] = int2uint(q[perm_at<DIMS>(i)]);
      any |= u[i];
    }
    // precision = width - (trailing zeros common to all coefficients), in [1, maxprec]
    uint32_t prec = any ? (uint32_t)P - (P == 64 ? ctz64((uint64_t)any) : (uint32_t)__ffs((int)any) - 1) : 0;
    prec = prec < prm.maxprec ? prec : prm.maxprec;
    prec = prec > 1 ? prec : 1;
    bw.put(prec - 1, TR::PBITS);
    bits += TR::PBITS;
    maxprec = prec;
  }
  to_planes<UInt, N>(u, sp);
  bits += encode_planes<N, P>(bw, prm.maxbits - bits, maxprec, sp);
  if (bits < prm.minbits) {
    bw.pad(prm.minbits - bits);
    bits = prm.minbits;
  }
  return bits;
}

template <int TYPE, int DIMS>
__device__ __forceinline__ uint32_t decode_block(typename Traits<TYPE>::Scalar (&v)[1 << (2 * DIMS)],
                                                 const Params& prm, BitReader& br,
                                                 typename PlaneWord<(1 << (2 * DIMS))>::type* sp)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int N = 1 << (2 * DIMS), P = TR::P;
  const bool reversible = prm.minexp < kMinExp;
  uint32_t bits = 0, maxprec = prm.maxprec;
  int emax = 0;
  bool reinterpret = false;

  if constexpr (TR::is_fp) {
    bits = 1;
    if (!br.get(1)) {
#pragma unroll
      for (int i = 0; i < N; i++)
        v[i] = (Scalar)0;
      return bits < prm.minbits ? prm.minbits : bits;  // the caller positions the next block
    }
    if (!reversible) {
      bits += TR::EBITS;
      emax = (int)br.get(TR::EBITS) - TR::EBIAS;
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, DIMS);
    }
    else {
      bits++;
      reinterpret = br.get(1) != 0;
      if (!reinterpret) {
        bits += TR::EBITS;
        emax = (int)br.get(TR::EBITS) - TR::EBIAS;
      }
    }
  }
  if (reversible) {
    maxprec = (uint32_t)br.get(TR::PBITS) + 1;
    bits += TR::PBITS;
  }

  int kstop;
  bits += decode_planes<N, P>(br, prm.maxbits - bits, maxprec, sp, kstop);
  if (bits < prm.minbits)
    bits = prm.minbits;

  UInt u[N];
  from_planes<UInt, N>(u, sp, kstop);
  Int q[N];
#pragma unroll
  for (int i = 0; i < N; i++)
    q[perm_at<DIMS>(i)] = uint2int(u[i]);
  if (!reversible)
    xform_inv<1, DIMS>(q);
  else
    xform_inv<3, DIMS>(q);

  if constexpr (TR::is_fp) {
    if (reinterpret) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        Int x = q[i];
        x = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        v[i] = FpBits<Scalar>::make((typename FpBits<Scalar>::U)x);
      }
    }
    else if (reversible && emax == -TR::EBIAS) {
#pragma unroll
      for (int i = 0; i < N; i++)
        v[i] = (Scalar)0;
    }
    else
      cast_inv<TR>(v, q, emax);
  }
  else {
#pragma unroll
    for (int i = 0; i < N; i++)
      v[i] = (Scalar)q[i];
  }
  return bits;
}

}  // namespace zb
