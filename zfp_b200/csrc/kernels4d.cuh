// kernels4d.cuh - 4-D blocks (4^4 = 256 values, 256-bit planes).
//
// A 256-value block does not fit a thread's registers, so this path keeps the block in
// thread-local arrays (L1-resident) and loops; bit planes are still produced by the 32x32
// register transpose, 8 groups of 32 coefficients at a time.  Same wire format as
// encode_many_ints / decode_many_ints (src/template/encode.c:135-176,208-236,
// decode.c:141-195,229-257), gather/scatter/pad as src/template/encode4.c, decode4.c,
// transform axis order x,y,z,w (encode4.c fwd_xform) and w,z,y,x for the inverse.
#pragma once

#include "kernels.cuh"

namespace zb {

constexpr int kThreads4 = 64;

struct BlockPos4 {
  int64_t offset;
  uint32_t ext[4];
  bool full;
};

__device__ __forceinline__ BlockPos4 locate4(const Geom& g, uint64_t b)
{
  BlockPos4 p;
  p.offset = 0;
  p.full = true;
  for (int d = 0; d < 4; d++) {
    uint64_t q = b / g.nb[d], c = b - q * g.nb[d];
    b = q;
    uint64_t org = 4 * c, left = g.n[d] - org;
    p.ext[d] = left < 4 ? (uint32_t)left : 4u;
    p.full &= left >= 4;
    p.offset += g.s[d] * (int64_t)org;
  }
  return p;
}

template <class Scalar>
__device__ void gather4(Scalar* v, const Scalar* data, const Geom& g, const BlockPos4& pos)
{
  const Scalar* p = data + pos.offset;
  for (int i = 0; i < 256; i++) {
    const uint32_t c0 = i & 3, c1 = (i >> 2) & 3, c2 = (i >> 4) & 3, c3 = (i >> 6) & 3;
    bool ok = c0 < pos.ext[0] && c1 < pos.ext[1] && c2 < pos.ext[2] && c3 < pos.ext[3];
    v[i] = ok ? __ldg(p + g.s[0] * c0 + g.s[1] * c1 + g.s[2] * c2 + g.s[3] * c3) : Scalar(0);
  }
  if (!pos.full)
    for (int d = 0; d < 4; d++) {
      const int st = 1 << (2 * d);
      if (pos.ext[d] < 4)
        for (int i = 0; i < 256; i++)
          if (((i >> (2 * d)) & 3) == 0)
            pad4(v[i], v[i + st], v[i + 2 * st], v[i + 3 * st], pos.ext[d]);
    }
}

template <class Scalar>
__device__ void scatter4(const Scalar* v, Scalar* data, const Geom& g, const BlockPos4& pos)
{
  Scalar* p = data + pos.offset;
  for (int i = 0; i < 256; i++) {
    const uint32_t c0 = i & 3, c1 = (i >> 2) & 3, c2 = (i >> 4) & 3, c3 = (i >> 6) & 3;
    if (c0 < pos.ext[0] && c1 < pos.ext[1] && c2 < pos.ext[2] && c3 < pos.ext[3])
      p[g.s[0] * c0 + g.s[1] * c1 + g.s[2] * c2 + g.s[3] * c3] = v[i];
  }
}

template <int KIND, class Int>
__device__ void xform4(Int* q, bool inverse)
{
  for (int a = 0; a < 4; a++) {
    const int axis = inverse ? 3 - a : a, st = 1 << (2 * axis);
    for (int i = 0; i < 256; i++)
      if (((i >> (2 * axis)) & 3) == 0)
        lift4<KIND>(q[i], q[i + st], q[i + 2 * st], q[i + 3 * st]);
  }
}

// planes as 32-bit pieces: piece g (coefficients 32g..32g+31) of plane k is pl[k*8 + g]
template <class UInt>
__device__ void to_planes4(const UInt* u, uint32_t* pl)
{
  constexpr int P = 8 * (int)sizeof(UInt);
  for (int h = 0; h < P / 32; h++)
    for (int g = 0; g < 8; g++) {
      uint32_t a[32];
#pragma unroll
      for (int i = 0; i < 32; i++)
        a[i] = (uint32_t)(u[32 * g + i] >> (32 * h));
      transpose32(a);
#pragma unroll
      for (int k = 0; k < 32; k++)
        pl[(32 * h + k) * 8 + g] = a[k];
    }
}

template <class UInt>
__device__ void from_planes4(UInt* u, const uint32_t* pl)
{
  constexpr int P = 8 * (int)sizeof(UInt);
  for (int i = 0; i < 256; i++)
    u[i] = 0;
  for (int h = 0; h < P / 32; h++)
    for (int g = 0; g < 8; g++) {
      uint32_t a[32];
#pragma unroll
      for (int k = 0; k < 32; k++)
        a[k] = pl[(32 * h + k) * 8 + g];
      transpose32(a);
#pragma unroll
      for (int i = 0; i < 32; i++)
        u[32 * g + i] |= (UInt)((UInt)a[i] << (32 * h));
    }
}

__device__ __forceinline__ uint64_t plane_word(const uint32_t* x, int w) { return (uint64_t)x[2 * w] | ((uint64_t)x[2 * w + 1] << 32); }

// index of the first one-bit at position >= n in a 256-bit plane, 256 if none
__device__ __forceinline__ uint32_t next_one(const uint32_t* x, uint32_t n)
{
  for (uint32_t w = n >> 5; w < 8; w++) {
    uint32_t r = x[w];
    if (w == (n >> 5)) r &= ~0u << (n & 31);
    if (r) return 32 * w + (uint32_t)__ffs((int)r) - 1;
  }
  return 256;
}

template <int P, class Writer>
__device__ uint32_t encode_planes4(Writer& bw, uint32_t budget, uint32_t maxprec, const uint32_t* pl)
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  for (int k = P - 1; bits && k >= kmin; k--) {
    const uint32_t* x = pl + k * 8;
    // verbatim part, 64 bits at a time
    uint32_t m = n < bits ? n : bits;
    bits -= m;
    for (int w = 0; m; w++) {
      uint32_t c = m < 64 ? m : 64;
      bw.put(plane_word(x, w) & lowmask64(c), c);
      m -= c;
    }
    while (bits && n < 256) {
      uint32_t nx = next_one(x, n);
      bw.put(nx < 256, 1);
      bits--;
      if (nx == 256) break;
      uint32_t zeros = nx - n;
      if (zeros > bits) zeros = bits;
      bw.pad(zeros);
      bits -= zeros;
      if (nx < 255 && bits) { bw.put(1, 1); bits--; }
      n = nx + 1;
    }
  }
  return budget - bits;
}

template <int P, bool ZERO = true>
__device__ uint32_t decode_planes4(BitReader& br, uint32_t budget, uint32_t maxprec, uint32_t* pl)
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  if (ZERO)
    for (int i = 0; i < P * 8; i++)
      pl[i] = 0;
  for (int k = P - 1; bits && k >= kmin; k--) {
    uint32_t* x = pl + k * 8;
    uint32_t m = n < bits ? n : bits;
    bits -= m;
    for (int w = 0; m; w++) {
      uint32_t c = m < 64 ? m : 64;
      uint64_t v = br.get(c);
      x[2 * w] = (uint32_t)v;
      x[2 * w + 1] = (uint32_t)(v >> 32);
      m -= c;
    }
    while (bits && n < 256) {
      bits--;
      if (!br.get(1)) break;
      uint32_t L = 255 - n;
      if (bits < L) L = bits;
      // scan up to L bits for a one
      bool found = false;
      while (L && !found) {
        uint32_t c = L < 64 ? L : 64;
        uint64_t t = br.peek(c);
        if (t) {
          uint32_t z = ctz64(t);
          br.skip(z + 1);
          bits -= z + 1;
          n += z;
          found = true;
        }
        else {
          br.skip(c);
          bits -= c;
          n += c;
          L -= c;
        }
      }
      x[n >> 5] |= 1u << (n & 31);
      n++;
    }
  }
  return budget - bits;
}

template <int TYPE, int OUT>
__global__ void __launch_bounds__(kThreads4)
encode4_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, void* __restrict__ out,
               uint64_t start_bit, uint32_t slot_words, uint16_t* __restrict__ lengths, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int P = TR::P;
  const uint64_t b = block0 + (uint64_t)blockIdx.x * kThreads4 + threadIdx.x;
  if (b >= block1) return;

  Scalar v[256];
  Int q[256];
  UInt u[256];
  uint32_t pl[P * 8];
  gather4(v, data, g, locate4(g, b));

  BitWriter<OUT == 1 ? 1 : 0> bw;
  if (OUT == 2) bw.init(out, (b - block0) * (uint64_t)slot_words * 64);
  else bw.init(out, start_bit + b * (uint64_t)prm.maxbits);

  const bool reversible = prm.minexp < kMinExp;
  uint32_t bits = 0, maxprec = prm.maxprec;
  bool coded = true;
  if constexpr (TR::is_fp) {
    // block exponent (same rules as block_emax in codec.cuh)
    const UInt absmask = ~(UInt)0 >> 1, infbits = (UInt)((1u << TR::EBITS) - 1) << TR::MANT;
    UInt mx = 0;
    for (int i = 0; i < 256; i++) {
      UInt a = FpBits<Scalar>::bits(v[i]) & absmask;
      a = a > infbits ? 0 : a;
      mx = a > mx ? a : mx;
    }
    const int E = (int)(mx >> TR::MANT);
    const int emax = mx == infbits ? 0 : E ? E - TR::EBIAS + 1 : (mx ? 1 - TR::EBIAS : -TR::EBIAS);
    const int se = P - 2 - emax;
    const Scalar s = pow2<Scalar>(se), sinv = pow2<Scalar>(emax - (P - 2));
    const bool overflow = se > (P == 32 ? 127 : 1023);
    if (!reversible) {
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 4);
      const uint32_t e = maxprec ? (uint32_t)(emax + TR::EBIAS) : 0;
      if (!e) {
        bw.put(0, 1);
        bits = 1;
        coded = false;
      }
      else {
        bits = 1 + TR::EBITS;
        bw.put(2 * (uint64_t)e + 1, bits);
        for (int i = 0; i < 256; i++)
          q[i] = overflow ? (Int)((UInt)1 << (P - 1)) : cvt_rz(s * v[i]);
      }
    }
    else {
      bool same = true;
      for (int i = 0; i < 256; i++) {
        if (emax != -TR::EBIAS) {
          q[i] = overflow ? (Int)((UInt)1 << (P - 1)) : cvt_rz(s * v[i]);
          Scalar back = sinv * cvt_rn(q[i], Scalar());
          same &= FpBits<Scalar>::bits(back) == FpBits<Scalar>::bits(v[i]);
        }
        else {
          q[i] = 0;
          same &= FpBits<Scalar>::bits(v[i]) == 0;
        }
      }
      if (same) {
        const uint32_t e = (uint32_t)(emax + TR::EBIAS);
        if (!e) {
          bw.put(0, 1);
          bw.flush();
          if (OUT == 2) lengths[b] = 1;
          return;
        }
        bw.put(1, 2);
        bw.put(e, TR::EBITS);
        bits = 2 + TR::EBITS;
      }
      else {
        for (int i = 0; i < 256; i++) {
          Int x = (Int)FpBits<Scalar>::bits(v[i]);
          q[i] = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        }
        bw.put(3, 2);
        bits = 2;
      }
    }
  }
  else
    for (int i = 0; i < 256; i++)
      q[i] = (Int)v[i];

  if (coded) {
    if (!reversible) {
      xform4<0>(q, false);
      for (int i = 0; i < 256; i++)
        u[i] = int2uint(q[c_perm4[i]]);
    }
    else {
      xform4<2>(q, false);
      UInt any = 0;
      for (int i = 0; i < 256; i++) {
        u[i] = int2uint(q[c_perm4[i]]);
        any |= u[i];
      }
      uint32_t prec = any ? (uint32_t)P - (P == 64 ? ctz64((uint64_t)any) : (uint32_t)__ffs((int)any) - 1) : 0;
      prec = prec < prm.maxprec ? prec : prm.maxprec;
      prec = prec > 1 ? prec : 1;
      bw.put(prec - 1, TR::PBITS);
      bits += TR::PBITS;
      maxprec = prec;
    }
    to_planes4(u, pl);
    bits += encode_planes4<P>(bw, prm.maxbits - bits, maxprec, pl);
  }
  if (bits < prm.minbits) {
    bw.pad(prm.minbits - bits);
    bits = prm.minbits;
  }
  bw.flush();
  if (OUT == 2) lengths[b] = (uint16_t)bits;
}

template <int TYPE, int OFFS>
__global__ void __launch_bounds__(kThreads4)
decode4_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, const void* __restrict__ in,
               uint64_t start_bit, const uint64_t* __restrict__ offsets, uint64_t block0, uint64_t block1,
               const uint16_t* __restrict__ lengths, uint32_t* __restrict__ check)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int P = TR::P;
  const uint64_t b_list = block0 + (uint64_t)blockIdx.x * kThreads4 + threadIdx.x;
  if (b_list >= block1) return;
  const uint64_t b = g.box ? box_block(g, b_list) : b_list;

  Scalar v[256];
  Int q[256];
  UInt u[256];
  uint32_t pl[P * 8];
  BitReader br;
  br.init(in, OFFS ? offsets[b] : start_bit + b * (uint64_t)prm.maxbits);

  const bool reversible = prm.minexp < kMinExp;
  uint32_t bits = 0, maxprec = prm.maxprec;
  int emax = 0;
  bool reinterpret = false, zero = false;
  if constexpr (TR::is_fp) {
    bits = 1;
    if (!br.get(1))
      zero = true;
    else if (!reversible) {
      bits += TR::EBITS;
      emax = (int)br.get(TR::EBITS) - TR::EBIAS;
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 4);
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
  if (zero) {
    for (int i = 0; i < 256; i++)
      v[i] = (Scalar)0;
  }
  else {
    if (reversible) {
      maxprec = (uint32_t)br.get(TR::PBITS) + 1;
      bits += TR::PBITS;
    }
    bits += decode_planes4<P>(br, prm.maxbits - bits, maxprec, pl);
    from_planes4(u, pl);
    for (int i = 0; i < 256; i++)
      q[c_perm4[i]] = uint2int(u[i]);
    if (!reversible) xform4<1>(q, true);
    else xform4<3>(q, true);
    if constexpr (TR::is_fp) {
      const Scalar sinv = pow2<Scalar>(emax - (P - 2));
      for (int i = 0; i < 256; i++) {
        if (reinterpret) {
          Int x = q[i];
          x = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
          v[i] = FpBits<Scalar>::make((typename FpBits<Scalar>::U)x);
        }
        else if (reversible && emax == -TR::EBIAS)
          v[i] = (Scalar)0;
        else
          v[i] = sinv * cvt_rn(q[i], Scalar());
      }
    }
    else
      for (int i = 0; i < 256; i++)
        v[i] = (Scalar)q[i];
  }
  if constexpr (OFFS == 1) {
    // the index the offsets came from must describe THIS stream: compare its length with the parsed one
    if (zero) bits = 1;
    if (bits < prm.minbits) bits = prm.minbits;
    if (check && lengths && bits != lengths[b])
      atomicOr(check, 1u);
  }
  scatter4(v, data, g, locate4(g, b));
}

// coded length of the 4-D block at the reader's position (index rebuild for foreign streams)
template <int TYPE>
__device__ uint32_t block_length4(BitReader& br, const Params& prm, uint32_t* pl)
{
  using TR = Traits<TYPE>;
  const bool reversible = prm.minexp < kMinExp;
  uint32_t bits = 0, maxprec = prm.maxprec;
  if constexpr (TR::is_fp) {
    bits = 1;
    if (!br.get(1))
      return bits < prm.minbits ? prm.minbits : bits;
    if (!reversible) {
      bits += TR::EBITS;
      const int emax = (int)br.get(TR::EBITS) - TR::EBIAS;
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 4);
    }
    else {
      bits++;
      if (!br.get(1)) {
        bits += TR::EBITS;
        br.skip(TR::EBITS);
      }
    }
  }
  if (reversible) {
    maxprec = (uint32_t)br.get(TR::PBITS) + 1;
    bits += TR::PBITS;
  }
  bits += decode_planes4<TR::P>(br, prm.maxbits - bits, maxprec, pl);
  return bits < prm.minbits ? prm.minbits : bits;
}

template <int TYPE>
__global__ void index_scan4_kernel(const void* __restrict__ in, uint64_t start_bit, uint64_t nblocks, Params prm,
                                   uint16_t* __restrict__ lengths)
{
  uint32_t pl[Traits<TYPE>::P * 8];
  uint64_t pos = start_bit;
  for (uint64_t b = 0; b < nblocks; b++) {
    BitReader br;
    br.init(in, pos);
    const uint32_t bits = block_length4<TYPE>(br, prm, pl);
    lengths[b] = (uint16_t)bits;
    pos += bits;
  }
}

}  // namespace zb
