// kernels4q.cuh - 4-D blocks (4^4 = 256 values) with FOUR LANES PER BLOCK.
//
// kernels4d.cuh keeps a whole 256-value block in one thread's local memory (6 KB of LDL / STL traffic per
// block, and a 64^4 array is only 65 536 threads: a fifth of one wave).  Here lane w of a quad owns the
// 3-D sub-block w of the 4-D block - 64 values in registers, the very code of the 3-D kernels for gather,
// cast and the lifts along x, y, z - and the quad meets in shared memory three times:
//   * the lift along w: values (position, w) exchanged so that each lane holds 16 positions x 4 w;
//   * the zfp coefficient order (codec4.c perm_4): coefficients scattered to their rank, each lane then
//     takes 64 consecutive ranks and transposes them into its 64-bit word of every 256-bit plane;
//   * the planes themselves: [plane][4 words], 32 bytes per plane, the layout the 256-bit embedded coder
//     of kernels4d.cuh (encode_planes4 / decode_planes4: src/template/encode.c:135-176, decode.c:141-195)
//     reads.  The coder is serial in the block and runs on lane 0 of the quad.
// One 2 KB buffer per block (1 KB for 32-bit types) serves all three.  All modes, all stream layouts of
// encode4_kernel / decode4_kernel, same wire format.
#pragma once

#include "kernels4d.cuh"

namespace zb {

#ifndef ZB_4Q_CTAS
#define ZB_4Q_CTAS 6  // CTAs of 64 threads per SM (168-register cap, what the 33 KB of shared memory per CTA allow): 64^4 rate 8 0.49 -> 0.40 ms against uncapped (228-255 registers, 4 CTAs)
#endif
constexpr int kThreads4q = 64;              // 16 blocks per CTA
constexpr int kBlocks4q = kThreads4q / 4;

template <int TYPE>
__host__ __device__ constexpr uint32_t block_bytes4q()
{
  return 256u * (uint32_t)sizeof(typename Traits<TYPE>::Int) + 32u;  // + 32: neighbouring blocks start 8 banks apart
}

template <class T>
__device__ __forceinline__ T quad_max(T v)
{
  const T a = __shfl_xor_sync(0xffffffffu, v, 1);
  v = a > v ? a : v;
  const T b = __shfl_xor_sync(0xffffffffu, v, 2);
  return b > v ? b : v;
}
__device__ __forceinline__ bool quad_all(bool p)
{
  const unsigned m = __ballot_sync(0xffffffffu, p);
  return ((m >> (threadIdx.x & 28)) & 0xfu) == 0xfu;
}

// largest finite magnitude of the block as raw bits (block_emax of codec.cuh, split so that the quad can reduce it)
template <class TR, bool EXACT>
__device__ __forceinline__ typename TR::UInt local_absmax(const typename TR::Scalar (&v)[64])
{
  using U = typename TR::UInt;
  const U absmask = ~(U)0 >> 1, infbits = (U)((1u << TR::EBITS) - 1) << TR::MANT;
  U m = 0;
#pragma unroll
  for (int i = 0; i < 64; i++) {
    U a = FpBits<typename TR::Scalar>::bits(v[i]) & absmask;
    if (EXACT) a = a > infbits ? 0 : a;
    m = a > m ? a : m;
  }
  return m;
}
template <class TR, bool EXACT>
__device__ __forceinline__ int emax_of(typename TR::UInt m)
{
  using U = typename TR::UInt;
  const U infbits = (U)((1u << TR::EBITS) - 1) << TR::MANT;
  const int E = (int)(m >> TR::MANT);
  if (EXACT && m == infbits) return 0 > 1 - TR::EBIAS ? 0 : 1 - TR::EBIAS;
  if (E) return E - TR::EBIAS + 1;
  return m ? 1 - TR::EBIAS : -TR::EBIAS;
}

// 3-D sub-block w of the 4-D block at `pos`: where it starts and what is valid (pad rule along w: a lane
// beyond the valid extent loads the sub-block a padded w line would copy, encode.c:8-27)
__device__ __forceinline__ BlockPos<3> sub_block(const Geom& g, const BlockPos4& pos, uint32_t w, bool& inside)
{
  const uint32_t ew = pos.ext[3];
  inside = w < ew;
  const uint32_t wsrc = inside ? w : (ew == 2 && w == 2) ? 1u : 0u;
  BlockPos<3> p;
  p.offset = pos.offset + g.s[3] * (int64_t)wsrc;
  p.ext[0] = pos.ext[0];
  p.ext[1] = pos.ext[1];
  p.ext[2] = pos.ext[2];
  p.full = pos.ext[0] == 4 && pos.ext[1] == 4 && pos.ext[2] == 4;
  return p;
}

template <int TYPE, int OUT, bool REV>
__global__ void __launch_bounds__(kThreads4q, ZB_4Q_CTAS)
encode4q_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, void* __restrict__ out,
                uint64_t start_bit, uint32_t slot_words, uint16_t* __restrict__ lengths, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int P = TR::P;
  extern __shared__ uint64_t smem_raw[];
  __shared__ uint8_t s_rank[256];  // rank of coefficient i4 = x + 4y + 16z + 64w in the zfp order (inverse of perm_4)
  for (int i = threadIdx.x; i < 256; i += kThreads4q)
    s_rank[c_perm4[i]] = (uint8_t)i;
  __syncthreads();

  const uint32_t quad = threadIdx.x >> 2, w = threadIdx.x & 3;
  const uint64_t b_raw = block0 + (uint64_t)blockIdx.x * kBlocks4q + quad;
  const bool valid = b_raw < block1;
  const uint64_t b = valid ? b_raw : block1 - 1;  // idle quads redo the last block and write nothing
  Int* X = reinterpret_cast<Int*>(reinterpret_cast<char*>(smem_raw) + quad * block_bytes4q<TYPE>());

  const BlockPos4 pos = locate4(g, b);
  bool inside;
  const BlockPos<3> sub = sub_block(g, pos, w, inside);
  Scalar v[64];
  gather<3>(v, data, g, sub);

  // ---- block header and cast (encodef.c:61-82, revencodef.c:6-80), decided for the whole block ----
  uint32_t hbits = 0, maxprec = prm.maxprec;
  uint64_t hval = 0;      // header bits (at most 2 + 11)
  bool coded = true, pad = true;
  Int q[64];
  if constexpr (TR::is_fp) {
    const int emax = emax_of<TR, REV>(quad_max(local_absmax<TR, REV>(v)));
    if (!REV) {
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 4);
      const uint32_t e = maxprec ? (uint32_t)(emax + TR::EBIAS) : 0;
      coded = e != 0;
      hbits = coded ? 1 + TR::EBITS : 1;
      hval = coded ? 2 * (uint64_t)e + 1 : 0;
      cast_fwd<TR>(q, v, emax);
    }
    else {
      bool same = true;
      if (emax != -TR::EBIAS) {
        Scalar back[64];
        cast_fwd<TR>(q, v, emax);
        cast_inv<TR>(back, q, emax);
#pragma unroll
        for (int i = 0; i < 64; i++)
          same &= FpBits<Scalar>::bits(back[i]) == FpBits<Scalar>::bits(v[i]);
      }
      else {
#pragma unroll
        for (int i = 0; i < 64; i++) {
          q[i] = 0;
          same &= FpBits<Scalar>::bits(v[i]) == 0;
        }
      }
      same = quad_all(same);
      if (same) {
        const uint32_t e = (uint32_t)(emax + TR::EBIAS);
        if (!e) {
          hbits = 1;
          hval = 0;
          coded = pad = false;  // a lone '0', no minbits padding on this path (revencodef.c:64-69)
        }
        else {
          hbits = 2 + TR::EBITS;
          hval = 1 | ((uint64_t)e << 2);
        }
      }
      else {
#pragma unroll
        for (int i = 0; i < 64; i++) {
          Int x = (Int)FpBits<Scalar>::bits(v[i]);
          q[i] = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        }
        hbits = 2;
        hval = 3;
      }
    }
  }
  else {
#pragma unroll
    for (int i = 0; i < 64; i++)
      q[i] = (Int)v[i];
  }

  // ---- transform: x, y, z in registers, w across the quad (encode4.c fwd_xform) -----------------------
  xform_fwd<REV ? 2 : 0, 3>(q);
#pragma unroll
  for (int i = 0; i < 64; i++)
    X[i * 4 + w] = q[i];
  __syncwarp();
  Int r[64];  // r[4 l + w'] = value (position 16 w + l, w')
#pragma unroll
  for (int l = 0; l < 16; l++) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      r[4 * l + k] = X[(16 * w + l) * 4 + k];
    lift4<REV ? 2 : 0>(r[4 * l], r[4 * l + 1], r[4 * l + 2], r[4 * l + 3]);
  }
  __syncwarp();
  // ---- negabinary, zfp order: coefficient (position i, w') has 4-D index i + 64 w' --------------------
  UInt* U = reinterpret_cast<UInt*>(X);
  UInt any = 0;
#pragma unroll
  for (int l = 0; l < 16; l++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const UInt c = int2uint(r[4 * l + k]);
      any |= c;
      U[s_rank[16 * w + l + 64 * k]] = c;
    }
  }
  __syncwarp();
  UInt u[64];
#pragma unroll
  for (int i = 0; i < 64; i++)
    u[i] = U[64 * w + i];
  __syncwarp();
  // ---- planes: lane w supplies word w of every 256-bit plane ------------------------------------------
  to_planes<0, UInt, 64, 4>(u, reinterpret_cast<uint64_t*>(X) + w);
  if (REV) {  // precision = width - trailing zeros common to all 256 coefficients (revencode.c:58-79)
    any |= __shfl_xor_sync(0xffffffffu, any, 1);
    any |= __shfl_xor_sync(0xffffffffu, any, 2);
  }
  __syncwarp();

  // ---- the embedded coder on 256-bit planes, serial in the block: lane 0 of the quad -------------------
  if (w == 0 && valid) {
    BitWriter<OUT == 1 ? 1 : 0> bw;
    if (OUT == 2) bw.init(out, (b - block0) * (uint64_t)slot_words * 64);
    else bw.init(out, start_bit + b * (uint64_t)prm.maxbits);
    uint32_t bits = hbits;
    if (hbits) bw.put(hval, hbits);
    if (coded) {
      if (REV) {
        uint32_t prec = any ? (uint32_t)P - (P == 64 ? ctz64((uint64_t)any) : (uint32_t)__ffs((int)any) - 1) : 0;
        prec = prec < prm.maxprec ? prec : prm.maxprec;
        prec = prec > 1 ? prec : 1;
        bw.put(prec - 1, TR::PBITS);
        bits += TR::PBITS;
        maxprec = prec;
      }
      bits += encode_planes4<P>(bw, prm.maxbits - bits, maxprec, reinterpret_cast<const uint32_t*>(X));
    }
    if (pad && bits < prm.minbits) {
      bw.pad(prm.minbits - bits);
      bits = prm.minbits;
    }
    bw.flush();
    if (OUT == 2) lengths[b] = (uint16_t)bits;
  }
}

template <int TYPE, int OFFS, bool REV>
__global__ void __launch_bounds__(kThreads4q, ZB_4Q_CTAS)
decode4q_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, const void* __restrict__ in,
                uint64_t start_bit, const uint64_t* __restrict__ offsets, uint64_t block0, uint64_t block1,
                const uint16_t* __restrict__ lengths, uint32_t* __restrict__ check)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int P = TR::P;
  extern __shared__ uint64_t smem_raw[];
  __shared__ uint8_t s_perm[256];  // perm_4: 4-D index of the coefficient with rank i
  for (int i = threadIdx.x; i < 256; i += kThreads4q)
    s_perm[i] = c_perm4[i];
  __syncthreads();

  const uint32_t quad = threadIdx.x >> 2, w = threadIdx.x & 3;
  const uint64_t b_raw = block0 + (uint64_t)blockIdx.x * kBlocks4q + quad;
  const bool valid = b_raw < block1;
  const uint64_t b_list = valid ? b_raw : block1 - 1;
  const uint64_t b = g.box ? box_block(g, b_list) : b_list;
  Int* X = reinterpret_cast<Int*>(reinterpret_cast<char*>(smem_raw) + quad * block_bytes4q<TYPE>());
  uint32_t* pl = reinterpret_cast<uint32_t*>(X);

  // planes the parse does not reach read as zero
#pragma unroll 4
  for (int i = w; i < P * 8; i += 4)
    pl[i] = 0;
  __syncwarp();

  // ---- lane 0: header and the 256-bit plane parse ------------------------------------------------------
  int emax = 0;
  uint32_t flags = 0;  // bit 0: all-zero block, bit 1: reversible reinterpret path
  if (w == 0) {
    BitReader br;
    br.init(in, OFFS ? offsets[b] : start_bit + b * (uint64_t)prm.maxbits);
    uint32_t bits = 0, maxprec = prm.maxprec;
    bool zero = false, reinterpret = false;
    if constexpr (TR::is_fp) {
      bits = 1;
      if (!br.get(1))
        zero = true;
      else if (!REV) {
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
    if (!zero) {
      if (REV) {
        maxprec = (uint32_t)br.get(TR::PBITS) + 1;
        bits += TR::PBITS;
      }
      bits += decode_planes4<P, false>(br, prm.maxbits - bits, maxprec, pl);
    }
    if (OFFS == 1) {  // the index the offsets came from must describe THIS stream
      if (zero) bits = 1;
      if (bits < prm.minbits) bits = prm.minbits;
      if (valid && check && lengths && bits != lengths[b])
        atomicOr(check, 1u);
    }
    flags = (zero ? 1u : 0u) | (reinterpret ? 2u : 0u);
  }
  emax = __shfl_sync(0xffffffffu, emax, threadIdx.x & 28);
  flags = __shfl_sync(0xffffffffu, flags, threadIdx.x & 28);
  __syncwarp();

  // ---- planes -> the lane's 64 consecutive ranks -> their places (position, w') ---------------------------
  UInt u[64];
  from_planes<0, UInt, 64, 4>(u, reinterpret_cast<const uint64_t*>(X) + w, 0);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 64; i++) {
    const uint32_t i4 = s_perm[64 * w + i];
    X[(i4 & 63) * 4 + (i4 >> 6)] = uint2int(u[i]);
  }
  __syncwarp();
  // ---- inverse transform: w across the quad, then z, y, x in registers (decode4.c inv_xform) --------------
  Int r[64];
#pragma unroll
  for (int l = 0; l < 16; l++) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      r[4 * l + k] = X[(16 * w + l) * 4 + k];
    lift4<REV ? 3 : 1>(r[4 * l], r[4 * l + 1], r[4 * l + 2], r[4 * l + 3]);
  }
  __syncwarp();
#pragma unroll
  for (int l = 0; l < 16; l++)
#pragma unroll
    for (int k = 0; k < 4; k++)
      X[(16 * w + l) * 4 + k] = r[4 * l + k];
  __syncwarp();
  Int q[64];
#pragma unroll
  for (int i = 0; i < 64; i++)
    q[i] = X[i * 4 + w];
  xform_inv<REV ? 3 : 1, 3>(q);

  Scalar v[64];
  if constexpr (TR::is_fp) {
    if (flags & 2u) {
#pragma unroll
      for (int i = 0; i < 64; i++) {
        Int x = q[i];
        x = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        v[i] = FpBits<Scalar>::make((typename FpBits<Scalar>::U)x);
      }
    }
    else if ((flags & 1u) || (REV && emax == -TR::EBIAS)) {
#pragma unroll
      for (int i = 0; i < 64; i++)
        v[i] = (Scalar)0;
    }
    else
      cast_inv<TR>(v, q, emax);
  }
  else {
#pragma unroll
    for (int i = 0; i < 64; i++)
      v[i] = (Scalar)q[i];
  }
  const BlockPos4 pos = locate4(g, b);
  bool inside;
  BlockPos<3> sub = sub_block(g, pos, w, inside);
  if (valid && inside)
    scatter<3>(v, data, g, sub);
}

}  // namespace zb
