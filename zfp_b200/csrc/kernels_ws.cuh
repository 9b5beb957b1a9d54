// kernels_ws.cuh - warp-specialised fixed-rate decode for 3-D blocks of 64-bit values.
//
// The one-thread-per-block decoder (decode_staged_kernel) is two very different programs run back to
// back by the same thread: the stream PARSE - a serial dependence chain per bit plane (position ->
// shared-memory load -> carries -> find-first-set -> next position) that needs some forty registers -
// and the TAIL - inverse plane transposes, 48 inverse lifts, cast and stores: ~3300 independent
// integer instructions on the block's 64 values, i.e. 128 registers of data.  Sized for the tail
// (168 registers) only 12 warps fit an SM, and while they walk the parse chain the issue slots stay
// half empty (ncu: 49-53 % issue utilisation, "wait" the top stall).
//
// Here the two programs run in different warps of one CTA with different register budgets
// (setmaxnreg): warps 0-3 parse (64 registers), warps 4-7 run the tail (192 registers); parse warp i
// feeds tail warp 4+i through a ring of 16-plane windows in shared memory guarded by mbarriers.  The
// parser never waits for the tail's results, so it runs ahead - into the next batch of 32 blocks -
// while the tail warp is still in its long straight-line code: 16 warps per SM instead of 12, and the
// latency-bound chains overlap with ALU-bound work of another warp on the same scheduler.
//
// Same stream semantics as decode_staged_kernel (reference src/template/decode.c, decodef.c); lossy
// fixed-rate parameters with word-aligned blocks only - everything else keeps the general kernels.
#pragma once

#include "kernels.cuh"

namespace zb {

// ---- mbarrier / register-budget helpers (PTX ISA: mbarrier, setmaxnreg) ---------------------------
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait for the phase with the given parity to complete; a protocol error traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity)
{
  uint32_t spins = 0;
  while (!mbar_try_wait(addr, parity))
    if (++spins > (1u << 22))
      __trap();
}
template <int REGS> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

#ifndef ZB_WS_SYNC
#define ZB_WS_SYNC 1
#endif
// ---- configuration -------------------------------------------------------------------------------
constexpr int kWsPairs = 4;                  // parse/tail warp pairs per CTA (one warpgroup each side)
constexpr int kWsThreads = 64 * kWsPairs;    // 256
constexpr int kWsRing = 5;                   // 16-plane windows in flight per pair
constexpr int kWsParseRegs = 64, kWsTailRegs = 192;  // 128 * (64 + 192) = the CTA's 256 * 128 registers
constexpr uint32_t kWsSlotBytes = 16 * 32 * 8;       // one window: 16 planes x 32 lanes x 64-bit plane words

// what the parser tells the tail about a window, next to the plane words
struct WsMeta {
  int8_t lowest[32];   // lowest plane this lane has stored so far (64 = none); planes below read as zero
  uint8_t n[32];       // significant coefficients so far (for the "coefficients 32..63 untouched" shortcut)
  int16_t emax[32];    // block exponent (first window of a batch); 0 for an all-zero block
  uint32_t last;       // no further window of this batch follows
  uint32_t pad[3];
};
static_assert(sizeof(WsMeta) == 144, "layout");

// shared memory of one pair: ring of windows + their metadata + the stream column of the batch being parsed
__host__ __device__ constexpr uint32_t ws_pair_bytes(uint32_t words)
{
  return kWsRing * (kWsSlotBytes + (uint32_t)sizeof(WsMeta)) + (words + (uint32_t)kReadSlack) * 32 * 4;
}
__host__ __device__ constexpr uint32_t ws_cta_bytes(uint32_t words)
{
  return kWsPairs * ws_pair_bytes(words) + kWsPairs * kWsRing * 2 * 8 /* mbarriers */ + 128 /* run table (unused) */;
}

// 16 planes (window Q = 0..3: bits 16Q .. 16Q+15 of every coefficient) -> 16-bit slice Q of the 64 words
template <int Q, int NEG>
__device__ __forceinline__ void from_planes_w16(uint64_t (&u)[64], const uint64_t* sp, int kstop, bool upper)
{
  constexpr int H = Q >> 1, W = Q & 1;
  uint32_t a0[16], a1[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint64_t x = (16 * Q + i >= kstop) ? sp[i * 32] : 0;
    a0[i] = (uint32_t)x;
    a1[i] = (uint32_t)(x >> 32);
  }
  constexpr uint32_t sel_lo = W ? 0x1054u : 0x7610u, sel_hi = W ? 0x3254u : 0x7632u;  // __byte_perm(a, word, sel)
  auto put = [](uint64_t& word, uint32_t a, uint32_t sel) {
    const uint32_t half = (uint32_t)(word >> (32 * H));
    set_half<H>(word, __byte_perm(a, half, sel));
  };
  transpose16x2<NEG>(a0);
#pragma unroll
  for (int l = 0; l < 16; l++) {
    put(u[l], a0[l], sel_lo);
    put(u[16 + l], a0[l], sel_hi);
  }
  if (upper) {
    transpose16x2<NEG>(a1);
#pragma unroll
    for (int l = 0; l < 16; l++) {
      put(u[32 + l], a1[l], sel_lo);
      put(u[48 + l], a1[l], sel_hi);
    }
  }
}

// One batch = 32 consecutive blocks of the stream.  Pair p of CTA c handles batches
// (c * kWsPairs + p) + j * gridDim.x * kWsPairs, j = 0, 1, ...
template <int TYPE>
__global__ void __launch_bounds__(kWsThreads, 2)
decode_ws_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, const uint64_t* __restrict__ in,
                 uint64_t start_bit, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int N = 64, P = 64, NEG = 2;
  static_assert(TR::P == 64, "64-bit types only");
  extern __shared__ uint64_t smem_raw[];
  const uint32_t words = prm.maxbits >> 5;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool parser = warp < kWsPairs;
  const uint32_t pair = parser ? warp : warp - kWsPairs;

  char* pair_base = reinterpret_cast<char*>(smem_raw) + pair * ws_pair_bytes(words);
  uint64_t* slots = reinterpret_cast<uint64_t*>(pair_base);
  WsMeta* metas = reinterpret_cast<WsMeta*>(pair_base + kWsRing * kWsSlotBytes);
  uint32_t* column = reinterpret_cast<uint32_t*>(pair_base + kWsRing * (kWsSlotBytes + sizeof(WsMeta)));
  const uint32_t bars = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<char*>(smem_raw) + kWsPairs * ws_pair_bytes(words)) +
                        pair * kWsRing * 16;  // full[s] at bars + 16 s, empty[s] at bars + 16 s + 8
  if (lane == 0 && parser) {
    for (int s = 0; s < kWsRing; s++) {
      mbar_init(bars + 16 * s, 32);
      mbar_init(bars + 16 * s + 8, 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint64_t nbatches = (block1 - block0 + 31) >> 5;
  const uint64_t stride = (uint64_t)gridDim.x * kWsPairs;
  uint32_t q = 0;  // windows produced / consumed so far by this pair

  // every pair of the CTA makes the same number of trips (pair 0's), so that the tail warps can meet at a named
  // barrier before each tail; a trip past the last batch parses the last batch again and stores nothing
  if (parser) {
    reg_dealloc<kWsParseRegs>();
    uint32_t* stage = column + lane;
    for (uint64_t t0 = (uint64_t)blockIdx.x * kWsPairs; t0 < nbatches; t0 += stride) {
      const uint64_t t = t0 + pair < nbatches ? t0 + pair : nbatches - 1;
      const uint64_t b_raw = block0 + t * 32 + lane;
      const uint64_t b = b_raw < block1 ? b_raw : block1 - 1;  // lanes past the end redo the last block (warp votes need 32 lanes)
      // the block's words into the lane's column, all loads in flight
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

      ColReader br;
      br.init(stage);
      // block header (decodef.c:10-24): '0' = all-zero block, else '1' + biased exponent
      uint32_t bits = 0, maxprec = prm.maxprec;
      int emax = 0;
      bool zero = false;
      if constexpr (TR::is_fp) {
        bits = 1;
        zero = !br.get(1);
        if (!zero) {
          bits += TR::EBITS;
          emax = (int)br.get(TR::EBITS) - TR::EBIAS;
          maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, 3);
        }
      }
      const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
      LockDecodeState st = { prm.maxbits - bits, 0, P, P, zero };
      for (int w = 3; w >= 0; w--) {
        const uint32_t s = q % kWsRing, ph = (q / kWsRing) & 1;
        mbar_wait(bars + 16 * s + 8, ph ^ 1);  // the tail warp has read the previous contents of this slot
        uint64_t* sp = slots + s * (kWsSlotBytes / 8) + lane;
        decode_planes_lockstep<N>(br, kmin, 16 * w, 16 * w, st, sp);
        const bool more = w > 0 && __any_sync(0xffffffffu, !st.done && st.k > kmin && st.bits != 0);
        WsMeta& m = metas[s];
        m.lowest[lane] = (int8_t)st.lowest;
        m.n[lane] = (uint8_t)st.n;
        m.emax[lane] = (int16_t)emax;
        if (lane == 0) m.last = more ? 0u : 1u;
        mbar_arrive(bars + 16 * s);
        q++;
        if (!more) break;
      }
    }
  }
  else {
    reg_alloc<kWsTailRegs>();
    for (uint64_t t0 = (uint64_t)blockIdx.x * kWsPairs; t0 < nbatches; t0 += stride) {
      const bool real = t0 + pair < nbatches;
      const uint64_t t = real ? t0 + pair : nbatches - 1;
      const uint64_t b_raw = block0 + t * 32 + lane;
      const bool valid = real && b_raw < block1;
      const uint64_t b = valid ? b_raw : block1 - 1;
      UInt u[N];
#pragma unroll
      for (int i = 0; i < N; i++)
        u[i] = (UInt)NegaWord<NEG>::w64;
      int emax = 0;
#pragma unroll 1
      for (int w = 3; w >= 0; w--) {
        const uint32_t s = q % kWsRing, ph = (q / kWsRing) & 1;
        mbar_wait(bars + 16 * s, ph);
        const uint64_t* sp = slots + s * (kWsSlotBytes / 8) + lane;
        const WsMeta& m = metas[s];
        const int kstop = m.lowest[lane];
        const bool upper = __any_sync(0xffffffffu, m.n[lane] > 32);
        const bool last = m.last != 0;
        if (w == 3) emax = m.emax[lane];
        switch (w) {
          case 3: from_planes_w16<3, NEG>(u, sp, kstop, upper); break;
          case 2: from_planes_w16<2, NEG>(u, sp, kstop, upper); break;
          case 1: from_planes_w16<1, NEG>(u, sp, kstop, upper); break;
          default: from_planes_w16<0, NEG>(u, sp, kstop, upper); break;
        }
        mbar_arrive(bars + 16 * s + 8);
        q++;
        if (last) break;
      }
#if ZB_WS_SYNC
      // the four tail warps of the CTA walk the long straight-line tail together: one instruction fetch serves four
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kWsPairs) : "memory");
#endif
      // inverse negabinary (the XOR half happened in the transposes), order, lifting, cast, scatter
      Int qv[N];
#pragma unroll
      for (int i = 0; i < N; i++)
        qv[perm_at<3>(i)] = (Int)(u[i] - (UInt)0xaaaaaaaaaaaaaaaaull);
      xform_inv<1, 3>(qv);
      Scalar v[N];
      if constexpr (TR::is_fp)
        cast_inv<TR>(v, qv, emax);
      else {
#pragma unroll
        for (int i = 0; i < N; i++)
          v[i] = (Scalar)qv[i];
      }
      if (valid) {
        const BlockPos<3> pos = locate<3>(g, b);
        scatter<3>(v, data, g, pos);
      }
    }
  }
}

}  // namespace zb
