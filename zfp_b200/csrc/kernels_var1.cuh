// kernels_var1.cuh - single-pass variable-rate encode for 3-D blocks (fixed precision / accuracy /
// reversible / expert parameters): encode, place and index in ONE kernel.
//
// The first variable-rate path (backend.cu, still used for 1-, 2- and 4-D) writes every block to a
// fixed-stride scratch slot, scans the 16-bit lengths with three kernels, zeroes the destination and
// copies the blocks into place with a fourth: ~3x the compressed bytes of extra traffic, up to 1 GiB
// of scratch per chunk and six launches per chunk.  Here a CTA (one tile of 64 / 128 blocks):
//   1. encodes its blocks into the lane-private shared-memory columns, as before;
//   2. concatenates them, bit-granular, in shared memory (the plane matrix is free by then: it is
//      exactly as large as the longest tile the column windows can hold);
//   3. learns where its tile starts from a decoupled look-back over the tiles before it (Merrill &
//      Garland's single-pass scan): every tile publishes its bit total and its last 64 bits as soon as
//      it has them, so a successor never waits for more than the aggregates;
//   4. writes its words shifted to the tile's bit phase with plain coalesced stores.  A word that
//      straddles two tiles belongs to the later one, which completes it with the predecessor's
//      published tail - no atomics, no pre-zeroed destination (stream_copy semantics of
//      include/zfp/bitstream.inl:412-424 without the read-modify-write).
// Blocks longer than the column window (2048 bits less the drain margin; rare outside noise at tight
// tolerances) only count their bits here and leave a zeroed hole of the right size; reencode_kernel
// then codes each of them again straight into its hole with the general bit writer.
#pragma once

#include "kernels.cuh"
#include "scan_util.cuh"

namespace zb {

struct Var1Status {
  unsigned long long state;  // bits 63..62: 0 nothing yet, 1 tile total, 2 inclusive end position; bits 61..0 the value
  unsigned long long tail;   // the last 64 bits of the stream up to the end of this tile (valid with state != 0)
};
struct Var1Overflow {
  uint64_t block, bit;  // block number and the position of its hole in the stream
};
constexpr unsigned long long kVar1Agg = 1ull << 62, kVar1Incl = 2ull << 62, kVar1Mask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// carry[0] = end position of everything before this launch, carry[1] = its last 64 bits; updated by the last
// tile of the launch.  status and ticket must be zero at launch; the overflow list accumulates over the launches
// of one array and reencode_kernel runs after the last of them (a hole may reach into the word a later
// launch's first tile writes).
template <int TYPE, bool REV>
__global__ void __launch_bounds__(EncCfg<TYPE>::threads, EncCfg<TYPE>::min_ctas(REV))
encode_var1_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, uint64_t* __restrict__ out,
                   uint16_t* __restrict__ lengths, uint64_t block0, uint64_t block1, Var1Status* __restrict__ status,
                   unsigned int* __restrict__ ticket, unsigned long long* __restrict__ carry, Var1Overflow* __restrict__ overflow,
                   unsigned int* __restrict__ overflow_count, unsigned int overflow_capacity)
{
  using TR = Traits<TYPE>;
  constexpr int N = 64, T = EncCfg<TYPE>::threads, W = T / 32;
  using PW = typename PlaneWord<N>::type;
  static_assert(sizeof(PW) == 8, "64 coefficients per plane");
  constexpr uint32_t kPlaneBytes = kStagedPlanes * 32 * (uint32_t)sizeof(PW);  // per warp
  constexpr uint32_t kSegWords = W * kPlaneBytes / 4;                          // 32-bit words of the tile buffer
  static_assert(kSegWords * 32 >= T * kVarStageWords * 32, "the tile buffer holds T full windows");
  extern __shared__ uint64_t smem_raw[];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_base, s_tail;
  // layout: [plane matrices of all warps = tile buffer][columns of all warps]
  uint32_t sp_off = (threadIdx.x >> 5) * kPlaneBytes + (threadIdx.x & 31) * (uint32_t)sizeof(PW);
  uint32_t stage_off = W * kPlaneBytes + (threadIdx.x >> 5) * (kVarStageWords * 32 * 4u) + (threadIdx.x & 31) * 4u;
  asm volatile("" : "+r"(sp_off), "+r"(stage_off));
  PW* sp = reinterpret_cast<PW*>(reinterpret_cast<char*>(smem_raw) + sp_off);
  uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + stage_off);
  uint32_t* seg = reinterpret_cast<uint32_t*>(smem_raw);

  if (threadIdx.x == 0)
    s_tile = atomicAdd(ticket, 1u);  // tiles are numbered in the order their CTAs start: a predecessor is never behind us in the queue
  __syncthreads();
  const uint32_t tile = s_tile;

  // ---- 1. encode into the column window -------------------------------------------------------------
  const uint64_t b_raw = block0 + (uint64_t)tile * T + threadIdx.x;
  const bool valid = b_raw < block1;
  const uint64_t b = valid ? b_raw : block1 - 1;
  uint32_t bits;
  bool big;
  {
    const BlockPos<3> pos = locate<3>(g, b);
    typename TR::Scalar v[N];
    gather<3>(v, data, g, pos);
    ColWriter bw;
    bw.init(stage, nullptr, kVarStageWords, true);
    bits = encode_block<TYPE, 3, REV>(v, prm, bw, sp);
    bw.finish_window();
    big = bw.drained != 0;
  }
  if (valid)
    lengths[b] = (uint16_t)bits;
  else
    bits = 0;

  // ---- 2. concatenate the tile in shared memory ---------------------------------------------------------
  uint32_t total;
  const uint32_t off = cta_excl_scan(bits, total);  // (synchronises the CTA: every warp is done with its plane matrix)
  // a tile with so many long blocks that even its holes do not fit the buffer leaves ALL its blocks to the clean-up
  const bool flood = total + 128 > kSegWords * 32;
  big = big || flood;
  const uint32_t seg_words = flood ? 4u : ((total + 31) >> 5) + 3;
  for (uint32_t i = threadIdx.x; i < seg_words; i += T)
    seg[i] = 0;
  __syncthreads();
  if (valid && !big) {
    const uint32_t nw = (bits + 31) >> 5, sh = off & 31;
    uint32_t* dst = seg + (off >> 5);
    for (uint32_t j = 0; j < nw; j++) {
      uint32_t v = stage[j * 32];
      if (j == nw - 1 && (bits & 31))
        v &= (1u << (bits & 31)) - 1;  // the coder may run a few bits past the block's end in its last word
      if (v) {
        atomicOr(dst + j, v << sh);
        if (sh)
          atomicOr(dst + j + 1, v >> (32 - sh));
      }
    }
  }
  __syncthreads();

  // ---- 3. where does the tile start?  decoupled look-back -------------------------------------------------
  if (threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    unsigned long long tail = 0;
    if (lane == 0) {
      if (total >= 64 && !flood) {
        const uint32_t p = total - 64, s = p & 31;
        const uint32_t t0 = seg[p >> 5], t1 = seg[(p >> 5) + 1], t2 = seg[(p >> 5) + 2];
        tail = (unsigned long long)__funnelshift_r(t0, t1, s) | ((unsigned long long)__funnelshift_r(t1, t2, s) << 32);
      }
      status[tile].tail = tail;
      __threadfence();
      atomicExch(&status[tile].state, kVar1Agg | total);
    }
    unsigned long long base, pred_tail;
    if (tile == 0) {
      base = carry[0];
      pred_tail = carry[1];
    }
    else {
      base = 0;
      pred_tail = 0;
      bool have_tail = false;
      int look = (int)tile - 1;  // nearest predecessor not yet accounted for
      for (;;) {
        const int idx = look - (int)lane;
        // lanes that reach before tile 0 add nothing; tile 0 itself is waited for until it has its END position
        // (it takes the launch's carry-in and never waits for anybody), so every walk ends at an inclusive entry
        // and only tile 0 ever reads `carry` - which the last tile overwrites when it is done
        unsigned long long st = kVar1Agg;
        if (idx >= 0) {
          do {
            st = ld_volatile_u64(&status[idx].state);
          } while ((st >> 62) == 0 || (idx == 0 && (st >> 62) != 2));
        }
        if (!have_tail) {  // the immediate predecessor's tail (lane 0 of the first round)
          __threadfence();
          const unsigned long long t = idx >= 0 ? ld_volatile_u64(&status[idx].tail) : 0;
          pred_tail = __shfl_sync(0xffffffffu, t, 0);
          have_tail = true;
        }
        const unsigned inclusive = __ballot_sync(0xffffffffu, (st >> 62) == 2);
        const int first = inclusive ? __ffs((int)inclusive) - 1 : 32;  // nearest tile with an end position
        unsigned long long v = (int)lane <= first ? (st & kVar1Mask) : 0;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
          v += __shfl_xor_sync(0xffffffffu, v, d);
        base += v;
        if (inclusive)
          break;
        look -= 32;
      }
    }
    if (lane == 0) {
      atomicExch(&status[tile].state, kVar1Incl | ((base + total) & kVar1Mask));
      s_base = base;
      s_tail = pred_tail;
      // the last tile of the launch hands the position and the tail to the next launch
      if ((uint64_t)(tile + 1) * T >= block1 - block0) {
        if (total < 64) {  // a short last tile: its tail reaches into the predecessor's bits
          const unsigned long long mine = total ? (unsigned long long)seg[0] | ((unsigned long long)seg[1] << 32) : 0;
          tail = total ? (mine << (64 - total)) | (pred_tail >> total) : pred_tail;
        }
        carry[1] = tail;
        __threadfence();
        carry[0] = base + total;
      }
    }
  }
  __syncthreads();
  const uint64_t base = s_base, pred_tail = s_tail;

  // ---- 4. blocks that outgrew their window: remember where their (zeroed) holes are ------------------------
  if (valid && big) {
    const unsigned int slot = atomicAdd(overflow_count, 1u);  // (a count beyond the capacity tells the host the list is incomplete)
    if (slot < overflow_capacity) {
      overflow[slot].block = b;
      overflow[slot].bit = base + off;
    }
  }

  // ---- 5. the tile's words, shifted to its bit phase --------------------------------------------------------
  {
    const uint32_t o = (uint32_t)(base & 63);
    const uint64_t w0 = base >> 6;
    const uint32_t nout = (o + total + 63) >> 6;              // words holding bits of this tile
    const bool partial_last = ((o + total) & 63) != 0;
    const bool array_end = block0 + (uint64_t)(tile + 1) * T >= g.nblocks;  // nobody comes after: we own the last partial word too
    const uint64_t* seg64 = reinterpret_cast<const uint64_t*>(seg);
    for (uint32_t k = threadIdx.x; k < nout; k += T) {
      if (k == nout - 1 && partial_last && !array_end)
        break;  // belongs to the next tile, which has our tail
      const uint64_t cur = flood ? 0 : seg64[k];  // (zero beyond the tile: the buffer was cleared three words past its end)
      uint64_t v = cur;
      if (o) {
        const uint64_t prev = k ? (flood ? 0 : seg64[k - 1]) : pred_tail;
        v = (cur << o) | (prev >> (64 - o));
      }
      out[w0 + k] = v;
    }
  }
}

// second encode of the blocks the single pass could not hold: general coder, OR-merged into the zeroed holes
template <int TYPE, bool REV>
__global__ void __launch_bounds__(kThreads)
reencode_kernel(const typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, void* __restrict__ out,
                const Var1Overflow* __restrict__ overflow, const unsigned int* __restrict__ overflow_count,
                unsigned int overflow_capacity)
{
  using TR = Traits<TYPE>;
  constexpr int N = 64;
  using PW = typename PlaneWord<N>::type;
  extern __shared__ uint64_t smem_raw[];
  PW* sp = reinterpret_cast<PW*>(smem_raw) + (threadIdx.x >> 5) * (TR::P * 32) + (threadIdx.x & 31);
  const unsigned int count = *overflow_count < overflow_capacity ? *overflow_count : overflow_capacity;
  for (unsigned int i = blockIdx.x * kThreads + threadIdx.x; i < count; i += gridDim.x * kThreads) {
    const uint64_t b = overflow[i].block;
    const BlockPos<3> pos = locate<3>(g, b);
    typename TR::Scalar v[N];
    gather<3>(v, data, g, pos);
    BitWriter<1> bw;
    bw.init(out, overflow[i].bit);
    encode_block<TYPE, 3, REV>(v, prm, bw, sp);
    bw.flush();
  }
}

}  // namespace zb
