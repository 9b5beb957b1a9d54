// kernels_ps.cuh - fixed-rate decode of 3-D blocks of 64-bit values with a PHASED register budget.
//
// decode_staged_kernel is two programs run back to back by the same thread: the stream parse of planes
// 63..32 (a serial chain per plane, about fifty registers) and everything after it (transposes, the
// remaining planes, inverse transform, stores: the block's 64 values = 128 registers of data).  Sized for
// the second (168 registers) only 12 warps fit an SM, and the kernel's time is close to proportional to
// 1 / warps (measured with shared-memory padding: 6 warps 7.96 ms, 12 warps 4.94 ms at 1024^3 rate 8).
//
// Here a CTA is kPsGroups warpgroups that loop over batches of 128 blocks, out of phase with each
// other.  A group gives registers back to the CTA's pool (setmaxnreg.dec) while it parses and takes them
// again (setmaxnreg.inc, behind a counter of large slots) before the register-heavy part, so the SM holds
// 16-20 warps on the registers 12 needed.  Same code per block as decode_staged_kernel (decode_block
// with the PS hooks); lossy fixed-rate parameters with word-aligned blocks only.
#pragma once

#include "kernels.cuh"

namespace zb {

#ifndef ZB_PS_GROUPS
#define ZB_PS_GROUPS 4
#endif
constexpr int kPsGroups = ZB_PS_GROUPS;
constexpr int kPsGroupThreads = ZB_PS_GROUP_THREADS;  // 128, or 256: two warpgroups that move together
constexpr int kPsThreads = kPsGroups * kPsGroupThreads;
// registers per thread: the launch gives 65536 / threads (rounded down to 8); pool = groups x that
//   4 groups: 128 at launch = 2 x 56 + 2 x 200;  5 groups: 96 at launch, 3 x 40 + 2 x 176 <= 480
//   (the parse phase compiles to 33 registers)
#ifndef ZB_PS_SMALL
#define ZB_PS_SMALL (ZB_PS_GROUPS == 4 ? 56 : 40)
#endif
#ifndef ZB_PS_BIG
#define ZB_PS_BIG (ZB_PS_GROUPS == 4 ? 200 : 176)
#endif
#ifndef ZB_PS_LARGE
#define ZB_PS_LARGE 2
#endif
constexpr int kPsLarge = ZB_PS_LARGE;  // groups the pool can hold at kPsBig next to the others at kPsSmall
constexpr int kPsSmall = ZB_PS_SMALL;
constexpr int kPsBig = ZB_PS_BIG;

template <int TYPE>
__global__ void __launch_bounds__(kPsThreads, 1)
decode_ps_kernel(typename Traits<TYPE>::Scalar* __restrict__ data, Geom g, Params prm, const uint64_t* __restrict__ in,
                 uint64_t start_bit, uint64_t block0, uint64_t block1)
{
  using TR = Traits<TYPE>;
  constexpr int DIMS = 3, N = 64;
  using PW = uint64_t;
  extern __shared__ uint64_t smem_raw[];
  __shared__ uint32_t run_table[32];
  if (threadIdx.x < 32)
    ColReader::fill_run_table(run_table, threadIdx.x);
  if (threadIdx.x == 0)
    *wg_large_slots() = kPsLarge;
  __syncthreads();
  const uint32_t words = prm.maxbits >> 5;
  const uint32_t warp_bytes = kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (words + kReadSlack) * 32 * 4;
  uint32_t sp_off = (threadIdx.x >> 5) * warp_bytes + (threadIdx.x & 31) * (uint32_t)sizeof(PW);
  uint32_t stage_off = (threadIdx.x >> 5) * warp_bytes + kStagedPlanes * 32 * (uint32_t)sizeof(PW) + (threadIdx.x & 31) * 4u;
  asm volatile("" : "+r"(sp_off), "+r"(stage_off));

  const uint64_t nbatches = (block1 - block0 + kPsGroupThreads - 1) / kPsGroupThreads;
  wg_reg_release<kPsSmall>();
  for (uint64_t batch = (uint64_t)blockIdx.x * kPsGroups + threadIdx.x / kPsGroupThreads; batch < nbatches; batch += (uint64_t)gridDim.x * kPsGroups) {
    PW* sp = reinterpret_cast<PW*>(reinterpret_cast<char*>(smem_raw) + sp_off);
    uint32_t* stage = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_raw) + stage_off);
    const uint64_t b_raw = block0 + batch * kPsGroupThreads + threadIdx.x % kPsGroupThreads;
    const bool valid = b_raw < block1;
    const uint64_t b_list = valid ? b_raw : block1 - 1;
    const uint64_t b = g.box ? box_block(g, b_list) : b_list;
    const uint4* src4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(in + (start_bit >> 6)) + b * (uint64_t)words);
#pragma unroll 4
    for (uint32_t w = 0; w < words; w += 4) {  // (the launcher checks: whole 128-bit groups, aligned)
      const uint4 v = __ldg(src4 + (w >> 2));
      stage[w * 32] = v.x;
      stage[(w + 1) * 32] = v.y;
      stage[(w + 2) * 32] = v.z;
      stage[(w + 3) * 32] = v.w;
    }
#pragma unroll
    for (int j = 0; j < kReadSlack; j++)
      stage[(words + j) * 32] = 0;

    ColReader br;
    br.init(stage);
    br.set_run_table(run_table);
    typename TR::Scalar v[N];
    decode_block<TYPE, DIMS, false, ColReader, kPsBig>(v, prm, br, sp);
    if (valid) {
      const BlockPos<DIMS> pos = locate<DIMS>(g, b);
      scatter<DIMS>(v, data, g, pos);
    }
    wg_leave_large<kPsSmall>();  // (also before leaving: registers of a finished group stay with it, not with the pool)
  }
}

inline size_t ps_cta_bytes(uint32_t words)
{
  return (size_t)(kPsThreads / 32) * (kStagedPlanes * 32 * sizeof(uint64_t) + (words + kReadSlack) * 32 * 4);
}

}  // namespace zb
