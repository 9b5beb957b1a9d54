// scan_util.cuh - warp / CTA prefix sums shared by the length scan (stream_kernels.cuh) and the single-pass
// variable-rate encoder (kernels_var1.cuh).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace zb {

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

}  // namespace zb
