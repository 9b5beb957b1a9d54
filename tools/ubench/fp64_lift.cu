// Microbenchmark: the 3-D INVERSE transform (48 lifts, decode.c:13-45 of the reference's template) on
// int64 registers against the same arithmetic on the FP64 pipe.  When every coefficient of a block is a
// multiple of 2^L with L >= 16 and nothing leaves the int64 range, all intermediates of the inverse
// lift are exactly representable doubles (at most six halvings, none of which truncates), so
// "y += w >> 1" is one DFMA and "w <<= 1; w -= y" another.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lift fp64_lift.cu && ./fp64_lift
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void inv_lift_i(int64_t& x, int64_t& y, int64_t& z, int64_t& w)
{
  using U = uint64_t;
  y = (int64_t)((U)y + (U)(w >> 1)); w = (int64_t)((U)w - (U)(y >> 1));
  y = (int64_t)((U)y + (U)w); w = (int64_t)((U)w << 1); w = (int64_t)((U)w - (U)y);
  z = (int64_t)((U)z + (U)x); x = (int64_t)((U)x << 1); x = (int64_t)((U)x - (U)z);
  y = (int64_t)((U)y + (U)z); z = (int64_t)((U)z << 1); z = (int64_t)((U)z - (U)y);
  w = (int64_t)((U)w + (U)x); x = (int64_t)((U)x << 1); x = (int64_t)((U)x - (U)w);
}
__device__ __forceinline__ void inv_lift_d(double& x, double& y, double& z, double& w)
{
  y = fma(w, 0.5, y); w = fma(y, -0.5, w);
  y += w; w = fma(w, 2.0, -y);
  z += x; x = fma(x, 2.0, -z);
  y += z; z = fma(z, 2.0, -y);
  w += x; x = fma(x, 2.0, -w);
}

template <int V>
__global__ void __launch_bounds__(128) inv_kernel(int64_t* data, int reps)
{
  int64_t* base = data + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 64;
  if (V == 0) {
    int64_t p[64];
#pragma unroll
    for (int i = 0; i < 64; i++) p[i] = base[i];
    for (int r = 0; r < reps; r++) {
#pragma unroll
      for (int a = 2; a >= 0; a--) {
        const int st = 1 << (2 * a);
#pragma unroll
        for (int i = 0; i < 64; i++)
          if (((i >> (2 * a)) & 3) == 0) inv_lift_i(p[i], p[i + st], p[i + 2 * st], p[i + 3 * st]);
      }
    }
#pragma unroll
    for (int i = 0; i < 64; i++) base[i] = p[i];
  }
  else {
    double p[64];
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) p[i] = (double)base[i];  // I2F.F64.S64
    for (int r = 0; r < reps; r++) {
      if (V == 2) {  // the weighted L1 bound that proves nothing leaves the int64 range
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          s0 = fma(fabs(p[i]), 1.0 + (i & 3), s0); s1 = fma(fabs(p[i + 1]), 1.5, s1);
          s2 = fma(fabs(p[i + 2]), 1.25, s2); s3 = fma(fabs(p[i + 3]), 1.875, s3);
        }
      }
#pragma unroll
      for (int a = 2; a >= 0; a--) {
        const int st = 1 << (2 * a);
#pragma unroll
        for (int i = 0; i < 64; i++)
          if (((i >> (2 * a)) & 3) == 0) inv_lift_d(p[i], p[i + st], p[i + 2 * st], p[i + 3 * st]);
      }
    }
    p[0] += (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int i = 0; i < 64; i++) base[i] = (int64_t)p[i];
  }
}

// an ALU-bound filler running next to the FP64 work: does the FP64 pipe issue beside a busy ALU pipe?
template <int MIX>
__global__ void __launch_bounds__(128) mix_kernel(uint64_t* out, uint32_t a0, int iters)
{
  uint32_t a[8];
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = a0 + i * 77u + threadIdx.x; d[i] = 1.0 + i + threadIdx.x; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MIX & 1) { a[i] = (a[i] ^ a[(i + 3) & 7]) + (a[(i + 5) & 7] >> 3); a[i] = __funnelshift_l(a[i], a[(i + 1) & 7], 7) & ~a[(i + 2) & 7]; }  // 4 ALU
      if (MIX & 2) { d[i] = fma(d[(i + 3) & 7], 0.5, d[i]); d[i] = fma(d[i], 2.0, -d[(i + 5) & 7]); }                                           // 2 FP64
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= a[i] ^ (uint64_t)__double_as_longlong(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run_inv(const char* name)
{
  const int ctas = 148 * 4, threads = 128, reps = 64;
  int64_t* d; cudaMalloc(&d, (size_t)ctas * threads * 64 * 8); cudaMemset(d, 0, (size_t)ctas * threads * 64 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  inv_kernel<V><<<ctas, threads>>>(d, reps); cudaDeviceSynchronize();
  cudaEventRecord(e0); inv_kernel<V><<<ctas, threads>>>(d, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warp_blocks = (double)ctas * threads / 32 * reps;
  double clks_per = (ms * 1e-3) * (clk * 1e3) * 148 * 4 / warp_blocks;
  printf("%-34s %.1f SMSP-clk per warp-wide 3-D inverse transform, %.3f ms\n", name, clks_per, ms);
  cudaFree(d);
}

template <int MIX>
void run_mix(const char* name)
{
  const int ctas = 148 * 4, threads = 128, iters = 4096;
  uint64_t* d; cudaMalloc(&d, (size_t)ctas * threads * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix_kernel<MIX><<<ctas, threads>>>(d, 1, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); mix_kernel<MIX><<<ctas, threads>>>(d, 1, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warp_iters = (double)ctas * threads / 32 * iters * 8;
  double clks_per = (ms * 1e-3) * (clk * 1e3) * 148 * 4 / warp_iters;
  printf("%-34s %.2f SMSP-clk per warp per group (4 ALU and / or 2 DFMA), %.3f ms\n", name, clks_per, ms);
  cudaFree(d);
}

int main()
{
  run_inv<0>("inverse 3-D, int64");
  run_inv<1>("inverse 3-D, fp64");
  run_inv<2>("inverse 3-D, fp64 + range bound");
  run_mix<1>("4 ALU");
  run_mix<2>("2 DFMA");
  run_mix<3>("4 ALU + 2 DFMA");
  return 0;
}
