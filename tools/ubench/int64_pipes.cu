// Microbenchmark: 64-bit integer building blocks of the lifting transform on the two integer pipes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int64_pipes int64_pipes.cu && ./int64_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define ILP 8

__device__ __forceinline__ uint64_t add64_wide(uint64_t a, uint64_t b, uint32_t one)
{
  // a + b: IMAD.WIDE.U32 adds b.lo (times 1) to the 64-bit a with the carry, then the high words
  uint64_t t;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"((uint32_t)b), "r"(one), "l"(a));
  uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32) + (uint32_t)(b >> 32);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ uint64_t sub64_wide(uint64_t a, uint64_t b, uint32_t ones)
{
  // a - b: a + b.lo * (2^32 - 1) = a - b.lo + (b.lo << 32); fix the high word with one 3-input add
  uint64_t t;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"((uint32_t)b), "r"(ones), "l"(a));
  uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32) - (uint32_t)b - (uint32_t)(b >> 32);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

template <int V>
__device__ __forceinline__ void lift(int64_t& x, int64_t& y, int64_t& z, int64_t& w, uint32_t one, uint32_t ones)
{
  using U = uint64_t;
  if (V == 0) {
    x = (int64_t)((U)x + (U)w); x >>= 1; w = (int64_t)((U)w - (U)x);
    z = (int64_t)((U)z + (U)y); z >>= 1; y = (int64_t)((U)y - (U)z);
    x = (int64_t)((U)x + (U)z); x >>= 1; z = (int64_t)((U)z - (U)x);
    w = (int64_t)((U)w + (U)y); w >>= 1; y = (int64_t)((U)y - (U)w);
    w = (int64_t)((U)w + (U)(y >> 1)); y = (int64_t)((U)y - (U)(w >> 1));
  }
  else {
    x = (int64_t)add64_wide((U)x, (U)w, one); x >>= 1; w = (int64_t)sub64_wide((U)w, (U)x, ones);
    z = (int64_t)add64_wide((U)z, (U)y, one); z >>= 1; y = (int64_t)sub64_wide((U)y, (U)z, ones);
    x = (int64_t)add64_wide((U)x, (U)z, one); x >>= 1; z = (int64_t)sub64_wide((U)z, (U)x, ones);
    w = (int64_t)add64_wide((U)w, (U)y, one); w >>= 1; y = (int64_t)sub64_wide((U)y, (U)w, ones);
    w = (int64_t)add64_wide((U)w, (U)(y >> 1), one); y = (int64_t)sub64_wide((U)y, (U)(w >> 1), ones);
  }
}

template <int KIND>
__global__ void k(uint64_t* out, uint32_t a0, uint32_t one, uint32_t ones)
{
  uint64_t x[ILP];
  const uint64_t y0 = ((uint64_t)a0 << 33) + threadIdx.x * 77u;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = ((uint64_t)(a0 + i) << 35) + i + threadIdx.x;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      const uint64_t y = x[(i + 3) % ILP] ^ y0;  // varies per element and iteration: nothing to hoist
      if (KIND == 0) x[i] += y;                                                   // IADD3 + IADD3.X
      if (KIND == 1) x[i] = add64_wide(x[i], y, one);                             // IMAD.WIDE + IADD
      if (KIND == 2) x[i] = sub64_wide(x[i], y, ones);                            // IMAD.WIDE + IADD3
      if (KIND == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((uint32_t)y), "r"(one));  // IMAD.WIDE alone
      if (KIND == 4) x[i] = (uint64_t)((int64_t)x[i] >> 1) + 0x4000000000000001ull * (i + 1);  // 64-bit asr + add
      if (KIND == 5) { uint32_t lo = (uint32_t)x[i], hi = (uint32_t)(x[i] >> 32);
                       asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo) : "r"(ones));
                       x[i] = (uint64_t)lo | ((uint64_t)hi << 32); }                 // IMAD.HI
      if (KIND == 6) { double d = __longlong_as_double((long long)x[i]);
                       long long r; asm volatile("cvt.rzi.s64.f64 %0, %1;" : "=l"(r) : "d"(d));
                       x[i] = (uint64_t)r ^ y; }                                     // F2I.S64.F64 (+2 LOP3)
      if (KIND == 7) { double d = __longlong_as_double((long long)x[i]);
                       asm volatile("mul.f64 %0, %0, %1;" : "+d"(d) : "d"(__longlong_as_double((long long)y)));
                       x[i] = (uint64_t)__double_as_longlong(d); }                   // DMUL
      if (KIND == 8) { uint32_t lo = (uint32_t)x[i], hi = (uint32_t)(x[i] >> 32), m = (uint32_t)y;
                       lo = max(max(lo, hi), m); x[i] = (uint64_t)lo | ((uint64_t)(hi + 1) << 32); }  // VIMNMX3?
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 48 lifts on a register-resident 4x4x4 block of int64, repeated
template <int V>
__global__ void lift_kernel(int64_t* data, uint32_t one, uint32_t ones, int reps)
{
  int64_t p[64];
  int64_t* base = data + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 64;
#pragma unroll
  for (int i = 0; i < 64; i++) p[i] = base[i];
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int st = 1 << (2 * a);
#pragma unroll
      for (int i = 0; i < 64; i++)
        if (((i >> (2 * a)) & 3) == 0)
          lift<V>(p[i], p[i + st], p[i + 2 * st], p[i + 3 * st], one, ones);
    }
  }
#pragma unroll
  for (int i = 0; i < 64; i++) base[i] = p[i];
}

template <int KIND>
void run(const char* name, double ops_per)
{
  uint64_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 8; warps <= 32; warps *= 2) {
    int ctas = 148, threads = warps * 32;
    k<KIND><<<ctas, threads>>>(d, 1, 1, 0xffffffffu); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<KIND><<<ctas, threads>>>(d, 1, 1, 0xffffffffu); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)148 * warps * ITER * ILP * ops_per;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s warps/SM %2d: %.3f ops/clk/SM (warp-wide ops, %d MHz assumed)\n", name, warps, inst / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
  }
  cudaFree(d);
}

template <int V>
void run_lift(const char* name)
{
  const int ctas = 148 * 4, threads = 128, reps = 64;
  int64_t* d; cudaMalloc(&d, (size_t)ctas * threads * 64 * 8); cudaMemset(d, 1, (size_t)ctas * threads * 64 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  lift_kernel<V><<<ctas, threads>>>(d, 1, 0xffffffffu, reps); cudaDeviceSynchronize();
  cudaEventRecord(e0); lift_kernel<V><<<ctas, threads>>>(d, 1, 0xffffffffu, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warp_blocks = (double)ctas * threads / 32 * reps;  // 3-D transforms per warp
  double clks_per = (ms * 1e-3) * (clk * 1e3) * 148 * 4 / warp_blocks;  // SMSP-clocks per warp-wide 3-D transform
  printf("%-28s %.1f SMSP-clk per warp-wide 3-D fwd transform (48 lifts), %.3f ms\n", name, clks_per, ms);
  cudaFree(d);
}

int main()
{
  run<0>("add64 IADD3+IADD3.X", 1); run<1>("add64 IMAD.WIDE+IADD", 1); run<2>("sub64 IMAD.WIDE+IADD3", 1);
  run<3>("IMAD.WIDE.U32", 1); run<4>("asr64 + add64", 1); run<5>("IMAD.HI.U32", 1); run<6>("F2I.S64.F64 (+2 LOP3)", 1);
  run<7>("DMUL", 1); run<8>("max3", 1);
  run_lift<0>("lift3d native"); run_lift<1>("lift3d imad.wide");
  return 0;
}
