// Microbenchmark: issue rate of the integer instruction classes the codec kernels are made of.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define ILP 8

template <int KIND>
__global__ void k(uint32_t* out, uint32_t a0, uint32_t b0, uint32_t c0)
{
  uint32_t x[ILP], y = b0 + threadIdx.x, z = c0;
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = a0 + i + threadIdx.x;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (KIND == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));                  // IADD3 2-src
      if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z)); // LOP3 3-src
      if (KIND == 2) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z)); // SHF 3-src
      if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));    // IMAD
      if (KIND == 4) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
                       else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z)); }  // IMAD + LOP3 mix
      if (KIND == 5) { if (i & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
                       else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z)); }  // IADD + LOP3 mix
      if (KIND == 6) asm volatile("{.reg .u32 t; add.cc.u32 %0, %0, %1; addc.u32 t, %0, %1; xor.b32 %0, %0, t;}" : "+r"(x[i]) : "r"(y)); // carry chain
      if (KIND == 7) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[i]) : "r"(y));            // PRMT
      if (KIND == 8) asm volatile("{.reg .u32 t; bfind.u32 t, %0; add.u32 %0, %0, t;}" : "+r"(x[i]));  // FLO (+IADD)
      if (KIND == 9) asm volatile("{.reg .u32 t; popc.b32 t, %0; add.u32 %0, %0, t;}" : "+r"(x[i]));   // POPC (+IADD)
      if (KIND == 10) { if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
                        else asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y)); }            // IMAD + IADD mix
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int ops_per)
{
  uint32_t* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {   // warps per SM
    int ctas = 148, threads = warps * 32;
    if (threads > 1024) { ctas *= threads / 1024; threads = 1024; }
    k<KIND><<<ctas, threads>>>(d, 1, 2, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<KIND><<<ctas, threads>>>(d, 1, 2, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)148 * warps * ITER * ILP * ops_per;   // warp-instructions
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-22s warps/SM %2d: %.3f warp-instr/clk/SM (assuming %d MHz)\n", name, warps, inst / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
  }
  cudaFree(d);
}

int main()
{
  run<0>("IADD (2 src)", 1); run<1>("LOP3 (3 src)", 1); run<2>("SHF (3 src)", 1); run<3>("IMAD", 1);
  run<4>("IMAD+LOP3 mix", 1); run<5>("IADD+LOP3 mix", 1); run<6>("ADD.CC+ADDC+XOR", 3); run<7>("PRMT", 1);
  run<8>("FLO+IADD", 2); run<9>("POPC+IADD", 2); run<10>("IMAD+IADD mix", 1);
  return 0;
}
