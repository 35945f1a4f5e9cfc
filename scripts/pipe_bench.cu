// Per-instruction throughput on sm_100a: ops / clk / SM for the instructions the softmax inner loops are made of.
// Each thread runs 8 independent dependency chains of one instruction; 4 blocks x 512 threads per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_bench scripts/pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum Op { EX2_F32, EX2_F16X2, EX2_BF16X2, FFMA, FFMA2, FADD2, FMNMX3, F2FP, LEA, IMAD, FMUL2, RCP };

template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, float c) {
  float a[8];
  uint32_t u[8];
  uint64_t w[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i) - 0.5f;
    u[i] = 0x3c003c00u + threadIdx.x + i;
    w[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f);
  }
  const uint64_t cc = ((uint64_t)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == EX2_F32) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == EX2_F16X2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(c));
      if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[i]) : "l"(cc));
      if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(cc));
      if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(cc));
      if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(a[(i + 1) & 7]));
      if (OP == F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
      if (OP == LEA) asm volatile("{.reg .u32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1;}" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, 8388608, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s += __float_as_uint(a[i]) + u[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static uint32_t* out;
static int sms;
static double ghz;

template <int OP>
static void run(const char* name, int per_instr) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2048, blocks = sms * 4;
  k<OP><<<blocks, 512>>>(out, 16, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP><<<blocks, 512>>>(out, iters, 0.999f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = (double)blocks * 512 * iters * 8;
  printf("%-12s %8.2f thread-instr/clk/SM  (%6.2f elem/clk/SM)  %.3f ms\n", name, instr / (ms * 1e-3) / (ghz * 1e9) / sms,
         instr * per_instr / (ms * 1e-3) / (ghz * 1e9) / sms, ms);
}

int main() {
  cudaMalloc(&out, 1 << 26);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  sms = p.multiProcessorCount;
  ghz = clk / 1e6;
  printf("%s: %d SMs, rates computed at the max clock %.3f GHz (actual clock may be lower under load)\n", p.name, sms, ghz);
  run<EX2_F32>("ex2.f32", 1);
  run<RCP>("rcp.f32", 1);
  run<EX2_F16X2>("ex2.f16x2", 2);
  run<EX2_BF16X2>("ex2.bf16x2", 2);
  run<FFMA>("fma.f32", 1);
  run<FFMA2>("fma.f32x2", 2);
  run<FADD2>("add.f32x2", 2);
  run<FMUL2>("mul.f32x2", 2);
  run<FMNMX3>("max3.f32", 2);
  run<F2FP>("cvt.bf16x2", 2);
  run<LEA>("shl+add", 1);
  run<IMAD>("mad.lo", 1);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
