// Standalone microbenchmark: what can the exp pipeline of this GPU sustain, and which softmax inner body gets closest?
//   k_ex2        : pure ex2.approx throughput (8 independent chains per thread)
//   k_softmax<V> : the attention kernel's inner body per element, operands from registers, one 16-byte st.shared per
//                  8 elements.   V = 0  scalar body (FFMA, FMNMX, MUFU, FADD, F2FP)            -- attention v1
//                                V = 1  packed body (FFMA2, FMNMX3, MUFU, FADD2, F2FP)
//                                V = 2  packed body, 1/4 of the pairs through the FMA-pipe polynomial (packed)
//                                V = 3  packed body, 3/8 of the pairs through the polynomial
//                                V = 4  packed body, 1/2 of the pairs through the polynomial
//                                V = 5  packed body without the row sum (tensor-core row sums)
//                                V = 6  packed body, no exponential at all (p = x): the cost of everything else
//                                V = 7  packed body, MUFU, but no 16-bit pack (stores raw words): is F2FP the limiter?
//                                V = 8 / 9 / 10  lazy reference (max only on polynomial lanes), no row sum (tensor-core row sums),
//                                                1/4 / 3/8 / 1/2 polynomial pairs
//                                V = 11 / 12     lazy reference, row sum in the body, 1/4 (= the attention kernel today) / 3/8 polynomial
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench scripts/mufu_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { unsigned w; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a)); return w; }
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void un2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// packed exp2 of two already scaled exponents (Cody-Waite + degree-3 minimax), FMA/ALU pipes only
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  un2(x2, x0, x1);
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  x2 = pk2(x0, x1);
  const uint64_t magic = pk2(12582912.0f, 12582912.0f);
  const uint64_t t2 = add2(x2, magic);
  const uint64_t f2 = sub2(x2, sub2(t2, magic));
  uint64_t p = fma2(f2, pk2(0.05517145f, 0.05517145f), pk2(0.24261084f, 0.24261084f));
  p = fma2(p, f2, pk2(0.69326097f, 0.69326097f));
  p = fma2(p, f2, pk2(0.99992812f, 0.99992812f));
  float t0, t1, q0, q1;
  un2(t2, t0, t1);
  un2(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

__global__ void __launch_bounds__(512) k_ex2(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]) - 1.0f;
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V> __device__ __forceinline__ bool use_poly(int pair) {
  if (V == 2 || V == 8 || V == 11) return (pair & 3) == 3;
  if (V == 3 || V == 9 || V == 12) return (pair & 7) == 2 || (pair & 7) == 5 || (pair & 7) == 7;
  if (V == 4 || V == 10) return (pair & 1) == 1;
  return false;
}

template <int WARPS, int V>
__global__ void __launch_bounds__(WARPS * 32) k_softmax(float* out, int iters, float c) {
  __shared__ uint4 sm[WARPS * 32 * 4];
  __shared__ float4 vin[WARPS * 32 * 8];  // stands in for TMEM: the scores are re-read every iteration
  for (int i = 0; i < 8; ++i) {
    const float b = -0.01f * ((threadIdx.x * 7 + i * 13) % 97);
    vin[i * WARPS * 32 + threadIdx.x] = make_float4(b, b - 0.1f, b - 0.2f, b - 0.3f);
  }
  __syncthreads();
  float v[32];
  float tmax = -1e30f, psum = 0.f, mc = 0.5f;
  uint64_t psum2 = pk2(0.f, 0.f);
  const uint64_t c2 = pk2(c, c);
  for (int it = 0; it < iters; ++it) {
    unsigned pk[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t a = (uint32_t)__cvta_generic_to_shared(&vin[i * WARPS * 32 + threadIdx.x]);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3]) : "r"(a));
    }
    if (V == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float s0 = v[i], s1 = v[i + 1];
        tmax = fmaxf(tmax, fmaxf(s0, s1));
        const float p0 = ex2(fmaf(s0, c, -mc)), p1 = ex2(fmaf(s1, c, -mc));
        pk[i >> 1] = pack(p0, p1);
        psum += p0 + p1;
      }
    } else {
      const uint64_t nmc2 = pk2(-mc, -mc);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        if (V < 8 || use_poly<V>(i >> 1)) tmax = max3(tmax, v[i], v[i + 1]);  // V >= 8: lazy reference, guard on polynomial lanes only
        const uint64_t x2 = fma2(pk2(v[i], v[i + 1]), c2, nmc2);
        float p0, p1;
        if (use_poly<V>(i >> 1)) {
          ex2_poly2(x2, p0, p1);
        } else if (V == 6) {
          un2(x2, p0, p1);
        } else {
          float x0, x1;
          un2(x2, x0, x1);
          p0 = ex2(x0);
          p1 = ex2(x1);
        }
        pk[i >> 1] = V == 7 ? (__float_as_uint(p0) ^ __float_as_uint(p1)) : pack(p0, p1);
        if (V != 5 && (V < 8 || V > 10)) psum2 = add2(psum2, pk2(p0, p1));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) sm[u * WARPS * 32 + threadIdx.x] = make_uint4(pk[u * 4], pk[u * 4 + 1], pk[u * 4 + 2], pk[u * 4 + 3]);
    mc += 1e-6f;
  }
  float a, b;
  un2(psum2, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = psum + a + b + tmax + (float)sm[(threadIdx.x * 4 + 1) % (WARPS * 128)].x;
}

static float* out;
static cudaEvent_t e0, e1;
static int sms;

template <int W, int V>
static void run(int B) {
  const int iters = 2048, blocks = sms * B;
  k_softmax<W, V><<<blocks, W * 32>>>(out, 8, 0.36f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_softmax<W, V><<<blocks, W * 32>>>(out, iters, 0.36f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("softmax body V=%d %2d warps/SM: %.3f Texp/s (%.3f ms)\n", V, W * B, (double)blocks * W * 32 * iters * 32 / ms / 1e9, ms);
}

int main() {
  cudaMalloc(&out, 1 << 26);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  sms = p.multiProcessorCount;
  printf("%s: %d SMs, max clock %.0f MHz; nominal MUFU peak 16/clk/SM = %.2f Texp/s\n", p.name, sms, clk / 1e3, 16.0 * sms * clk * 1e3 / 1e12);
  {
    const int iters = 4096, blocks = sms * 4;
    k_ex2<<<blocks, 512>>>(out, 16); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_ex2<<<blocks, 512>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("pure ex2            : %.3f Texp/s (%.3f ms)\n", (double)blocks * 512 * iters * 8 / ms / 1e9, ms);
  }
  run<4, 0>(1); run<4, 0>(2); run<4, 0>(4); run<8, 0>(4);
  run<4, 1>(2); run<4, 1>(4); run<8, 1>(4);
  run<4, 2>(2); run<4, 2>(4); run<8, 2>(4);
  run<4, 3>(2); run<4, 3>(4); run<8, 3>(4);
  run<4, 4>(2); run<4, 4>(4); run<8, 4>(4);
  run<4, 5>(2); run<4, 5>(4); run<8, 5>(4);
  run<4, 6>(4); run<8, 6>(4);
  run<4, 7>(4); run<8, 7>(4);
  run<4, 11>(4); run<4, 12>(4); run<4, 8>(4); run<4, 9>(4); run<4, 10>(4); run<8, 11>(4); run<8, 8>(4); run<8, 9>(4); run<8, 10>(4);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
