// Shared helpers for the sgb200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sgb200.h"

namespace sg {

// ---- error plumbing (thread-local last-error string, returned through the C ABI) -------------
void set_error(const char* fmt, ...);
int launch_status(const char* what);  // cudaGetLastError -> SG_OK / SG_ERR_LAUNCH
int num_sms();                        // SM count of the current device (persistent-kernel grid sizing)

#define SG_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ::sg::set_error(__VA_ARGS__);  \
      return SG_ERR_ARG;             \
    }                                \
  } while (0)

static inline cudaStream_t as_stream(sg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Per-device caches are indexed by the CUDA ordinal of the calling thread's current device (one rank = one GPU is the
// normal case, but nothing here is wrong for a process that drives several devices).
constexpr int SG_MAX_DEVICES = 64;
static inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < SG_MAX_DEVICES ? dev : 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-kernel attribute: set once for each (kernel, device).
template <auto Kernel>
static inline int set_max_smem(int bytes, const char* what) {
  static bool done[SG_MAX_DEVICES] = {};
  const int dev = current_device();
  if (done[dev]) return SG_OK;
  cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d B smem): %s", what, bytes, cudaGetErrorString(e));
    return SG_ERR_LAUNCH;
  }
  done[dev] = true;
  return SG_OK;
}
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// Every kernel of the sampling step is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins
//   [prologue that touches no global memory: barrier init, TMEM allocation, tensor-map prefetch]
//   pdl_wait();                 // ALL threads: the preceding grid has completed and its writes are visible
//   pdl_launch_dependents();    // the next grid's CTAs may take the SM slots this grid frees at its tail
// so the launch latency and prologue of kernel k+1 overlap the last wave of kernel k (in a captured graph the edges
// become programmatic dependencies).  Rule: no global-memory access of any kind before pdl_wait().
// Measured (B200): 5 % lower latency per CFG step at batch 1-8 (1.06 -> 1.01 ms), break-even around batch 64, and 2-3 %
// SLOWER at batch 256-512 when every launch carries the attribute -- long kernels gain nothing from the early launch and
// each programmatic edge costs a few microseconds -- so by default only grids of at most 4 x #SM CTAs are launched that way
// (sg_set_pdl: 2; 1 = every launch, 0 = none: everything fully serialised and the waits are no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_mode();  // abi.cu (sg_set_pdl, thread-local): 0 off, 1 every launch, 2 (default) only grids of at most pdl_max_ctas() CTAs
int pdl_max_ctas();

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = pdl_mode();
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  cfg.numAttrs = (mode == 1 || (mode == 2 && ctas <= pdl_max_ctas())) ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through launch_status()
}

// ---- activation-type load/store ---------------------------------------------------------------
template <typename T>
struct ActIO;

template <>
struct ActIO<float> {
  static __device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
  static __device__ __forceinline__ void store2(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
  }
};

template <>
struct ActIO<__nv_bfloat16> {
  static __device__ __forceinline__ void store4(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
    __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo);
    v.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = v;
  }
  static __device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

template <>
struct ActIO<__half> {
  static __device__ __forceinline__ void store4(__half* p, float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b);
    __half2 hi = __floats2half2_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo);
    v.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = v;
  }
  static __device__ __forceinline__ void store2(__half* p, float a, float b) {
    *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b);
  }
};

// pack two fp32 into one 32-bit word of 16-bit values (lo = a), runtime format
__device__ __forceinline__ uint32_t pack16(float a, float b, int dtype) {
  if (dtype == SG_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  uint32_t w;  // saturating: a raw fp16 value beyond +-65504 must not become inf (it feeds a normalisation)
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  return w;
}
__device__ __forceinline__ float2 unpack16(uint32_t w, int dtype) {
  if (dtype == SG_BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}
// ---- operands of the split-TF32 (fp32-accurate tensor-core) engine ----
// tcgen05 kind::tf32 reads the top 19 bits of each fp32 container, i.e. it multiplies trunc19(x).  Two operand forms:
//   weights (packed once):   hi = rna_tf32(w), lo = rna_tf32(w - hi)                      (sg_split_tf32 with hi != NULL)
//   activations:             hi = the fp32 tensor ITSELF, lo = rna_tf32(x - trunc19(x))   (tf32_lo below)
// x - trunc19(x) is exact (13 significant bits); rounding it to TF32 leaves |x - (trunc19(x) + lo)| <= 2^-22 |x|, unbiased.
// The activation form costs one extra 4-byte store in the kernel that produces x and no separate pass over it.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_lo(float x) {
  return tf32_rna(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}
__device__ __forceinline__ float4 tf32_lo4(float a, float b, float c, float d) {
  return make_float4(tf32_lo(a), tf32_lo(b), tf32_lo(c), tf32_lo(d));
}

// store 4 consecutive values to an fp32 and/or a second destination at element offset `off`: 16-bit copies (dtype
// SG_BF16 / SG_F16) or, dtype == SG_F32, the TF32 low parts of the values
__device__ __forceinline__ void store4_dual(float* o32, void* o16, int dtype, int64_t off, float a, float b, float c,
                                            float d) {
  if (o32) *reinterpret_cast<float4*>(o32 + off) = make_float4(a, b, c, d);
  if (o16 && dtype == SG_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(o16) + off) = tf32_lo4(a, b, c, d);
  } else if (o16) {
    uint2 v;
    v.x = pack16(a, b, dtype);
    v.y = pack16(c, d, dtype);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(o16) + off) = v;
  }
}

// exact (erf) GELU, as torch.nn.GELU() default
// nn.GELU() (exact, erf-based; /root/reference/src/diff_modules.py:84,91).  erff() costs ~40 instructions with two
// divergent branches, which made every GELU-carrying "memory-bound" kernel compute-bound (gn_apply: 3.3 TB/s with
// GELU, 5.4 TB/s without).  Branch-free form: erfc(z) = (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z = |x| / sqrt 2
// (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7), gelu(x) = x/2 * (x >= 0 ? 2 - erfc(z) : erfc(z)) -- no cancellation on
// the negative side.  Measured against fp64 over [-12, 12]: max abs error 4.2e-7 (torch's own fp32 GELU: 1.2e-6).
// packed fp32 arithmetic of sm_100 (two IEEE lanes per instruction: the same results as the scalar forms at half the
// issue slots); a pair lives in one 64-bit register
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void un2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// Every operation is an explicitly rounded intrinsic so that the scalar and the packed form below give bit-identical
// results (kernels pick one or the other depending on how many values a thread holds).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = __fmul_rn(fabsf(x), 0.70710678118654752440f);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(__fmaf_rn(0.3275911f, z, 1.0f)));
  float p = __fmaf_rn(1.061405429f, t, -1.453152027f);
  p = __fmaf_rn(p, t, 1.421413741f);
  p = __fmaf_rn(p, t, -0.284496736f);
  p = __fmaf_rn(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(__fmul_rn(z, z), -1.4426950408889634f)));
  const float E = __fmul_rn(__fmul_rn(p, t), e);  // erfc(|x| / sqrt 2)
  // gelu(x) = x Phi(x) = relu(x) - |x|/2 * erfc(|x| / sqrt 2)   (x >= 0: x - x E / 2;  x < 0: x E / 2)
  return __fmaf_rn(__fmul_rn(fabsf(x), -0.5f), E, fmaxf(x, 0.f));
}
// two values at once on the packed instructions: 15 issue slots + 4 MUFU per pair instead of 26 + 4
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t a2 = pk2(fabsf(x0), fabsf(x1));
  const uint64_t z2 = mul2(a2, pk2(0.70710678118654752440f, 0.70710678118654752440f));
  float d0, d1, t0, t1;
  un2(fma2(pk2(0.3275911f, 0.3275911f), z2, pk2(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t2 = pk2(t0, t1);
  uint64_t p = fma2(pk2(1.061405429f, 1.061405429f), t2, pk2(-1.453152027f, -1.453152027f));
  p = fma2(p, t2, pk2(1.421413741f, 1.421413741f));
  p = fma2(p, t2, pk2(-0.284496736f, -0.284496736f));
  p = fma2(p, t2, pk2(0.254829592f, 0.254829592f));
  float a0, a1, e0, e1;
  un2(mul2(mul2(z2, z2), pk2(-1.4426950408889634f, -1.4426950408889634f)), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const uint64_t E2 = mul2(mul2(p, t2), pk2(e0, e1));
  const uint64_t nh2 = mul2(a2, pk2(-0.5f, -0.5f));
  un2(fma2(nh2, E2, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}

// GELU for the 16-bit engines' GroupNorm-apply passes (4 bytes of traffic per element: the erfc form above made them
// issue-bound at 72 % of the HBM rate).  gelu(x) = x Phi(x) = x / (1 + 2^(x r(x^2))), where -x r(x^2) log2(e) is an odd
// degree-9 polynomial fitted (minimax, scripts/fit_gelu.py) to logit Phi(x): 8 packed instructions + 4 MUFU per pair
// instead of 15 + 4.  |error| <= 3.5e-6 absolute over all x against fp64 (fp32 evaluation included) -- 1/70 of an fp16 ulp
// at |y| = 0.5; r > 0 and x r(x^2) is monotonic, so large |x| saturate to x and -0 without a clamp.  The fp32 engines
// keep the erfc form (4e-7).
__device__ __forceinline__ void gelu_logistic2(float& x0, float& x1) {
  const uint64_t x2 = pk2(x0, x1);
  const uint64_t t2 = mul2(x2, x2);
  uint64_t p = fma2(pk2(-3.22899314e-06f, -3.22899314e-06f), t2, pk2(8.82382083e-05f, 8.82382083e-05f));
  p = fma2(p, t2, pk2(0.000360274193f, 0.000360274193f));
  p = fma2(p, t2, pk2(-0.105226688f, -0.105226688f));
  p = fma2(p, t2, pk2(-2.30204535f, -2.30204535f));
  float a0, a1, e0, e1, r0, r1;
  un2(mul2(p, x2), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  un2(add2(pk2(e0, e1), pk2(1.0f, 1.0f)), a0, a1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(a0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(a1));
  un2(mul2(x2, pk2(r0, r1)), x0, x1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-tile GroupNorm partial sums.
// `rowstat` is shared memory [BM][2] holding (sum, sumsq) of each tile row over this tile's
// columns.  Rows of the tile are M-consecutive pixels; a tile holds BM/HW whole samples when
// HW < BM, else it lies inside one sample.  Writes partials[sample][p][2] deterministically.
//   P       partials per sample,  p_base  = partial slot of this tile for a sample it covers
template <int BM>
__device__ __forceinline__ void write_tile_partials(const float (*rowstat)[2], int tid, int m0, int M, int HW,
                                                    float* partials, int P, int tile_n, int n_tiles) {
  const int group = HW < BM ? HW : BM;  // rows per sample inside this tile
  const int groups = BM / group;
  if (tid < groups) {
    const int r0 = tid * group;
    if (m0 + r0 < M) {
      float s = 0.f, q = 0.f;
      for (int r = 0; r < group; ++r) {
        s += rowstat[r0 + r][0];
        q += rowstat[r0 + r][1];
      }
      const int sample = (m0 + r0) / HW;
      const int tile_in_sample = HW < BM ? 0 : ((m0 % HW) / BM);
      const int p = tile_in_sample * n_tiles + tile_n;
      partials[((int64_t)sample * P + p) * 2 + 0] = s;
      partials[((int64_t)sample * P + p) * 2 + 1] = q;
    }
  }
}

// ---- Philox4x32-10 + Box-Muller ------------------------------------------------------------------
struct Philox4 {
  uint32_t v[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
// four N(0,1) draws for elements [4*q, 4*q+4) of (sample, step)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t sample, uint32_t step, uint32_t q) {
  Philox4 r = philox4x32_10(q, step, (uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
  const float k = 2.3283064365386963e-10f;  // 2^-32
  const float u0 = ((float)r.v[0] + 0.5f) * k, u1 = ((float)r.v[1] + 0.5f) * k;
  const float u2 = ((float)r.v[2] + 0.5f) * k, u3 = ((float)r.v[3] + 0.5f) * k;
  const float r0 = sqrtf(-2.0f * logf(fmaxf(u0, 1e-30f))), r1 = sqrtf(-2.0f * logf(fmaxf(u2, 1e-30f)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

}  // namespace sg
