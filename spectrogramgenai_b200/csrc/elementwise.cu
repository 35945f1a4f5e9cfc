// Memory-bound kernels of the sampling path: CFG lerp + posterior update (K6), uint8 tail (K8),
// Philox x_T, step counter, MaxPool2d(2) (K3a) and bilinear-upsample + skip-concat (K3b).
// All are HBM-roofline kernels: 128-bit coalesced accesses, no shared memory, one pass.
#include "common.cuh"

namespace sg {

// ------------------------------------------------------------------------------------------------
// K6  (/root/reference/src/diff_modules.py:426-439)
//   eps = lerp(eps_u, eps_c, s)            ATen (cpu/LerpKernel.cpp lerp_vec): fma(s < 0.5 ? s : s - 1, eps_c - eps_u,
//                                          s < 0.5 ? eps_u : eps_c) -- one fused multiply-add
//   x   = c1 * (x - c2 * eps) + c3 * z     every product / sum rounded separately (torch runs them as
//                                          separate elementwise ops), hence the __f*_rn intrinsics
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lerp_aten(float u, float c, float w) {
  const float d = __fsub_rn(c, u);
  const bool small = fabsf(w) < 0.5f;
  return __fmaf_rn(small ? w : __fsub_rn(w, 1.0f), d, small ? u : c);
}
__device__ __forceinline__ float posterior(float x, float e, float c1, float c2, float c3, float z) {
  return __fadd_rn(__fmul_rn(c1, __fsub_rn(x, __fmul_rn(c2, e))), __fmul_rn(c3, z));
}

__global__ void __launch_bounds__(256) cfg_update_kernel(float* __restrict__ x, const float* __restrict__ eps, int n,
                                                         int E4, float cfg, const float* __restrict__ coef, int T,
                                                         const int32_t* __restrict__ step,
                                                         const float* __restrict__ noise, uint64_t seed,
                                                         int64_t sample_base) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = (int64_t)n * E4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i = *step;
  const float c1 = coef[i * 3 + 0], c2 = coef[i * 3 + 1], c3 = coef[i * 3 + 2];
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* e4 = reinterpret_cast<const float4*>(eps);
  float4 xv = x4[idx];
  float4 ec = __ldcs(e4 + idx);
  float4 e = ec;
  if (cfg > 0.0f) {
    const float4 eu = __ldcs(e4 + total + idx);
    e.x = lerp_aten(eu.x, ec.x, cfg);
    e.y = lerp_aten(eu.y, ec.y, cfg);
    e.z = lerp_aten(eu.z, ec.z, cfg);
    e.w = lerp_aten(eu.w, ec.w, cfg);
  }
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i > 1) {
    if (noise != nullptr) {
      z = __ldcs(reinterpret_cast<const float4*>(noise) + (int64_t)(T - i) * total + idx);
    } else {
      const int64_t sample = idx / E4;
      z = philox_normal4(seed, (uint64_t)(sample_base + sample), (uint32_t)i, (uint32_t)(idx - sample * E4));
    }
  }
  xv.x = posterior(xv.x, e.x, c1, c2, c3, z.x);
  xv.y = posterior(xv.y, e.y, c1, c2, c3, z.y);
  xv.z = posterior(xv.z, e.z, c1, c2, c3, z.z);
  xv.w = posterior(xv.w, e.w, c1, c2, c3, z.w);
  reinterpret_cast<float4*>(x)[idx] = xv;
}

__global__ void step_advance_kernel(int32_t* step) {
  pdl_wait();
  pdl_launch_dependents();
  *step -= 1;
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ x, int n, int E4, uint64_t seed,
                                                            int64_t sample_base, int step_tag) {
  const int64_t total = (int64_t)n * E4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t sample = idx / E4;
  reinterpret_cast<float4*>(x)[idx] =
      philox_normal4(seed, (uint64_t)(sample_base + sample), (uint32_t)step_tag, (uint32_t)(idx - sample * E4));
}

// K8 (:440-441): clamp(-1,1) -> +1 -> /2 -> *255 -> truncating cast
__device__ __forceinline__ uint32_t quant_u8(float v) {
  v = fminf(fmaxf(v, -1.0f), 1.0f);
  v = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), 255.0f);
  return (uint32_t)(int)v;  // v in [0,255]; NaN -> 0 like the CPU cast of clamp(NaN)... (NaN never occurs in parity runs)
}
// the un-clamped form the denoise-trajectory dumps use (:672-675): (x+1)/2*255 truncated to an integer and reduced mod 256
// (what the CPU / CUDA float -> uint8 cast of torch does for values that fit an int32)
__device__ __forceinline__ uint32_t wrap_u8(float v) {
  v = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), 255.0f);
  return (uint32_t)__float2int_rz(v) & 255u;
}
template <bool WRAP> __device__ __forceinline__ uint32_t to_u8(float v) { return WRAP ? wrap_u8(v) : quant_u8(v); }
template <bool WRAP>
__global__ void __launch_bounds__(256) to_uint8_kernel(const float* __restrict__ x, int64_t count4,
                                                       uint32_t* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[idx];
  out[idx] = to_u8<WRAP>(v.x) | (to_u8<WRAP>(v.y) << 8) | (to_u8<WRAP>(v.z) << 16) | (to_u8<WRAP>(v.w) << 24);
}
template <bool WRAP>
__global__ void to_uint8_tail_kernel(const float* __restrict__ x, int64_t begin, int64_t count, uint8_t* out) {
  const int64_t idx = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) out[idx] = (uint8_t)to_u8<WRAP>(x[idx]);
}

// K3a  MaxPool2d(2)  (:100).  NHWC, one thread = 4 channels of one output pixel.
__global__ void __launch_bounds__(256) maxpool2_kernel(const float* __restrict__ in, int64_t total4, int Ho, int Wo,
                                                       int C4, float* __restrict__ o32, void* __restrict__ o16,
                                                       int dtype) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = (int)(idx % C4);
  int64_t p = idx / C4;
  const int wo = (int)(p % Wo);
  p /= Wo;
  const int ho = (int)(p % Ho);
  const int64_t r = p / Ho;
  const int W = Wo * 2;
  const float4* src = reinterpret_cast<const float4*>(in) + ((r * (Ho * 2) + ho * 2) * W + wo * 2) * C4 + c4;
  const float4 a = __ldg(src), b = __ldg(src + C4), c = __ldg(src + (int64_t)W * C4),
               d = __ldg(src + (int64_t)W * C4 + C4);
  const float m0 = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
  const float m1 = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
  const float m2 = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
  const float m3 = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
  store4_dual(o32, o16, dtype, idx * 4, m0, m1, m2, m3);
}

// K3b  Upsample(x2, bilinear, align_corners=True) + cat([skip, x], dim=1)   (:120, :132-133)
// One thread = 4*V channels of one output pixel of the concatenated tensor (V = 2 when Cx and Cs are multiples of 8:
// two independent 16-byte loads per source in flight, half the index arithmetic, one 16-byte 16-bit store).
template <int V>
__global__ void __launch_bounds__(256) upsample_cat_kernel(const float* __restrict__ x, const float* __restrict__ skip,
                                                           unsigned per_row, int skip_rows, int h, int w, int Cx4, int Cs4,
                                                           float sh, float sw, float* __restrict__ o32,
                                                           void* __restrict__ o16, int dtype) {
  // grid = (chunks, rows): 32-bit index arithmetic inside a row (the 64-bit div / mod chain of the flat version cost
  // more than the loads); per_row counts units of V float4's
  pdl_wait();
  pdl_launch_dependents();
  const unsigned row = blockIdx.y;
  const unsigned Ct = (unsigned)(Cx4 + Cs4) / V;
  const unsigned H = 2 * h, W = 2 * w;
  const float4* xrow = reinterpret_cast<const float4*>(x) + (size_t)row * h * w * Cx4;
  const float4* srow = reinterpret_cast<const float4*>(skip) + (size_t)(row % (unsigned)skip_rows) * H * W * Cs4;
  const size_t obase = (size_t)row * per_row * (4 * V);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const unsigned p = i / Ct;
    const unsigned c4 = (i - p * Ct) * V;
    const unsigned ho = p / W, wo = p - ho * W;
    float4 v[V];
    if (c4 < (unsigned)Cs4) {
      const float4* src = srow + (size_t)p * Cs4 + c4;
#pragma unroll
      for (int u = 0; u < V; ++u) v[u] = __ldcs(src + u);
    } else {
      const unsigned cx = c4 - Cs4;
      // align_corners=True: src = dst * (in-1)/(out-1)
      const float fy = sh * (float)ho, fx = sw * (float)wo;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float hy = 1.0f - ly, hx = 1.0f - lx;
      const float4* base = xrow + cx;
      float4 a[V], b[V], c[V], d[V];
#pragma unroll
      for (int u = 0; u < V; ++u) {
        a[u] = __ldg(base + (size_t)(y0 * w + x0) * Cx4 + u);
        b[u] = __ldg(base + (size_t)(y0 * w + x1) * Cx4 + u);
        c[u] = __ldg(base + (size_t)(y1 * w + x0) * Cx4 + u);
        d[u] = __ldg(base + (size_t)(y1 * w + x1) * Cx4 + u);
      }
#pragma unroll
      for (int u = 0; u < V; ++u) {
        v[u].x = hy * (hx * a[u].x + lx * b[u].x) + ly * (hx * c[u].x + lx * d[u].x);
        v[u].y = hy * (hx * a[u].y + lx * b[u].y) + ly * (hx * c[u].y + lx * d[u].y);
        v[u].z = hy * (hx * a[u].z + lx * b[u].z) + ly * (hx * c[u].z + lx * d[u].z);
        v[u].w = hy * (hx * a[u].w + lx * b[u].w) + ly * (hx * c[u].w + lx * d[u].w);
      }
    }
    const size_t off = obase + (size_t)i * (4 * V);
    if constexpr (V == 2) {
      if (o32) {
        __stcs(reinterpret_cast<float4*>(o32 + off), v[0]);
        __stcs(reinterpret_cast<float4*>(o32 + off + 4), v[1]);
      }
      if (o16 && dtype == SG_F32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(o16) + off) = tf32_lo4(v[0].x, v[0].y, v[0].z, v[0].w);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(o16) + off + 4) = tf32_lo4(v[1].x, v[1].y, v[1].z, v[1].w);
      } else if (o16) {
        uint4 wv;
        wv.x = pack16(v[0].x, v[0].y, dtype);
        wv.y = pack16(v[0].z, v[0].w, dtype);
        wv.z = pack16(v[1].x, v[1].y, dtype);
        wv.w = pack16(v[1].z, v[1].w, dtype);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o16) + off) = wv;
      }
    } else {
      store4_dual(o32, o16, dtype, off, v[0].x, v[0].y, v[0].z, v[0].w);
    }
  }
}

// Cx == Cs, multiples of 8: one thread = one output pixel x (8 skip channels k, 8 upsampled channels k), so no warp mixes
// "copy" lanes with "gather" lanes (the generic kernel runs both branches in every warp with half the lanes idle).
__global__ void __launch_bounds__(256) upsample_cat_paired_kernel(const float* __restrict__ x, const float* __restrict__ skip,
                                                                  unsigned per_row, int skip_rows, int h, int w, int C8,
                                                                  float sh, float sw, float* __restrict__ o32,
                                                                  void* __restrict__ o16, int dtype) {
  // per_row = pixels * C8 thread-units; C8 = Cs / 8 = Cx / 8
  pdl_wait();
  pdl_launch_dependents();
  const unsigned row = blockIdx.y;
  const unsigned W = 2 * w, HW = 4u * h * w;
  const float4* xrow = reinterpret_cast<const float4*>(x) + (size_t)row * h * w * (C8 * 2);
  const float4* srow = reinterpret_cast<const float4*>(skip) + (size_t)(row % (unsigned)skip_rows) * HW * (C8 * 2);
  const size_t obase = (size_t)row * HW * (C8 * 16);  // elements
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const unsigned p = i / (unsigned)C8, k = i - p * (unsigned)C8;
    const unsigned ho = p / W, wo = p - ho * W;
    const float4* src = srow + ((size_t)p * C8 + k) * 2;
    const float4 s0 = __ldcs(src), s1 = __ldcs(src + 1);
    const float fy = sh * (float)ho, fx = sw * (float)wo;  // align_corners=True: src = dst * (in-1)/(out-1)
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const float4* base = xrow + k * 2;
    float4 v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float4 a = __ldg(base + (size_t)(y0 * w + x0) * (C8 * 2) + u), b = __ldg(base + (size_t)(y0 * w + x1) * (C8 * 2) + u);
      const float4 c = __ldg(base + (size_t)(y1 * w + x0) * (C8 * 2) + u), d = __ldg(base + (size_t)(y1 * w + x1) * (C8 * 2) + u);
      v[u].x = hy * (hx * a.x + lx * b.x) + ly * (hx * c.x + lx * d.x);
      v[u].y = hy * (hx * a.y + lx * b.y) + ly * (hx * c.y + lx * d.y);
      v[u].z = hy * (hx * a.z + lx * b.z) + ly * (hx * c.z + lx * d.z);
      v[u].w = hy * (hx * a.w + lx * b.w) + ly * (hx * c.w + lx * d.w);
    }
    const size_t off_s = obase + ((size_t)p * (2 * C8) + k) * 8, off_x = off_s + (size_t)C8 * 8;
    if (o32) {
      __stcs(reinterpret_cast<float4*>(o32 + off_s), s0);
      __stcs(reinterpret_cast<float4*>(o32 + off_s + 4), s1);
      __stcs(reinterpret_cast<float4*>(o32 + off_x), v[0]);
      __stcs(reinterpret_cast<float4*>(o32 + off_x + 4), v[1]);
    }
    if (o16 && dtype == SG_F32) {
      float* ol = reinterpret_cast<float*>(o16);
      *reinterpret_cast<float4*>(ol + off_s) = tf32_lo4(s0.x, s0.y, s0.z, s0.w);
      *reinterpret_cast<float4*>(ol + off_s + 4) = tf32_lo4(s1.x, s1.y, s1.z, s1.w);
      *reinterpret_cast<float4*>(ol + off_x) = tf32_lo4(v[0].x, v[0].y, v[0].z, v[0].w);
      *reinterpret_cast<float4*>(ol + off_x + 4) = tf32_lo4(v[1].x, v[1].y, v[1].z, v[1].w);
    } else if (o16) {
      uint4 ws, wx;
      ws.x = pack16(s0.x, s0.y, dtype); ws.y = pack16(s0.z, s0.w, dtype);
      ws.z = pack16(s1.x, s1.y, dtype); ws.w = pack16(s1.z, s1.w, dtype);
      wx.x = pack16(v[0].x, v[0].y, dtype); wx.y = pack16(v[0].z, v[0].w, dtype);
      wx.z = pack16(v[1].x, v[1].y, dtype); wx.w = pack16(v[1].z, v[1].w, dtype);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o16) + off_s) = ws;
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o16) + off_x) = wx;
    }
  }
}

// ---- forward-only training helpers (the backward pass is not built) ----
// Diffusion.noise_images (:404-409): x_t = sqrt(ah[t]) * x + sqrt(1 - ah[t]) * eps, un-fused like the reference's
// two multiplies and one add; eps injected or drawn from the Philox stream (and written back, the method returns it)
constexpr int NOISE_TAG = 1 << 20;  // Philox step tag of the training-noise stream (sampling uses 1..T)
__global__ void __launch_bounds__(256) noise_images_kernel(const float* __restrict__ x, const int64_t* __restrict__ t,
                                                           const float* __restrict__ alpha_hat, int T, int n, int E4,
                                                           const float* __restrict__ eps_in, uint64_t seed,
                                                           int64_t sample_base, float* __restrict__ x_t,
                                                           float* __restrict__ eps_out) {
  const int64_t total = (int64_t)n * E4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t sample = idx / E4;
  int64_t ti = t[sample];
  ti = ti < 0 ? ti + T : ti;  // python-style negative index, as alpha_hat[t] would take it
  const float ah = alpha_hat[ti];
  const float sa = __fsqrt_rn(ah), sb = __fsqrt_rn(__fsub_rn(1.0f, ah));
  const float4 xv = reinterpret_cast<const float4*>(x)[idx];
  float4 e;
  if (eps_in) e = reinterpret_cast<const float4*>(eps_in)[idx];
  else e = philox_normal4(seed, (uint64_t)(sample_base + sample), (uint32_t)NOISE_TAG, (uint32_t)(idx - sample * E4));
  float4 o;
  o.x = __fadd_rn(__fmul_rn(sa, xv.x), __fmul_rn(sb, e.x));
  o.y = __fadd_rn(__fmul_rn(sa, xv.y), __fmul_rn(sb, e.y));
  o.z = __fadd_rn(__fmul_rn(sa, xv.z), __fmul_rn(sb, e.z));
  o.w = __fadd_rn(__fmul_rn(sa, xv.w), __fmul_rn(sb, e.w));
  reinterpret_cast<float4*>(x_t)[idx] = o;
  if (eps_out) reinterpret_cast<float4*>(eps_out)[idx] = e;
}

// EMA.update_average (:37-40): ma = ma * beta + (1 - beta) * cur, two rounded products and a rounded sum
__global__ void __launch_bounds__(256) ema_update_kernel(float* __restrict__ ma, const float* __restrict__ cur, int64_t n,
                                                         float beta, float one_minus_beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ma[i] = __fadd_rn(__fmul_rn(ma[i], beta), __fmul_rn(one_minus_beta, cur[i]));
}

// nn.MSELoss (:478), deterministic: MSE_BLOCKS partial sums in double (fixed assignment of elements to blocks), then a
// fixed-order final sum by one block
constexpr int MSE_BLOCKS = 1024;
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int64_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(a[i], b[i]);
    s += (double)d * (double)d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int i = 0; i < 8; ++i) tsum += red[i];
    partial[blockIdx.x] = tsum;
  }
}
__global__ void __launch_bounds__(256) mse_final_kernel(const double* __restrict__ partial, int blocks, int64_t n,
                                                        float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < blocks; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(red[0] / (double)n);
}

// Weight repack (once per load_state_dict): Conv2d / Linear fp32 [Cout][Cin][taps] -> [taps][Cout][Cin] in the operand
// dtype, K (= Cin) contiguous -- the B-operand layout of sg_igemm.  One thread per output element; rounding is
// round-to-nearest-even (what torch's .to(bfloat16 / float16) does), fp16 overflow goes to inf (no saturation).
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                                           void* __restrict__ out, int dtype) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_tap = (int64_t)Cout * Cin;
  if (idx >= per_tap * taps) return;
  const int tap = (int)(idx / per_tap);
  const int64_t oc = idx - (int64_t)tap * per_tap;  // co * Cin + ci
  const float v = w[oc * taps + tap];
  if (dtype == SG_F32) {
    reinterpret_cast<float*>(out)[idx] = v;
  } else if (dtype == SG_BF16) {
    reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(v);
  } else {
    reinterpret_cast<__half*>(out)[idx] = __float2half_rn(v);
  }
}

// fp32 -> (hi, lo) with hi = tf32(x) (round to nearest, ties away) and lo = tf32(x - hi): x - hi is exact in fp32, so
// hi + lo carries 21-22 of x's 24 mantissa bits and both parts are exactly representable TF32 operands.
// hi == NULL: the activation form (common.cuh) -- x itself is the high operand and only lo = tf32_lo(x) is written.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                                         float4* __restrict__ lo, int64_t n4) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldcs(x + i);
  if (hi == nullptr) {
    lo[i] = tf32_lo4(v.x, v.y, v.z, v.w);
    return;
  }
  float4 h, l;
  h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
  l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
  hi[i] = h;
  lo[i] = l;
}

}  // namespace sg

using namespace sg;

template <bool WRAP>
static int to_uint8_launch(const float* x, int64_t count, uint8_t* out, sg_stream_t stream, const char* what) {
  if (count == 0) return SG_OK;
  const int64_t c4 = count / 4;
  if (c4 > 0)
    to_uint8_kernel<WRAP><<<cdiv(c4, 256), 256, 0, as_stream(stream)>>>(x, c4, reinterpret_cast<uint32_t*>(out));
  if (count % 4) to_uint8_tail_kernel<WRAP><<<1, 32, 0, as_stream(stream)>>>(x, c4 * 4, count, out);
  return launch_status(what);
}

extern "C" {

int sg_cfg_update(float* x, const float* eps, int n, int E, float cfg_scale, const float* coef, int T,
                  const int32_t* step, const float* noise, uint64_t seed, int64_t sample_base, sg_stream_t stream) {
  SG_REQUIRE(x && eps && coef && step, "sg_cfg_update: null pointer");
  SG_REQUIRE(n > 0 && E > 0 && E % 4 == 0 && T > 1, "sg_cfg_update: bad shape n=%d E=%d T=%d", n, E, T);
  const int64_t total = (int64_t)n * (E / 4);
  launch_k(cfg_update_kernel, dim3(cdiv(total, 256)), dim3(256), 0, as_stream(stream), x, eps, n, E / 4, cfg_scale, coef, T, step,
           noise, seed, sample_base);
  return launch_status("sg_cfg_update");
}

int sg_step_advance(int32_t* step, sg_stream_t stream) {
  SG_REQUIRE(step, "sg_step_advance: null pointer");
  launch_k(step_advance_kernel, dim3(1), dim3(1), 0, as_stream(stream), step);
  return launch_status("sg_step_advance");
}

int sg_philox_normal(float* x, int n, int E, uint64_t seed, int64_t sample_base, int step_tag, sg_stream_t stream) {
  SG_REQUIRE(x && n > 0 && E > 0 && E % 4 == 0, "sg_philox_normal: bad arguments");
  const int64_t total = (int64_t)n * (E / 4);
  philox_normal_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, n, E / 4, seed, sample_base, step_tag);
  return launch_status("sg_philox_normal");
}

int sg_to_uint8(const float* x, int64_t count, uint8_t* out, sg_stream_t stream) {
  SG_REQUIRE(x && out && count >= 0, "sg_to_uint8: bad arguments");
  return to_uint8_launch<false>(x, count, out, stream, "sg_to_uint8");
}

int sg_noise_images(const float* x, const int64_t* t, const float* alpha_hat, int T, int n, int E, const float* eps_in,
                    uint64_t seed, int64_t sample_base, float* x_t, float* eps_out, sg_stream_t stream) {
  SG_REQUIRE(x && t && alpha_hat && x_t, "sg_noise_images: null pointer");
  SG_REQUIRE(n > 0 && E > 0 && E % 4 == 0 && T > 0, "sg_noise_images: bad shape n=%d E=%d T=%d", n, E, T);
  const int64_t total = (int64_t)n * (E / 4);
  noise_images_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, t, alpha_hat, T, n, E / 4, eps_in, seed,
                                                                       sample_base, x_t, eps_out);
  return launch_status("sg_noise_images");
}

int sg_ema_update(float* ma, const float* cur, int64_t n, float beta, float one_minus_beta, sg_stream_t stream) {
  SG_REQUIRE(ma && cur && n >= 0, "sg_ema_update: bad arguments");
  if (n == 0) return SG_OK;
  ema_update_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(ma, cur, n, beta, one_minus_beta);
  return launch_status("sg_ema_update");
}

int sg_mse_scratch_doubles(void) { return MSE_BLOCKS; }

int sg_mse(const float* a, const float* b, int64_t n, double* scratch, float* out, sg_stream_t stream) {
  SG_REQUIRE(a && b && scratch && out && n > 0, "sg_mse: bad arguments");
  int blocks = (int)(cdiv(n, 256) < MSE_BLOCKS ? cdiv(n, 256) : MSE_BLOCKS);
  mse_partial_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a, b, n, scratch);
  mse_final_kernel<<<1, 256, 0, as_stream(stream)>>>(scratch, blocks, n, out);
  return launch_status("sg_mse");
}

int sg_pack_weights(const float* w, int Cout, int Cin, int taps, void* out, int out_dtype, sg_stream_t stream) {
  SG_REQUIRE(w && out, "sg_pack_weights: null pointer");
  SG_REQUIRE(Cout > 0 && Cin > 0 && taps > 0, "sg_pack_weights: bad shape Cout=%d Cin=%d taps=%d", Cout, Cin, taps);
  SG_REQUIRE(out_dtype == SG_F32 || out_dtype == SG_BF16 || out_dtype == SG_F16, "sg_pack_weights: bad dtype %d", out_dtype);
  const int64_t total = (int64_t)Cout * Cin * taps;
  pack_weights_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(w, Cout, Cin, taps, out, out_dtype);
  return launch_status("sg_pack_weights");
}

int sg_split_tf32(const float* x, float* hi, float* lo, int64_t n, sg_stream_t stream) {
  SG_REQUIRE(x && lo && n > 0 && n % 4 == 0, "sg_split_tf32: null pointer or n=%lld not a positive multiple of 4", (long long)n);
  SG_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
             "sg_split_tf32: buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  SG_REQUIRE(cdiv(n4, 256) < (1ll << 31) - 1, "sg_split_tf32: too many elements");
  launch_k(split_tf32_kernel, dim3(cdiv(n4, 256)), dim3(256), 0, as_stream(stream), reinterpret_cast<const float4*>(x),
           reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo), n4);
  return launch_status("sg_split_tf32");
}

int sg_to_uint8_wrap(const float* x, int64_t count, uint8_t* out, sg_stream_t stream) {
  SG_REQUIRE(x && out && count >= 0, "sg_to_uint8_wrap: bad arguments");
  return to_uint8_launch<true>(x, count, out, stream, "sg_to_uint8_wrap");
}

int sg_maxpool2(const float* in, int rows, int H, int W, int C, float* out_f32, void* out_act, int act_dtype,
                sg_stream_t stream) {
  SG_REQUIRE(in && (out_f32 || out_act), "sg_maxpool2: null pointer");
  SG_REQUIRE(rows > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "sg_maxpool2: bad shape");
  SG_REQUIRE(!out_act || act_dtype == SG_BF16 || act_dtype == SG_F16 || act_dtype == SG_F32, "sg_maxpool2: bad act_dtype");
  const int64_t total4 = (int64_t)rows * (H / 2) * (W / 2) * (C / 4);
  launch_k(maxpool2_kernel, dim3(cdiv(total4, 256)), dim3(256), 0, as_stream(stream), in, total4, H / 2, W / 2, C / 4, out_f32,
           out_act, act_dtype);
  return launch_status("sg_maxpool2");
}

int sg_upsample_cat(const float* x, const float* skip, int rows, int skip_rows, int h, int w, int Cx, int Cs,
                    float* out_f32, void* out_act, int act_dtype, sg_stream_t stream) {
  SG_REQUIRE(skip_rows > 0 && rows % skip_rows == 0, "sg_upsample_cat: rows=%d must be a multiple of skip_rows=%d", rows, skip_rows);
  SG_REQUIRE(x && skip && (out_f32 || out_act), "sg_upsample_cat: null pointer");
  SG_REQUIRE(rows > 0 && h >= 1 && w >= 1 && Cx % 4 == 0 && Cs % 4 == 0, "sg_upsample_cat: bad shape");
  SG_REQUIRE(!out_act || act_dtype == SG_BF16 || act_dtype == SG_F16 || act_dtype == SG_F32, "sg_upsample_cat: bad act_dtype");
  const int64_t row4 = (int64_t)(2 * h) * (2 * w) * ((Cx + Cs) / 4);  // float4's per row
  SG_REQUIRE(row4 < (1ll << 31) && rows <= 65535, "sg_upsample_cat: row too large / too many rows");
  // torch: scale = (in - 1) / (out - 1) in fp32 (area_pixel_compute_scale, align_corners=True)
  const float sh = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
  const float sw = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
  const int v = (Cx % 8 == 0 && Cs % 8 == 0) ? 2 : 1;
  const int64_t per_row = row4 / v;
  int chunks = (int)cdiv(per_row, 256 * 2);  // two elements per thread
  const int want = (int)cdiv(148 * 8, rows);
  if (chunks > want) chunks = want > 1 ? want : 1;
  dim3 grid((unsigned)chunks, (unsigned)rows);
  if (v == 2 && Cx == Cs) {
    const int64_t units = (int64_t)(2 * h) * (2 * w) * (Cs / 8);
    int ch = (int)cdiv(units, 256 * 2);
    if (ch > want) ch = want > 1 ? want : 1;
    launch_k(upsample_cat_paired_kernel, dim3((unsigned)ch, (unsigned)rows), dim3(256), 0, as_stream(stream), x, skip,
             (unsigned)units, skip_rows, h, w, Cs / 8, sh, sw, out_f32, out_act, act_dtype);
  } else if (v == 2)
    launch_k(upsample_cat_kernel<2>, grid, dim3(256), 0, as_stream(stream), x, skip, (unsigned)per_row, skip_rows, h, w, Cx / 4,
             Cs / 4, sh, sw, out_f32, out_act, act_dtype);
  else
    launch_k(upsample_cat_kernel<1>, grid, dim3(256), 0, as_stream(stream), x, skip, (unsigned)per_row, skip_rows, h, w, Cx / 4,
             Cs / 4, sh, sw, out_f32, out_act, act_dtype);
  return launch_status("sg_upsample_cat");
}

}  // extern "C"
