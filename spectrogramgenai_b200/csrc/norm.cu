// K2: GroupNorm(1, C) finalize + apply (+GELU / +residual / +time-embedding), and LayerNorm over C.
// Both are single-pass, HBM-bound kernels (1 read + 1..2 writes of the activation, 128-bit accesses).
#include "common.cuh"

namespace sg {

// ------------------------------------------------------------------------------------------------
// GroupNorm(num_groups=1) over (C, H, W) of each row (/root/reference/src/diff_modules.py:83,:86;
// torch eps 1e-5, biased variance).  The conv kernels leave per-tile (sum, sum-of-squares) partials;
// each block re-reduces its row's P partials in fp64 (deterministic, P <= a few thousand floats from L2),
// then streams its slice of the row.
// grid = (chunks, rows); one thread handles 4 consecutive channels of one pixel per iteration.
// ------------------------------------------------------------------------------------------------
// RAW16: the raw tensor is fp16 (tensor-core modes).  MODE >= 0 fixes the mode at compile time and lets a thread keep
// the folded per-channel scale / shift of its 8 channels in registers (valid when the grid stride is a multiple of C,
// which the launcher checks); MODE = -1 is the general runtime-mode path.
// The fp16 raw tensor a GroupNorm reads is only as good as fp16's range: GroupNorm is scale invariant in the reference,
// fp16 is not.  The statistics come from the fp32 accumulators, so the mean square of the row is known exactly here:
// outside [2^-20, 2^20] (rms outside [2^-10, 2^10]: values within a few 10 sigma of saturation at 65504, or so small that
// the 2^-24 subnormal spacing costs relative precision) the kernel raises *range_flag and the host re-runs the plan with
// fp32 raw tensors (engine.UNetPlan / Diffusion.sample).  An exactly zero row is fine.
__device__ __forceinline__ void flag_fp16_range(double mean_sq, int32_t* range_flag) {
  if (range_flag && (mean_sq > 1048576.0 || (mean_sq > 0.0 && mean_sq < 9.5367431640625e-07) || mean_sq != mean_sq))
    *range_flag = 1;
}

template <bool RAW16, int MODE>
__global__ void __launch_bounds__(256, (RAW16 && (MODE == 0 || MODE == 1)) ? 3 : 2) gn_apply_kernel(const void* __restrict__ raw_v, const float* __restrict__ partials,
                                                       int P, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int64_t per_row4, int C4,
                                                       int raw_rows, int mode_rt, const float* __restrict__ residual,
                                                       const float* __restrict__ emb, int emb_stride,
                                                       float* __restrict__ o32, void* __restrict__ o16, int dtype,
                                                       int32_t* __restrict__ range_flag) {
  __shared__ double red[2][8];
  __shared__ float stat[2];
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.y;
  const int rrow = row % raw_rows;  // raw / partials row (the label-independent prefix is stored once for both halves)
  const int mode = MODE >= 0 ? MODE : mode_rt;
  {
    double s = 0.0, q = 0.0;
    const float2* pp = reinterpret_cast<const float2*>(partials) + (int64_t)rrow * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const float2 v = pp[i];
      s += (double)v.x;
      q += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = s;
      red[1][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ts = 0.0, tq = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
        ts += red[0][i];
        tq += red[1][i];
      }
      const double cnt = (double)per_row4 * 4.0;
      const double mean = ts / cnt;
      double var = tq / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      stat[0] = (float)mean;
      stat[1] = (float)(1.0 / sqrt(var + 1e-5));
      if (RAW16) flag_fp16_range(tq / cnt, range_flag);
    }
    __syncthreads();
  }
  const float mean = stat[0], rstd = stat[1];
  const int64_t base4 = (int64_t)row * per_row4;    // output / residual row
  const int64_t rbase4 = (int64_t)rrow * per_row4;  // raw row
  const float4* res4 = residual ? reinterpret_cast<const float4*>(residual) + base4 : nullptr;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const float4* e4 = emb ? reinterpret_cast<const float4*>(emb + (int64_t)row * emb_stride) : nullptr;
  auto norm4 = [&](float4 v, float4 r, int c4) {  // 4 consecutive channels starting at channel 4*c4
    const float4 g = __ldg(g4 + c4), b = __ldg(b4 + c4);
    const float4 e = e4 ? __ldg(e4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 y;
    if constexpr (RAW16) {
      // the SAME folded arithmetic as the fixed-channel fast path below (y = v * sc + sh), so that a sample's result
      // does not depend on which of the two paths the grid geometry of its batch selects
      // explicit intrinsics: both paths must round identically whatever the compiler would contract
      const float s0 = __fmul_rn(rstd, g.x), s1 = __fmul_rn(rstd, g.y), s2 = __fmul_rn(rstd, g.z), s3 = __fmul_rn(rstd, g.w);
      y.x = __fmaf_rn(v.x, s0, __fadd_rn(__fmaf_rn(-mean, s0, b.x), mode == 0 ? e.x : 0.f));
      y.y = __fmaf_rn(v.y, s1, __fadd_rn(__fmaf_rn(-mean, s1, b.y), mode == 0 ? e.y : 0.f));
      y.z = __fmaf_rn(v.z, s2, __fadd_rn(__fmaf_rn(-mean, s2, b.z), mode == 0 ? e.z : 0.f));
      y.w = __fmaf_rn(v.w, s3, __fadd_rn(__fmaf_rn(-mean, s3, b.w), mode == 0 ? e.w : 0.f));
    } else {
      y.x = (v.x - mean) * rstd * g.x + b.x;
      y.y = (v.y - mean) * rstd * g.y + b.y;
      y.z = (v.z - mean) * rstd * g.z + b.z;
      y.w = (v.w - mean) * rstd * g.w + b.w;
    }
    if (mode == 2) {
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    if (mode >= 1) {
      if constexpr (RAW16) {  // 16-bit engines (both fp16-raw paths use the same form: results do not depend on the path)
        gelu_logistic2(y.x, y.y);
        gelu_logistic2(y.z, y.w);
      } else {
        gelu_erf2(y.x, y.y);
        gelu_erf2(y.z, y.w);
      }
    }
    if (e4 && (!RAW16 || mode >= 1)) {
      y.x += e.x; y.y += e.y; y.z += e.z; y.w += e.w;
    }
    return y;
  };
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if constexpr (RAW16 && MODE >= 0) {
    // fp16 raw, compile-time mode, grid stride a multiple of C/8: this thread always sees the same 8 channels, so
    // y = v * sc + sh with sc = rstd * gamma, sh = beta - mean * sc (+ emb when nothing follows the affine) is one FMA
    constexpr int U = 4;
    const uint4* h8 = reinterpret_cast<const uint4*>(raw_v) + rbase4 / 2;
    const int64_t per_row8 = per_row4 / 2;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4 = (int)((2 * first) % C4);
    float sc[8], sh[8], ev[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 g = __ldg(g4 + c4 + h), b = __ldg(b4 + c4 + h);
      const float4 e = e4 ? __ldg(e4 + c4 + h) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w}, ee[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[h * 4 + j] = __fmul_rn(rstd, gg[j]);
        sh[h * 4 + j] = __fadd_rn(__fmaf_rn(-mean, sc[h * 4 + j], bb[j]), MODE == 0 ? ee[j] : 0.f);
        ev[h * 4 + j] = ee[j];
      }
    }
    // MODE 1 (GELU, 2 bytes in / 2 bytes out) is software-pipelined: the loads of the next U units are issued before the
    // current ones are normalised -- the kernel was latency-bound (60 % DRAM utilisation, every warp waiting on its own
    // loads between compute phases): 4.98 -> 5.42 TB/s.  The write-heavy modes measured 8 % slower with the prefetch.
    constexpr bool PIPE = MODE == 1;
    uint4 h[U], hn[U];
    if constexpr (PIPE) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = first + u * stride;
        if (i < per_row8) h[u] = __ldcs(h8 + i);
      }
    }
    for (int64_t i0 = first; i0 < per_row8; i0 += stride * U) {
      float4 r[U][2];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if constexpr (PIPE) {
          if (i + stride * U < per_row8) hn[u] = __ldcs(h8 + i + stride * U);
        } else {
          if (i < per_row8) h[u] = __ldcs(h8 + i);
        }
        if (MODE == 2 && i < per_row8) {
          r[u][0] = __ldcs(res4 + 2 * i);
          r[u][1] = __ldcs(res4 + 2 * i + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i >= per_row8) break;
        const float2 a0 = unpack16(h[u].x, SG_F16), a1 = unpack16(h[u].y, SG_F16);
        const float2 a2 = unpack16(h[u].z, SG_F16), a3 = unpack16(h[u].w, SG_F16);
        float y[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = __fmaf_rn(y[j], sc[j], sh[j]);
        if (MODE == 2) {
          y[0] += r[u][0].x; y[1] += r[u][0].y; y[2] += r[u][0].z; y[3] += r[u][0].w;
          y[4] += r[u][1].x; y[5] += r[u][1].y; y[6] += r[u][1].z; y[7] += r[u][1].w;
        }
        if (MODE >= 1) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) gelu_logistic2(y[j], y[j + 1]);
          if (e4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] += ev[j];
          }
        }
        const int64_t off = (base4 + 2 * i) * 4;
        if (o32) {
          *reinterpret_cast<float4*>(o32 + off) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(o32 + off + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
        if (o16) {  // 8 values = one 16-byte store
          uint4 w;
          w.x = pack16(y[0], y[1], dtype);
          w.y = pack16(y[2], y[3], dtype);
          w.z = pack16(y[4], y[5], dtype);
          w.w = pack16(y[6], y[7], dtype);
          *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o16) + off) = w;
        }
      }
      if constexpr (PIPE) {
#pragma unroll
        for (int u = 0; u < U; ++u) h[u] = hn[u];
      }
    }
  } else if constexpr (RAW16) {
    // fp16 raw: one thread = 8 consecutive channels = one 16-byte load; U independent loads in flight per thread
    constexpr int U = 4;
    const uint4* h8 = reinterpret_cast<const uint4*>(raw_v) + rbase4 / 2;
    const int64_t per_row8 = per_row4 / 2;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < per_row8; i0 += stride * U) {
      uint4 h[U];
      float4 r[U][2];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < per_row8) {
          h[u] = __ldcs(h8 + i);
          if (mode == 2) {
            r[u][0] = __ldcs(res4 + 2 * i);
            r[u][1] = __ldcs(res4 + 2 * i + 1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i >= per_row8) break;
        const int c4 = (int)((2 * i) % C4);
        const float2 a0 = unpack16(h[u].x, SG_F16), a1 = unpack16(h[u].y, SG_F16);
        const float2 a2 = unpack16(h[u].z, SG_F16), a3 = unpack16(h[u].w, SG_F16);
        const float4 y0 = norm4(make_float4(a0.x, a0.y, a1.x, a1.y), r[u][0], c4);
        const float4 y1 = norm4(make_float4(a2.x, a2.y, a3.x, a3.y), r[u][1], c4 + 1);
        const int64_t off = (base4 + 2 * i) * 4;
        if (o32) {
          *reinterpret_cast<float4*>(o32 + off) = y0;
          *reinterpret_cast<float4*>(o32 + off + 4) = y1;
        }
        if (o16) {  // 8 values = one 16-byte store
          uint4 w;
          w.x = pack16(y0.x, y0.y, dtype);
          w.y = pack16(y0.z, y0.w, dtype);
          w.z = pack16(y1.x, y1.y, dtype);
          w.w = pack16(y1.z, y1.w, dtype);
          *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(o16) + off) = w;
        }
      }
    }
  } else {
    // 4 independent 16-byte loads in flight per thread (the single-load loop was latency-bound at 57 % of HBM peak)
    constexpr int U = 4;
    const float4* r4 = reinterpret_cast<const float4*>(raw_v) + rbase4;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < per_row4; i0 += stride * U) {
      float4 v[U], r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < per_row4) {
          v[u] = __ldcs(r4 + i);
          if (mode == 2) r[u] = __ldcs(res4 + i);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i >= per_row4) break;
        const float4 y = norm4(v[u], r[u], (int)(i % C4));
        store4_dual(o32, o16, dtype, (base4 + i) * 4, y.x, y.y, y.z, y.w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GELU(GroupNorm(raw) + cat([skip, upsample2x(x)])) -> 16 bit: the second GroupNorm of Up's first DoubleConv
// (:132-134 feeding the residual DoubleConv :88-91), with the residual RECOMPUTED from its sources instead of read
// from an fp32 copy of the concatenated tensor.  Materialising that copy cost a 2.1 GB write in upsample_cat plus a
// 2.1 GB read here at up3 (n = 512); the sources are 0.8 GB and mostly L2-resident.  Same bilinear expression as
// upsample_cat_kernel (align_corners=True).  One thread = 8 channels of one pixel; fp16 raw only (tensor-core modes).
// ------------------------------------------------------------------------------------------------
constexpr int VCAT_MAXC = 512;  // channels of the concatenated tensor (up1: 256 + 256)

__global__ void __launch_bounds__(256, 4) gn_apply_vcat_kernel(const uint4* __restrict__ raw8, const float* __restrict__ partials,
                                                            int P, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int HW, int W2, int C8,
                                                            int Cs8, const float* __restrict__ x,
                                                            const float* __restrict__ skip, int skip_rows, int h, int w,
                                                            float sh, float sw, void* __restrict__ o16, int dtype,
                                                            int32_t* __restrict__ range_flag) {
  __shared__ double red[2][8];
  __shared__ float stat[2];
  __shared__ __align__(16) float s_sc[VCAT_MAXC], s_sf[VCAT_MAXC];
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.y;
  const int64_t per_row8 = (int64_t)HW * C8;
  {
    double s = 0.0, q = 0.0;
    const float2* pp = reinterpret_cast<const float2*>(partials) + (int64_t)row * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const float2 v = pp[i];
      s += (double)v.x;
      q += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = s;
      red[1][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ts = 0.0, tq = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
        ts += red[0][i];
        tq += red[1][i];
      }
      const double cnt = (double)per_row8 * 8.0;
      const double mean = ts / cnt;
      double var = tq / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      stat[0] = (float)mean;
      stat[1] = (float)(1.0 / sqrt(var + 1e-5));
      flag_fp16_range(tq / cnt, range_flag);
    }
    __syncthreads();
  }
  const float mean = stat[0], rstd = stat[1];
  const int Cx8 = C8 - Cs8;
  // folded scale / shift of every channel (same rounding as gn_apply_kernel) in shared memory: a thread re-reads its
  // 16 channels each iteration (8 LDS.128) instead of holding 32 registers, which doubles the resident warps
  for (int c = threadIdx.x; c < C8 * 8; c += blockDim.x) {
    const float scv = __fmul_rn(rstd, __ldg(gamma + c));
    s_sc[c] = scv;
    s_sf[c] = __fmaf_rn(-mean, scv, __ldg(beta + c));
  }
  __syncthreads();
  const uint4* rrow = raw8 + (int64_t)row * per_row8;
  uint4* orow = reinterpret_cast<uint4*>(o16) + (int64_t)row * per_row8;
  const float4* skip4 = reinterpret_cast<const float4*>(skip) + (int64_t)(row % skip_rows) * HW * (Cs8 * 2);
  const float4* x4 = reinterpret_cast<const float4*>(x) + (int64_t)row * h * w * (Cx8 * 2);
  // One thread = one pixel x (8 skip channels k, then the 8 upsampled channels Cs8 + k): every lane runs both the skip
  // load and the bilinear gather (a warp that mixed "skip lanes" and "x lanes" executed both branches with half its
  // lanes idle).  32-bit index arithmetic; the launcher makes the grid stride a multiple of Cs8 (= Cx8, checked
  // there), so a thread keeps its channels.
  const unsigned nu = (unsigned)HW * (unsigned)Cs8, stride = gridDim.x * blockDim.x;
  const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned k = first % (unsigned)Cs8;
  // y = GELU(GN(raw) + r) for 8 channels starting at channel c0 -> one 16-byte store
  auto finish = [&](const uint4 hraw, const float (&r)[8], unsigned c0, size_t e) {
    const float2 a0 = unpack16(hraw.x, SG_F16), a1 = unpack16(hraw.y, SG_F16);
    const float2 a2 = unpack16(hraw.z, SG_F16), a3 = unpack16(hraw.w, SG_F16);
    float y[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
    const float4 c0v = *reinterpret_cast<const float4*>(s_sc + c0), c1v = *reinterpret_cast<const float4*>(s_sc + c0 + 4);
    const float4 f0v = *reinterpret_cast<const float4*>(s_sf + c0), f1v = *reinterpret_cast<const float4*>(s_sf + c0 + 4);
    const float sc[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
    const float sf[8] = {f0v.x, f0v.y, f0v.z, f0v.w, f1v.x, f1v.y, f1v.z, f1v.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = __fmaf_rn(y[j], sc[j], sf[j]) + r[j];
#pragma unroll
    for (int j = 0; j < 8; j += 2) gelu_logistic2(y[j], y[j + 1]);
    uint4 wv;
    wv.x = pack16(y[0], y[1], dtype);
    wv.y = pack16(y[2], y[3], dtype);
    wv.z = pack16(y[4], y[5], dtype);
    wv.w = pack16(y[6], y[7], dtype);
    orow[e] = wv;
  };
  unsigned p = first / (unsigned)Cs8;
  const unsigned dp = stride / (unsigned)Cs8;
  // software pipeline (as gn_apply_kernel's GELU mode): the raw loads of the next pixel are issued before this one is
  // normalised
  uint4 n_s = make_uint4(0, 0, 0, 0), n_x = n_s;
  if (first < nu) {
    n_s = __ldcs(rrow + (size_t)p * C8 + k);
    n_x = __ldcs(rrow + (size_t)p * C8 + k + Cs8);
  }
  for (unsigned i = first; i < nu; i += stride, p += dp) {
    const size_t e_s = (size_t)p * C8 + k, e_x = e_s + Cs8;  // units of 8 channels inside the row
    const uint4 h_s = n_s, h_x = n_x;
    if (i + stride < nu) {
      n_s = __ldcs(rrow + (size_t)(p + dp) * C8 + k);
      n_x = __ldcs(rrow + (size_t)(p + dp) * C8 + k + Cs8);
    }
    const unsigned ho = p / (unsigned)W2, wo = p - ho * (unsigned)W2;
    const float fy = sh * (float)ho, fx = sw * (float)wo;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const float4* xa = x4 + (size_t)(y0 * w + x0) * (Cx8 * 2) + k * 2;
    const float4* xb = x4 + (size_t)(y0 * w + x1) * (Cx8 * 2) + k * 2;
    const float4* xc = x4 + (size_t)(y1 * w + x0) * (Cx8 * 2) + k * 2;
    const float4* xd = x4 + (size_t)(y1 * w + x1) * (Cx8 * 2) + k * 2;
    {
      const float4 a = __ldg(skip4 + ((size_t)p * Cs8 + k) * 2), b = __ldg(skip4 + ((size_t)p * Cs8 + k) * 2 + 1);
      const float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      finish(h_s, r, k * 8, e_s);
    }
    {
      float r[8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 a = __ldg(xa + u), b = __ldg(xb + u), c = __ldg(xc + u), d = __ldg(xd + u);
        r[u * 4 + 0] = hy * (hx * a.x + lx * b.x) + ly * (hx * c.x + lx * d.x);
        r[u * 4 + 1] = hy * (hx * a.y + lx * b.y) + ly * (hx * c.y + lx * d.y);
        r[u * 4 + 2] = hy * (hx * a.z + lx * b.z) + ly * (hx * c.z + lx * d.z);
        r[u * 4 + 3] = hy * (hx * a.w + lx * b.w) + ly * (hx * c.w + lx * d.w);
      }
      finish(h_x, r, (Cs8 + k) * 8, e_x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over C (:57, :59; eps 1e-5).  C/4 lanes (at most 32) hold one token as float4's, so a warp instruction
// moves 512 contiguous bytes (the first version, one warp per token with 2 values per lane, reached 3.0 TB/s); four
// passes are unrolled so that every lane has four independent 16-byte loads in flight.  Two-pass mean / variance
// like torch, reductions by xor-shuffles inside the token's lane group.
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int64_t M,
                                                        float* __restrict__ o32, void* __restrict__ o16, int dtype) {
  constexpr int LPT = C / 4 < 32 ? C / 4 : 32;  // lanes per token: 16 (C = 64), 32 (C = 128, 256)
  constexpr int NV = C / (4 * LPT);             // float4's per lane: 1, 1, 2
  constexpr int TPW = 32 / LPT;                 // tokens per warp pass: 2, 1, 1
  constexpr int U = 4;                          // passes in flight
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPT, tsel = lane / LPT;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t tok0 = wid * (U * TPW) + tsel;
  float4 gm[NV], bt[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    gm[j] = __ldg(reinterpret_cast<const float4*>(gamma) + j * LPT + sub);
    bt[j] = __ldg(reinterpret_cast<const float4*>(beta) + j * LPT + sub);
  }
  float4 v[U][NV];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t tok = tok0 + u * TPW;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      v[u][j] = tok < M ? __ldcs(reinterpret_cast<const float4*>(in + tok * C) + j * LPT + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t tok = tok0 + u * TPW;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += (v[u][j].x + v[u][j].y) + (v[u][j].z + v[u][j].w);
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float d0 = v[u][j].x - mean, d1 = v[u][j].y - mean, d2 = v[u][j].z - mean, d3 = v[u][j].w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / C) + 1e-5f);
    if (tok < M) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float y0 = (v[u][j].x - mean) * rstd * gm[j].x + bt[j].x;
        const float y1 = (v[u][j].y - mean) * rstd * gm[j].y + bt[j].y;
        const float y2 = (v[u][j].z - mean) * rstd * gm[j].z + bt[j].z;
        const float y3 = (v[u][j].w - mean) * rstd * gm[j].w + bt[j].w;
        store4_dual(o32, o16, dtype, tok * C + (j * LPT + sub) * 4, y0, y1, y2, y3);
      }
    }
  }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_gn_apply(const void* raw, int raw_dtype, const float* partials, int P, const float* gamma, const float* beta,
                int rows, int raw_rows, int HW, int C, int mode, const float* residual, const float* emb, int emb_stride, float* out_f32,
                void* out_act, int act_dtype, int32_t* range_flag, sg_stream_t stream) {
  SG_REQUIRE(raw_dtype == SG_F32 || raw_dtype == SG_F16, "sg_gn_apply: raw_dtype must be SG_F32 or SG_F16");
  SG_REQUIRE(raw && partials && gamma && beta && (out_f32 || out_act), "sg_gn_apply: null pointer");
  SG_REQUIRE(rows > 0 && HW > 0 && C % 4 == 0 && P > 0, "sg_gn_apply: bad shape rows=%d HW=%d C=%d P=%d", rows, HW, C, P);
  SG_REQUIRE(raw_dtype == SG_F32 || C % 8 == 0, "sg_gn_apply: fp16 raw needs C %% 8 == 0 (C=%d)", C);
  SG_REQUIRE(raw_rows > 0 && rows % raw_rows == 0, "sg_gn_apply: rows=%d must be a multiple of raw_rows=%d", rows, raw_rows);
  SG_REQUIRE(mode >= 0 && mode <= 2, "sg_gn_apply: mode %d", mode);
  SG_REQUIRE(mode != 2 || residual, "sg_gn_apply: mode 2 needs a residual");
  SG_REQUIRE(!emb || emb_stride % 4 == 0, "sg_gn_apply: emb stride must be a multiple of 4");
  SG_REQUIRE(!out_act || act_dtype == SG_BF16 || act_dtype == SG_F16 || (act_dtype == SG_F32 && raw_dtype == SG_F32),
             "sg_gn_apply: out_act is 16 bit, or (fp32 raw only) the fp32 TF32-low-part tensor");
  const int64_t per_row4 = (int64_t)HW * (C / 4);
  // enough blocks to fill 148 SMs x 8 resident blocks, at most one float4 per thread per pass
  int chunks = cdiv(per_row4, 256 * 4);
  const int want = cdiv(148 * 8, rows);
  if (chunks > want) chunks = want > 1 ? want : 1;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, rows);
  cudaStream_t st = as_stream(stream);
#define SG_GN_LAUNCH(R16, MD)                                                                                          \
  launch_k(gn_apply_kernel<R16, MD>, grid, dim3(256), 0, st, raw, partials, P, gamma, beta, per_row4, C / 4, raw_rows, mode, \
           residual, emb, emb_stride, out_f32, out_act, act_dtype, range_flag)
  if (raw_dtype == SG_F16) {
    // fixed-channel fast path: every thread of the grid-stride loop must land on the same 8 channels each pass
    const bool fixed = ((int64_t)chunks * 256 * 8) % C == 0;
    if (fixed && mode == 0) SG_GN_LAUNCH(true, 0);
    else if (fixed && mode == 1) SG_GN_LAUNCH(true, 1);
    else if (fixed && mode == 2) SG_GN_LAUNCH(true, 2);
    else SG_GN_LAUNCH(true, -1);
  } else {
    SG_GN_LAUNCH(false, -1);
  }
#undef SG_GN_LAUNCH
  return launch_status("sg_gn_apply");
}

int sg_gn_apply_vcat(const void* raw, const float* partials, int P, const float* gamma, const float* beta, int rows,
                     const float* x, const float* skip, int skip_rows, int h, int w, int Cx, int Cs, void* out_act,
                     int act_dtype, int32_t* range_flag, sg_stream_t stream) {
  SG_REQUIRE(raw && partials && gamma && beta && x && skip && out_act, "sg_gn_apply_vcat: null pointer");
  SG_REQUIRE(rows > 0 && h >= 1 && w >= 1 && P > 0 && Cx % 8 == 0 && Cs % 8 == 0 && Cx > 0 && Cs > 0,
             "sg_gn_apply_vcat: bad shape rows=%d h=%d w=%d Cx=%d Cs=%d", rows, h, w, Cx, Cs);
  SG_REQUIRE(skip_rows > 0 && rows % skip_rows == 0, "sg_gn_apply_vcat: rows=%d must be a multiple of skip_rows=%d", rows, skip_rows);
  SG_REQUIRE(act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_gn_apply_vcat: out_act needs a 16-bit dtype");
  const int H2 = 2 * h, W2 = 2 * w, C = Cx + Cs;
  const int64_t per_row8 = (int64_t)H2 * W2 * (C / 8);
  SG_REQUIRE(per_row8 < (1ll << 31), "sg_gn_apply_vcat: row too large");
  int chunks = cdiv(per_row8 / 2, 256);
  const int want = cdiv(148 * 8, rows);
  if (chunks > want) chunks = want > 1 ? want : 1;
  if (chunks < 1) chunks = 1;
  SG_REQUIRE(C <= VCAT_MAXC, "sg_gn_apply_vcat: %d channels > %d", C, VCAT_MAXC);
  SG_REQUIRE(Cx == Cs, "sg_gn_apply_vcat: Cx=%d != Cs=%d (a thread pairs skip channel k with upsampled channel k)", Cx, Cs);
  // the kernel keeps a thread on the same channels: the grid stride (chunks * 256 threads) must be a multiple of Cs/8
  while (((int64_t)chunks * 256) % (Cs / 8) != 0) ++chunks;
  // torch: scale = (in - 1) / (out - 1) in fp32 (area_pixel_compute_scale, align_corners=True), as sg_upsample_cat
  const float sh = (H2 > 1) ? (float)(h - 1) / (float)(H2 - 1) : 0.f;
  const float sw = (W2 > 1) ? (float)(w - 1) / (float)(W2 - 1) : 0.f;
  launch_k(gn_apply_vcat_kernel, dim3(chunks, rows), dim3(256), 0, as_stream(stream), reinterpret_cast<const uint4*>(raw),
           partials, P, gamma, beta, H2 * W2, W2, C / 8, Cs / 8, x, skip, skip_rows, h, w, sh, sw, out_act, act_dtype, range_flag);
  return launch_status("sg_gn_apply_vcat");
}

int sg_layernorm(const float* in, const float* gamma, const float* beta, int64_t M, int C, void* out_act, int act_dtype,
                 sg_stream_t stream) {
  SG_REQUIRE(in && gamma && beta && out_act, "sg_layernorm: null pointer");
  SG_REQUIRE(M > 0, "sg_layernorm: M=%lld", (long long)M);
  float* o32 = act_dtype == SG_F32 ? reinterpret_cast<float*>(out_act) : nullptr;
  void* o16 = act_dtype == SG_F32 ? nullptr : out_act;
  const int tok_per_block = 8 * 4 * (C == 64 ? 2 : 1);  // 8 warps x 4 passes x tokens per pass
  const int blocks = cdiv(M, tok_per_block);
  cudaStream_t s = as_stream(stream);
  switch (C) {
    case 64: launch_k(layernorm_kernel<64>, dim3(blocks), dim3(256), 0, s, in, gamma, beta, M, o32, o16, act_dtype); break;
    case 128: launch_k(layernorm_kernel<128>, dim3(blocks), dim3(256), 0, s, in, gamma, beta, M, o32, o16, act_dtype); break;
    case 256: launch_k(layernorm_kernel<256>, dim3(blocks), dim3(256), 0, s, in, gamma, beta, M, o32, o16, act_dtype); break;
    default: SG_REQUIRE(false, "sg_layernorm: C=%d not in {64,128,256}", C);
  }
  return launch_status("sg_layernorm");
}

}  // extern "C"
