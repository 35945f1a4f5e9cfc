// fp32 CUDA-core ("SIMT engine") convolution / linear kernels: the fp32-accurate path (rel-L2 <= 1e-4
// against the fp32 reference) and the two memory-bound edge convolutions that never go to tensor cores
// (inc.double_conv.0 with K = 9*c_in <= 36, and outc with N = c_out <= 8).
#include "common.cuh"

namespace sg {

int igemm_tc(const sg_igemm_args* a, cudaStream_t stream);  // igemm_tc.cu

// ------------------------------------------------------------------------------------------------
// inc.double_conv.0  (/root/reference/src/diff_modules.py:82 via :144): direct 3x3, NCHW fp32 in,
// NHWC fp32 out (64 channels) + GroupNorm partials.
// ------------------------------------------------------------------------------------------------
// The 64 x c_in x 3 x 3 weights (<= 9216 B) are passed BY VALUE as a __grid_constant__ kernel parameter, i.e. they live in
// the constant bank of this launch: with one pixel x all 64 output channels per thread every weight index is a
// compile-time constant, so each FMA takes its weight straight from the constant bank (no shared-memory staging, no load
// instructions; the first version was bound by LDS.128 wavefronts at 1.4 ms), four at a time by LDCU.128 in the [k][co]
// order sg_conv_in repacks them to.  A launch carries its own copy: two streams / two models never share state.
template <int CIN>
struct ConvInW {
  float w[CIN * 9 * 64];  // [k = ci*9 + dy*3 + dx][co]
};

template <int CIN, bool RAW16>
__global__ void __launch_bounds__(128, 4) conv_in_kernel(const __grid_constant__ ConvInW<CIN> cw,
                                                         const float* __restrict__ x, int n_src, int S,
                                                         void* __restrict__ raw_v, float* __restrict__ partials) {
  constexpr int K = CIN * 9;
  __shared__ float red[2][4];
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int HW = S * S;
  const int blocks_per_row = HW / 128;
  const int row = blockIdx.x / blocks_per_row;
  const int p = (blockIdx.x % blocks_per_row) * 128 + tid;
  const int h = p / S, wq = p % S;
  const float* xs = x + (int64_t)(row % n_src) * CIN * HW;
  float in[K];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int hh = h + dy - 1, ww = wq + dx - 1;
        in[ci * 9 + dy * 3 + dx] = (hh >= 0 && hh < S && ww >= 0 && ww < S) ? __ldg(xs + ci * HW + hh * S + ww) : 0.f;
      }
  float* dst = reinterpret_cast<float*>(raw_v) + ((int64_t)row * HW + p) * 64;
  uint16_t* dst16 = reinterpret_cast<uint16_t*>(raw_v) + ((int64_t)row * HW + p) * 64;
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {  // 16 channels at a time keeps the accumulators + patch under 128 registers
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(in[k], cw.w[k * 64 + c0 + j], acc[j]);
    if constexpr (RAW16) {  // statistics from the fp32 values, storage in (saturating) fp16: 2 x 16 bytes
#pragma unroll
      for (int j8 = 0; j8 < 2; ++j8) {
        uint4 w;
        w.x = pack16(acc[j8 * 8 + 0], acc[j8 * 8 + 1], SG_F16);
        w.y = pack16(acc[j8 * 8 + 2], acc[j8 * 8 + 3], SG_F16);
        w.z = pack16(acc[j8 * 8 + 4], acc[j8 * 8 + 5], SG_F16);
        w.w = pack16(acc[j8 * 8 + 6], acc[j8 * 8 + 7], SG_F16);
        __stcs(reinterpret_cast<uint4*>(dst16 + c0 + j8 * 8), w);
      }
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        __stcs(reinterpret_cast<float4*>(dst + c0 + j4 * 4),
               make_float4(acc[j4 * 4], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3]));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      s += acc[j];
      q += acc[j] * acc[j];
    }
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = s;
    red[1][tid >> 5] = q;
  }
  __syncthreads();
  if (tid == 0) {
    float* pp = partials + (int64_t)blockIdx.x * 2;
    pp[0] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    pp[1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  }
}

// ------------------------------------------------------------------------------------------------
// K1 (SIMT engine): fp32 implicit GEMM.  BM=128 pixels x BN=64 output channels x BK=16 input channels
// of one tap per k-block; 256 threads, 8x4 outputs per thread, register-prefetch double buffering.
// ------------------------------------------------------------------------------------------------
constexpr int SBM = 128, SBN = 64, SBK = 16, SLDA = 130, SLDB = 68;

__global__ void __launch_bounds__(256) igemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ Wt,
                                                         const float* __restrict__ bias,
                                                         const float* __restrict__ residual, float* __restrict__ out,
                                                         float* __restrict__ partials, int64_t M, int H, int W, int Cin,
                                                         int Cout, int taps, int act, int P) {
  __shared__ float As[2][SBK][SLDA];
  __shared__ __align__(16) float Bs[2][SBK][SLDB];
  __shared__ float rowstat[SBM][2];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;
  const int HW = H * W;
  // ---- loader mapping: thread -> (pixel lm / lm+64, 4-channel quad kq) ----
  const int lm = tid >> 2, kq = tid & 3;
  int ph[2], pw[2];
  int64_t pbase[2];
  bool pvalid[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t m = m0 + lm + 64 * j;
    pvalid[j] = m < M;
    const int64_t mm = pvalid[j] ? m : 0;
    const int pix = (int)(mm % HW);
    ph[j] = pix / W;
    pw[j] = pix % W;
    pbase[j] = mm * Cin;  // element offset of the centre pixel, channel 0
  }
  const int cblocks = Cin / SBK;
  const int nk = taps * cblocks;
  float4 ra[2], rb;
  auto load = [&](int kb) {
    const int tap = kb / cblocks, c0 = (kb % cblocks) * SBK;
    int dy = 0, dx = 0;
    if (taps == 9) {
      dy = tap / 3 - 1;
      dx = tap % 3 - 1;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int hh = ph[j] + dy, ww = pw[j] + dx;
      const bool ok = pvalid[j] && hh >= 0 && hh < H && ww >= 0 && ww < W;
      ra[j] = ok ? __ldg(reinterpret_cast<const float4*>(A + pbase[j] + ((int64_t)dy * W + dx) * Cin + c0 + kq * 4))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rb = __ldg(reinterpret_cast<const float4*>(Wt + ((int64_t)tap * Cout + n0 + lm) * Cin + c0 + kq * 4));
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      As[buf][kq * 4 + 0][lm + 64 * j] = ra[j].x;
      As[buf][kq * 4 + 1][lm + 64 * j] = ra[j].y;
      As[buf][kq * 4 + 2][lm + 64 * j] = ra[j].z;
      As[buf][kq * 4 + 3][lm + 64 * j] = ra[j].w;
    }
    Bs[buf][kq * 4 + 0][lm] = rb.x;
    Bs[buf][kq * 4 + 1][lm] = rb.y;
    Bs[buf][kq * 4 + 2][lm] = rb.z;
    Bs[buf][kq * 4 + 3][lm] = rb.w;
  };
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load(0);
  stash(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load(kb + 1);
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[buf][k][ty + 16 * i];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    if (kb + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }
  // ---- epilogue: + bias, GELU, + residual, store, GroupNorm partials ----
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + tx * 4));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty + 16 * i;
    const int64_t m = m0 + r;
    float v0 = acc[i][0] + bv.x, v1 = acc[i][1] + bv.y, v2 = acc[i][2] + bv.z, v3 = acc[i][3] + bv.w;
    if (act == SG_ACT_GELU) {
      v0 = gelu_erf(v0); v1 = gelu_erf(v1); v2 = gelu_erf(v2); v3 = gelu_erf(v3);
    }
    float s = 0.f, q = 0.f;
    if (m < M) {
      const int64_t off = m * Cout + n0 + tx * 4;
      if (residual) {
        const float4 rr = __ldg(reinterpret_cast<const float4*>(residual + off));
        v0 += rr.x; v1 += rr.y; v2 += rr.z; v3 += rr.w;
      }
      if (act == SG_ACT_RELU_POST) {
        v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
      }
      *reinterpret_cast<float4*>(out + off) = make_float4(v0, v1, v2, v3);
      s = v0 + v1 + v2 + v3;
      q = v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
    }
    if (partials) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (tx == 0) {
        rowstat[r][0] = s;
        rowstat[r][1] = q;
      }
    }
  }
  if (partials) {
    __syncthreads();
    write_tile_partials<SBM>(rowstat, tid, m0, M, HW, partials, P, blockIdx.y, gridDim.y);
  }
}

// ------------------------------------------------------------------------------------------------
// outc (:166, :195): 1x1 conv 64 -> c_out (+bias); NHWC fp32 in, NCHW fp32 out.  AI ~ 4 FLOP/B: HBM-bound.
// One thread per pixel: 256 B contiguous read, coalesced per-channel writes.
// ------------------------------------------------------------------------------------------------
constexpr int CO_LD = 68;  // padded row (floats): 16-byte aligned, and 8 consecutive rows start 4 banks apart
__global__ void __launch_bounds__(128) conv_out_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ b, int64_t total, int HW, int c_out,
                                                       float* __restrict__ eps) {
  // tile = 128 pixels x 64 channels (32 KB), staged through shared memory so that the global read is one fully
  // coalesced stream (the first version had every thread walk its own 256-byte pixel row: half of each 32-byte
  // sector was wasted, 1.7 TB/s); then one thread per pixel, coalesced per-channel writes.
  __shared__ __align__(16) float ws[8][64];
  __shared__ float bs[8];
  __shared__ __align__(16) float tile[128 * CO_LD];
  for (int i = threadIdx.x; i < c_out * 64; i += blockDim.x) ws[i / 64][i % 64] = w[i];
  if (threadIdx.x < c_out) bs[threadIdx.x] = b[threadIdx.x];
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const float4* src = reinterpret_cast<const float4*>(in + m0 * 64);
  const int64_t left = total - m0;
  const int npx = left < 128 ? (int)left : 128;
#pragma unroll 4
  for (int i = threadIdx.x; i < npx * 16; i += 128) {
    const float4 v = __ldcs(src + i);
    *reinterpret_cast<float4*>(&tile[(i >> 4) * CO_LD + (i & 15) * 4]) = v;
  }
  __syncthreads();
  if ((int)threadIdx.x >= npx) return;
  const int64_t m = m0 + threadIdx.x;
  float acc[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) acc[co] = 0.f;
  const float* trow = &tile[threadIdx.x * CO_LD];
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) {
    const float4 v = *reinterpret_cast<const float4*>(trow + k4 * 4);
#pragma unroll
    for (int co = 0; co < 8; ++co) {
      if (co < c_out) {
        const float4 ww = *reinterpret_cast<const float4*>(&ws[co][k4 * 4]);
        acc[co] = fmaf(v.x, ww.x, acc[co]);
        acc[co] = fmaf(v.y, ww.y, acc[co]);
        acc[co] = fmaf(v.z, ww.z, acc[co]);
        acc[co] = fmaf(v.w, ww.w, acc[co]);
      }
    }
  }
  const int64_t row = m / HW, p = m % HW;
#pragma unroll
  for (int co = 0; co < 8; ++co)
    if (co < c_out) eps[(row * c_out + co) * HW + p] = acc[co] + bs[co];
}

static inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace sg

using namespace sg;

extern "C" {

int sg_conv_in_partials(int S) { return S * S / 128; }

int sg_conv_in(const float* x, int n_src, int c_in, int S, const float* w, int rows, void* raw, int raw_dtype,
               float* partials, sg_stream_t stream) {
  SG_REQUIRE(raw_dtype == SG_F32 || raw_dtype == SG_F16, "sg_conv_in: raw_dtype must be SG_F32 or SG_F16");
  SG_REQUIRE(x && w && raw && partials, "sg_conv_in: null pointer");
  SG_REQUIRE(c_in >= 1 && c_in <= 4, "sg_conv_in: c_in=%d not in 1..4", c_in);
  SG_REQUIRE(pow2(S) && S >= 16 && rows > 0 && n_src > 0, "sg_conv_in: S=%d must be a power of two >= 16", S);
  const int64_t ntiles = (int64_t)rows * (S * S / 128);
  SG_REQUIRE(ntiles < (1ll << 31), "sg_conv_in: too many tiles");
  cudaStream_t s = as_stream(stream);
  const int blocks = (int)ntiles;
#define SG_CONV_IN(CI)                                                                                                \
  do {                                                                                                                \
    ConvInW<CI> cw;                                                                                                   \
    for (int co = 0; co < 64; ++co)                                                                                   \
      for (int k = 0; k < CI * 9; ++k) cw.w[k * 64 + co] = w[co * CI * 9 + k]; /* w: HOST pointer, [co][ci][3][3] */  \
    if (raw_dtype == SG_F16)                                                                                          \
      launch_k(conv_in_kernel<CI, true>, dim3(blocks), dim3(128), 0, s, cw, x, n_src, S, raw, partials);              \
    else                                                                                                              \
      launch_k(conv_in_kernel<CI, false>, dim3(blocks), dim3(128), 0, s, cw, x, n_src, S, raw, partials);             \
  } while (0)
  switch (c_in) {
    case 1: SG_CONV_IN(1); break;
    case 2: SG_CONV_IN(2); break;
    case 3: SG_CONV_IN(3); break;
    default: SG_CONV_IN(4); break;
  }
#undef SG_CONV_IN
  return launch_status("sg_conv_in");
}

int sg_igemm_partials(int engine, int H, int W, int Cout) {
  const int HW = H * W;
  const int mt = HW >= 128 ? HW / 128 : 1;
  int bn = 64;
  if (engine == SG_ENGINE_TC) bn = (Cout % 128 == 0) ? 128 : 64;
  return mt * (Cout / bn);
}

int sg_igemm(const sg_igemm_args* a, sg_stream_t stream) {
  SG_REQUIRE(a, "sg_igemm: null args");
  SG_REQUIRE(a->a && a->w && (a->out_f32 || a->out_act), "sg_igemm: null pointer");
  SG_REQUIRE(a->taps == 1 || a->taps == 9, "sg_igemm: taps=%d", a->taps);
  SG_REQUIRE(a->rows > 0 && pow2(a->H) && pow2(a->W), "sg_igemm: rows=%d H=%d W=%d (H, W must be powers of two)",
             a->rows, a->H, a->W);
  SG_REQUIRE(a->Cout % 64 == 0, "sg_igemm: Cout=%d %% 64 != 0", a->Cout);
  if (a->engine == SG_ENGINE_TC) return igemm_tc(a, as_stream(stream));
  SG_REQUIRE(a->engine == SG_ENGINE_SIMT, "sg_igemm: engine %d", a->engine);
  SG_REQUIRE(a->act_dtype == SG_F32 && a->out_dtype == 0, "sg_igemm: the SIMT engine computes and stores in fp32");
  SG_REQUIRE(a->Cin % 16 == 0, "sg_igemm: Cin=%d %% 16 != 0", a->Cin);
  float* out = a->out_f32 ? a->out_f32 : reinterpret_cast<float*>(a->out_act);
  const int64_t M = (int64_t)a->rows * a->H * a->W;
  const int P = sg_igemm_partials(SG_ENGINE_SIMT, a->H, a->W, a->Cout);
  dim3 grid(cdiv(M, SBM), a->Cout / SBN);
  igemm_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(a->a),
                                                         reinterpret_cast<const float*>(a->w), a->bias, a->residual,
                                                         out, a->partials, M, a->H, a->W, a->Cin, a->Cout, a->taps,
                                                         a->act, P);
  return launch_status("sg_igemm(simt)");
}

int sg_conv_out(const float* in, const float* w, const float* b, int rows, int HW, int c_out, float* eps,
                sg_stream_t stream) {
  SG_REQUIRE(in && w && b && eps, "sg_conv_out: null pointer");
  SG_REQUIRE(rows > 0 && HW > 0 && c_out >= 1 && c_out <= 8, "sg_conv_out: c_out=%d not in 1..8", c_out);
  const int64_t total = (int64_t)rows * HW;
  conv_out_kernel<<<cdiv(total, 128), 128, 0, as_stream(stream)>>>(in, w, b, total, HW, c_out, eps);
  return launch_status("sg_conv_out");
}

}  // extern "C"
