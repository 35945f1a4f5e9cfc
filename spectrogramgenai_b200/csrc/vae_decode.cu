// DiffusionVAE decode tail (/root/reference/src/diff_modules.py:702-706): the pieces around the two igemm launches.
//   sg_vq_quantize   clamp(-1,1) + VQEmbeddingEMA.forward, eval path (:290-318): nearest of 512 four-dimensional
//                    codewords for every group of 4 CONSECUTIVE elements of the NCHW latent (reshape(-1, 4) at :292)
//   sg_dec_in_proj   Decoder.in_proj (:330,:340): 1x1 conv 4 -> 512 + bias, NCHW fp32 in, NHWC fp32 / 16-bit out
//   sg_tconv2_u8     Decoder.strided_t_conv_2 (:336,:350) fused with the image tail (:704-705): the second
//                    ConvTranspose2d(512, 1, k2, s2) is four 512-long dot products per input pixel; its input is the
//                    UN-SHUFFLED output of the first transposed conv ([a][pixel][(b, co)], written by sg_igemm as a
//                    Linear 512 -> 2 x 1024), so the 128 x 128 x 512 intermediate is never re-laid-out
// The 1x1 / 3x3 residual convolutions and the first ConvTranspose2d (a Linear per input pixel) run on sg_igemm with the
// SG_ACT_RELU_POST epilogue.  All of this is ~0.08 % of the sampling FLOPs: the kernels are written for exactness
// against the CPU reference, not for the last GB/s.
#include "common.cuh"

namespace sg {

constexpr int VQ_D = 4;

// torch.cdist(p=2) on these shapes takes the matmul route (_euclidean_dist): d^2 = sum_k (-2 x_k) e_k + |x|^2 + |e|^2,
// clamped at 0, then sqrt; the reference then squares (-d) again (:294).  argmin keeps the first minimum.
__global__ void __launch_bounds__(256) vq_quantize_kernel(const float* __restrict__ x, int64_t groups,
                                                          const float* __restrict__ codebook, int n_codes, int clamp,
                                                          float* __restrict__ quantized, int32_t* __restrict__ indices) {
  extern __shared__ float4 cb[];  // [n_codes] codewords, then [n_codes] squared norms
  float* cn = reinterpret_cast<float*>(cb + n_codes);
  for (int i = threadIdx.x; i < n_codes; i += blockDim.x) {
    const float4 e = reinterpret_cast<const float4*>(codebook)[i];
    cb[i] = e;
    cn[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y)), __fmul_rn(e.z, e.z)), __fmul_rn(e.w, e.w));
  }
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  float4 v = reinterpret_cast<const float4*>(x)[g];
  if (clamp) {
    v.x = fminf(fmaxf(v.x, -1.f), 1.f);
    v.y = fminf(fmaxf(v.y, -1.f), 1.f);
    v.z = fminf(fmaxf(v.z, -1.f), 1.f);
    v.w = fminf(fmaxf(v.w, -1.f), 1.f);
  }
  const float xn = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fmul_rn(v.z, v.z)), __fmul_rn(v.w, v.w));
  const float a0 = -2.f * v.x, a1 = -2.f * v.y, a2 = -2.f * v.z, a3 = -2.f * v.w;
  float best = INFINITY;
  int bi = 0;
  for (int c = 0; c < n_codes; ++c) {
    const float4 e = cb[c];
    float acc = a0 * e.x;
    acc = fmaf(a1, e.y, acc);
    acc = fmaf(a2, e.z, acc);
    acc = fmaf(a3, e.w, acc);
    acc = __fadd_rn(acc, xn);
    acc = __fadd_rn(acc, cn[c]);
    const float d = sqrtf(fmaxf(acc, 0.f));
    const float d2 = __fmul_rn(d, d);
    if (d2 < best) {
      best = d2;
      bi = c;
    }
  }
  const float4 q = cb[bi];
  // straight-through expression of the reference, x + (q - x) (:313): not always bit-identical to q
  float4 o;
  o.x = __fadd_rn(v.x, __fsub_rn(q.x, v.x));
  o.y = __fadd_rn(v.y, __fsub_rn(q.y, v.y));
  o.z = __fadd_rn(v.z, __fsub_rn(q.z, v.z));
  o.w = __fadd_rn(v.w, __fsub_rn(q.w, v.w));
  reinterpret_cast<float4*>(quantized)[g] = o;
  if (indices) indices[g] = bi;
}

// in_proj: thread = 4 output channels of one pixel.  z is NCHW [n, 4, S, S]; out NHWC [n, S, S, 512].
__global__ void __launch_bounds__(256) dec_in_proj_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                          const float* __restrict__ b, int64_t total4, int HW, int Cout4,
                                                          float* __restrict__ o32, void* __restrict__ o16, int dtype) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = (int)(idx % Cout4);
  const int64_t pix = idx / Cout4;
  const int64_t n = pix / HW, p = pix % HW;
  const float* zp = z + n * VQ_D * HW + p;
  const float z0 = __ldg(zp), z1 = __ldg(zp + HW), z2 = __ldg(zp + 2 * HW), z3 = __ldg(zp + 3 * HW);
  float y[4];
  const float4 bv = __ldg(reinterpret_cast<const float4*>(b) + c4);
  const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + c4 * 4 + j);  // w is [512, 4]
    float acc = z0 * wv.x;
    acc = fmaf(z1, wv.y, acc);
    acc = fmaf(z2, wv.z, acc);
    acc = fmaf(z3, wv.w, acc);
    y[j] = acc + bb[j];
  }
  store4_dual(o32, o16, dtype, idx * 4, y[0], y[1], y[2], y[3]);
}

// (x + 1) / 2 * 255 -> uint8 exactly as the CPU cast does it: truncate to int32, keep the low byte (no clamp, :704-705)
__device__ __forceinline__ uint8_t image_u8(float y) {
  const float v = __fmul_rn(__fmul_rn(__fadd_rn(y, 1.0f), 0.5f), 255.0f);
  return (uint8_t)((uint32_t)__float2int_rz(v) & 255u);
}

// One warp = one (input pixel m, a, b): the 512 channels of the first transposed conv's output pixel (2h+a, 2w+b),
// read from T[a][m][b*512 + co]; four dot products with w2[co][a2][b2] give the 2 x 2 output pixels.
template <typename TIn>
__global__ void __launch_bounds__(256) tconv2_u8_kernel(const TIn* __restrict__ t, int dtype, int64_t M, int S, int C,
                                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                                        uint8_t* __restrict__ out_u8, float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= M * 4) return;
  const int ab = (int)(wid & 3);
  const int64_t m = wid >> 2;
  const int a = ab >> 1, b = ab & 1;
  const TIn* src = t + ((int64_t)a * M + m) * (2 * C) + (int64_t)b * C;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int co = lane * 4; co < C; co += 128) {
    float v[4];
    if constexpr (sizeof(TIn) == 4) {
      const float4 f = *reinterpret_cast<const float4*>(src + co);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
      const uint2 h = *reinterpret_cast<const uint2*>(src + co);
      const float2 lo = unpack16(h.x, dtype), hi = unpack16(h.y, dtype);
      v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w2) + co + j);  // w2 is [C, 1, 2, 2]
      acc[0] = fmaf(v[j], wv.x, acc[0]);
      acc[1] = fmaf(v[j], wv.y, acc[1]);
      acc[2] = fmaf(v[j], wv.z, acc[2]);
      acc[3] = fmaf(v[j], wv.w, acc[3]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  if (lane < 4) {
    const float y = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + __ldg(b2);
    const int a2 = lane >> 1, bb2 = lane & 1;
    const int64_t HW = (int64_t)S * S;
    const int64_t n = m / HW;
    const int p = (int)(m % HW);
    const int h = p / S, w = p % S;
    const int64_t OS = 4 * (int64_t)S;
    const int64_t off = (n * OS + (4 * h + 2 * a + a2)) * OS + (4 * w + 2 * b + bb2);
    if (out_u8) out_u8[off] = image_u8(y);
    if (out_f32) out_f32[off] = y;
  }
}

}  // namespace sg

using namespace sg;

extern "C" {

int sg_vq_quantize(const float* x, int64_t count, const float* codebook, int n_codes, int clamp, float* quantized,
                   int32_t* indices, sg_stream_t stream) {
  SG_REQUIRE(x && codebook && quantized, "sg_vq_quantize: null pointer");
  SG_REQUIRE(count > 0 && count % VQ_D == 0, "sg_vq_quantize: count=%lld must be a positive multiple of 4", (long long)count);
  SG_REQUIRE(n_codes > 0 && n_codes <= 2048, "sg_vq_quantize: n_codes=%d not in 1..2048", n_codes);
  const int64_t groups = count / VQ_D;
  const size_t smem = (size_t)n_codes * (sizeof(float4) + sizeof(float));
  vq_quantize_kernel<<<cdiv(groups, 256), 256, smem, as_stream(stream)>>>(x, groups, codebook, n_codes, clamp, quantized,
                                                                         indices);
  return launch_status("sg_vq_quantize");
}

int sg_dec_in_proj(const float* z, const float* w, const float* b, int n, int S, int Cout, float* out_f32, void* out_act,
                   int act_dtype, sg_stream_t stream) {
  SG_REQUIRE(z && w && b && (out_f32 || out_act), "sg_dec_in_proj: null pointer");
  SG_REQUIRE(n > 0 && S > 0 && Cout % 4 == 0, "sg_dec_in_proj: bad shape n=%d S=%d Cout=%d", n, S, Cout);
  SG_REQUIRE(!out_act || act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_dec_in_proj: out_act needs a 16-bit dtype");
  const int64_t total4 = (int64_t)n * S * S * (Cout / 4);
  dec_in_proj_kernel<<<cdiv(total4, 256), 256, 0, as_stream(stream)>>>(z, w, b, total4, S * S, Cout / 4, out_f32, out_act,
                                                                      act_dtype);
  return launch_status("sg_dec_in_proj");
}

int sg_tconv2_u8(const void* t, int t_dtype, int n, int S, int C, const float* w2, const float* b2, uint8_t* out_u8,
                 float* out_f32, sg_stream_t stream) {
  SG_REQUIRE(t && w2 && b2 && (out_u8 || out_f32), "sg_tconv2_u8: null pointer");
  SG_REQUIRE(n > 0 && S > 0 && C % 128 == 0, "sg_tconv2_u8: bad shape n=%d S=%d C=%d", n, S, C);
  const int64_t M = (int64_t)n * S * S;
  const int64_t warps = M * 4;
  const unsigned blocks = (unsigned)cdiv(warps, 8);
  if (t_dtype == SG_F32)
    tconv2_u8_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(t), t_dtype, M, S, C, w2,
                                                                  b2, out_u8, out_f32);
  else
    tconv2_u8_kernel<uint16_t><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint16_t*>(t), t_dtype, M, S,
                                                                     C, w2, b2, out_u8, out_f32);
  return launch_status("sg_tconv2_u8");
}

}  // extern "C"
