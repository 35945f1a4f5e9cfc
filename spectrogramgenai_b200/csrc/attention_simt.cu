// K4 (SIMT engine): fp32 streaming-softmax self-attention core.  Never materialises the L x L matrix
// (the reference's nn.MultiheadAttention does, and additionally head-averages and discards it:
// /root/reference/src/diff_modules.py:69).  One thread owns one query row (q and the output accumulator
// stay in registers); K/V tiles are staged in shared memory and read as warp-wide broadcasts.
#include "common.cuh"

namespace sg {

int attention_tc(const void* qkv, void* out, int rows, int L, int C, int heads, int act_dtype,
                 cudaStream_t stream);  // attention_tc.cu

template <int D>
__global__ void __launch_bounds__(128) attention_simt_kernel(const float* __restrict__ qkv, float* __restrict__ o32,
                                                             void* __restrict__ o16, int dtype, int L, int C,
                                                             float scale) {
  constexpr int KT = 64;
  __shared__ __align__(16) float Ks[KT][D];
  __shared__ __align__(16) float Vs[KT][D];
  const int tid = threadIdx.x;
  const int head = blockIdx.y;
  const int64_t row = blockIdx.z;
  const int qi = blockIdx.x * 128 + tid;
  const bool active = qi < L;
  const int64_t C3 = 3 * (int64_t)C;
  const float* base = qkv + row * L * C3 + head * D;
  float q[D], o[D];
#pragma unroll
  for (int d4 = 0; d4 < D / 4; ++d4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)qi * C3) + d4);
    q[d4 * 4 + 0] = v.x * scale; q[d4 * 4 + 1] = v.y * scale; q[d4 * 4 + 2] = v.z * scale; q[d4 * 4 + 3] = v.w * scale;
    o[d4 * 4 + 0] = o[d4 * 4 + 1] = o[d4 * 4 + 2] = o[d4 * 4 + 3] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  for (int kt = 0; kt < L; kt += KT) {
    const int kmax = min(KT, L - kt);
    __syncthreads();
    for (int i = tid; i < kmax * (D / 4); i += 128) {
      const int j = i / (D / 4), d4 = i % (D / 4);
      const float* src = base + (int64_t)(kt + j) * C3;
      reinterpret_cast<float4*>(&Ks[j][0])[d4] = __ldg(reinterpret_cast<const float4*>(src + C) + d4);
      reinterpret_cast<float4*>(&Vs[j][0])[d4] = __ldg(reinterpret_cast<const float4*>(src + 2 * C) + d4);
    }
    __syncthreads();
    if (!active) continue;
    for (int j0 = 0; j0 < kmax; j0 += 8) {
      float s[8];
      float tmax = mx;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float a = 0.f;
        if (j0 + jj < kmax) {
#pragma unroll
          for (int d4 = 0; d4 < D / 4; ++d4) {
            const float4 k = reinterpret_cast<const float4*>(&Ks[j0 + jj][0])[d4];
            a = fmaf(q[d4 * 4 + 0], k.x, a);
            a = fmaf(q[d4 * 4 + 1], k.y, a);
            a = fmaf(q[d4 * 4 + 2], k.z, a);
            a = fmaf(q[d4 * 4 + 3], k.w, a);
          }
        } else {
          a = -INFINITY;
        }
        s[jj] = a;
        tmax = fmaxf(tmax, a);
      }
      if (tmax > mx) {
        const float corr = expf(mx - tmax);
        l *= corr;
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] *= corr;
        mx = tmax;
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (j0 + jj < kmax) {
          const float p = expf(s[jj] - mx);
          l += p;
#pragma unroll
          for (int d4 = 0; d4 < D / 4; ++d4) {
            const float4 v = reinterpret_cast<const float4*>(&Vs[j0 + jj][0])[d4];
            o[d4 * 4 + 0] = fmaf(p, v.x, o[d4 * 4 + 0]);
            o[d4 * 4 + 1] = fmaf(p, v.y, o[d4 * 4 + 1]);
            o[d4 * 4 + 2] = fmaf(p, v.z, o[d4 * 4 + 2]);
            o[d4 * 4 + 3] = fmaf(p, v.w, o[d4 * 4 + 3]);
          }
        }
      }
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    const int64_t off = (row * L + qi) * C + head * D;
#pragma unroll
    for (int d4 = 0; d4 < D / 4; ++d4)
      store4_dual(o32, o16, dtype, off + d4 * 4, o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv,
                  o[d4 * 4 + 3] * inv);
  }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_attention(const void* qkv, void* out, int rows, int L, int C, int heads, int engine, int act_dtype,
                            sg_stream_t stream) {
  SG_REQUIRE(qkv && out, "sg_attention: null pointer");
  SG_REQUIRE(rows > 0 && L > 0 && heads > 0 && C % heads == 0, "sg_attention: bad shape rows=%d L=%d C=%d heads=%d",
             rows, L, C, heads);
  const int d = C / heads;
  SG_REQUIRE(d == 16 || d == 32 || d == 64, "sg_attention: head dim %d not in {16,32,64}", d);
  if (engine == SG_ENGINE_TC) return attention_tc(qkv, out, rows, L, C, heads, act_dtype, as_stream(stream));
  SG_REQUIRE(engine == SG_ENGINE_SIMT, "sg_attention: engine %d", engine);
  SG_REQUIRE(act_dtype == SG_F32 || act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_attention: bad out dtype");
  SG_REQUIRE(rows <= 65535 && heads <= 65535, "sg_attention: grid too large");
  dim3 grid(cdiv(L, 128), heads, rows);
  const float scale = 1.0f / sqrtf((float)d);
  const float* q = reinterpret_cast<const float*>(qkv);
  float* o32 = act_dtype == SG_F32 ? reinterpret_cast<float*>(out) : nullptr;
  void* o16 = act_dtype == SG_F32 ? nullptr : out;
  cudaStream_t s = as_stream(stream);
  if (d == 16) attention_simt_kernel<16><<<grid, 128, 0, s>>>(q, o32, o16, act_dtype, L, C, scale);
  else if (d == 32) attention_simt_kernel<32><<<grid, 128, 0, s>>>(q, o32, o16, act_dtype, L, C, scale);
  else attention_simt_kernel<64><<<grid, 128, 0, s>>>(q, o32, o16, act_dtype, L, C, scale);
  return launch_status("sg_attention(simt)");
}
