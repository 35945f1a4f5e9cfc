// K5: sinusoidal timestep encoding + label embedding + SiLU + all six emb_layer projections, one launch.
// Replaces UNet.pos_encoding (/root/reference/src/diff_modules.py:168-173), `t += label_emb(y)` (:214-215)
// and the `emb_layer` (SiLU -> Linear(256, Cout)) of every Down / Up block (:105-108, :126-129).
#include "common.cuh"

namespace sg {

constexpr int TE_ROWS = 8;    // batch rows per block: a slice of the 896 x 256 projection matrix is read once per 8 rows
constexpr int TE_SLICES = 4;  // blockIdx.y: interleaved slices of the output channels (more CTAs than SMs at n = 512)

__global__ void __launch_bounds__(256) time_embed_kernel(const float* __restrict__ t, const int32_t* __restrict__ step,
                                                         const int64_t* __restrict__ y,
                                                         const float* __restrict__ inv_freq,
                                                         const float* __restrict__ label, int num_classes,
                                                         const float* __restrict__ w_emb,
                                                         const float* __restrict__ b_emb, int emb_total, int rows,
                                                         float* __restrict__ temb, float* __restrict__ emb) {
  __shared__ __align__(16) float act[TE_ROWS][256];  // SiLU(temb)
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TE_ROWS;
  for (int rr = 0; rr < TE_ROWS; ++rr) {
    const int row = r0 + rr;
    float v = 0.f;
    if (row < rows) {
      const float tv = t ? t[row] : (float)(*step);
      const int k = tid & 127;
      const float arg = tv * inv_freq[k];  // fp32 product, as `t.repeat(...) * inv_freq` (:171-172)
      v = (tid < 128) ? sinf(arg) : cosf(arg);
      if (y) {
        const int64_t cls = y[row];
        if (cls >= 0 && cls < num_classes) v += label[cls * 256 + tid];
      }
      if (blockIdx.y == 0) temb[(int64_t)row * 256 + tid] = v;
      v = v / (1.0f + expf(-v));  // SiLU
    }
    act[rr][tid] = v;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  // lane -> the row whose total it ends up with after the transposing reduction below
  const int my_rr = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  for (int j = blockIdx.y * 8 + warp; j < emb_total; j += 8 * TE_SLICES) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_emb + (int64_t)j * 256) + lane);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w_emb + (int64_t)j * 256) + 32 + lane);
    const float bj = __ldg(b_emb + j);
    float s[TE_ROWS];
#pragma unroll
    for (int rr = 0; rr < TE_ROWS; ++rr) {
      const float4 a0 = reinterpret_cast<const float4*>(&act[rr][0])[lane];
      const float4 a1 = reinterpret_cast<const float4*>(&act[rr][0])[32 + lane];
      s[rr] = a0.x * w0.x + a0.y * w0.y + a0.z * w0.z + a0.w * w0.w + a1.x * w1.x + a1.y * w1.y + a1.z * w1.z +
              a1.w * w1.w;
    }
    // 8 row sums over 32 lanes in 9 shuffles: each step halves the values a lane carries (the same pairing tree for
    // every row, so a row's result does not depend on its position in the block)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool hi = lane & 16;
      const float send = hi ? s[i] : s[i + 4], keep = hi ? s[i + 4] : s[i];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool hi = lane & 8;
      const float send = hi ? s[i] : s[i + 2], keep = hi ? s[i + 2] : s[i];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
      const bool hi = lane & 4;
      const float send = hi ? s[0] : s[1], keep = hi ? s[1] : s[0];
      s[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 2);
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
    if ((lane & 3) == 0 && r0 + my_rr < rows) emb[(int64_t)(r0 + my_rr) * emb_total + j] = s[0] + bj;
  }
}

}  // namespace sg

using namespace sg;

extern "C" int sg_time_embed(const float* t, const int32_t* step, const int64_t* y, const float* inv_freq,
                             const float* label, int num_classes, const float* w_emb, const float* b_emb,
                             int emb_total, int rows, float* temb, float* emb, sg_stream_t stream) {
  SG_REQUIRE((t || step) && inv_freq && w_emb && b_emb && temb && emb, "sg_time_embed: null pointer");
  SG_REQUIRE(!y || label, "sg_time_embed: labels given without a label table");
  SG_REQUIRE(rows > 0 && emb_total > 0, "sg_time_embed: bad shape");
  launch_k(time_embed_kernel, dim3((unsigned)cdiv(rows, TE_ROWS), TE_SLICES), dim3(256), 0, as_stream(stream), t, step, y, inv_freq, label, num_classes, w_emb,
                                                                        b_emb, emb_total, rows, temb, emb);
  return launch_status("sg_time_embed");
}
