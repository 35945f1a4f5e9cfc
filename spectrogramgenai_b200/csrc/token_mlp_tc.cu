// Fused per-token halves of SelfAttention (/root/reference/src/diff_modules.py:52-72) around the attention core,
// tcgen05 engine, C = 64 (sa5 / sa6 at 64x64 latents: 80 % of all attention tokens):
//
//   sg_ln_inproj :  qkv = LayerNorm(x) Win^T + bin                                         (:67 self.ln, :69 in_proj)
//   sg_attn_tail :  a = att Wo^T + bo + x;  h = GELU(LayerNorm(a) W1^T + b1);  out = h W2^T + b2 + a   (:69-71)
//
// Unfused, the tail is four launches (out_proj, LayerNorm, FFN1, FFN2) that move the [M, C] activation eleven times
// (9.7 GB at sa6, n = 512) and the head two launches (2.7 GB + 1.6 GB written); fused they read att (bf16) and x
// (fp32) once and write out once (2.7 GB), resp. read x and write qkv (2.7 GB).
//
// One CTA owns a tile of 128 tokens at a time (persistent over tiles).  Thread r of the four warps IS token r: it is
// TMEM lane r, so bias / residual / LayerNorm / GELU of a token are pure register math (LayerNorm needs no
// shuffles).  All global traffic is TMA: the fp32 tiles are loaded / stored as two SWIZZLE_128B boxes of 32 floats
// so that a thread walking its own 256-byte row is bank-conflict free, the 16-bit tiles are SWIZZLE_128B K-major
// UMMA operands.  The A operand of every GEMM after the first is written by the row threads straight into the
// operand tile (same swizzle the TMA would have produced).  Weights (3 x 8 KB) stay in shared memory for the
// lifetime of the CTA.  A tile is a serial chain (load -> GEMM -> epilogue -> GEMM -> ...); three (tail) / two
// (in_proj) CTAs per SM overlap each other's chains.
#include "tc_common.cuh"

namespace sg {
namespace tc {

constexpr int TM = 128;          // tokens per tile
constexpr int TC = 64;           // channels
constexpr int A_TILE = TM * TC * 2;       // 16 KB: [128 x 64] 16-bit, one SWIZZLE_128B atom column
constexpr int X_TILE = TM * TC * 4;       // 32 KB: [128 x 64] fp32 as two [128 x 32] SWIZZLE_128B boxes
constexpr int W_TILE = TC * TC * 2;       // 8 KB : [64 x 64] 16-bit weights (rows = output features, K-major)

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int DT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t w;
  if constexpr (DT == SG_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  return w;
}

// this thread's fp32 row (64 floats) of a [128 x 64] tile stored as two SWIZZLE_128B [128 x 32] boxes
__device__ __forceinline__ uint32_t xrow_chunk_addr(uint32_t tile, int r, int c /*16-byte chunk 0..15*/) {
  return tile + (uint32_t)(c >> 3) * (TM * 128) + (uint32_t)r * 128u + ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4);
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// LayerNorm of the 64 values a thread holds (two-pass like torch, eps 1e-5), packed to 16 bit and written as row r
// of a SWIZZLE_128B K-major operand tile.
template <int DT>
__device__ __forceinline__ void ln_row_to_operand(const float (&a)[TC], const float* __restrict__ s_gamma,
                                                  const float* __restrict__ s_beta, uint32_t tile, int r) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < TC; ++j) s += a[j];
  const float mean = s * (1.0f / TC);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < TC; ++j) {
    const float d = a[j] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(q * (1.0f / TC) + 1e-5f);
  const uint32_t row = tile + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
  for (int c = 0; c < 8; ++c) {  // 8 channels = one 16-byte chunk
    const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c * 8), g1 = *reinterpret_cast<const float4*>(s_gamma + c * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c * 8), b1 = *reinterpret_cast<const float4*>(s_beta + c * 8 + 4);
    const float y0 = fmaf((a[c * 8 + 0] - mean) * rstd, g0.x, b0.x), y1 = fmaf((a[c * 8 + 1] - mean) * rstd, g0.y, b0.y);
    const float y2 = fmaf((a[c * 8 + 2] - mean) * rstd, g0.z, b0.z), y3 = fmaf((a[c * 8 + 3] - mean) * rstd, g0.w, b0.w);
    const float y4 = fmaf((a[c * 8 + 4] - mean) * rstd, g1.x, b1.x), y5 = fmaf((a[c * 8 + 5] - mean) * rstd, g1.y, b1.y);
    const float y6 = fmaf((a[c * 8 + 6] - mean) * rstd, g1.z, b1.z), y7 = fmaf((a[c * 8 + 7] - mean) * rstd, g1.w, b1.w);
    sts128u(row + ((((uint32_t)c) ^ ((uint32_t)r & 7u)) << 4), pack2<DT>(y0, y1), pack2<DT>(y2, y3), pack2<DT>(y4, y5),
            pack2<DT>(y6, y7));
  }
}

// D[tmem, 128 x N] = A[smem 128 x 64, SW128 K-major] * B[smem N x 64, SW128 K-major]^T : four K = 16 steps
__device__ __forceinline__ void gemm_k64(uint32_t tmem_d, uint32_t a_tile, uint32_t b_tile, uint32_t idesc) {
  const uint64_t ad = make_desc_k128(a_tile), bd = make_desc_k128(b_tile);
#pragma unroll
  for (int k = 0; k < TC / 16; ++k) umma_ss(tmem_d, ad + 2 * k, bd + 2 * k, idesc, k != 0);
}

struct TailParams {
  const float* bo;
  const float* ln_g;
  const float* ln_b;
  const float* b1;
  const float* b2;
  int64_t M;
  int ntiles;
  uint32_t idesc;
};

// ------------------------------------------------------------------------------------------------------------------
// sg_attn_tail
// ------------------------------------------------------------------------------------------------------------------
constexpr int TAIL_SMEM = 1024 + 3 * W_TILE + A_TILE + X_TILE + 5 * TC * 4 + 128;

template <int DT>
__global__ void __launch_bounds__(128, 3)
attn_tail_kernel(const __grid_constant__ CUtensorMap tm_att, const __grid_constant__ CUtensorMap tm_x,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_wo,
                 const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                 const TailParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sW = smem;                    // Wo | W1 | W2
  uint8_t* sA = sW + 3 * W_TILE;         // operand tile: att, then LN(a), then GELU(.)
  uint8_t* sX = sA + A_TILE;             // x tile (fp32), later the output tile
  float* sPar = reinterpret_cast<float*>(sX + X_TILE);  // bo | ln_g | ln_b | b1 | b2
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 5 * TC);
  uint64_t* w_full = bars;
  uint64_t* in_full = bars + 1;
  uint64_t* mma_done = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool leader = tid == 0;
  if (leader) {
    prefetch_tensormap(&tm_att);
    prefetch_tensormap(&tm_x);
    prefetch_tensormap(&tm_out);
    mbar_init(w_full, 1);
    mbar_init(in_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 5 * TC; i += 128) {
    const float* src = i < TC ? p.bo : (i < 2 * TC ? p.ln_g : (i < 3 * TC ? p.ln_b : (i < 4 * TC ? p.b1 : p.b2)));
    sPar[i] = src[i & (TC - 1)];
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<128>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t aA = smem_u32(sA), aX = smem_u32(sX), aW = smem_u32(sW);
  const int r = tid;

  // TMA / MMA instructions are issued by one elected lane of the CONVERGENT warp 0 (descriptors stay in uniform registers)
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(w_full, 3 * W_TILE);
    tma_load_2d(sW, &tm_wo, w_full, 0, 0);
    tma_load_2d(sW + W_TILE, &tm_w1, w_full, 0, 0);
    tma_load_2d(sW + 2 * W_TILE, &tm_w2, w_full, 0, 0);
  }
  uint32_t in_ph = 0, mma_ph = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int m0 = tile * TM;
    if (warp == 0) {
      if (elect_one()) {
        tma_store_wait_read();  // the previous tile's output (staged in sX) has left shared memory
        mbar_arrive_expect_tx(in_full, A_TILE + X_TILE);
        tma_load_2d(sA, &tm_att, in_full, 0, m0);
        tma_load_2d(sX, &tm_x, in_full, 0, m0);
        tma_load_2d(sX + TM * 128, &tm_x, in_full, 32, m0);
      }
      if (first) mbar_wait_spin(w_full, 0);
      mbar_wait_spin(in_full, in_ph);
      tc_fence_after();
      if (elect_one()) {
        gemm_k64(tmem_base, aA, aW, p.idesc);  // att Wo^T
        umma_commit(mma_done);
      }
      __syncwarp();
    }
    first = false;
    // ---- a = att Wo^T + bo + x ; LN(a) -> operand ----
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    mbar_wait(in_full, in_ph);  // the x tile is visible to this thread (acquire on the TMA barrier)
    in_ph ^= 1u;
    float a[TC];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      tmem_ld32(t_row + h * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 xv = lds128(xrow_chunk_addr(aX, r, h * 8 + c));
        const float4 bv = *reinterpret_cast<const float4*>(sPar + h * 32 + c * 4);
        a[h * 32 + c * 4 + 0] = __uint_as_float(v[c * 4 + 0]) + bv.x + xv.x;
        a[h * 32 + c * 4 + 1] = __uint_as_float(v[c * 4 + 1]) + bv.y + xv.y;
        a[h * 32 + c * 4 + 2] = __uint_as_float(v[c * 4 + 2]) + bv.z + xv.z;
        a[h * 32 + c * 4 + 3] = __uint_as_float(v[c * 4 + 3]) + bv.w + xv.w;
      }
    }
    ln_row_to_operand<DT>(a, sPar + TC, sPar + 2 * TC, aA, r);  // the GEMM that read sA has completed (mma_done)
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        gemm_k64(tmem_base + 64, aA, aW + W_TILE, p.idesc);  // LN(a) W1^T
        umma_commit(mma_done);
      }
      __syncwarp();
    }
    // ---- h = GELU(. + b1) -> operand ----
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    {
      const uint32_t row = aA + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(t_row + 64 + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 8 channels per 16-byte chunk
          const float4 b0 = *reinterpret_cast<const float4*>(sPar + 3 * TC + h * 32 + c * 8);
          const float4 b1v = *reinterpret_cast<const float4*>(sPar + 3 * TC + h * 32 + c * 8 + 4);
          float g0 = __uint_as_float(v[c * 8 + 0]) + b0.x, g1 = __uint_as_float(v[c * 8 + 1]) + b0.y;
          float g2 = __uint_as_float(v[c * 8 + 2]) + b0.z, g3 = __uint_as_float(v[c * 8 + 3]) + b0.w;
          float g4 = __uint_as_float(v[c * 8 + 4]) + b1v.x, g5 = __uint_as_float(v[c * 8 + 5]) + b1v.y;
          float g6 = __uint_as_float(v[c * 8 + 6]) + b1v.z, g7 = __uint_as_float(v[c * 8 + 7]) + b1v.w;
          gelu_erf2(g0, g1);
          gelu_erf2(g2, g3);
          gelu_erf2(g4, g5);
          gelu_erf2(g6, g7);
          sts128u(row + ((((uint32_t)(h * 4 + c)) ^ ((uint32_t)r & 7u)) << 4), pack2<DT>(g0, g1), pack2<DT>(g2, g3),
                  pack2<DT>(g4, g5), pack2<DT>(g6, g7));
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        gemm_k64(tmem_base, aA, aW + 2 * W_TILE, p.idesc);  // h W2^T
        umma_commit(mma_done);
      }
      __syncwarp();
    }
    // ---- out = . + b2 + a -> fp32 tile (over this thread's own x row) -> TMA store ----
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      tmem_ld32(t_row + h * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 bv = *reinterpret_cast<const float4*>(sPar + 4 * TC + h * 32 + c * 4);
        float4 o;
        o.x = __uint_as_float(v[c * 4 + 0]) + bv.x + a[h * 32 + c * 4 + 0];
        o.y = __uint_as_float(v[c * 4 + 1]) + bv.y + a[h * 32 + c * 4 + 1];
        o.z = __uint_as_float(v[c * 4 + 2]) + bv.z + a[h * 32 + c * 4 + 2];
        o.w = __uint_as_float(v[c * 4 + 3]) + bv.w + a[h * 32 + c * 4 + 3];
        sts128(xrow_chunk_addr(aX, r, h * 8 + c), o);
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tma_store_2d(&tm_out, sX, 0, m0);
      tma_store_2d(&tm_out, sX + TM * 128, 32, m0);
      tma_store_commit();
    }
  }
  if (warp == 0 && elect_one()) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<128>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// sg_ln_inproj : qkv[M, 192] = LayerNorm(x) Win^T + bin
// ------------------------------------------------------------------------------------------------------------------
struct InprojParams {
  const float* ln_g;
  const float* ln_b;
  const float* bias;  // [192]
  int64_t M;
  int ntiles;
  uint32_t idesc;     // M128 x N192
};
constexpr int INPROJ_SMEM = 1024 + 3 * W_TILE + X_TILE + A_TILE + 5 * TC * 4 + 128;

template <int DT>
__global__ void __launch_bounds__(128, 2)
ln_inproj_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_qkv, const InprojParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sW = smem;             // Win: [192 x 64] 16-bit, SWIZZLE_128B (24 KB)
  uint8_t* sX = sW + 3 * W_TILE;  // x tile (fp32, 32 KB); with sA the 48 KB staging of the three 16-bit output boxes
  uint8_t* sA = sX + X_TILE;      // LN(x) operand tile (16 KB)
  float* sPar = reinterpret_cast<float*>(sA + A_TILE);  // ln_g | ln_b | bias[192]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 5 * TC);
  uint64_t* w_full = bars;
  uint64_t* in_full = bars + 1;
  uint64_t* mma_done = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool leader = tid == 0;
  if (leader) {
    prefetch_tensormap(&tm_x);
    prefetch_tensormap(&tm_qkv);
    mbar_init(w_full, 1);
    mbar_init(in_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 5 * TC; i += 128) sPar[i] = i < TC ? p.ln_g[i] : (i < 2 * TC ? p.ln_b[i - TC] : p.bias[i - 2 * TC]);
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t aA = smem_u32(sA), aX = smem_u32(sX), aW = smem_u32(sW);
  const int r = tid;

  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(w_full, 3 * W_TILE);
    tma_load_2d(sW, &tm_w, w_full, 0, 0);
  }
  uint32_t in_ph = 0, mma_ph = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int m0 = tile * TM;
    if (warp == 0 && elect_one()) {
      tma_store_wait_read();  // the previous tile's qkv boxes (staged in sX | sA) have left shared memory
      mbar_arrive_expect_tx(in_full, X_TILE);
      tma_load_2d(sX, &tm_x, in_full, 0, m0);
      tma_load_2d(sX + TM * 128, &tm_x, in_full, 32, m0);
    }
    __syncthreads();  // nobody writes sA (LN output) before the previous stores have drained
    mbar_wait(in_full, in_ph);
    in_ph ^= 1u;
    float a[TC];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 xv = lds128(xrow_chunk_addr(aX, r, c));
      a[c * 4 + 0] = xv.x;
      a[c * 4 + 1] = xv.y;
      a[c * 4 + 2] = xv.z;
      a[c * 4 + 3] = xv.w;
    }
    ln_row_to_operand<DT>(a, sPar, sPar + TC, aA, r);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (first) mbar_wait_spin(w_full, 0);
      tc_fence_after();
      if (elect_one()) {
        gemm_k64(tmem_base, aA, aW, p.idesc);  // one M128 x N192 accumulator
        umma_commit(mma_done);
      }
      __syncwarp();
    }
    first = false;
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    // qkv rows -> three [128 x 64] 16-bit SWIZZLE_128B boxes at sX, sX + 16 KB, sX + 32 KB (= sA: its GEMM is complete)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const uint32_t row = aX + (uint32_t)b * A_TILE + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(t_row + b * 64 + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = *reinterpret_cast<const float4*>(sPar + 2 * TC + b * 64 + h * 32 + c * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(sPar + 2 * TC + b * 64 + h * 32 + c * 8 + 4);
          sts128u(row + ((((uint32_t)(h * 4 + c)) ^ ((uint32_t)r & 7u)) << 4),
                  pack2<DT>(__uint_as_float(v[c * 8 + 0]) + b0.x, __uint_as_float(v[c * 8 + 1]) + b0.y),
                  pack2<DT>(__uint_as_float(v[c * 8 + 2]) + b0.z, __uint_as_float(v[c * 8 + 3]) + b0.w),
                  pack2<DT>(__uint_as_float(v[c * 8 + 4]) + b1.x, __uint_as_float(v[c * 8 + 5]) + b1.y),
                  pack2<DT>(__uint_as_float(v[c * 8 + 6]) + b1.z, __uint_as_float(v[c * 8 + 7]) + b1.w));
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tma_store_2d(&tm_qkv, sX, 0, m0);
      tma_store_2d(&tm_qkv, sX + A_TILE, 64, m0);
      tma_store_2d(&tm_qkv, sX + 2 * A_TILE, 128, m0);
      tma_store_commit();
    }
  }
  if (warp == 0 && elect_one()) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<256>(tmem_base);
  }
}

static int tmap2d(CUtensorMap* out, int dtype, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner,
                  uint32_t box_rows) {
  const uint64_t esz = dtype == SG_F32 ? 4 : 2;
  const uint64_t dims[2] = {inner, rows};
  const uint64_t strides[1] = {inner * esz};
  const uint32_t box[2] = {box_inner, box_rows};
  return make_tmap(out, dtype, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <typename K>
static int set_smem(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%d B smem): %s", what, bytes, cudaGetErrorString(e));
    return SG_ERR_LAUNCH;
  }
  return SG_OK;
}

}  // namespace tc
}  // namespace sg

using namespace sg;
using namespace sg::tc;

extern "C" {

int sg_attn_tail(const void* att, const float* x, const void* wo, const float* bo, const float* ln_g, const float* ln_b,
                 const void* w1, const float* b1, const void* w2, const float* b2, int64_t M, int C, float* out,
                 int act_dtype, sg_stream_t stream) {
  SG_REQUIRE(att && x && wo && bo && ln_g && ln_b && w1 && b1 && w2 && b2 && out, "sg_attn_tail: null pointer");
  SG_REQUIRE(C == TC, "sg_attn_tail: C=%d (the fused kernel is built for C = 64; use the unfused launches otherwise)", C);
  SG_REQUIRE(act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_attn_tail: act_dtype must be SG_BF16 or SG_F16");
  SG_REQUIRE(M > 0 && M < (1ll << 31) - TM, "sg_attn_tail: M=%lld", (long long)M);
  CUtensorMap tm_att, tm_x, tm_out, tm_wo, tm_w1, tm_w2;
  int rc;
  if ((rc = tmap2d(&tm_att, act_dtype, att, TC, (uint64_t)M, TC, TM))) return rc;
  if ((rc = tmap2d(&tm_x, SG_F32, x, TC, (uint64_t)M, 32, TM))) return rc;
  if ((rc = tmap2d(&tm_out, SG_F32, out, TC, (uint64_t)M, 32, TM))) return rc;
  if ((rc = tmap2d(&tm_wo, act_dtype, wo, TC, TC, TC, TC))) return rc;
  if ((rc = tmap2d(&tm_w1, act_dtype, w1, TC, TC, TC, TC))) return rc;
  if ((rc = tmap2d(&tm_w2, act_dtype, w2, TC, TC, TC, TC))) return rc;
  TailParams p;
  p.bo = bo; p.ln_g = ln_g; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2;
  p.M = M;
  p.ntiles = (int)cdiv(M, TM);
  p.idesc = make_idesc(act_dtype, 128, TC, 0, 0);
  const int grid = p.ntiles < 3 * num_sms() ? p.ntiles : 3 * num_sms();
  cudaStream_t s = as_stream(stream);
  if (act_dtype == SG_BF16) {
    static bool cfg = false;
    if (!cfg) {
      if ((rc = set_smem(attn_tail_kernel<SG_BF16>, TAIL_SMEM, "sg_attn_tail"))) return rc;
      cfg = true;
    }
    attn_tail_kernel<SG_BF16><<<grid, 128, TAIL_SMEM, s>>>(tm_att, tm_x, tm_out, tm_wo, tm_w1, tm_w2, p);
  } else {
    static bool cfg = false;
    if (!cfg) {
      if ((rc = set_smem(attn_tail_kernel<SG_F16>, TAIL_SMEM, "sg_attn_tail"))) return rc;
      cfg = true;
    }
    attn_tail_kernel<SG_F16><<<grid, 128, TAIL_SMEM, s>>>(tm_att, tm_x, tm_out, tm_wo, tm_w1, tm_w2, p);
  }
  return launch_status("sg_attn_tail");
}

int sg_ln_inproj(const float* x, const float* ln_g, const float* ln_b, const void* w_in, const float* b_in, int64_t M,
                 int C, void* qkv, int act_dtype, sg_stream_t stream) {
  SG_REQUIRE(x && ln_g && ln_b && w_in && b_in && qkv, "sg_ln_inproj: null pointer");
  SG_REQUIRE(C == TC, "sg_ln_inproj: C=%d (the fused kernel is built for C = 64; use the unfused launches otherwise)", C);
  SG_REQUIRE(act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_ln_inproj: act_dtype must be SG_BF16 or SG_F16");
  SG_REQUIRE(M > 0 && M < (1ll << 31) - TM, "sg_ln_inproj: M=%lld", (long long)M);
  CUtensorMap tm_x, tm_w, tm_qkv;
  int rc;
  if ((rc = tmap2d(&tm_x, SG_F32, x, TC, (uint64_t)M, 32, TM))) return rc;
  if ((rc = tmap2d(&tm_w, act_dtype, w_in, TC, 3 * TC, TC, 3 * TC))) return rc;
  if ((rc = tmap2d(&tm_qkv, act_dtype, qkv, 3 * TC, (uint64_t)M, TC, TM))) return rc;
  InprojParams p;
  p.ln_g = ln_g; p.ln_b = ln_b; p.bias = b_in;
  p.M = M;
  p.ntiles = (int)cdiv(M, TM);
  p.idesc = make_idesc(act_dtype, 128, 3 * TC, 0, 0);
  const int grid = p.ntiles < 2 * num_sms() ? p.ntiles : 2 * num_sms();
  cudaStream_t s = as_stream(stream);
  if (act_dtype == SG_BF16) {
    static bool cfg = false;
    if (!cfg) {
      if ((rc = set_smem(ln_inproj_kernel<SG_BF16>, INPROJ_SMEM, "sg_ln_inproj"))) return rc;
      cfg = true;
    }
    ln_inproj_kernel<SG_BF16><<<grid, 128, INPROJ_SMEM, s>>>(tm_x, tm_w, tm_qkv, p);
  } else {
    static bool cfg = false;
    if (!cfg) {
      if ((rc = set_smem(ln_inproj_kernel<SG_F16>, INPROJ_SMEM, "sg_ln_inproj"))) return rc;
      cfg = true;
    }
    ln_inproj_kernel<SG_F16><<<grid, 128, INPROJ_SMEM, s>>>(tm_x, tm_w, tm_qkv, p);
  }
  return launch_status("sg_ln_inproj");
}

}  // extern "C"
