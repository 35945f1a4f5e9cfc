// Fused per-token halves of SelfAttention (/root/reference/src/diff_modules.py:52-72) around the attention core,
// tcgen05 engine, C = 64 (sa5 / sa6 at 64x64 latents: 80 % of all attention tokens) and C = 128 (sa1 / sa4):
//
//   sg_ln_inproj :  qkv = LayerNorm(x) Win^T + bin                                         (:67 self.ln, :69 in_proj)
//   sg_attn_tail :  a = att Wo^T + bo + x;  h = GELU(LayerNorm(a) W1^T + b1);  out = h W2^T + b2 + a   (:69-71)
//
// Unfused, the tail is four launches (out_proj, LayerNorm, FFN1, FFN2) that move the [M, C] activation eleven times
// (9.7 GB at sa6, n = 512) and the head two launches (2.7 GB + 1.6 GB written); fused they read att (16 bit) and x
// (fp32) once and write out once (2.7 GB), resp. read x and write qkv (2.7 GB).
//
// One CTA owns a tile of 128 tokens at a time (persistent over tiles).  Thread r of the four warps IS token r: it is
// TMEM lane r, so bias / residual / LayerNorm / GELU of a token are pure register math (LayerNorm needs no
// shuffles).  All global traffic is TMA: the fp32 tiles are loaded / stored as C/32 SWIZZLE_128B boxes of 32 floats
// so that a thread walking its own row is bank-conflict free, the 16-bit tiles are SWIZZLE_128B K-major UMMA operands
// (C/64 atom columns of 64 channels).  The A operand of every GEMM after the first is written by the row threads
// straight into the operand tile (same swizzle the TMA would have produced).  Weights stay in shared memory for the
// lifetime of the CTA (24 KB at C = 64, 96 KB at C = 128).  A tile is a serial chain (load -> GEMM -> epilogue ->
// GEMM -> ...); at C = 64 three (tail) / two (in_proj) CTAs per SM overlap each other's chains, at C = 128 the
// weights leave room for one.
#include "tc_common.cuh"

namespace sg {
namespace tc {

constexpr int TM = 128;  // tokens per tile

template <int C>
struct Tok {
  static constexpr int KB = C / 64;             // 64-channel k-blocks = SWIZZLE_128B atom columns of a 16-bit tile
  static constexpr int A_ATOM = TM * 128;       // 16 KB: [128 x 64] 16-bit
  static constexpr int A_TILE = KB * A_ATOM;    // [128 x C] 16-bit
  static constexpr int XB = C / 32;             // fp32 boxes of 32 floats
  static constexpr int X_TILE = XB * TM * 128;  // [128 x C] fp32
  static constexpr int W_KB = C * 128;          // one k-block of a [C x C] weight: C rows x 128 bytes
  static constexpr int W_TILE = KB * W_KB;      // [C x C] 16-bit
  static constexpr int WIN_KB = 3 * C * 128;    // one k-block of Win [3C x C]
  static constexpr int WIN_TILE = KB * WIN_KB;
  static constexpr bool TAIL_DIRECT = C == 128;  // block output stored through per-warp transposition scratch, no TMA staging
  static constexpr int TAIL_SCR = TAIL_DIRECT ? 4 * 4096 : 0;  // 4 warps x [32 rows x 32 floats]
  static constexpr int TAIL_SMEM = 1024 + 3 * W_TILE + A_TILE + X_TILE + TAIL_SCR + 5 * C * 4 + 128;
  static constexpr int INPROJ_XBUF = C == 64 ? 2 : 1;  // x tiles in flight (C = 64: the next tile's x is prefetched)
  static constexpr bool INPROJ_DIRECT = C == 128;  // epilogue stores qkv through per-warp transposition scratch, no TMA staging
  static constexpr int INPROJ_SCR = INPROJ_DIRECT ? 4 * 4096 : 0;  // 4 warps x [32 rows x 64 cols] 16 bit
  static constexpr int INPROJ_SMEM = 1024 + WIN_TILE + INPROJ_XBUF * X_TILE + A_TILE + INPROJ_SCR + 5 * C * 4 + 128;
  static constexpr int TAIL_TMEM = 2 * C;               // two accumulators [128 x C]: 128 / 256 columns
  static constexpr int INPROJ_TMEM = C == 64 ? 256 : 512;  // one accumulator [128 x 3C]
  static constexpr int TAIL_CTAS = C == 64 ? 3 : 1;
  static constexpr int INPROJ_CTAS = C == 64 ? 2 : 1;
};

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

// 16-byte chunk c (4 floats) of this thread's fp32 row in a [128 x C] tile stored as C/32 SWIZZLE_128B [128 x 32] boxes
__device__ __forceinline__ uint32_t xrow_chunk_addr(uint32_t tile, int r, int c) {
  return tile + (uint32_t)(c >> 3) * (TM * 128) + (uint32_t)r * 128u + ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4);
}
// 16-byte chunk c (8 x 16 bit) of this thread's row in a [128 x C] K-major operand tile (C/64 atom columns)
__device__ __forceinline__ uint32_t arow_chunk_addr(uint32_t tile, int r, int c) {
  return tile + (uint32_t)(c >> 3) * (TM * 128) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u +
         ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4);
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

// LayerNorm of the C values a thread holds (two-pass like torch, eps 1e-5), packed to 16 bit and written as row r
// of a SWIZZLE_128B K-major operand tile.
template <int C, int DT>
__device__ __forceinline__ void ln_row_to_operand(const float (&a)[C], const float* __restrict__ s_gamma,
                                                  const float* __restrict__ s_beta, uint32_t tile, int r) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) s += a[j];
  const float mean = s * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const float d = a[j] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(q * (1.0f / C) + 1e-5f);
#pragma unroll
  for (int c = 0; c < C / 8; ++c) {  // 8 channels = one 16-byte chunk
    const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + c * 8), g1 = *reinterpret_cast<const float4*>(s_gamma + c * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(s_beta + c * 8), b1 = *reinterpret_cast<const float4*>(s_beta + c * 8 + 4);
    const float y0 = fmaf((a[c * 8 + 0] - mean) * rstd, g0.x, b0.x), y1 = fmaf((a[c * 8 + 1] - mean) * rstd, g0.y, b0.y);
    const float y2 = fmaf((a[c * 8 + 2] - mean) * rstd, g0.z, b0.z), y3 = fmaf((a[c * 8 + 3] - mean) * rstd, g0.w, b0.w);
    const float y4 = fmaf((a[c * 8 + 4] - mean) * rstd, g1.x, b1.x), y5 = fmaf((a[c * 8 + 5] - mean) * rstd, g1.y, b1.y);
    const float y6 = fmaf((a[c * 8 + 6] - mean) * rstd, g1.z, b1.z), y7 = fmaf((a[c * 8 + 7] - mean) * rstd, g1.w, b1.w);
    sts128u(arow_chunk_addr(tile, r, c), pack2<DT>(y0, y1), pack2<DT>(y2, y3), pack2<DT>(y4, y5), pack2<DT>(y6, y7));
  }
}

// D[tmem, 128 x N] (+)= A[smem 128 x C, SW128 K-major] * B[smem rows [row0, row0 + N) of a [ROWS x C] weight]^T.
// The weight is stored as C/64 k-blocks of ROWS x 128 bytes; called by one elected lane.
template <int C>
__device__ __forceinline__ void gemm_kc(uint32_t tmem_d, uint32_t a_tile, uint32_t b_tile, int b_kb_bytes, int row0,
                                        uint32_t idesc) {
#pragma unroll
  for (int kb = 0; kb < C / 64; ++kb) {
    const uint64_t ad = make_desc_k128(a_tile + kb * (TM * 128));
    const uint64_t bd = make_desc_k128(b_tile + kb * b_kb_bytes + row0 * 128);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
  }
}

constexpr int OUTC_C = 64;  // the fused output conv follows the C = 64 block sa6
struct TailParams {
  const float* bo;
  const float* ln_g;
  const float* ln_b;
  const float* b1;
  const float* b2;
  int64_t M;
  int ntiles;
  uint32_t idesc;
  // fused outc (sg_attn_tail_outc): eps[row, k, pix] = outc_b[k] + sum_c out[token, c] * outc_w[k, c]
  float* out;     // direct-store epilogue (C = 128): fp32 [M, C]
  float* eps;     // NCHW [rows, c_out, HW]
  int c_out;      // 1..4
  int log_hw;     // HW is a power of two
  int write_out;  // also write the fp32 token tensor (debug taps)
  // 1x1 output conv weights [4][OUTC_C] (rows >= c_out zero) followed by the 4 biases, passed BY VALUE: kernel parameters
  // live in the constant bank, every thread reads the same word at the same time, so they are FFMA constant-bank operands
  // (no load instructions) -- and, unlike a __constant__ array rewritten per call, belong to this launch alone
  float outc[4 * 64 + 4];
};

template <int C, int DT, bool OUTC>
__global__ void __launch_bounds__(128, Tok<C>::TAIL_CTAS)
attn_tail_kernel(const __grid_constant__ CUtensorMap tm_att, const __grid_constant__ CUtensorMap tm_x,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_wo,
                 const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                 const TailParams p) {
  using T = Tok<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sW = smem;                  // Wo | W1 | W2
  uint8_t* sA = sW + 3 * T::W_TILE;    // operand tile: att, then LN(a), then GELU(.)
  uint8_t* sX = sA + T::A_TILE;        // x tile (fp32), later the output tile
  uint8_t* sScr = sX + T::X_TILE;      // C = 128: per-warp transposition scratch of the direct-store epilogue
  float* sPar = reinterpret_cast<float*>(sScr + T::TAIL_SCR);  // bo | ln_g | ln_b | b1 | b2
  constexpr bool DIRECT = T::TAIL_DIRECT && !OUTC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 5 * C);
  uint64_t* w_full = bars;
  uint64_t* att_full = bars + 1;
  uint64_t* mma_done = bars + 2;
  uint64_t* x_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    prefetch_tensormap(&tm_att);
    prefetch_tensormap(&tm_x);
    prefetch_tensormap(&tm_out);
    mbar_init(w_full, 1);
    mbar_init(att_full, 1);
    mbar_init(x_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 5 * C; i += 128) {
    const float* src = i < C ? p.bo : (i < 2 * C ? p.ln_g : (i < 3 * C ? p.ln_b : (i < 4 * C ? p.b1 : p.b2)));
    sPar[i] = src[i & (C - 1)];
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<T::TAIL_TMEM>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // every activation access (TMA loads included) follows this point; parameters read above are immutable weights
  pdl_launch_dependents();
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t aA = smem_u32(sA), aX = smem_u32(sX), aW = smem_u32(sW);
  const int r = tid;

  // TMA / MMA instructions are issued by one elected lane of the CONVERGENT warp 0 (descriptors stay in uniform registers)
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(w_full, 3 * T::W_TILE);
#pragma unroll
    for (int kb = 0; kb < T::KB; ++kb) {
      tma_load_2d(sW + kb * T::W_KB, &tm_wo, w_full, kb * 64, 0);
      tma_load_2d(sW + T::W_TILE + kb * T::W_KB, &tm_w1, w_full, kb * 64, 0);
      tma_load_2d(sW + 2 * T::W_TILE + kb * T::W_KB, &tm_w2, w_full, kb * 64, 0);
    }
  }
  auto load_att = [&](int tile) {  // one elected lane
    mbar_arrive_expect_tx(att_full, T::A_TILE);
#pragma unroll
    for (int kb = 0; kb < T::KB; ++kb) tma_load_2d(sA + kb * T::A_ATOM, &tm_att, att_full, kb * 64, tile * TM);
  };
  auto load_x = [&](int tile) {
    mbar_arrive_expect_tx(x_full, T::X_TILE);
#pragma unroll
    for (int j = 0; j < T::XB; ++j) tma_load_2d(sX + j * (TM * 128), &tm_x, x_full, j * 32, tile * TM);
  };
  // When the block output is not written (fused outc), sX is never an output staging buffer: the next tile's x is
  // fetched as soon as this tile's x has been read, and its att as soon as the last GEMM has released sA.
  const bool prefetch = DIRECT || (OUTC && !p.write_out);
  uint32_t in_ph = 0, mma_ph = 0;
  bool first = true;
  if (prefetch && warp == 0 && (int)blockIdx.x < p.ntiles) {
    if (elect_one()) {
      load_att(blockIdx.x);
      load_x(blockIdx.x);
    }
    __syncwarp();
  }
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int m0 = tile * TM;
    const int next = tile + (int)gridDim.x;
    if (warp == 0) {
      if (!prefetch && elect_one()) {
        tma_store_wait_read();  // the previous tile's output (staged in sX) has left shared memory
        load_att(tile);
        load_x(tile);
      }
      if (first) mbar_wait_spin(w_full, 0);
      mbar_wait_spin(att_full, in_ph);
      tc_fence_after();
      if (elect_one()) {
        gemm_kc<C>(tmem_base, aA, aW, T::W_KB, 0, p.idesc);  // att Wo^T
        umma_commit(mma_done);
      }
      __syncwarp();
    }
    first = false;
    // ---- a = att Wo^T + bo + x ; LN(a) -> operand ----
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    mbar_wait(x_full, in_ph);  // the x tile is visible to this thread (acquire on the TMA barrier)
    in_ph ^= 1u;
    float a[C];
#pragma unroll
    for (int h = 0; h < C / 32; ++h) {
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
    ln_row_to_operand<C, DT>(a, sPar + C, sPar + 2 * C, aA, r);  // the GEMM that read sA has completed (mma_done)
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        gemm_kc<C>(tmem_base + C, aA, aW + T::W_TILE, T::W_KB, 0, p.idesc);  // LN(a) W1^T
        umma_commit(mma_done);
        if (prefetch && next < p.ntiles) load_x(next);  // every thread has read its x row (barrier above)
      }
      __syncwarp();
    }
    // ---- h = GELU(. + b1) -> operand ----
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < C / 32; ++h) {
      uint32_t v[32];
      tmem_ld32(t_row + C + h * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // 8 channels per 16-byte chunk
        const float4 b0 = *reinterpret_cast<const float4*>(sPar + 3 * C + h * 32 + c * 8);
        const float4 b1v = *reinterpret_cast<const float4*>(sPar + 3 * C + h * 32 + c * 8 + 4);
        float g0 = __uint_as_float(v[c * 8 + 0]) + b0.x, g1 = __uint_as_float(v[c * 8 + 1]) + b0.y;
        float g2 = __uint_as_float(v[c * 8 + 2]) + b0.z, g3 = __uint_as_float(v[c * 8 + 3]) + b0.w;
        float g4 = __uint_as_float(v[c * 8 + 4]) + b1v.x, g5 = __uint_as_float(v[c * 8 + 5]) + b1v.y;
        float g6 = __uint_as_float(v[c * 8 + 6]) + b1v.z, g7 = __uint_as_float(v[c * 8 + 7]) + b1v.w;
        gelu_erf2(g0, g1);
        gelu_erf2(g2, g3);
        gelu_erf2(g4, g5);
        gelu_erf2(g6, g7);
        sts128u(arow_chunk_addr(aA, r, h * 4 + c), pack2<DT>(g0, g1), pack2<DT>(g2, g3), pack2<DT>(g4, g5), pack2<DT>(g6, g7));
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        gemm_kc<C>(tmem_base, aA, aW + 2 * T::W_TILE, T::W_KB, 0, p.idesc);  // h W2^T
        umma_commit(mma_done);
      }
      __syncwarp();
    }
    // ---- out = . + b2 + a -> fp32 tile (over this thread's own x row) -> TMA store [and / or the fused outc] ----
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    if (prefetch && next < p.ntiles && warp == 0) {
      if (elect_one()) load_att(next);  // the last GEMM has completed: sA is free
      __syncwarp();
    }
    const bool stage = !DIRECT && (!OUTC || p.write_out);
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
    if constexpr (OUTC) {
      e0 = p.outc[4 * OUTC_C + 0];
      e1 = p.outc[4 * OUTC_C + 1];
      e2 = p.outc[4 * OUTC_C + 2];
      e3 = p.outc[4 * OUTC_C + 3];
    }
#pragma unroll
    for (int h = 0; h < C / 32; ++h) {
      uint32_t v[32];
      tmem_ld32(t_row + h * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 bv = *reinterpret_cast<const float4*>(sPar + 4 * C + h * 32 + c * 4);
        float4 o;
        o.x = __uint_as_float(v[c * 4 + 0]) + bv.x + a[h * 32 + c * 4 + 0];
        o.y = __uint_as_float(v[c * 4 + 1]) + bv.y + a[h * 32 + c * 4 + 1];
        o.z = __uint_as_float(v[c * 4 + 2]) + bv.z + a[h * 32 + c * 4 + 2];
        o.w = __uint_as_float(v[c * 4 + 3]) + bv.w + a[h * 32 + c * 4 + 3];
        if constexpr (OUTC) {
          const int ch = h * 32 + c * 4;
          const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            e0 = fmaf(ov[u], p.outc[0 * OUTC_C + ch + u], e0);
            e1 = fmaf(ov[u], p.outc[1 * OUTC_C + ch + u], e1);
            e2 = fmaf(ov[u], p.outc[2 * OUTC_C + ch + u], e2);
            e3 = fmaf(ov[u], p.outc[3 * OUTC_C + ch + u], e3);
          }
        }
        if (stage) sts128(xrow_chunk_addr(aX, r, h * 8 + c), o);
        if constexpr (DIRECT) {  // this warp's [32 rows x 128 B] scratch, 16-byte chunks XOR-swizzled by the row
          const uint32_t scr = smem_u32(sScr) + (uint32_t)warp * 4096u;
          sts128(scr + (uint32_t)(tid & 31) * 128u + ((((uint32_t)c) ^ ((uint32_t)tid & 7u)) << 4), o);
        }
      }
      if constexpr (DIRECT) {  // 32 channels of 32 tokens -> global, four complete 128-byte lines per store instruction
        const uint32_t scr = smem_u32(sScr) + (uint32_t)warp * 4096u;
        const int lane = tid & 31, rrow = lane >> 3, rchunk = lane & 7;
        __syncwarp();
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
          const int row = i2 * 4 + rrow;
          const float4 w4 = lds128(scr + (uint32_t)row * 128u + (((uint32_t)rchunk ^ ((uint32_t)row & 7u)) << 4));
          const int64_t m = (int64_t)m0 + warp * 32 + row;
          if (m < p.M) *reinterpret_cast<float4*>(p.out + m * C + h * 32 + rchunk * 4) = w4;
        }
        __syncwarp();
      }
    }
    if constexpr (OUTC) {
      const int64_t m = (int64_t)m0 + r;
      if (m < p.M) {  // consecutive threads = consecutive pixels of one sample: 128-byte stores per output channel
        const int64_t row = m >> p.log_hw;
        const int64_t pix = m - (row << p.log_hw);
        float* dst = p.eps + ((row * p.c_out) << p.log_hw) + pix;
        const int64_t hw = (int64_t)1 << p.log_hw;
        dst[0] = e0;
        if (p.c_out > 1) dst[hw] = e1;
        if (p.c_out > 2) dst[2 * hw] = e2;
        if (p.c_out > 3) dst[3 * hw] = e3;
      }
    }
    tc_fence_before();  // every thread's tcgen05.ld of this accumulator precedes the next tile's first GEMM
    if (stage) fence_proxy_async();
    __syncthreads();
    if (stage && warp == 0 && elect_one()) {
#pragma unroll
      for (int j = 0; j < T::XB; ++j) tma_store_2d(&tm_out, sX + j * (TM * 128), j * 32, m0);
      tma_store_commit();
    }
  }
  if (warp == 0 && elect_one()) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<T::TAIL_TMEM>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// sg_ln_inproj : qkv[M, 3C] = LayerNorm(x) Win^T + bin
// ------------------------------------------------------------------------------------------------------------------
struct InprojParams {
  const float* ln_g;
  const float* ln_b;
  const float* bias;  // [3C]
  uint16_t* qkv;      // direct-store epilogue (C = 128)
  int64_t M;
  int ntiles;
  uint32_t idesc_a, idesc_b;  // M128 x N(first MMA) and, at C = 128, M128 x N128 for rows [256, 384) of Win
};

template <int C, int DT>
__global__ void __launch_bounds__(128, Tok<C>::INPROJ_CTAS)
ln_inproj_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_qkv,
                 const InprojParams p) {
  using T = Tok<C>;
  constexpr int N = 3 * C;
  constexpr int NA = N <= 256 ? N : 256;  // columns of the first MMA (an MMA is at most 256 wide)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sW = smem;                // Win: C/64 k-blocks of [3C x 64] 16-bit, SWIZZLE_128B
  constexpr int NBUF = T::INPROJ_XBUF;
  uint8_t* sX = sW + T::WIN_TILE;    // x tile(s) (fp32); the consumed tile + sA stage the 3C/64 16-bit output boxes
  uint8_t* sA = sX + NBUF * T::X_TILE;  // LN(x) operand tile
  uint8_t* sScr = sA + T::A_TILE;    // C = 128: per-warp transposition scratch of the direct-store epilogue
  float* sPar = reinterpret_cast<float*>(sScr + T::INPROJ_SCR);  // ln_g | ln_b | bias[3C]
  constexpr bool DIRECT = T::INPROJ_DIRECT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 5 * C);
  uint64_t* w_full = bars;
  uint64_t* in_full = bars + 1;  // [NBUF]
  uint64_t* mma_done = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  static_assert(T::X_TILE + T::A_TILE == (N / 64) * T::A_ATOM, "output staging = x tile + operand tile");
  constexpr int XBOXES = T::X_TILE / T::A_ATOM;  // output boxes staged in the x tile; the rest go to sA

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    prefetch_tensormap(&tm_x);
    prefetch_tensormap(&tm_qkv);
    mbar_init(w_full, 1);
    mbar_init(&in_full[0], 1);
    mbar_init(&in_full[1], 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 5 * C; i += 128) sPar[i] = i < C ? p.ln_g[i] : (i < 2 * C ? p.ln_b[i - C] : p.bias[i - 2 * C]);
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<T::INPROJ_TMEM>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // every activation access (TMA loads included) follows this point; parameters read above are immutable weights
  pdl_launch_dependents();
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t aA = smem_u32(sA), aW = smem_u32(sW);
  const int r = tid;
  auto load_x = [&](int tile, int buf) {  // one elected lane
    mbar_arrive_expect_tx(&in_full[buf], T::X_TILE);
#pragma unroll
    for (int j = 0; j < T::XB; ++j) tma_load_2d(sX + buf * T::X_TILE + j * (TM * 128), &tm_x, &in_full[buf], j * 32, tile * TM);
  };

  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(w_full, T::WIN_TILE);
#pragma unroll
    for (int kb = 0; kb < T::KB; ++kb) {
      tma_load_2d(sW + kb * T::WIN_KB, &tm_w, w_full, kb * 64, 0);                         // rows [0, NA)
      if (N > NA) tma_load_2d(sW + kb * T::WIN_KB + NA * 128, &tm_w2, w_full, kb * 64, NA);  // rows [NA, N)
    }
  }
  uint32_t mma_ph = 0;
  bool first = true;
  if ((NBUF == 2 || DIRECT) && warp == 0 && elect_one() && (int)blockIdx.x < p.ntiles) load_x(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    const int m0 = tile * TM;
    const int buf = NBUF == 2 ? (it & 1) : 0;
    const uint32_t aX = smem_u32(sX + buf * T::X_TILE);
    if (!DIRECT && warp == 0 && elect_one()) {
      tma_store_wait_read();  // the previous tile's qkv boxes (staged in its x tile | sA) have left shared memory
      if (NBUF == 2) {
        // prefetch: the other buffer staged the previous tile's output and is free now; the load overlaps this tile
        if (tile + (int)gridDim.x < p.ntiles) load_x(tile + gridDim.x, buf ^ 1);
      } else {
        load_x(tile, 0);
      }
    }
    if (!DIRECT) __syncthreads();  // nobody writes sA (LN output) before the previous stores have drained
    mbar_wait(&in_full[buf], NBUF == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u));
    {
      float a[C];
#pragma unroll
      for (int c = 0; c < C / 4; ++c) {
        const float4 xv = lds128(xrow_chunk_addr(aX, r, c));
        a[c * 4 + 0] = xv.x;
        a[c * 4 + 1] = xv.y;
        a[c * 4 + 2] = xv.z;
        a[c * 4 + 3] = xv.w;
      }
      ln_row_to_operand<C, DT>(a, sPar, sPar + C, aA, r);
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (first) mbar_wait_spin(w_full, 0);
      tc_fence_after();
      if (elect_one()) {
        gemm_kc<C>(tmem_base, aA, aW, T::WIN_KB, 0, p.idesc_a);                   // columns [0, NA)
        if (N > NA) gemm_kc<C>(tmem_base + NA, aA, aW, T::WIN_KB, NA, p.idesc_b);  // columns [NA, N)
        umma_commit(mma_done);
        // DIRECT: every thread has read its x row (barrier above) and nothing is staged in sX: fetch the next tile now
        if (DIRECT && tile + (int)gridDim.x < p.ntiles) load_x(tile + gridDim.x, 0);
      }
      __syncwarp();
    }
    first = false;
    mbar_wait(mma_done, mma_ph);
    mma_ph ^= 1u;
    tc_fence_after();
    if constexpr (DIRECT) {
      // qkv rows -> global, 64 columns at a time through this warp's [32 rows x 128 B] scratch (16-byte chunks XOR-swizzled
      // by the row so that both the row-per-thread writes and the 4-rows-per-instruction reads are conflict free): every
      // global store instruction writes four complete 128-byte lines
      const uint32_t scr = smem_u32(sScr) + (uint32_t)warp * 4096u;
      const int lane = tid & 31;
      const int rrow = lane >> 3, rchunk = lane & 7;  // read side: 4 rows x 8 chunks per instruction
#pragma unroll 1
      for (int b = 0; b < N / 64; ++b) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld32(t_row + b * 64 + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 b0 = *reinterpret_cast<const float4*>(sPar + 2 * C + b * 64 + h * 32 + c * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sPar + 2 * C + b * 64 + h * 32 + c * 8 + 4);
            const uint32_t chunk = (uint32_t)(h * 4 + c);
            sts128u(scr + (uint32_t)lane * 128u + ((chunk ^ ((uint32_t)lane & 7u)) << 4),
                    pack2<DT>(__uint_as_float(v[c * 8 + 0]) + b0.x, __uint_as_float(v[c * 8 + 1]) + b0.y),
                    pack2<DT>(__uint_as_float(v[c * 8 + 2]) + b0.z, __uint_as_float(v[c * 8 + 3]) + b0.w),
                    pack2<DT>(__uint_as_float(v[c * 8 + 4]) + b1.x, __uint_as_float(v[c * 8 + 5]) + b1.y),
                    pack2<DT>(__uint_as_float(v[c * 8 + 6]) + b1.z, __uint_as_float(v[c * 8 + 7]) + b1.w));
          }
        }
        __syncwarp();
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
          const int row = i2 * 4 + rrow;
          uint4 w4;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w4.x), "=r"(w4.y), "=r"(w4.z), "=r"(w4.w)
                       : "r"(scr + (uint32_t)row * 128u + (((uint32_t)rchunk ^ ((uint32_t)row & 7u)) << 4)));
          const int64_t m = (int64_t)m0 + warp * 32 + row;
          if (m < p.M) *reinterpret_cast<uint4*>(p.qkv + m * N + b * 64 + rchunk * 8) = w4;
        }
        __syncwarp();
      }
      tc_fence_before();  // the accumulator reads above precede the next tile's MMAs
      __syncthreads();    // ... and every warp is done with sA / the accumulator before warp 0 moves on
    } else {
      // qkv rows -> 3C/64 [128 x 64] 16-bit SWIZZLE_128B boxes in the consumed x tile, the last ones in sA (its GEMM is complete)
  #pragma unroll 1
      for (int b = 0; b < N / 64; ++b) {
        const uint32_t box = b < XBOXES ? aX + (uint32_t)b * T::A_ATOM : aA + (uint32_t)(b - XBOXES) * T::A_ATOM;
  #pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld32(t_row + b * 64 + h * 32, v);
          tmem_ld_wait();
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 b0 = *reinterpret_cast<const float4*>(sPar + 2 * C + b * 64 + h * 32 + c * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(sPar + 2 * C + b * 64 + h * 32 + c * 8 + 4);
            sts128u(arow_chunk_addr(box, r, h * 4 + c),
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
  #pragma unroll
        for (int b = 0; b < N / 64; ++b)
          tma_store_2d(&tm_qkv, b < XBOXES ? sX + buf * T::X_TILE + b * T::A_ATOM : sA + (b - XBOXES) * T::A_ATOM, b * 64, m0);
        tma_store_commit();
      }
    }
  }
  if (warp == 0 && elect_one()) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<T::INPROJ_TMEM>(tmem_base);
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

template <int C, int DT, bool OUTC>
static int launch_tail(const void* att, const float* x, const void* wo, const void* w1, const void* w2, float* out,
                       TailParams p, int act_dtype, cudaStream_t s) {
  using T = Tok<C>;
  CUtensorMap tm_att, tm_x, tm_out, tm_wo, tm_w1, tm_w2;
  int rc;
  if ((rc = tmap2d(&tm_att, act_dtype, att, C, (uint64_t)p.M, 64, TM))) return rc;
  if ((rc = tmap2d(&tm_x, SG_F32, x, C, (uint64_t)p.M, 32, TM))) return rc;
  if ((rc = tmap2d(&tm_out, SG_F32, out ? out : x, C, (uint64_t)p.M, 32, TM))) return rc;  // unused when out is NULL
  if ((rc = tmap2d(&tm_wo, act_dtype, wo, C, C, 64, C))) return rc;
  if ((rc = tmap2d(&tm_w1, act_dtype, w1, C, C, 64, C))) return rc;
  if ((rc = tmap2d(&tm_w2, act_dtype, w2, C, C, 64, C))) return rc;
  p.idesc = make_idesc(act_dtype, 128, C, 0, 0);
  const int per_sm = T::TAIL_CTAS * num_sms();
  const int grid = p.ntiles < per_sm ? p.ntiles : per_sm;
  if ((rc = set_max_smem<attn_tail_kernel<C, DT, OUTC>>(T::TAIL_SMEM, "sg_attn_tail"))) return rc;
  launch_k(attn_tail_kernel<C, DT, OUTC>, dim3(grid), dim3(128), T::TAIL_SMEM, s, tm_att, tm_x, tm_out, tm_wo, tm_w1, tm_w2, p);
  return launch_status("sg_attn_tail");
}

template <int C, int DT>
static int launch_inproj(const float* x, const void* w_in, void* qkv, InprojParams p, int act_dtype, cudaStream_t s) {
  using T = Tok<C>;
  constexpr int N = 3 * C, NA = N <= 256 ? N : 256;
  CUtensorMap tm_x, tm_w, tm_w2, tm_qkv;
  int rc;
  if ((rc = tmap2d(&tm_x, SG_F32, x, C, (uint64_t)p.M, 32, TM))) return rc;
  if ((rc = tmap2d(&tm_w, act_dtype, w_in, C, N, 64, NA))) return rc;
  if ((rc = tmap2d(&tm_w2, act_dtype, w_in, C, N, 64, N > NA ? N - NA : NA))) return rc;
  if ((rc = tmap2d(&tm_qkv, act_dtype, qkv, N, (uint64_t)p.M, 64, TM))) return rc;
  p.idesc_a = make_idesc(act_dtype, 128, NA, 0, 0);
  p.idesc_b = make_idesc(act_dtype, 128, N > NA ? N - NA : NA, 0, 0);
  const int per_sm = T::INPROJ_CTAS * num_sms();
  const int grid = p.ntiles < per_sm ? p.ntiles : per_sm;
  if ((rc = set_max_smem<ln_inproj_kernel<C, DT>>(T::INPROJ_SMEM, "sg_ln_inproj"))) return rc;
  launch_k(ln_inproj_kernel<C, DT>, dim3(grid), dim3(128), T::INPROJ_SMEM, s, tm_x, tm_w, tm_w2, tm_qkv, p);
  return launch_status("sg_ln_inproj");
}

}  // namespace tc
}  // namespace sg

using namespace sg;
using namespace sg::tc;

extern "C" {

static int attn_tail_impl(const void* att, const float* x, const void* wo, const float* bo, const float* ln_g,
                          const float* ln_b, const void* w1, const float* b1, const void* w2, const float* b2, int64_t M,
                          int C, float* out, const float* outc_w, const float* outc_b, int c_out, int HW, float* eps,
                          int act_dtype, sg_stream_t stream, const char* what) {
  SG_REQUIRE(att && x && wo && bo && ln_g && ln_b && w1 && b1 && w2 && b2, "%s: null pointer", what);
  SG_REQUIRE(act_dtype == SG_BF16 || act_dtype == SG_F16, "%s: act_dtype must be SG_BF16 or SG_F16", what);
  SG_REQUIRE(M > 0 && M < (1ll << 31) - TM, "%s: M=%lld", what, (long long)M);
  TailParams p;
  p.bo = bo; p.ln_g = ln_g; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2;
  p.M = M;
  p.ntiles = (int)cdiv(M, TM);
  p.idesc = 0;
  p.out = out; p.eps = eps; p.c_out = c_out; p.log_hw = 0; p.write_out = out != nullptr;
  cudaStream_t s = as_stream(stream);
  if (eps) {
    SG_REQUIRE(C == 64, "%s: C=%d (the fused output conv follows the C = 64 block sa6)", what, C);
    SG_REQUIRE(outc_w && outc_b && c_out >= 1 && c_out <= 4, "%s: c_out=%d not in 1..4 or null weights", what, c_out);
    SG_REQUIRE(HW > 0 && (HW & (HW - 1)) == 0 && HW >= TM && M % HW == 0, "%s: HW=%d must be a power of two >= %d dividing M", what, HW, TM);
    while ((1 << p.log_hw) < HW) ++p.log_hw;
    for (int i = 0; i < 4 * OUTC_C + 4; ++i) p.outc[i] = 0.f;
    for (int k = 0; k < c_out; ++k) {
      for (int c = 0; c < C; ++c) p.outc[k * OUTC_C + c] = outc_w[k * C + c];  // HOST pointers (see sgb200.h)
      p.outc[4 * OUTC_C + k] = outc_b[k];
    }
    if (act_dtype == SG_BF16) return launch_tail<64, SG_BF16, true>(att, x, wo, w1, w2, out, p, act_dtype, s);
    return launch_tail<64, SG_F16, true>(att, x, wo, w1, w2, out, p, act_dtype, s);
  }
  SG_REQUIRE(out, "%s: null output", what);
  SG_REQUIRE(C == 64 || C == 128, "%s: C=%d (the fused kernel is built for C = 64 and 128; use the unfused launches otherwise)", what, C);
  if (C == 64) {
    if (act_dtype == SG_BF16) return launch_tail<64, SG_BF16, false>(att, x, wo, w1, w2, out, p, act_dtype, s);
    return launch_tail<64, SG_F16, false>(att, x, wo, w1, w2, out, p, act_dtype, s);
  }
  if (act_dtype == SG_BF16) return launch_tail<128, SG_BF16, false>(att, x, wo, w1, w2, out, p, act_dtype, s);
  return launch_tail<128, SG_F16, false>(att, x, wo, w1, w2, out, p, act_dtype, s);
}

int sg_attn_tail(const void* att, const float* x, const void* wo, const float* bo, const float* ln_g, const float* ln_b,
                 const void* w1, const float* b1, const void* w2, const float* b2, int64_t M, int C, float* out,
                 int act_dtype, sg_stream_t stream) {
  return attn_tail_impl(att, x, wo, bo, ln_g, ln_b, w1, b1, w2, b2, M, C, out, nullptr, nullptr, 0, 0, nullptr, act_dtype,
                        stream, "sg_attn_tail");
}

int sg_attn_tail_outc(const void* att, const float* x, const void* wo, const float* bo, const float* ln_g,
                      const float* ln_b, const void* w1, const float* b1, const void* w2, const float* b2, int64_t M,
                      int C, float* out, const float* outc_w, const float* outc_b, int c_out, int HW, float* eps,
                      int act_dtype, sg_stream_t stream) {
  SG_REQUIRE(eps, "sg_attn_tail_outc: null eps");
  return attn_tail_impl(att, x, wo, bo, ln_g, ln_b, w1, b1, w2, b2, M, C, out, outc_w, outc_b, c_out, HW, eps, act_dtype,
                        stream, "sg_attn_tail_outc");
}

int sg_ln_inproj(const float* x, const float* ln_g, const float* ln_b, const void* w_in, const float* b_in, int64_t M,
                 int C, void* qkv, int act_dtype, sg_stream_t stream) {
  SG_REQUIRE(x && ln_g && ln_b && w_in && b_in && qkv, "sg_ln_inproj: null pointer");
  SG_REQUIRE(C == 64 || C == 128, "sg_ln_inproj: C=%d (the fused kernel is built for C = 64 and 128; use the unfused launches otherwise)", C);
  SG_REQUIRE(act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_ln_inproj: act_dtype must be SG_BF16 or SG_F16");
  SG_REQUIRE(M > 0 && M < (1ll << 31) - TM, "sg_ln_inproj: M=%lld", (long long)M);
  InprojParams p;
  p.ln_g = ln_g; p.ln_b = ln_b; p.bias = b_in;
  p.qkv = reinterpret_cast<uint16_t*>(qkv);
  p.M = M;
  p.ntiles = (int)cdiv(M, TM);
  p.idesc_a = p.idesc_b = 0;
  cudaStream_t s = as_stream(stream);
  if (C == 64) {
    if (act_dtype == SG_BF16) return launch_inproj<64, SG_BF16>(x, w_in, qkv, p, act_dtype, s);
    return launch_inproj<64, SG_F16>(x, w_in, qkv, p, act_dtype, s);
  }
  if (act_dtype == SG_BF16) return launch_inproj<128, SG_BF16>(x, w_in, qkv, p, act_dtype, s);
  return launch_inproj<128, SG_F16>(x, w_in, qkv, p, act_dtype, s);
}

}  // extern "C"
