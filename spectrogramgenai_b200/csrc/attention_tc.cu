// K4 (tensor-core engine): flash-style multi-head self-attention on tcgen05 / TMEM, fed by TMA.
//
// Replaces the scaled-dot-product inside nn.MultiheadAttention (/root/reference/src/diff_modules.py:56,:69)
// without ever materialising the L x L matrix (the reference materialises it AND head-averages it, then
// throws the average away).
//
// Work decomposition: the tokens of all batch rows form one flat sequence of M = rows*L tokens (NHWC makes
// the reference's view/swapaxes free).  One CTA = one tile of 128 consecutive query tokens x one head.
//   * L >= 128: the tile lies inside one batch row; it loops over that row's L/128 key tiles.
//   * L <  128: the tile holds 128/L whole batch rows; a single key tile (the same tokens) is used with a
//     block-diagonal mask.
// Per key tile:   S = Q K^T          tcgen05.mma  M=128 (queries) x N=128 (keys) x K=d      -> TMEM
//                 P = exp2(S*c - m*c) softmax warps: one query row per thread, two passes over TMEM
//                                     (row max, then exp / row sum / 16-bit pack) -> swizzled smem
//                 O_j = P V           tcgen05.mma  M=128 x N=d x K=128 keys, V consumed MN-major -> TMEM
//                 o = o*alpha + O_j   running output, max and sum stay in registers (fp32)
// TMEM use is 128 columns (O_j aliases the dead S columns), so up to four CTAs share an SM and overlap each
// other's MMA / MUFU / TMA phases; inside a CTA the phases are serial.
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = softmax + epilogue.
#include <stdlib.h>

#include "tc_common.cuh"

namespace sg {
namespace tc {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int P_BYTES = ATT_BM * ATT_BN * 2;  // 32 KB: two SWIZZLE_128B atoms of 64 keys

// generic K-/MN-major descriptor for tiles whose rows are one swizzle span of `row_bytes` (32 / 64 / 128)
__device__ __forceinline__ uint64_t make_desc_rows(uint32_t saddr, int row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                          // LBO: single atom along the other dimension -> unused
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;     // SBO: 8 rows
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// two fp32 -> one packed 16-bit pair (lo = a), format fixed at compile time: one F2FP instruction
template <int DT>
__device__ __forceinline__ uint32_t pack_pair(float a, float b) {
  uint32_t w;
  if constexpr (DT == SG_BF16) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  return w;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max rel. error 7.5e-5 -- far below
// the 2^-9 / 2^-12 rounding P receives anyway).  Used for a fraction of the elements so that the MUFU pipe
// (16 ex2 / clk / SM), which bounds d = 16 attention, is not the only unit producing probabilities.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517145f, 0.24261084f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.99992812f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

struct AttGeom {
  int64_t M;          // rows * L tokens
  int L, logL, C, heads;
  int nkv;            // key tiles per query tile
  float c;            // softmax scale * log2(e)
  uint32_t tile_bytes;  // bytes of one TMA box (d*2 * min(128, M))
  uint32_t idesc_s, idesc_o, idesc_1;
  int act_dtype;
  float redo_log2;  // largest tolerated (tile max - reference max) * c before the tile is recomputed
};

constexpr int ONES_BYTES = 2048;  // a [16 x 64] 16-bit K-major tile of 1.0: B operand of the row-sum MMA
template <int D>
constexpr int att_smem_bytes() {
  return 1024 + ATT_BM * D * 2 /*Q*/ + 2 * 2 * ATT_BN * D * 2 /*K,V x 2 stages*/ + P_BYTES + ONES_BYTES + 256;
}

template <int D, int DT, int POLY>
__global__ void __launch_bounds__(192, (D == 64 ? 2 : (D == 32 ? 3 : 4)))
attention_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttGeom g, uint16_t* __restrict__ out) {
  constexpr int ROWB = D * 2;              // bytes per token row of a head slice = the swizzle span
  constexpr int TILE = ATT_BN * ROWB;      // one Q / K / V tile
  constexpr int NACC = D == 64 ? 2 : 4;    // independent P V accumulators (NACC * D <= 128 TMEM columns)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE;        // [2 stages]
  uint8_t* sV = sK + 2 * TILE;    // [2 stages]
  uint8_t* sP = sV + 2 * TILE;    // 1024-aligned: TILE is a multiple of 4096
  uint8_t* sOnes = sP + P_BYTES;  // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_read = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * ATT_BM;
  const int64_t kv0 = g.L >= ATT_BN ? (m0 >> g.logL) << g.logL : m0;  // first key token of this tile's row(s)

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);  // one elected lane per softmax warp
    mbar_init(o_full, 1);
    mbar_init(o_read, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  {
    const uint32_t one2 = DT == SG_BF16 ? 0x3F803F80u : 0x3C003C00u;  // two 1.0 values
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = one2;
    fence_proxy_async();
  }
  if (g.M < ATT_BM) {
    // tiny problems: the TMA box is clamped to M rows, so clear the tiles once (0 * stale-NaN would poison P V)
    for (int i = threadIdx.x; i < 5 * TILE / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: Q once, then the K / V ring =====
      mbar_arrive_expect_tx(q_full, g.tile_bytes);
      tma_load_2d(sQ, &tm, q_full, head * D, (int)m0);
      for (int j = 0; j < g.nkv; ++j) {
        const int s = j & 1;
        mbar_wait_spin(&kv_empty[s], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[s], 2 * g.tile_bytes);
        const int tok = (int)(kv0 + (int64_t)j * ATT_BN);
        tma_load_2d(sK + s * TILE, &tm, &kv_full[s], g.C + head * D, tok);
        tma_load_2d(sV + s * TILE, &tm, &kv_full[s], 2 * g.C + head * D, tok);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      mbar_wait_spin(q_full, 0);
      for (int j = 0; j < g.nkv; ++j) {
        const int s = j & 1;
        mbar_wait_spin(&kv_full[s], (uint32_t)(j >> 1) & 1u);
        if (j > 0) mbar_wait_spin(o_read, (uint32_t)(j - 1) & 1u);  // O_{j-1} (aliasing S) has been consumed
        tc_fence_after();
        // S = Q K^T : K-major A and B, K = D in steps of 16 (32 bytes inside the swizzle span)
        const uint64_t qd = make_desc_rows(smem_u32(sQ), ROWB);
        const uint64_t kd = make_desc_rows(smem_u32(sK + s * TILE), ROWB);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_ss(tmem_base, qd + 2 * k, kd + 2 * k, g.idesc_s, k != 0);
        umma_commit(s_full);
        // O_j = P V : A = P (K-major, two 64-key SWIZZLE_128B atoms), B = V consumed MN-major (d contiguous)
        mbar_wait_spin(p_ready, (uint32_t)j & 1u);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP);
        const uint32_t va = smem_u32(sV + s * TILE);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          const uint64_t pd = make_desc_k128(pa + (k >> 2) * (ATT_BM * 128) + (k & 3) * 32);
          const uint64_t vd = make_desc_rows(va + k * 16 * ROWB, ROWB);
          // NACC independent accumulators (columns [a*D, (a+1)*D) of the dead S region): the eight N = D MMAs of
          // one key tile would otherwise form a dependent, latency-bound chain on a single accumulator
          umma_ss(tmem_base + (uint32_t)((k % NACC) * D), pd, vd, g.idesc_o, k >= NACC);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(o_full);
      }
    }
  } else {
    // ===== softmax + epilogue: thread = one query row (TMEM lane quadrant = warp % 4) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int64_t tok = m0 + r;
    const bool row_valid = tok < g.M;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool masked = g.L < ATT_BN;  // block-diagonal tile (several batch rows) and/or ragged tail
    const int64_t my_row = tok >> g.logL;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m_ref = -INFINITY, l = 0.f;  // reference maximum of the exponent; running row sum (relative to m_ref)
    const uint32_t p_row = smem_u32(sP) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    for (int j = 0; j < g.nkv; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int64_t key0 = kv0 + (int64_t)j * ATT_BN;
      // ---- ONE sweep over S: p = exp2(s*c - m_ref*c) against the maximum known BEFORE this tile, and the tile
      // maximum as a by-product.  Exact algebra (o, l are rescaled afterwards); the sweep is only repeated when the
      // tile maximum exceeds the reference by more than redo_log2 (overflow guard; always on the first tile).
      float tmax, psum;
      bool redo;
      do {
        const float mc = m_ref * g.c;
        tmax = -INFINITY;
        psum = 0.f;
#pragma unroll 1
        for (int cch = 0; cch < 4; ++cch) {
          uint32_t v[32];
          tmem_ld32(t_row + cch * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
            bool k0 = true, k1 = true;
            if (masked) {
              const int64_t kt = key0 + cch * 32 + i;
              k0 = !(row_valid && (kt >= g.M || (kt >> g.logL) != my_row));
              k1 = !(row_valid && (kt + 1 >= g.M || ((kt + 1) >> g.logL) != my_row));
              if (!k0) s0 = -INFINITY;
              if (!k1) s1 = -INFINITY;
            }
            tmax = fmaxf(tmax, fmaxf(s0, s1));
            float p0, p1;
            if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
              p0 = ex2_poly(fmaf(s0, g.c, -mc));
              p1 = ex2_poly(fmaf(s1, g.c, -mc));
            } else {
              p0 = ex2(fmaf(s0, g.c, -mc));
              p1 = ex2(fmaf(s1, g.c, -mc));
            }
            if (masked) {
              if (!k0) p0 = 0.f;
              if (!k1) p1 = 0.f;
            }
            pk[i >> 1] = pack_pair<DT>(p0, p1);
            psum += p0 + p1;
          }
          // 32 keys = 64 bytes = four 16-byte chunks jj = cch*4 .. cch*4+3 of this row's 256-byte P row
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int jj = cch * 4 + u;
            const uint32_t addr = p_row + (uint32_t)(jj >> 3) * (ATT_BM * 128) + (uint32_t)(((jj & 7) ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[u * 4]), "r"(pk[u * 4 + 1]),
                         "r"(pk[u * 4 + 2]), "r"(pk[u * 4 + 3])
                         : "memory");
          }
        }
        const bool over = (tmax - m_ref) * g.c > g.redo_log2;  // also true while m_ref == -inf
        redo = __any_sync(0xffffffffu, over);
        if (over) {
          const float a0 = ex2((m_ref - tmax) * g.c);  // 0 on the first tile
          l *= a0;
#pragma unroll
          for (int i = 0; i < D; ++i) o[i] *= a0;
          m_ref = tmax;
        }
      } while (redo);
      tc_fence_before();    // our tcgen05.ld of S precede the MMA that overwrites those columns
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // ---- O_j and the row sum, both relative to m_ref ----
      mbar_wait(o_full, (uint32_t)j & 1u);
      tc_fence_after();
#pragma unroll
      for (int cch = 0; cch < NACC * D / 32; ++cch) {
        uint32_t v[32];
        tmem_ld32(t_row + cch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[(cch * 32 + i) % D] += __uint_as_float(v[i]);
      }
      l += psum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_read);
      // ---- raise the reference to the new running maximum (exact rescale of o, l) ----
      if (tmax > m_ref) {
        const float a1 = ex2((m_ref - tmax) * g.c);
        l *= a1;
#pragma unroll
        for (int i = 0; i < D; ++i) o[i] *= a1;
        m_ref = tmax;
      }
    }
    if (row_valid) {
      const float inv = 1.0f / l;
      uint16_t* dst = out + tok * g.C + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) {
        uint4 w;
        w.x = pack_pair<DT>(o[i] * inv, o[i + 1] * inv);
        w.y = pack_pair<DT>(o[i + 2] * inv, o[i + 3] * inv);
        w.z = pack_pair<DT>(o[i + 4] * inv, o[i + 5] * inv);
        w.w = pack_pair<DT>(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<128>(tmem_base);
  }
}

template <int D, int DT, int POLY>
static int launch_att(const CUtensorMap& tm, const AttGeom& g, uint16_t* out, dim3 grid, cudaStream_t stream) {
  constexpr int smem = att_smem_bytes<D>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<D, DT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("sg_attention(tc): cudaFuncSetAttribute(%d B smem): %s", smem, cudaGetErrorString(e));
      return SG_ERR_LAUNCH;
    }
    configured = true;
  }
  attention_tc_kernel<D, DT, POLY><<<grid, 192, smem, stream>>>(tm, g, out);
  return launch_status("sg_attention(tc)");
}

// =====================================================================================================
// v2 (L >= 256): one CTA = TWO tiles of 128 queries (A, B) of one head against the same K/V stream.
//   * two softmax warp groups (4 warps each) ping-pong: while group A runs exp on S_A(j) the tensor core
//     computes S_B(j) / P_A V / S_A(j+1), so the MUFU pipe always has a second warp per scheduler to issue from;
//   * O accumulates in TMEM across key tiles (tcgen05.mma accumulate); the softmax reference maximum is only
//     raised -- and O rescaled through tcgen05.ld/st -- when the tile maximum exceeds it by more than 2^8
//     ("lazy rescale"), so a key tile costs ONE mbarrier wait per thread and no per-tile O round trip;
//   * TMEM loads are register double-buffered (the load of chunk c+1 is in flight while chunk c is processed).
// TMEM: S_A [0,128) S_B [128,256) O_A [256,256+D) O_B [320,320+D) -> 512 columns, one CTA per SM.
// Warp roles (320 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = softmax A, 6..9 = softmax B.
// =====================================================================================================
template <int D>
constexpr int att2_smem_bytes() {
  return 1024 + 2 * ATT_BM * D * 2 /*Q_A,Q_B*/ + 3 * 2 * ATT_BN * D * 2 /*K,V x 3 stages*/ + 2 * P_BYTES + 256;
}

template <int D, int DT>
__global__ void __launch_bounds__(320, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tm, const AttGeom g, uint16_t* __restrict__ out) {
  constexpr int ROWB = D * 2;
  constexpr int TILE = ATT_BN * ROWB;
  constexpr int KVS = 3;
  constexpr float RESCALE_LOG2 = 8.0f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                  // [2][TILE]
  uint8_t* sKV = sQ + 2 * TILE;        // [KVS][K | V][TILE]
  uint8_t* sP = sKV + KVS * 2 * TILE;  // [2][P_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [3]
  uint64_t* kv_empty = bars + 4;  // [3]
  uint64_t* s_full = bars + 7;    // [2]
  uint64_t* p_ready = bars + 9;   // [2]
  uint64_t* o_done = bars + 11;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * (2 * ATT_BM);
  const int64_t kv0 = (m0 >> g.logL) << g.logL;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tm);
    mbar_init(q_full, 1);
    for (int s = 0; s < KVS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&p_ready[x], 128);
      mbar_init(&o_done[x], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_arrive_expect_tx(q_full, 2 * g.tile_bytes);
      tma_load_2d(sQ, &tm, q_full, head * D, (int)m0);
      tma_load_2d(sQ + TILE, &tm, q_full, head * D, (int)m0 + ATT_BM);
      for (int j = 0; j < g.nkv; ++j) {
        const int s = j % KVS;
        mbar_wait_spin(&kv_empty[s], ((uint32_t)(j / KVS) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[s], 2 * g.tile_bytes);
        const int tok = (int)(kv0 + (int64_t)j * ATT_BN);
        tma_load_2d(sKV + (2 * s) * TILE, &tm, &kv_full[s], g.C + head * D, tok);
        tma_load_2d(sKV + (2 * s + 1) * TILE, &tm, &kv_full[s], 2 * g.C + head * D, tok);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      auto issue_s = [&](int x, int s) {  // S_x = Q_x K^T
        const uint64_t qd = make_desc_rows(smem_u32(sQ + x * TILE), ROWB);
        const uint64_t kd = make_desc_rows(smem_u32(sKV + (2 * s) * TILE), ROWB);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_ss(tmem_base + x * 128, qd + 2 * k, kd + 2 * k, g.idesc_s, k != 0);
        umma_commit(&s_full[x]);
      };
      mbar_wait_spin(q_full, 0);
      mbar_wait_spin(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < g.nkv; ++j) {
        const int s = j % KVS;
        for (int x = 0; x < 2; ++x) {
          mbar_wait_spin(&p_ready[x], (uint32_t)j & 1u);
          tc_fence_after();
          // O_x (+)= P_x V : A = P (K-major, two 64-key atoms), B = V consumed MN-major
          const uint32_t pa = smem_u32(sP + x * P_BYTES);
          const uint32_t va = smem_u32(sKV + (2 * s + 1) * TILE);
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k) {
            const uint64_t pd = make_desc_k128(pa + (k >> 2) * (ATT_BM * 128) + (k & 3) * 32);
            const uint64_t vd = make_desc_rows(va + k * 16 * ROWB, ROWB);
            umma_ss(tmem_base + 256 + x * 64, pd, vd, g.idesc_o, (j | k) != 0);
          }
          umma_commit(&o_done[x]);
          if (x == 1) umma_commit(&kv_empty[s]);
          if (j + 1 < g.nkv) {
            const int s1 = (j + 1) % KVS;
            if (x == 0) {
              mbar_wait_spin(&kv_full[s1], (uint32_t)((j + 1) / KVS) & 1u);
              tc_fence_after();
            }
            issue_s(x, s1);
          }
        }
      }
    }
  } else {
    // ===== softmax groups =====
    const int x = (warp - 2) >> 2;  // 0 = tile A, 1 = tile B
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int64_t tok = m0 + x * ATT_BM + r;
    const uint32_t t_s = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * 128);
    const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + x * 64);
    const uint32_t p_row = smem_u32(sP + x * P_BYTES) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < g.nkv; ++j) {
      mbar_wait(&s_full[x], (uint32_t)j & 1u);
      tc_fence_after();
      // ---- pass 1: row max (register double-buffered TMEM loads) ----
      uint32_t va[32], vb[32];
      float tmax = -INFINITY;
      tmem_ld32(t_s, va);
      tmem_ld_wait();
      tmem_ld32(t_s + 32, vb);
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(va[i]));
      tmem_ld_wait();
      tmem_ld32(t_s + 64, va);
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(vb[i]));
      tmem_ld_wait();
      tmem_ld32(t_s + 96, vb);
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(va[i]));
      tmem_ld_wait();
      tmem_ld32(t_s, va);  // first chunk of pass 2 already in flight
#pragma unroll
      for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(vb[i]));
      // ---- reference maximum: raise it (and rescale O, l) only when exceeded by more than 2^8 ----
      if (j == 0) {
        m_used = tmax;
      } else {
        const bool need = (tmax - m_used) * g.c > RESCALE_LOG2;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2((m_used - tmax) * g.c) : 1.0f;
          if (need) m_used = tmax;
          l *= alpha;
          tmem_ld_wait();  // the in-flight S load must land before this warp issues other tcgen05.ld
          // PV(j-1) has completed (its commit precedes S(j)'s in the issuer's stream), so O is quiescent
#pragma unroll
          for (int cch = 0; cch < D / 16; ++cch) {
            uint32_t ov[16];
            tmem_ld16(t_o + cch * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st16(t_o + cch * 16, ov);
          }
          tmem_st_wait();
        }
      }
      const float mc = m_used * g.c;
      // ---- pass 2: p = exp2(s*c - m*c), row sum, 16-bit pack, swizzled st.shared ----
      float psum = 0.f;
      auto emit = [&](const uint32_t (&v)[32], int cch) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(v[i]), g.c, -mc));
          const float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), g.c, -mc));
          pk[i >> 1] = pack_pair<DT>(p0, p1);
          psum += p0 + p1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = cch * 4 + u;
          const uint32_t addr = p_row + (uint32_t)(jj >> 3) * (ATT_BM * 128) + (uint32_t)(((jj & 7) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[u * 4]), "r"(pk[u * 4 + 1]),
                       "r"(pk[u * 4 + 2]), "r"(pk[u * 4 + 3])
                       : "memory");
        }
      };
      tmem_ld_wait();
      tmem_ld32(t_s + 32, vb);
      emit(va, 0);
      tmem_ld_wait();
      tmem_ld32(t_s + 64, va);
      emit(vb, 1);
      tmem_ld_wait();
      tmem_ld32(t_s + 96, vb);
      emit(va, 2);
      tmem_ld_wait();
      emit(vb, 3);
      l += psum;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&p_ready[x]);
    }
    // ---- epilogue: O / l ----
    mbar_wait(&o_done[x], (uint32_t)(g.nkv - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint16_t* dst = out + tok * g.C + head * D;
#pragma unroll
    for (int cch = 0; cch < D / 16; ++cch) {
      uint32_t ov[16];
      tmem_ld16(t_o + cch * 16, ov);
      tmem_ld_wait();
      uint4 w0, w1;
      w0.x = pack_pair<DT>(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
      w0.y = pack_pair<DT>(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
      w0.z = pack_pair<DT>(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
      w0.w = pack_pair<DT>(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
      w1.x = pack_pair<DT>(__uint_as_float(ov[8]) * inv, __uint_as_float(ov[9]) * inv);
      w1.y = pack_pair<DT>(__uint_as_float(ov[10]) * inv, __uint_as_float(ov[11]) * inv);
      w1.z = pack_pair<DT>(__uint_as_float(ov[12]) * inv, __uint_as_float(ov[13]) * inv);
      w1.w = pack_pair<DT>(__uint_as_float(ov[14]) * inv, __uint_as_float(ov[15]) * inv);
      *reinterpret_cast<uint4*>(dst + cch * 16) = w0;
      *reinterpret_cast<uint4*>(dst + cch * 16 + 8) = w1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int D, int DT>
static int launch_att2(const CUtensorMap& tm, const AttGeom& g, uint16_t* out, dim3 grid, cudaStream_t stream) {
  constexpr int smem = att2_smem_bytes<D>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc2_kernel<D, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("sg_attention(tc2): cudaFuncSetAttribute(%d B smem): %s", smem, cudaGetErrorString(e));
      return SG_ERR_LAUNCH;
    }
    configured = true;
  }
  attention_tc2_kernel<D, DT><<<grid, 320, smem, stream>>>(tm, g, out);
  return launch_status("sg_attention(tc2)");
}


// =====================================================================================================
// v3 (L >= 256): the occupancy of v1 (128 TMEM columns, four CTAs per SM) with the short hand-off chain of v2.
// The key tile is BN = 128 - D keys, so S (BN columns) and the running output O (D columns) live side by side in
// the CTA's 128 TMEM columns:
//   * O accumulates in TMEM across key tiles (no per-tile O round trip); the exponent reference m_ref is only
//     raised -- with O rescaled through tcgen05.ld/st and the sweep repeated -- when a tile maximum exceeds it by
//     more than redo_log2 (always on the first tile, rare afterwards);
//   * the issuer launches S(j+1) = Q K(j+1)^T the moment the softmax warps have released S(j), BEFORE P(j) V(j),
//     so the next sweep never waits for the P V phase: one mbarrier wait per key tile per thread;
//   * one sweep over S per tile: p = exp2(s*c - m_ref*c), row sum, tile max, 16-bit pack, swizzled st.shared.
// The last key tile of a row may be ragged (L is not a multiple of BN): keys >= L get p = 0.
// =====================================================================================================
template <int D>
struct Att3 {
  static constexpr int BN = 128 - D;               // keys per tile: 112 / 96 / 64
  static constexpr int ROWB = D * 2;
  static constexpr int QT = ATT_BM * ROWB;         // Q tile bytes
  static constexpr int KVT = ((BN * ROWB + 1023) / 1024) * 1024;  // K / V stage bytes (1024-aligned)
  static constexpr int SMEM = 1024 + QT + 4 * KVT + P_BYTES + 256;
};

template <int D, int DT>
__global__ void __launch_bounds__(192, (D == 64 ? 2 : (D == 32 ? 3 : 4)))
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const AttGeom g, uint16_t* __restrict__ out) {
  using A = Att3<D>;
  constexpr int BN = A::BN, ROWB = A::ROWB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + A::QT;          // [2 stages]
  uint8_t* sV = sK + 2 * A::KVT;     // [2 stages]
  uint8_t* sP = sV + 2 * A::KVT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * ATT_BM;
  const int64_t kv0 = (m0 >> g.logL) << g.logL;
  const int nkv = (g.L + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_arrive_expect_tx(q_full, (uint32_t)A::QT);
      tma_load_2d(sQ, &tmQ, q_full, head * D, (int)m0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait_spin(&kv_empty[s], ((uint32_t)(j >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[s], 2u * BN * ROWB);
        const int tok = (int)(kv0 + (int64_t)j * BN);
        tma_load_2d(sK + s * A::KVT, &tmKV, &kv_full[s], g.C + head * D, tok);
        tma_load_2d(sV + s * A::KVT, &tmKV, &kv_full[s], 2 * g.C + head * D, tok);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      auto issue_s = [&](int s) {
        const uint64_t qd = make_desc_rows(smem_u32(sQ), ROWB);
        const uint64_t kd = make_desc_rows(smem_u32(sK + s * A::KVT), ROWB);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_ss(tmem_base, qd + 2 * k, kd + 2 * k, g.idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait_spin(q_full, 0);
      mbar_wait_spin(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait_spin(p_ready, (uint32_t)j & 1u);  // S(j) released, P(j) written
        tc_fence_after();
        if (j + 1 < nkv) {
          mbar_wait_spin(&kv_full[s ^ 1], (uint32_t)((j + 1) >> 1) & 1u);
          tc_fence_after();
          issue_s(s ^ 1);  // next scores first: the softmax warps never wait for the P V phase
        }
        const uint32_t pa = smem_u32(sP);
        const uint32_t va = smem_u32(sV + s * A::KVT);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint64_t pd = make_desc_k128(pa + (k >> 2) * (ATT_BM * 128) + (k & 3) * 32);
          const uint64_t vd = make_desc_rows(va + k * 16 * ROWB, ROWB);
          umma_ss(tmem_base + BN, pd, vd, g.idesc_o, (j | k) != 0);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(o_done);
      }
    }
  } else {
    // ===== softmax + epilogue =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int64_t tok = m0 + r;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_o = t_row + BN;
    const uint32_t p_row = smem_u32(sP) + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int nvalid = g.L - j * BN;  // keys of this tile that belong to the row (>= BN except on the last tile)
      bool waited_o = (j == 0);
      float tmax, psum;
      bool redo;
      do {
        const float mc = m_ref * g.c;
        tmax = -INFINITY;
        psum = 0.f;
        // one chunk = NC columns starting at c0: exp, sum, max, pack, store
        auto chunk = [&](const uint32_t* v, int c0, int nc) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            if (i < nc) {
              float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
              if (c0 + i >= nvalid) s0 = -INFINITY;
              if (c0 + i + 1 >= nvalid) s1 = -INFINITY;
              tmax = fmaxf(tmax, fmaxf(s0, s1));
              const float p0 = ex2(fmaf(s0, g.c, -mc)), p1 = ex2(fmaf(s1, g.c, -mc));
              pk[i >> 1] = pack_pair<DT>(p0, p1);
              psum += p0 + p1;
            }
          }
          if (!waited_o) {  // P(j-1) V(j-1) must have finished reading the P buffer (and O is quiescent)
            mbar_wait(o_done, (uint32_t)(j - 1) & 1u);
            tc_fence_after();
            waited_o = true;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (u * 8 < nc) {
              const int jj = (c0 >> 3) + u;
              const uint32_t addr = p_row + (uint32_t)(jj >> 3) * (ATT_BM * 128) + (uint32_t)(((jj & 7) ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[u * 4]), "r"(pk[u * 4 + 1]),
                           "r"(pk[u * 4 + 2]), "r"(pk[u * 4 + 3])
                           : "memory");
            }
          }
        };
#pragma unroll 1
        for (int c0 = 0; c0 + 32 <= BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          chunk(v, c0, 32);
        }
        if constexpr (BN % 32 != 0) {
          uint32_t v16[16];
          tmem_ld16(t_row + (BN / 32) * 32, v16);
          tmem_ld_wait();
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v16[i];
          chunk(v, (BN / 32) * 32, 16);
        }
        const bool over = (tmax - m_ref) * g.c > g.redo_log2;  // also true while m_ref == -inf
        redo = __any_sync(0xffffffffu, over);
        if (redo) {
          const float a0 = over ? ex2((m_ref - tmax) * g.c) : 1.0f;  // 0 on the first tile
          if (over) m_ref = tmax;
          l *= a0;
          if (j > 0) {
#pragma unroll
            for (int cch = 0; cch < D / 16; ++cch) {
              uint32_t ov[16];
              tmem_ld16(t_o + cch * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * a0);
              tmem_st16(t_o + cch * 16, ov);
            }
            tmem_st_wait();
          }
        }
      } while (redo);
      l += psum;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l ----
    mbar_wait(o_done, (uint32_t)(nkv - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint16_t* dst = out + tok * g.C + head * D;
#pragma unroll
    for (int cch = 0; cch < D / 16; ++cch) {
      uint32_t ov[16];
      tmem_ld16(t_o + cch * 16, ov);
      tmem_ld_wait();
      uint4 w0, w1;
      w0.x = pack_pair<DT>(__uint_as_float(ov[0]) * inv, __uint_as_float(ov[1]) * inv);
      w0.y = pack_pair<DT>(__uint_as_float(ov[2]) * inv, __uint_as_float(ov[3]) * inv);
      w0.z = pack_pair<DT>(__uint_as_float(ov[4]) * inv, __uint_as_float(ov[5]) * inv);
      w0.w = pack_pair<DT>(__uint_as_float(ov[6]) * inv, __uint_as_float(ov[7]) * inv);
      w1.x = pack_pair<DT>(__uint_as_float(ov[8]) * inv, __uint_as_float(ov[9]) * inv);
      w1.y = pack_pair<DT>(__uint_as_float(ov[10]) * inv, __uint_as_float(ov[11]) * inv);
      w1.z = pack_pair<DT>(__uint_as_float(ov[12]) * inv, __uint_as_float(ov[13]) * inv);
      w1.w = pack_pair<DT>(__uint_as_float(ov[14]) * inv, __uint_as_float(ov[15]) * inv);
      *reinterpret_cast<uint4*>(dst + cch * 16) = w0;
      *reinterpret_cast<uint4*>(dst + cch * 16 + 8) = w1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<128>(tmem_base);
  }
}

template <int D, int DT>
static int launch_att3(const void* qkv, const AttGeom& g0, int act_dtype, uint16_t* out, cudaStream_t stream) {
  using A = Att3<D>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc3_kernel<D, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, A::SMEM);
    if (e != cudaSuccess) {
      set_error("sg_attention(tc3): cudaFuncSetAttribute(%d B smem): %s", A::SMEM, cudaGetErrorString(e));
      return SG_ERR_LAUNCH;
    }
    configured = true;
  }
  AttGeom g = g0;
  g.idesc_s = make_idesc(act_dtype, 128, A::BN, 0, 0);
  CUtensorMap tmQ, tmKV;
  const uint64_t dims[2] = {(uint64_t)3 * g.C, (uint64_t)g.M};
  const uint64_t strides[1] = {(uint64_t)3 * g.C * 2};
  const uint32_t boxq[2] = {(uint32_t)D, 128u};
  const uint32_t boxkv[2] = {(uint32_t)D, (uint32_t)A::BN};
  const CUtensorMapSwizzle sw = D == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (D == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  int rc = make_tmap(&tmQ, act_dtype, 2, qkv, dims, strides, boxq, sw);
  if (rc) return rc;
  rc = make_tmap(&tmKV, act_dtype, 2, qkv, dims, strides, boxkv, sw);
  if (rc) return rc;
  dim3 grid((unsigned)(g.M / ATT_BM), (unsigned)g.heads);
  attention_tc3_kernel<D, DT><<<grid, 192, A::SMEM, stream>>>(tmQ, tmKV, g, out);
  return launch_status("sg_attention(tc3)");
}

// fraction of exponentials evaluated by the FMA-pipe polynomial: SGB200_ATTN_POLY = 0 (none), 4 (1/4), 2 (1/2)
static int attention_poly() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGB200_ATTN_POLY");
    v = e ? atoi(e) : 0;  // measured: no gain on B200 (the kernel is not MUFU-throughput bound), so off by default
    if (v != 0 && v != 2 && v != 4) v = 0;
  }
  return v;
}

static int attention_version() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("SGB200_ATTN");
    // 1 = v1 (always used for L < 256), 2 = two-tile ping-pong, 3 = TMEM-resident O, 0 = auto (default):
    // measured on B200 (rows=128): d=16 L=4096: v1 2.93 ms, v2 3.57, v3 3.12;  d=32 L=1024: v1 0.315, v3 0.267
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 3) v = 0;
  }
  return v;
}

}  // namespace tc

int attention_tc(
const void* qkv, void* out, int rows, int L, int C, int heads, int act_dtype, cudaStream_t stream) {
  using namespace tc;
  SG_REQUIRE(act_dtype == SG_BF16 || act_dtype == SG_F16, "sg_attention(tc): act_dtype must be SG_BF16 or SG_F16");
  SG_REQUIRE(L > 0 && (L & (L - 1)) == 0, "sg_attention(tc): L=%d must be a power of two", L);
  SG_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "sg_attention(tc): qkv / out must be 16-byte aligned");
  const int d = C / heads;
  AttGeom g;
  g.M = (int64_t)rows * L;
  SG_REQUIRE(g.M < (1ll << 31), "sg_attention(tc): too many tokens");
  g.L = L;
  g.logL = 0;
  while ((1 << g.logL) < L) ++g.logL;
  g.C = C;
  g.heads = heads;
  g.nkv = L >= ATT_BN ? L / ATT_BN : 1;
  g.c = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
  const uint32_t box_rows = (uint32_t)(g.M < 128 ? g.M : 128);
  g.tile_bytes = box_rows * (uint32_t)d * 2u;
  g.idesc_s = make_idesc(act_dtype, 128, ATT_BN, 0, 0);
  g.idesc_o = make_idesc(act_dtype, 128, d, 0, 1);  // B = V is MN-major
  g.idesc_1 = make_idesc(act_dtype, 128, 16, 0, 0);
  g.redo_log2 = act_dtype == SG_BF16 ? 60.0f : 13.0f;  // p <= 2^60 (bf16/fp32 range) / 2^13 (fp16 max 65504)
  g.act_dtype = act_dtype;
  CUtensorMap tm;
  const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)g.M};
  const uint64_t strides[1] = {(uint64_t)3 * C * 2};
  const uint32_t box[2] = {(uint32_t)d, box_rows};
  const CUtensorMapSwizzle sw = d == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (d == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  int rc = make_tmap(&tm, act_dtype, 2, qkv, dims, strides, box, sw);
  if (rc) return rc;
  SG_REQUIRE(heads <= 65535, "sg_attention(tc): too many heads");
  dim3 grid((unsigned)cdiv(g.M, ATT_BM), (unsigned)heads);
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  if (L >= 2 * ATT_BM && (attention_version() == 3 || (attention_version() == 0 && d >= 32))) {
    if (act_dtype == SG_BF16) {
      if (d == 16) return launch_att3<16, SG_BF16>(qkv, g, act_dtype, o, stream);
      if (d == 32) return launch_att3<32, SG_BF16>(qkv, g, act_dtype, o, stream);
      return launch_att3<64, SG_BF16>(qkv, g, act_dtype, o, stream);
    }
    if (d == 16) return launch_att3<16, SG_F16>(qkv, g, act_dtype, o, stream);
    if (d == 32) return launch_att3<32, SG_F16>(qkv, g, act_dtype, o, stream);
    return launch_att3<64, SG_F16>(qkv, g, act_dtype, o, stream);
  }
  if (L >= 2 * ATT_BM && attention_version() == 2) {
    dim3 grid2((unsigned)(g.M / (2 * ATT_BM)), (unsigned)heads);
    if (act_dtype == SG_BF16) {
      if (d == 16) return launch_att2<16, SG_BF16>(tm, g, o, grid2, stream);
      if (d == 32) return launch_att2<32, SG_BF16>(tm, g, o, grid2, stream);
      return launch_att2<64, SG_BF16>(tm, g, o, grid2, stream);
    }
    if (d == 16) return launch_att2<16, SG_F16>(tm, g, o, grid2, stream);
    if (d == 32) return launch_att2<32, SG_F16>(tm, g, o, grid2, stream);
    return launch_att2<64, SG_F16>(tm, g, o, grid2, stream);
  }
  const int poly = attention_poly();
#define SG_ATT_DISPATCH(DD, TT)                                                  \
  do {                                                                           \
    if (poly == 2) return launch_att<DD, TT, 2>(tm, g, o, grid, stream);         \
    if (poly == 4) return launch_att<DD, TT, 4>(tm, g, o, grid, stream);         \
    return launch_att<DD, TT, 0>(tm, g, o, grid, stream);                        \
  } while (0)
  if (act_dtype == SG_BF16) {
    if (d == 16) SG_ATT_DISPATCH(16, SG_BF16);
    if (d == 32) SG_ATT_DISPATCH(32, SG_BF16);
    SG_ATT_DISPATCH(64, SG_BF16);
  }
  if (d == 16) SG_ATT_DISPATCH(16, SG_F16);
  if (d == 32) SG_ATT_DISPATCH(32, SG_F16);
  SG_ATT_DISPATCH(64, SG_F16);
#undef SG_ATT_DISPATCH
}

}  // namespace sg
