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
// Two kernels live here: v8 (L >= 128, d = 32 / 64; four softmax warps, named-barrier hand-offs, packed fp32 body,
// part of the exponentials on the FMA pipe) and v1 (the first tcgen05 kernel: 192 threads, warp 0 = TMA producer,
// warp 1 = MMA issuer, warps 2..5 = softmax; used for L < 128, where a tile holds several batch rows and needs the
// block-diagonal mask).  d = 16 with L >= 128 (sa5 / sa6, 85 % of the attention time) runs v12, attention_tc12.cu.  The measured history of the other variants
// is in the dispatch comment of attention_tc() below and in profiles/README.md.
#include <stdlib.h>

#include <type_traits>

#include "attention_common.cuh"

namespace sg {
namespace tc {

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
  // heads are the FASTEST grid dimension: the head slices of a token (d * 2 = 32..128 bytes) share 128-byte lines, so
  // the heads of one query tile must run together to be served from L2 (with heads slowest, ncu showed 4x the
  // algorithmic DRAM reads at sa6: every head pass re-fetched all of qkv)
  const int head = blockIdx.x;
  const int64_t m0 = ((int64_t)blockIdx.z * 32768 + blockIdx.y) * ATT_BM;
  if (m0 >= g.M) return;  // ragged last grid.z slice (more than 32768 query tiles, not a multiple of it); nothing was touched yet
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
  pdl_wait();  // every activation access (TMA loads included) follows this point; parameters read above are immutable weights
  pdl_launch_dependents();

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
  if (int rc = set_max_smem<attention_tc_kernel<D, DT, POLY>>(smem, "sg_attention(tc)")) return rc;
  launch_k(attention_tc_kernel<D, DT, POLY>, grid, dim3(192), smem, stream, tm, g, out);
  return launch_status("sg_attention(tc)");
}

template <int D>
constexpr int att4_smem_bytes() {
  return 1024 + ATT_BM * D * 2 /*Q*/ + 2 * 2 * ATT_BN * D * 2 /*K,V x 2 stages*/ + 256;  // P lives in TMEM
}

// =====================================================================================================
// v8 (L >= 128): no dedicated producer / issuer warps at all.  ncu on v1/v5: the two single-lane roles poll their
// mbarriers in spin loops that make up 57 % of all issued instructions and share schedulers with the softmax warps.
// Here the CTA is just the four softmax warps (128 threads, four CTAs per SM, up to 128 registers); the hand-offs
// are named barriers: warps 1..3 bar.arrive and go on to wait for the product, warp 0 bar.sync's and its lane 0
// issues the tcgen05.mma's (and the next TMA loads) in line.  Nothing spins while the others compute.
//   per key tile:  max S_j | sweep S_j -> P_j | bar 1 | [P_j V_j] | wait o_full | o += O_j | bar 2 | [TMA j+2, S_{j+1}] | wait s_full
// Inner body as v5 (FFMA2 / FADD2 / FMNMX3, 16-column double-buffered TMEM reads, POLY/8 pairs on the FMA pipe).
// P never touches shared memory (round 2): the packed probabilities go back into TMEM with tcgen05.st -- columns
// [64, 128) of the 128 S columns, swept from the last chunk to the first so that chunk ch's eight P columns 64 + 8 ch ..
// lie inside the already consumed range [16 ch, 128) -- and P_j V_j takes its A operand from TMEM, its NACC accumulators
// from columns [0, 64).  Without the 32 KB P tile the CTA needs 41 KB (d = 32) / 81 KB (d = 64) of shared memory: four /
// two CTAs per SM instead of three / one (sa1 1.38 -> 1.21 ms, sa2 0.33 -> 0.19 ms).  S does not survive the sweep, so
// the exponent reference is raised to the tile maximum FIRST (one extra pass of TMEM loads and 3-input maxima, exact
// rescale of o, l): every p <= 1, no overflow guard, no second sweep.
// =====================================================================================================
template <int D, int DT, int POLY, int NACC>
__global__ void __launch_bounds__(128, (D == 64 ? 2 : 4))
attention_tc8_kernel(const __grid_constant__ CUtensorMap tm, const AttGeom g, uint16_t* __restrict__ out) {
  constexpr int ROWB = D * 2;
  constexpr int TILE = ATT_BN * ROWB;
  static_assert(NACC * D <= 64 && (ATT_BN / 16) % NACC == 0, "accumulators live in S columns [0, 64), P in [64, 128)");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE;      // [2 stages]
  uint8_t* sV = sK + 2 * TILE;  // [2 stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;  // [2]
  uint64_t* s_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heads are the FASTEST grid dimension: the head slices of a token (d * 2 = 32..128 bytes) share 128-byte lines, so
  // the heads of one query tile must run together to be served from L2 (with heads slowest, ncu showed 4x the
  // algorithmic DRAM reads at sa6: every head pass re-fetched all of qkv)
  const int head = blockIdx.x;
  const int64_t m0 = ((int64_t)blockIdx.z * 32768 + blockIdx.y) * ATT_BM;
  if (m0 >= g.M) return;  // ragged last grid.z slice (more than 32768 query tiles, not a multiple of it); nothing was touched yet
  const int64_t kv0 = (m0 >> g.logL) << g.logL;
  const int nkv = g.L / ATT_BN;
  const bool leader = threadIdx.x == 0;

  if (leader) {
    prefetch_tensormap(&tm);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(s_full, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc<128>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // every activation access (TMA loads included) follows this point; parameters read above are immutable weights
  pdl_launch_dependents();

  // Issue helpers: called by ALL lanes of warp 0 (convergent); one elected lane executes the TMA / MMA instructions.
  const uint64_t q_desc = make_desc_rows(smem_u32(sQ), ROWB);
  auto load_tile = [&](int t) {
    const int s = t & 1;
    const int tok = (int)(kv0 + (int64_t)t * ATT_BN);
    if (elect_one()) {
      mbar_arrive_expect_tx(&kv_full[s], 2 * g.tile_bytes);
      tma_load_2d(sK + s * TILE, &tm, &kv_full[s], g.C + head * D, tok);
      tma_load_2d(sV + s * TILE, &tm, &kv_full[s], 2 * g.C + head * D, tok);
    }
  };
  auto issue_s = [&](int t) {  // S = Q K_t^T
    mbar_wait_spin(&kv_full[t & 1], (uint32_t)(t >> 1) & 1u);
    tc_fence_after();
    const uint64_t kd = make_desc_rows(smem_u32(sK + (t & 1) * TILE), ROWB);
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < D / 16; ++k) umma_ss(tmem_base, q_desc + 2 * k, kd + 2 * k, g.idesc_s, k != 0);
      umma_commit(s_full);
    }
  };
  auto issue_pv = [&](int t) {  // O_t = P V_t : A = P in TMEM (columns [64, 128), eight per 16-key slab), B = V MN-major
    tc_fence_after();
    const uint64_t vd = make_desc_rows(smem_u32(sV + (t & 1) * TILE), ROWB);
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < ATT_BN / 16; ++k)
        umma_ts(tmem_base + (uint32_t)((k % NACC) * D), tmem_base + 64u + (uint32_t)(k * 8), vd + (uint64_t)(k * 16 * ROWB / 16),
                g.idesc_o, k >= NACC);
      umma_commit(o_full);
    }
  };
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, g.tile_bytes);
      tma_load_2d(sQ, &tm, q_full, head * D, (int)m0);
    }
    load_tile(0);
    if (nkv > 1) load_tile(1);
    mbar_wait_spin(q_full, 0);
    issue_s(0);
    __syncwarp();
  }

  const int r = warp * 32 + lane;
  const int64_t tok = m0 + r;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint64_t o2[D / 2];
#pragma unroll
  for (int i = 0; i < D / 2; ++i) o2[i] = 0ull;  // two +0.0f
  float m_ref = 0.f, l = 0.f;
  const uint64_t c2 = pk2(g.c, g.c);
  // row maximum of the current S tile
  auto tile_max = [&]() {
    float mx = -INFINITY;
#pragma unroll 1
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t v[16];
      tmem_ld16(t_row + ch * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; e += 2) mx = max3(mx, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
    }
    return mx;
  };
  for (int j = 0; j < nkv; ++j) {
    const uint32_t ph = (uint32_t)j & 1u;
    mbar_wait(s_full, ph);
    tc_fence_after();
    // raise the reference to the tile maximum first (exact rescale of o, l; a0 = 0 on the first tile, where o = l = 0)
    const float tmax = tile_max();
    if (tmax > m_ref || j == 0) {
      const float a0 = j == 0 ? 0.f : ex2((m_ref - tmax) * g.c);
      l *= a0;
      const uint64_t a2 = pk2(a0, a0);
#pragma unroll
      for (int i = 0; i < D / 2; ++i) o2[i] = mul2(o2[i], a2);
      m_ref = tmax;
    }
    const float nmc = -m_ref * g.c;
    const uint64_t nmc2 = pk2(nmc, nmc);
    uint64_t psum2 = 0ull;
    {
      uint32_t va[16], vb[16];
      tmem_ld16(t_row + 7 * 16, va);
      tmem_ld_wait();
      auto chunk = [&](const uint32_t(&v)[16], int ch) {  // 16 columns: 8 pairs -> eight packed P columns
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const uint64_t x2 = fma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nmc2);
          float p0, p1;
          if (pair_is_poly<POLY>(i >> 1)) {
            ex2_poly2(x2, p0, p1);
          } else {
            float x0, x1;
            un2(x2, x0, x1);
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          pk[i >> 1] = pack_pair<DT>(p0, p1);
          psum2 = add2(psum2, pk2(p0, p1));
        }
        tmem_st8(t_row + 64u + (uint32_t)(ch * 8), pk);
      };
#pragma unroll
      for (int ch = 7; ch > 0; ch -= 2) {  // last chunk first: P only overwrites S columns that have been read
        tmem_ld16(t_row + (ch - 1) * 16, vb);  // in flight while chunk ch is processed
        chunk(va, ch);
        tmem_ld_wait();
        if (ch - 2 >= 0) tmem_ld16(t_row + (ch - 2) * 16, va);
        chunk(vb, ch - 1);
        if (ch - 2 >= 0) tmem_ld_wait();
      }
      tmem_st_wait();
    }
    tc_fence_before();  // our tcgen05.ld of S and tcgen05.st of P precede the MMAs that read P and overwrite those columns
    if (warp == 0) {
      named_bar_sync<1, 128>();
      issue_pv(j);
      __syncwarp();
    } else {
      named_bar_arrive<1, 128>();
    }
    float ps0, ps1;
    un2(psum2, ps0, ps1);
    l += ps0 + ps1;
    mbar_wait(o_full, ph);
    tc_fence_after();
    if constexpr (NACC * D >= 32) {
#pragma unroll
      for (int cch = 0; cch < NACC * D / 32; ++cch) {
        uint32_t v[32];
        tmem_ld32(t_row + cch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int e = ((cch * 32 + i) % D) >> 1;
          o2[e] = add2(o2[e], pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        }
      }
    } else {
      uint32_t v[16];
      tmem_ld16(t_row, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; i += 2) o2[i >> 1] = add2(o2[i >> 1], pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
    }
    tc_fence_before();
    if (j + 1 < nkv) {
      if (warp == 0) {
        named_bar_sync<2, 128>();
        // o_full(j) has been observed: every MMA that read K/V stage j&1 is complete -> it can be refilled with tile j+2
        issue_s(j + 1);
        if (j + 2 < nkv) load_tile(j + 2);
        __syncwarp();
      } else {
        named_bar_arrive<2, 128>();
      }
    }
  }
  const float inv = 1.0f / l;
  uint16_t* dst = out + tok * g.C + head * D;
#pragma unroll
  for (int i = 0; i < D; i += 8) {
    float f[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) un2(o2[(i >> 1) + u], f[2 * u], f[2 * u + 1]);
    uint4 w;
    w.x = pack_pair<DT>(f[0] * inv, f[1] * inv);
    w.y = pack_pair<DT>(f[2] * inv, f[3] * inv);
    w.z = pack_pair<DT>(f[4] * inv, f[5] * inv);
    w.w = pack_pair<DT>(f[6] * inv, f[7] * inv);
    *reinterpret_cast<uint4*>(dst + i) = w;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<128>(tmem_base);
  }
}

template <int D, int DT, int POLY, int NACC>
static int launch_att8(const CUtensorMap& tm, const AttGeom& g, uint16_t* out, dim3 grid, cudaStream_t stream) {
  constexpr int smem = att4_smem_bytes<D>();
  if (int rc = set_max_smem<attention_tc8_kernel<D, DT, POLY, NACC>>(smem, "sg_attention(tc8)")) return rc;
  launch_k(attention_tc8_kernel<D, DT, POLY, NACC>, grid, dim3(128), smem, stream, tm, g, out);
  return launch_status("sg_attention(tc8)");
}

}  // namespace tc

int attention_tc12(const void* qkv, const tc::AttGeom& g, uint16_t* out, dim3 grid, cudaStream_t stream);  // attention_tc12.cu

int attention_tc(const void* qkv, void* out, int rows, int L, int C, int heads, int act_dtype, cudaStream_t stream) {
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
  g.idesc_ol = make_idesc(act_dtype, 128, d + 16, 0, 1);
  g.redo_log2 = act_dtype == SG_BF16 ? 60.0f : 13.0f;  // p <= 2^60 (bf16/fp32 range) / 2^13 (fp16 max 65504)
  g.l_max = act_dtype == SG_BF16 ? 1.152921504606847e18f : 16777216.0f;  // 2^60 / 2^24 (an overflowed fp16 p is +inf)
  g.act_dtype = act_dtype;
  CUtensorMap tm;
  const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)g.M};
  const uint64_t strides[1] = {(uint64_t)3 * C * 2};
  const uint32_t box[2] = {(uint32_t)d, box_rows};
  const CUtensorMapSwizzle sw = d == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (d == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  int rc = make_tmap(&tm, act_dtype, 2, qkv, dims, strides, box, sw);
  if (rc) return rc;
  SG_REQUIRE(heads <= 65535, "sg_attention(tc): too many heads");
  const int64_t tiles = cdiv(g.M, ATT_BM);
  dim3 grid((unsigned)heads, (unsigned)(tiles < 32768 ? tiles : 32768), (unsigned)cdiv(tiles, 32768));
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  // Kernel choice (measured on B200, bf16, ms; v2-v7 / v9 - v11 were experiments, see profiles/README.md):
  //                           v1     v8 (P in smem)  v8 (P in TMEM)  v12
  //   d=16 L=4096 rows=1024   -      -               -               12.9    v12 (attention_tc12.cu: double-buffered S, output and P
  //   d=16 L=1024 rows=1024   -      1.22            -               0.90    resident in TMEM, reference-free exponent) for d = 16;
  //   d=32 L=1024 rows=1024   -      1.38            1.21            -       v8 (named-barrier hand-offs, 1/4 polynomial exp2, P in
  //   d=64 L=256  rows=1024   -      0.33            0.19            -       TMEM, two / one accumulators) for d = 32 / 64
  // L < 128: v1, whose tile holds 128 / L batch rows under a block-diagonal mask.
#define SG_ATT_BY_DTYPE(CALL_BF16, CALL_F16) \
  do {                                       \
    if (act_dtype == SG_BF16) return CALL_BF16; \
    return CALL_F16;                         \
  } while (0)
  if (L >= ATT_BN && d == 16) return attention_tc12(qkv, g, o, grid, stream);
  if (L >= ATT_BN) {
    if (d == 32) SG_ATT_BY_DTYPE((launch_att8<32, SG_BF16, 2, 2>(tm, g, o, grid, stream)), (launch_att8<32, SG_F16, 2, 2>(tm, g, o, grid, stream)));
    SG_ATT_BY_DTYPE((launch_att8<64, SG_BF16, 2, 1>(tm, g, o, grid, stream)), (launch_att8<64, SG_F16, 2, 1>(tm, g, o, grid, stream)));
  }
  if (d == 16) SG_ATT_BY_DTYPE((launch_att<16, SG_BF16, 0>(tm, g, o, grid, stream)), (launch_att<16, SG_F16, 0>(tm, g, o, grid, stream)));
  if (d == 32) SG_ATT_BY_DTYPE((launch_att<32, SG_BF16, 0>(tm, g, o, grid, stream)), (launch_att<32, SG_F16, 0>(tm, g, o, grid, stream)));
  SG_ATT_BY_DTYPE((launch_att<64, SG_BF16, 0>(tm, g, o, grid, stream)), (launch_att<64, SG_F16, 0>(tm, g, o, grid, stream)));
#undef SG_ATT_BY_DTYPE
}

}  // namespace sg
