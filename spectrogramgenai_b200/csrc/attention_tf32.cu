// K4 (fp32-accurate tensor-core engine): flash-style multi-head self-attention with SPLIT TF32 operands.
//
// Replaces the scaled-dot-product inside nn.MultiheadAttention (/root/reference/src/diff_modules.py:56,:69) at the
// reference's own precision: the reference computes in fp32, and a plain TF32 contraction (10-bit mantissa) misses the
// 1e-4 parity bar by 10x.  Every fp32 operand x is therefore split into hi = tf32(x) and
// lo = tf32(x - hi) (sg_split_tf32), and each product is evaluated as three kind::tf32 MMAs into ONE fp32 TMEM accumulator
//   A B  ~=  A_hi B_hi + A_hi B_lo + A_lo B_hi                       (dropped term lo x lo: 2^-22 relative)
// so that S = Q K^T and O = P V carry ~21 mantissa bits.  The probabilities are split by the softmax threads themselves.
//
// One CTA = 128 consecutive query tokens x one head; key tiles of 64 tokens.  Per key tile:
//   S   = Q K^T      3 x d/8 tcgen05.mma kind::tf32 (M128 x N64 x K8), K-major operands           -> TMEM cols [0, 64)
//   m   = max(m, max_j S)   one query row per thread; o, l rescaled exactly when the maximum rises
//   P   = exp2(S c - m c)   fp32; P_hi written over S in place, P_lo into TMEM cols [64, 128) (tcgen05.st)
//   O_j = P V        3 x 8 tcgen05.mma (M128 x N=d x K8); A = P from TMEM, B = V^T, K-major       -> TMEM cols [128, 128+d)
//   o   = o + O_j           running output / maximum / sum in registers
// P never touches shared memory: as swizzled fp32 tiles (round 2's first cut) the probabilities cost 64 KB of stores and
// 96 KB of tensor-core operand reads per key tile, i.e. the kernel ran at the shared-memory bandwidth.
// The lo x hi / hi x lo products are issued BEFORE hi x hi: the tensor core's fp32 accumulation truncates, so the small
// terms are added while the accumulator is still small.
// V as the B operand of P V would be MN-major, which kind::tf32 only accepts in a dedicated 32-byte-base swizzle; instead
// sg_attn_prep_tf32 (one pass over the in_proj output) writes, next to the (hi, lo) split of q | k, the split of V
// TRANSPOSED per (batch row, head): vt [rows * heads * d, L], so that a [d x 64 keys] K-major tile is two TMA boxes.
// Warp roles (192 threads, as the first 16-bit kernel): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = softmax.
// Requires L >= 128 (power of two): shorter sequences (a handful of tokens per row, < 1 % of the FLOPs) stay on the CUDA-core
// kernel.  This engine is the accuracy mode (1 CTA per SM, serial phases): several times the CUDA-core kernel, not a
// roofline kernel -- the throughput path is attention_tc.cu.
#include "attention_common.cuh"

namespace sg {
namespace tc {

template <int D>
struct AttF {
  static constexpr int BM = 128, BN = 64;
  static constexpr int ROWB = D * 4 < 128 ? D * 4 : 128;  // bytes of one swizzle row of a head slice
  static constexpr int NATOM = D * 4 / ROWB;              // swizzle atoms along the head dimension (d = 64: 2)
  static constexpr int AE = ROWB / 4;                     // floats per atom row
  static constexpr int Q_ATOM = BM * ROWB, KV_ATOM = BN * ROWB;
  static constexpr int Q_TILE = NATOM * Q_ATOM;           // one of hi / lo
  static constexpr int KV_TILE = NATOM * KV_ATOM;         // K tile [64 keys x d]; the V^T tile [d x 64 keys] = two
  static constexpr int VT_ATOM = D * 128;                 // [d x 32 keys] SWIZZLE_128B atoms has the same size
  static_assert(2 * VT_ATOM == KV_TILE, "K and V^T tiles have the same size");
  // d = 16 (sa5 / sa6: 85 % of the attention FLOPs): one K / V stage keeps the CTA under half of the SM's shared memory,
  // and two resident CTAs overlap each other's serial chain -- worth more than a prefetched stage; d = 64 has no room for two
  // d = 16 (sa5 / sa6: 85 % of the attention FLOPs): 160 TMEM columns and 49 KB of shared memory per CTA, so three
  // resident CTAs overlap each other's serial chain (S -> sweep -> P V -> O read-back)
  static constexpr int STAGES = D == 64 ? 1 : 2;
  static constexpr int CTAS = D == 16 ? 3 : (D == 32 ? 2 : 1);
  static constexpr int KV_STAGE = 4 * KV_TILE;            // K_hi, K_lo, Vt_hi, Vt_lo
  static constexpr int O_COLS = D <= 32 ? 32 : 64;        // second TMEM allocation (the first: S / P_hi | P_lo, 128 columns)
  static constexpr int SMEM = 1024 + 2 * Q_TILE + STAGES * KV_STAGE + 256;
  static_assert(CTAS * (SMEM + 1024) <= 228 * 1024 && CTAS * (128 + O_COLS) <= 512, "smem / TMEM budget");
};

struct AttFGeom {
  int64_t M;  // rows * L tokens
  int L, logL, C, heads;
  int nkv;    // key tiles (64 tokens) per query tile
  float c;    // softmax scale * log2(e)
  uint32_t q_bytes, kv_bytes;  // bytes the TMA boxes of Q (hi + lo) / one K,V stage deliver
  uint32_t idesc_s, idesc_o;
};

// operand tile whose rows are one swizzle span of `row_bytes` (64 / 128); lbo_bytes = distance to the next atom along
// the OTHER dimension (MN-major B operand wider than one atom), 0 = single atom
__device__ __forceinline__ uint64_t make_desc_f(uint32_t saddr, int row_bytes, uint32_t lbo_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes ? (lbo_bytes >> 4) : 1) << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;  // 8-row groups
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <int D>
__global__ void __launch_bounds__(192, AttF<D>::CTAS)
attention_tf32_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                      const __grid_constant__ CUtensorMap tmKh, const __grid_constant__ CUtensorMap tmKl,
                      const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                      const AttFGeom g, float* __restrict__ out) {
  using A = AttF<D>;
  constexpr int ROWB = A::ROWB, NATOM = A::NATOM, AE = A::AE, STAGES = A::STAGES, BN = A::BN, BM = A::BM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                             // [hi | lo] x NATOM atoms
  uint8_t* sKV = sQ + 2 * A::Q_TILE;              // [STAGES] x [K_hi | K_lo | V_hi | V_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * A::KV_STAGE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* o_read = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x;  // heads fastest: the head slices of a token share cache lines
  const int64_t m0 = ((int64_t)blockIdx.z * 32768 + blockIdx.y) * BM;
  if (m0 >= g.M) return;
  const int64_t brow = m0 >> g.logL;        // batch row of this query tile (L >= 128: a tile lies inside one row)
  const int64_t kv0 = brow << g.logL;       // its first key token
  const int vt_row0 = (int)((brow * g.heads + head) * D);  // first row of this (batch row, head) in vt [rows*heads*d, L]

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQh);
    prefetch_tensormap(&tmQl);
    prefetch_tensormap(&tmKh);
    prefetch_tensormap(&tmKl);
    prefetch_tensormap(&tmVh);
    prefetch_tensormap(&tmVl);
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
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot + 1)), "n"(A::O_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot[0];  // S (P_hi in place) | P_lo
  const uint32_t tmem_o = tmem_slot[1];
  pdl_wait();  // every activation access (TMA loads included) follows this point
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: Q (hi, lo) once, then the K / V ring =====
      mbar_arrive_expect_tx(q_full, g.q_bytes);
#pragma unroll
      for (int a = 0; a < NATOM; ++a) {
        tma_load_2d(sQ + a * A::Q_ATOM, &tmQh, q_full, head * D + a * AE, (int)m0);
        tma_load_2d(sQ + A::Q_TILE + a * A::Q_ATOM, &tmQl, q_full, head * D + a * AE, (int)m0);
      }
      for (int j = 0; j < g.nkv; ++j) {
        const int s = j % STAGES;
        mbar_wait_spin(&kv_empty[s], ((uint32_t)(j / STAGES) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[s], g.kv_bytes);
        const int tok = (int)(kv0 + (int64_t)j * BN);
        uint8_t* st = sKV + s * A::KV_STAGE;
#pragma unroll
        for (int a = 0; a < NATOM; ++a) {
          tma_load_2d(st + 0 * A::KV_TILE + a * A::KV_ATOM, &tmKh, &kv_full[s], g.C + head * D + a * AE, tok);
          tma_load_2d(st + 1 * A::KV_TILE + a * A::KV_ATOM, &tmKl, &kv_full[s], g.C + head * D + a * AE, tok);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {  // V^T: [d x 32 keys] boxes at key offset j * 64 + a * 32 of this row's sequence
          tma_load_2d(st + 2 * A::KV_TILE + a * A::VT_ATOM, &tmVh, &kv_full[s], j * BN + a * 32, vt_row0);
          tma_load_2d(st + 3 * A::KV_TILE + a * A::VT_ATOM, &tmVl, &kv_full[s], j * BN + a * 32, vt_row0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, descriptors stay in uniform registers, one elected
    // lane issues -- inside an `if (lane == 0)` region every MMA costs ~15 dependent single-lane instructions, and this
    // kernel issues 30 small MMAs per key tile =====
    mbar_wait_spin(q_full, 0);
    const uint64_t qd0 = make_desc_f(smem_u32(sQ), ROWB, 0);          // + Q_TILE / 16: the lo part; + Q_ATOM / 16: next atom
    for (int j = 0; j < g.nkv; ++j) {
      const int s = j % STAGES;
      mbar_wait_spin(&kv_full[s], (uint32_t)(j / STAGES) & 1u);
      tc_fence_after();
      const uint32_t st = smem_u32(sKV + s * A::KV_STAGE);
      const uint64_t kd0 = make_desc_f(st, ROWB, 0);                    // K_hi; + KV_TILE / 16: K_lo
      const uint64_t vd0 = make_desc_k128(st + 2 * A::KV_TILE);         // Vt_hi; + KV_TILE / 16: Vt_lo
      // S = Q K^T = Ql Kh + Qh Kl + Qh Kh: K-major A and B, K = d in steps of 8 floats (32 bytes inside the swizzle span).
      // The S columns are free: p_ready(j-1) was observed before P_{j-1} V was issued.
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint64_t qd = qd0 + (uint64_t)((t == 0 ? A::Q_TILE : 0) >> 4);
          const uint64_t kd = kd0 + (uint64_t)((t == 1 ? A::KV_TILE : 0) >> 4);
#pragma unroll
          for (int a = 0; a < NATOM; ++a) {
#pragma unroll
            for (int k = 0; k < AE / 8; ++k) {
              umma_ss_tf32(tmem_base, qd + (uint64_t)((a * A::Q_ATOM) >> 4) + 2 * k, kd + (uint64_t)((a * A::KV_ATOM) >> 4) + 2 * k,
                           g.idesc_s, acc);
              acc = 1;
            }
          }
        }
        umma_commit(s_full);
      }
      __syncwarp();
      // O_j = P V = Pl Vh + Ph Vl + Ph Vh: A = P in TMEM (8 columns per K step), B = V^T K-major in two 32-key SWIZZLE_128B atoms
      mbar_wait_spin(p_ready, (uint32_t)j & 1u);
      if (j > 0) mbar_wait_spin(o_read, (uint32_t)(j - 1) & 1u);  // O_{j-1} has been added to the running output
      tc_fence_after();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t pa = tmem_base + (uint32_t)(t == 0 ? BN : 0);
          const uint64_t vd = vd0 + (uint64_t)((t == 1 ? A::KV_TILE : 0) >> 4);
#pragma unroll
          for (int k = 0; k < BN / 8; ++k) {
            umma_ts_tf32(tmem_o, pa + (uint32_t)(k * 8), vd + (uint64_t)(((k >> 2) * A::VT_ATOM + (k & 3) * 32) >> 4),
                         g.idesc_o, acc);
            acc = 1;
          }
        }
        umma_commit(&kv_empty[s]);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===== softmax + epilogue: thread = one query row (TMEM lane quadrant = warp % 4) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int64_t tok = m0 + r;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m_ref = -INFINITY, l = 0.f;  // running maximum of the scores; running row sum relative to it
    const uint32_t t_o = tmem_o + ((uint32_t)(q * 32) << 16);
    for (int j = 0; j < g.nkv; ++j) {
      mbar_wait_spin(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      // ---- pass 1: the tile maximum; raise the reference first (exact rescale of o, l), so that every p <= 1 ----
      float tmax = -INFINITY;
#pragma unroll
      for (int cch = 0; cch < BN / 16; ++cch) {
        uint32_t v[16];
        tmem_ld16(t_row + cch * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i += 2) tmax = max3(tmax, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      }
      if (tmax > m_ref) {
        float a0;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"((m_ref - tmax) * g.c));  // 0 on the first tile
        l *= a0;
#pragma unroll
        for (int i = 0; i < D; ++i) o[i] *= a0;
        m_ref = tmax;
      }
      // ---- pass 2: p = exp2(s c - m c); P_hi over S in place, P_lo next to it ----
      const float mc = m_ref * g.c;
      float psum = 0.f;
#pragma unroll
      for (int cch = 0; cch < BN / 16; ++cch) {
        uint32_t v[16], pl[16];
        tmem_ld16(t_row + cch * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          float p;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(__uint_as_float(v[u]), g.c, -mc)));
          psum += p;
          // hi = p rounded to TF32 with two integer instructions (p >= 0; cvt.rna.tf32 would put two more conversions
          // per element on the pipe the exponential already occupies); lo = p - hi is exact, the MMA reads its top 19 bits
          v[u] = (__float_as_uint(p) + 0x1000u) & 0xFFFFE000u;
          pl[u] = __float_as_uint(p - __uint_as_float(v[u]));
        }
        tmem_st16(t_row + cch * 16, v);
        tmem_st16(t_row + BN + cch * 16, pl);
      }
      tmem_st_wait();
      tc_fence_before();  // our tcgen05.ld of S / tcgen05.st of P precede the MMAs that read P and overwrite S
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // ---- O_j and the row sum, both relative to m_ref ----
      mbar_wait_spin(o_full, (uint32_t)j & 1u);
      tc_fence_after();
#pragma unroll
      for (int cch = 0; cch < D / 16; ++cch) {
        uint32_t v[16];
        tmem_ld16(t_o + cch * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[cch * 16 + i] += __uint_as_float(v[i]);
      }
      l += psum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_read);
    }
    {
      const float inv = 1.0f / l;
      float* dst = out + tok * g.C + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<128>(tmem_base);
    tmem_dealloc<A::O_COLS>(tmem_o);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// sg_attn_prep_tf32: in_proj output qkv fp32 [M, 3C] -> split q | k (hi, lo) [M, 2C] and split, per-(row, head) transposed
// V: vt (hi, lo) [rows * heads * d, L].  One block = 32 tokens of one batch row; 256 threads.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_prep_tf32_kernel(const float* __restrict__ qkv, float* __restrict__ qk_hi,
                                                             float* __restrict__ qk_lo, float* __restrict__ vt_hi,
                                                             float* __restrict__ vt_lo, int L, int C) {
  extern __shared__ float sv[];  // [C][33]: the V slab, transposed
  pdl_wait();
  pdl_launch_dependents();
  const int tiles_per_row = L / 32;
  const int brow = blockIdx.x / tiles_per_row, t0 = (blockIdx.x % tiles_per_row) * 32;
  const int64_t tok0 = (int64_t)brow * L + t0;
  const int C4 = C / 4;
  // q | k: 32 tokens x 2C floats, float4 per thread, coalesced in and out
  for (int i = threadIdx.x; i < 32 * 2 * C4; i += blockDim.x) {
    const int tk = i / (2 * C4), c4 = i % (2 * C4);
    const float4 v = __ldcs(reinterpret_cast<const float4*>(qkv + (tok0 + tk) * 3 * C) + c4);
    float4 h, l;
    h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
    reinterpret_cast<float4*>(qk_hi + (tok0 + tk) * 2 * C)[c4] = h;
    reinterpret_cast<float4*>(qk_lo + (tok0 + tk) * 2 * C)[c4] = l;
  }
  // v: [32 tokens x C] -> smem [C][33] (conflict-free column writes), then rows of 32 tokens per channel
  for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) {
    const int tk = i / C, c = i % C;
    sv[c * 33 + tk] = __ldcs(qkv + (tok0 + tk) * 3 * C + 2 * C + c);
  }
  __syncthreads();
  // channel c of batch row brow = head c / d, component c % d -> vt row (brow * heads + head) * d + c % d = brow * C + c
  for (int i = threadIdx.x; i < 32 * C; i += blockDim.x) {
    const int c = i / 32, tk = i % 32;
    const float v = sv[c * 33 + tk];
    const float h = tf32_rna(v);
    const int64_t o = ((int64_t)brow * C + c) * L + t0 + tk;
    vt_hi[o] = h;
    vt_lo[o] = tf32_rna(v - h);
  }
}

template <int D>
static int launch_att_tf32(const float* qk_hi, const float* qk_lo, const float* vt_hi, const float* vt_lo, float* out,
                           AttFGeom g, cudaStream_t stream) {
  using A = AttF<D>;
  const CUtensorMapSwizzle sw = A::ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const uint64_t dims[2] = {(uint64_t)2 * g.C, (uint64_t)g.M};
  const uint64_t strides[1] = {(uint64_t)2 * g.C * 4};
  const uint32_t qbox[2] = {(uint32_t)A::AE, (uint32_t)A::BM}, kbox[2] = {(uint32_t)A::AE, (uint32_t)A::BN};
  const uint64_t vdims[2] = {(uint64_t)g.L, (uint64_t)(g.M / g.L) * g.C};  // [rows * heads * d, L], L contiguous
  const uint64_t vstrides[1] = {(uint64_t)g.L * 4};
  const uint32_t vbox[2] = {32u, (uint32_t)D};
  CUtensorMap tmQh, tmQl, tmKh, tmKl, tmVh, tmVl;
  int rc;
  if ((rc = make_tmap(&tmQh, SG_F32, 2, qk_hi, dims, strides, qbox, sw))) return rc;
  if ((rc = make_tmap(&tmQl, SG_F32, 2, qk_lo, dims, strides, qbox, sw))) return rc;
  if ((rc = make_tmap(&tmKh, SG_F32, 2, qk_hi, dims, strides, kbox, sw))) return rc;
  if ((rc = make_tmap(&tmKl, SG_F32, 2, qk_lo, dims, strides, kbox, sw))) return rc;
  if ((rc = make_tmap(&tmVh, SG_F32, 2, vt_hi, vdims, vstrides, vbox, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap(&tmVl, SG_F32, 2, vt_lo, vdims, vstrides, vbox, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  g.q_bytes = 2u * A::Q_TILE;
  g.kv_bytes = (uint32_t)A::KV_STAGE;
  g.idesc_s = make_idesc_tf32(128, A::BN, 0);
  g.idesc_o = make_idesc_tf32(128, D, 0);
  if ((rc = set_max_smem<attention_tf32_kernel<D>>(A::SMEM, "sg_attention_tf32"))) return rc;
  const int64_t tiles = g.M / A::BM;
  dim3 grid((unsigned)g.heads, (unsigned)(tiles < 32768 ? tiles : 32768), (unsigned)cdiv(tiles, 32768));
  launch_k(attention_tf32_kernel<D>, grid, dim3(192), (size_t)A::SMEM, stream, tmQh, tmQl, tmKh, tmKl, tmVh, tmVl, g, out);
  return launch_status("sg_attention_tf32");
}

}  // namespace tc
}  // namespace sg

using namespace sg;
using namespace sg::tc;

static int att_tf32_check(const char* what, int rows, int L, int C, int heads) {
  SG_REQUIRE(rows > 0 && L >= 128 && (L & (L - 1)) == 0 && heads > 0 && heads <= 65535 && C % heads == 0,
             "%s: bad shape rows=%d L=%d C=%d heads=%d (L must be a power of two >= 128)", what, rows, L, C, heads);
  const int d = C / heads;
  SG_REQUIRE(d == 16 || d == 32 || d == 64, "%s: head dim %d not in {16,32,64}", what, d);
  SG_REQUIRE((int64_t)rows * L < (1ll << 31), "%s: too many tokens", what);
  return SG_OK;
}

extern "C" int sg_attn_prep_tf32(const float* qkv, float* qk_hi, float* qk_lo, float* vt_hi, float* vt_lo, int rows, int L,
                                 int C, int heads, sg_stream_t stream) {
  SG_REQUIRE(qkv && qk_hi && qk_lo && vt_hi && vt_lo, "sg_attn_prep_tf32: null pointer");
  if (int rc = att_tf32_check("sg_attn_prep_tf32", rows, L, C, heads)) return rc;
  SG_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(qk_hi) | reinterpret_cast<uintptr_t>(qk_lo)) & 15) == 0,
             "sg_attn_prep_tf32: buffers must be 16-byte aligned");
  const size_t smem = (size_t)C * 33 * 4;  // <= 33 KB
  launch_k(attn_prep_tf32_kernel, dim3((unsigned)(rows * (L / 32))), dim3(256), smem, as_stream(stream), qkv, qk_hi, qk_lo,
           vt_hi, vt_lo, L, C);
  return launch_status("sg_attn_prep_tf32");
}

extern "C" int sg_attention_tf32(const float* qk_hi, const float* qk_lo, const float* vt_hi, const float* vt_lo, float* out,
                                 int rows, int L, int C, int heads, sg_stream_t stream) {
  SG_REQUIRE(qk_hi && qk_lo && vt_hi && vt_lo && out, "sg_attention_tf32: null pointer");
  if (int rc = att_tf32_check("sg_attention_tf32", rows, L, C, heads)) return rc;
  SG_REQUIRE(((reinterpret_cast<uintptr_t>(qk_hi) | reinterpret_cast<uintptr_t>(qk_lo) | reinterpret_cast<uintptr_t>(vt_hi) |
               reinterpret_cast<uintptr_t>(vt_lo) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "sg_attention_tf32: operands must be 16-byte aligned");
  const int d = C / heads;
  AttFGeom g;
  g.M = (int64_t)rows * L;
  g.L = L;
  g.logL = 0;
  while ((1 << g.logL) < L) ++g.logL;
  g.C = C;
  g.heads = heads;
  g.nkv = L / 64;
  g.c = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
  cudaStream_t s = as_stream(stream);
  if (d == 16) return launch_att_tf32<16>(qk_hi, qk_lo, vt_hi, vt_lo, out, g, s);
  if (d == 32) return launch_att_tf32<32>(qk_hi, qk_lo, vt_hi, vt_lo, out, g, s);
  return launch_att_tf32<64>(qk_hi, qk_lo, vt_hi, vt_lo, out, g, s);
}
