// K1 (tensor-core engine): implicit-GEMM 3x3 convolution / linear layer on tcgen05 + TMEM, fed by TMA.
//
//   out[m, co] = sum_{tap, ci} act[pixel(m) + tap, ci] * w[tap, co, ci]         M = rows*H*W, N = Cout, K = taps*Cin
//
// * A operand: the NHWC activation tensor is described by ONE rank-4 tensor map {C, W, H, rows}.  For k-block
//   (tap, 64-channel slice) the producer issues a single TMA box load {64, bw, bh, bn} at
//   (c0, w0+dx, h0+dy, n0); the conv's zero padding is TMA's out-of-bounds zero fill, so there is no im2col
//   buffer, no halo logic and no predication anywhere.  The box lands as 128 rows x 128 B (SWIZZLE_128B),
//   which is exactly the canonical K-major UMMA operand tile.
// * B operand: weights pre-packed [tap][Cout][Cin] (16-bit), rank-2 map, box {64, BN}.
// * One elected lane of the MMA warp issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block into each of the tile's two
//   TMEM accumulators; tcgen05.commit releases the smem stage back to the producer and finally signals the epilogue.
// * Epilogue warps read TMEM (tcgen05.ld 32x32b: one accumulator row per thread), apply bias / GELU /
//   residual, write fp32 and/or 16-bit outputs, and emit deterministic GroupNorm(1,C) partial sums.
// Persistent CTAs, one per SM, 256-pixel work tiles, double-buffered TMEM accumulators: see igemm_tc2_kernel below (warp
// roles, the split-TF32 passes of the fp32 engine and the slab pipeline of the Cout = 64 layers).
#include <stdlib.h>

#include "tc_common.cuh"

namespace sg {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    if (e == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
  }
  return fn;
}

int make_tmap(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims, const uint64_t* strides,
              const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return SG_ERR_LAUNCH;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides[i];
  }
  const CUtensorMapDataType dt = dtype == SG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : (dtype == SG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return SG_ERR_ARG;
  }
  return SG_OK;
}

constexpr int BM = 128;        // UMMA M (one TMEM lane per output row)
constexpr int BK = 64;         // one 128-byte swizzle span of 16-bit channels per k-block
constexpr int A_BYTES = BM * BK * 2;

struct IgemmGeom {
  int64_t M;
  int H, W, HW;
  int Cout, taps, cblocks;  // cblocks = Cin / bke
  int bke;                  // channels per k-block: one 128-byte swizzle span = 64 (16-bit operands) or 32 (tf32)
  int passes;               // 1 (16-bit operands) or 3 (split tf32: A_hi B_hi + A_hi B_lo + A_lo B_hi)
  int n_tiles;              // Cout / BN
  int tiles_per_sample;     // HW / 128 when HW >= 128, else 0
  int samples_per_tile;     // 128 / HW when HW < 128, else 0
  uint32_t tx_bytes;        // bytes one k-block's two TMA boxes deliver
  uint32_t idesc;
};
struct IgemmEpi {
  const float* bias;
  const float* residual;
  float* out_f32;
  void* out_act;
  float* partials;
  int act, P, act_dtype;  // act: sg_act
};

// Coalesced epilogue of one 128-row x BN accumulator, executed by ONE group of four warps (TMEM lane quadrant =
// warp % 4).  Each warp owns a 32-row x 32-column block per step: thread = accumulator row for the TMEM read, bias,
// GELU and the GroupNorm row statistics; the block is then transposed through a per-warp smem scratch so that every
// global access instruction touches 4 rows x 128 contiguous bytes (4 transactions instead of 32) for the residual
// read and for the fp32 / 16-bit stores.
constexpr int EPI_LD = 36;                          // scratch row stride in floats (144 B: conflict-free 128-bit access)
constexpr int EPI_SCRATCH = 32 * EPI_LD * 4;        // bytes per warp
constexpr int MAX_COUT = 1024;                      // bias staged in smem
constexpr int BIAS_BYTES = MAX_COUT * 4;
template <int BN>
__device__ __forceinline__ void epilogue_coalesced(uint32_t tmem_acc, int64_t m0, int n0, int tile_n, int warp, int lane,
                                                   int group_tid, int bar_id, const IgemmGeom& g, const IgemmEpi& ep,
                                                   float (*rowstat)[2], float* scratch, const float* s_bias) {
  const int q = warp & 3;
  const int r = q * 32 + lane;
  const bool valid = m0 + r < g.M;
  float s_sum = 0.f, s_sq = 0.f;
  float s_even = 0.f, s_odd = 0.f, q_even = 0.f, q_odd = 0.f;  // generic path: per-parity chains (see below)
  const int trow = lane >> 3, tcol = (lane & 7) * 4;  // transposed-domain role of this lane
  // Fast path for the two hottest epilogues -- raw fp16 conv output + GroupNorm statistics, and 16-bit Linear output
  // with bias (+GELU) -- on full tiles: packed fp32 statistics, the 16-bit values (not fp32) go through the transposing
  // scratch, no per-row predicates / residual slots.  (The generic loop below costs ~7 instructions per element on
  // only eight epilogue warps and bounds the Cout = 64 convolutions and the 16-bit Linear layers.)
  const bool fast = ep.out_act && !ep.out_f32 && !ep.residual && ep.act != SG_ACT_RELU_POST && m0 + BM <= g.M;
  if (fast) {
    uint32_t* sc32 = reinterpret_cast<uint32_t*>(scratch);  // [32 rows][20 words]: 64 B of 16-bit data per row + pad
    uint16_t* outp = reinterpret_cast<uint16_t*>(ep.out_act) + (m0 + q * 32) * (int64_t)g.Cout + n0;
    const bool has_bias = ep.bias != nullptr, gelu = ep.act == SG_ACT_GELU, stats = ep.partials != nullptr;
    uint64_t s2 = 0ull, q2 = 0ull;
    const int prow = lane >> 2, piece = lane & 3;  // store role: 8 rows x four 16-byte pieces per instruction
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (has_bias) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + n0 + c * 32 + j4 * 4);
          f[j4 * 4 + 0] += b4.x; f[j4 * 4 + 1] += b4.y; f[j4 * 4 + 2] += b4.z; f[j4 * 4 + 3] += b4.w;
        }
      }
      if (gelu) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) gelu_erf2(f[j], f[j + 1]);
      }
      if (stats) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          uint64_t p;
          asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(f[j]), "f"(f[j + 1]));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s2) : "l"(p));
          asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q2) : "l"(p));
        }
      }
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        uint4 w;
        w.x = pack16(f[j4 * 8 + 0], f[j4 * 8 + 1], ep.act_dtype);
        w.y = pack16(f[j4 * 8 + 2], f[j4 * 8 + 3], ep.act_dtype);
        w.z = pack16(f[j4 * 8 + 4], f[j4 * 8 + 5], ep.act_dtype);
        w.w = pack16(f[j4 * 8 + 6], f[j4 * 8 + 7], ep.act_dtype);
        *reinterpret_cast<uint4*>(sc32 + lane * 20 + j4 * 4) = w;
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int row = it * 8 + prow;
        const uint4 w = *reinterpret_cast<const uint4*>(sc32 + row * 20 + piece * 4);
        *reinterpret_cast<uint4*>(outp + (int64_t)row * g.Cout + c * 32 + piece * 8) = w;
      }
      __syncwarp();
    }
    if (stats) {
      float a0, a1, b0, b1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(s2));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(q2));
      s_sum = __fadd_rn(a0, a1);
      s_sq = __fadd_rn(b0, b1);
    }
  }
#pragma unroll 1
  for (int c = 0; c < (fast ? 0 : BN / 32); ++c) {
    const int nb = n0 + c * 32;
    // residual loads first: their HBM latency overlaps the TMEM read, the bias / GELU math and the transposition
    float4 rsd[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + trow;
      rsd[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.residual && m0 + q * 32 + rr < g.M)
        rsd[it] = __ldg(reinterpret_cast<const float4*>(ep.residual + (m0 + q * 32 + rr) * g.Cout + nb + tcol));
    }
    uint32_t v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 b4 = *reinterpret_cast<const float4*>(s_bias + nb + j4 * 4);  // smem broadcast (zeros if no bias)
      f[j4 * 4 + 0] = __uint_as_float(v[j4 * 4 + 0]) + b4.x;
      f[j4 * 4 + 1] = __uint_as_float(v[j4 * 4 + 1]) + b4.y;
      f[j4 * 4 + 2] = __uint_as_float(v[j4 * 4 + 2]) + b4.z;
      f[j4 * 4 + 3] = __uint_as_float(v[j4 * 4 + 3]) + b4.w;
    }
    if (ep.act == SG_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) gelu_erf2(f[j], f[j + 1]);
    }
    if (ep.partials && valid) {
      // same summation order as the packed fast path (even / odd columns in two chains, combined at the end), so that
      // a sample's GroupNorm statistics do not depend on whether its tile took the fast or the generic path
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        s_even = __fadd_rn(s_even, f[j]);
        s_odd = __fadd_rn(s_odd, f[j + 1]);
        q_even = __fmaf_rn(f[j], f[j], q_even);
        q_odd = __fmaf_rn(f[j + 1], f[j + 1], q_odd);
      }
    }
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
      *reinterpret_cast<float4*>(scratch + lane * EPI_LD + j4 * 4) =
          make_float4(f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
    __syncwarp();
    float4 o[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) o[it] = *reinterpret_cast<const float4*>(scratch + (it * 4 + trow) * EPI_LD + tcol);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int64_t m = m0 + q * 32 + it * 4 + trow;
      if (m < g.M) {
        const int64_t off = m * g.Cout + nb + tcol;
        float4 v4 = make_float4(o[it].x + rsd[it].x, o[it].y + rsd[it].y, o[it].z + rsd[it].z, o[it].w + rsd[it].w);
        if (ep.act == SG_ACT_RELU_POST) v4 = make_float4(fmaxf(v4.x, 0.f), fmaxf(v4.y, 0.f), fmaxf(v4.z, 0.f), fmaxf(v4.w, 0.f));
        if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + off) = v4;
        if (ep.out_act) {
          uint2 w;
          w.x = pack16(v4.x, v4.y, ep.act_dtype);
          w.y = pack16(v4.z, v4.w, ep.act_dtype);
          *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ep.out_act) + off) = w;
        }
      }
    }
    __syncwarp();
  }
  if (!fast) {
    s_sum = __fadd_rn(s_even, s_odd);
    s_sq = __fadd_rn(q_even, q_odd);
  }
  if (ep.partials) {
    // Deterministic GroupNorm partials without a serial tail: xor-shuffle within the sample's rows of this warp,
    // then (samples spanning several warps) a 4-entry smem combine.  group = rows of one sample inside the tile.
    const int group = g.HW < BM ? g.HW : BM;
    const int span = group < 32 ? group : 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (o < span) {
        s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
        s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
      }
    }
    if (group < 32) {
      if ((lane & (group - 1)) == 0 && valid) {
        const int64_t sample = (m0 + r) / g.HW;
        float* pp = ep.partials + (sample * ep.P + tile_n) * 2;
        pp[0] = s_sum;
        pp[1] = s_sq;
      }
    } else {
      if (lane == 0) {
        rowstat[q][0] = s_sum;
        rowstat[q][1] = s_sq;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // this group's four warps only
      const int wpg = group >> 5;  // warps per sample: 1, 2 or 4
      if (group_tid < 4 / wpg && m0 + group_tid * group < g.M) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < wpg; ++w) {
          a += rowstat[group_tid * wpg + w][0];
          b += rowstat[group_tid * wpg + w][1];
        }
        const int64_t row0 = m0 + group_tid * group;
        const int64_t sample = row0 / g.HW;
        const int tile_in_sample = g.HW < BM ? 0 : (int)((row0 % g.HW) / BM);
        float* pp = ep.partials + (sample * ep.P + tile_in_sample * g.n_tiles + tile_n) * 2;
        pp[0] = a;
        pp[1] = b;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // rowstat is rewritten by the next tile
    }
  }
}

// =====================================================================================================
// v2: persistent CTAs, 256 x BN output tiles (two M=128 accumulators share every B k-block, so the
// L2 -> smem operand traffic per MMA drops by 25 %), TMEM double-buffered so that the epilogue of tile i
// overlaps the main loop of tile i+1, and TWO epilogue warp groups (one per accumulator) with coalesced stores.
// One CTA per SM (<= 215 KB smem, up to 512 TMEM columns).  Warp roles (320 threads): 0 = TMA producer,
// 1 = TMEM allocator + MMA issuer, 2..5 = epilogue of rows [0,128), 6..9 = epilogue of rows [128,256).
// =====================================================================================================
// SLAB (Cout = 64 3x3 layers on 64- / 32-pixel-wide images, 16-bit operands): those layers are bound by L2 -> smem operand
// traffic (tensor pipe 36 %: a 256 x 64 tile re-reads its A tile for each of the nine taps and only 64 output channels
// amortise it).  A stage is then (64-channel slice, column shift dx): ONE TMA box of the tile's image rows plus a halo
// row above and below, shifted by dx -- {64, W, 256/W + 2, 1} -- serves the three taps (dy = -1, 0, +1) as views that
// start dy image rows further down (row pitch W x 128 B = 8 / 4 KB: a multiple of the 1024-byte swizzle atom, so the
// views are ordinary descriptors), with the three taps' weight tiles behind it: 72 KB per 24 MMAs instead of 120 KB.
// SLAB = 2 (Cin = 64 as well: the 64 -> 64 layers): the nine 8 KB weight tiles are loaded ONCE per CTA and stay resident, a
// stage is the A slab alone -- a third of the L2 reads of SLAB = 1 were weight tiles every CTA re-fetched for every tile.
// What is left (ncu: tensor pipe 47-50 %, smem->tensor 56 %): an M128 x N64 x K16 MMA is 32 clocks of math but reads 4 KB
// of A + 2 KB of B through the shared-memory port the TMA fills share -- N = 64 gives the A bytes too little reuse; an L2
// prefetch of the next tile's slab measured 6 % slower (the layer is not latency-bound).
constexpr int SLAB_A_BYTES = 48 * 1024;  // (256 / W + 2) rows x W pixels x 128 B: 48 KB (W = 64) / 40 KB (W = 32)
template <int BN, int SLAB = 0>
struct V2 {
  static constexpr int STAGES = SLAB ? 2 : (BN == 128 ? 3 : 4);  // 72 KB (48 KB) / 48 KB / 40 KB per stage
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = SLAB == 2 ? SLAB_A_BYTES : (SLAB == 1 ? SLAB_A_BYTES + 3 * B_BYTES : 2 * A_BYTES + B_BYTES);
  static constexpr int RES_B = SLAB == 2 ? 9 * B_BYTES : 0;  // resident weight tiles, behind the stage ring
  static constexpr int TMEM_COLS = 4 * BN;  // 2 buffers x 2 accumulators: 512 (BN=128) / 256 (BN=64)
  static constexpr int THREADS = 320;
  static constexpr int RING = STAGES * STAGE + RES_B;  // bytes in front of the barriers
  static_assert(1024 + RING + 256 + 2 * BM * 2 * 4 + 8 * EPI_SCRATCH + BIAS_BYTES <= 227 * 1024, "smem budget");
  static constexpr int SMEM = 1024 + RING + 256 + 2 * BM * 2 * 4 + 8 * EPI_SCRATCH + BIAS_BYTES;
};

__device__ __forceinline__ void tile_coords(const IgemmGeom& g, int mt, int& cn, int& ch, int& cw) {
  if (g.taps == 1) {
    cn = 0; ch = 0; cw = mt * BM;  // linear layer: the map is {Cin, M, 1, 1}
  } else if (g.tiles_per_sample > 0) {
    cn = mt / g.tiles_per_sample;
    const int p0 = (mt % g.tiles_per_sample) * BM;
    ch = p0 / g.W;
    cw = p0 % g.W;
  } else {
    cn = mt * g.samples_per_tile; ch = 0; cw = 0;
  }
}

// KIND = 0: kind::f16 (bf16 / fp16 operands), one pass over K.  KIND = 1: the fp32-accurate engine -- kind::tf32 on SPLIT
// operands (x = hi + lo, both tf32-representable fp32 tensors written by sg_split_tf32): the k-loop runs three times over
// (tap, channel block) with the operand maps (A_lo, B_hi), (A_hi, B_lo), (A_hi, B_hi), all accumulating into the same
// fp32 TMEM tile, so the product carries ~21 mantissa bits (the dropped lo x lo term is 2^-22 relative) at 1/6 of the
// 16-bit rate.  A k-block is the same 128-byte swizzle span either way (64 x 16 bit or 32 x fp32) and one MMA consumes
// 32 bytes of it (K = 16 or K = 8), so tiles, descriptors and the pipeline are identical.
template <int BN, int KIND, int SLAB>
__global__ void __launch_bounds__(320, 1) igemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB,
                                                           const __grid_constant__ CUtensorMap tmA_lo,
                                                           const __grid_constant__ CUtensorMap tmB_lo, const IgemmGeom g,
                                                           const IgemmEpi ep, const int num_tiles) {
  using K = V2<BN, SLAB>;
  constexpr int STAGES = K::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* res_b = smem + STAGES * K::STAGE;  // SLAB == 2: the nine resident weight tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + K::RING);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;   // [2]
  uint64_t* b_full = tmem_empty + 2;      // SLAB == 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  float(*rowstat)[2] = reinterpret_cast<float(*)[2]>(smem + K::RING + 256);           // [2 groups][128][2]
  float* scratch_all = reinterpret_cast<float*>(smem + K::RING + 256 + 2 * BM * 2 * 4);  // [8 warps]
  float* s_bias = scratch_all + 8 * (EPI_SCRATCH / 4);
  for (int i = threadIdx.x; i < g.Cout; i += blockDim.x) s_bias[i] = ep.bias ? __ldg(ep.bias + i) : 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk1 = g.taps * g.cblocks;  // k-blocks of one pass
  const int nk = SLAB ? 3 * g.cblocks : nk1 * g.passes;  // SLAB: stages = (channel slice, column shift)

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (KIND == 1) {
      prefetch_tensormap(&tmA_lo);
      prefetch_tensormap(&tmB_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 256);
    }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<K::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // (the bias above is an immutable weight; every activation access follows this point)
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
    uint32_t it = 0;  // running k-block counter across tiles
    if constexpr (SLAB == 2) {  // weights are immutable: loaded once, resident for every tile of this CTA
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full, (uint32_t)K::RES_B);
#pragma unroll
        for (int t = 0; t < 9; ++t) tma_load_2d(res_b + t * K::B_BYTES, &tmB, b_full, 0, t * g.Cout);
      }
      __syncwarp();
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt2 = tile / g.n_tiles, n0 = (tile % g.n_tiles) * BN;
      int cn0, ch0, cw0, cn1, ch1, cw1;
      tile_coords(g, 2 * mt2, cn0, ch0, cw0);
      tile_coords(g, 2 * mt2 + 1, cn1, ch1, cw1);
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait_spin(&empty[s], ((it / STAGES) & 1u) ^ 1u);
        if constexpr (SLAB) {
          const int c0 = (kb / 3) * BK, dxi = kb % 3;  // dx = dxi - 1; the slab starts one image row above the tile
          uint8_t* st = smem + s * K::STAGE;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[s], g.tx_bytes);
            tma_load_4d(st, &tmA, &full[s], c0, dxi - 1, ch0 - 1, cn0);
            if constexpr (SLAB == 1) {
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi)
                tma_load_2d(st + SLAB_A_BYTES + dyi * K::B_BYTES, &tmB, &full[s], c0, (dyi * 3 + dxi) * g.Cout + n0);
            }
          }
          __syncwarp();
          continue;
        }
        const int pass = KIND == 1 ? kb / nk1 : 0, kk = kb - pass * nk1;
        const int tap = kk / g.cblocks, c0 = (kk % g.cblocks) * g.bke;
        int dy = 0, dx = 0;
        if (g.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
        // small terms first (a_lo b_hi, a_hi b_lo, then a_hi b_hi): the tensor core's fp32 accumulation truncates
        // (measured: a uniform relative shrink of ~2.5e-8 per accumulating MMA), so only the last pass adds to a
        // full-size accumulator
        const CUtensorMap* ma = (KIND == 1 && pass == 0) ? &tmA_lo : &tmA;
        const CUtensorMap* mb = (KIND == 1 && pass == 1) ? &tmB_lo : &tmB;
        uint8_t* st = smem + s * K::STAGE;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], g.tx_bytes);
          tma_load_4d(st, ma, &full[s], c0, cw0 + dx, ch0 + dy, cn0);
          tma_load_4d(st + A_BYTES, ma, &full[s], c0, cw1 + dx, ch1 + dy, cn1);
          tma_load_2d(st + 2 * A_BYTES, mb, &full[s], c0, tap * g.Cout + n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, descriptors in uniform registers, one elected lane issues =====
    uint32_t it = 0, lt = 0;
    if constexpr (SLAB == 2) {
      mbar_wait_spin(b_full, 0);
      tc_fence_after();
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1u;
      mbar_wait_spin(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator pair
      tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * (2 * BN), acc1 = acc0 + BN;
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait_spin(&full[s], (it / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * K::STAGE);
        if constexpr (SLAB) {
          const uint32_t pitch = (uint32_t)g.W * 128u;  // bytes of one image row of the slab
          if (elect_one()) {
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const uint64_t a0 = make_desc_k128(sa + dyi * pitch), a1 = make_desc_k128(sa + dyi * pitch + A_BYTES);
              // SLAB == 2: one channel slice, so kb IS the column shift; tap = dy * 3 + dx
              const uint64_t bd = make_desc_k128(SLAB == 2 ? smem_u32(res_b) + (uint32_t)((dyi * 3 + kb) * K::B_BYTES)
                                                           : sa + SLAB_A_BYTES + dyi * K::B_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_ss(acc0, a0 + 2 * k, bd + 2 * k, g.idesc, (kb | dyi | k) != 0);
                umma_ss(acc1, a1 + 2 * k, bd + 2 * k, g.idesc, (kb | dyi | k) != 0);
              }
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          continue;
        }
        const uint64_t a0 = make_desc_k128(sa), a1 = make_desc_k128(sa + A_BYTES);
        const uint64_t bd = make_desc_k128(sa + 2 * A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per 128-byte swizzle span
            if constexpr (KIND == 1) {
              umma_ss_tf32(acc0, a0 + 2 * k, bd + 2 * k, g.idesc, (kb | k) != 0);
              umma_ss_tf32(acc1, a1 + 2 * k, bd + 2 * k, g.idesc, (kb | k) != 0);
            } else {
              umma_ss(acc0, a0 + 2 * k, bd + 2 * k, g.idesc, (kb | k) != 0);
              umma_ss(acc1, a1 + 2 * k, bd + 2 * k, g.idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full[buf]);
      __syncwarp();
    }
  } else {
    // ===== epilogue: group 0 (warps 2..5) drains accumulator 0, group 1 (warps 6..9) accumulator 1 =====
    const int grp = (warp - 2) >> 2;
    const int group_tid = (int)threadIdx.x - 64 - grp * 128;
    float* scratch = scratch_all + (warp - 2) * (EPI_SCRATCH / 4);
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const uint32_t buf = lt & 1u;
      const int mt2 = tile / g.n_tiles, tile_n = tile % g.n_tiles;
      mbar_wait_spin(&tmem_full[buf], (lt >> 1) & 1u);
      tc_fence_after();
      const int64_t m0 = (int64_t)(2 * mt2 + grp) * BM;
      if (m0 < g.M)
        epilogue_coalesced<BN>(tmem_base + buf * (2 * BN) + grp * BN, m0, tile_n * BN, tile_n, warp, lane, group_tid,
                               1 + grp, g, ep, rowstat + grp * BM, scratch, s_bias);
      tc_fence_before();
      mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<K::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int KIND, int SLAB = 0>
static int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_lo,
                   const IgemmGeom& g, const IgemmEpi& ep, cudaStream_t stream) {
  constexpr int smem = V2<BN, SLAB>::SMEM;
  if (int rc = set_max_smem<igemm_tc2_kernel<BN, KIND, SLAB>>(smem, "sg_igemm(tc2)")) return rc;
  const int64_t mt2 = cdiv(g.M, 2 * BM);
  const int64_t tiles = mt2 * g.n_tiles;
  if (tiles >= (1ll << 31)) {
    set_error("sg_igemm(tc2): too many tiles");
    return SG_ERR_ARG;
  }
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  launch_k(igemm_tc2_kernel<BN, KIND, SLAB>, dim3(grid), dim3(V2<BN, SLAB>::THREADS), smem, stream, tmA, tmB, tmA_lo, tmB_lo,
           g, ep, (int)tiles);
  return launch_status("sg_igemm(tc2)");
}

}  // namespace tc

// A operand maps of one activation tensor (hi or lo part): NHWC rank-4 {C, W, H, rows} for the 3x3 convs, {Cin, M, 1, 1}
// for Linear layers; box = one 128-byte channel span x 128 pixels.
static int make_a_map(CUtensorMap* tm, const sg_igemm_args* a, const void* base, int dtype, int esz, int bke, int64_t M,
                      uint32_t* box_rows, bool slab = false) {
  using tc::make_tmap;
  const uint64_t cb = (uint64_t)a->Cin * esz;
  if (slab) {  // the 256-pixel tile's image rows + one halo row above and below, full width (shifted by dx at load time)
    const uint32_t th = 256u / (uint32_t)a->W + 2u;
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->rows};
    const uint64_t strides[3] = {cb, cb * a->W, cb * a->W * a->H};
    const uint32_t box[4] = {(uint32_t)bke, (uint32_t)a->W, th, 1};
    *box_rows = (uint32_t)a->W * th;
    return make_tmap(tm, dtype, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (a->taps == 1) {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)M, 1, 1};
    const uint64_t strides[3] = {cb, cb * M, cb * M};
    const uint32_t box[4] = {(uint32_t)bke, (uint32_t)(M < 128 ? M : 128), 1, 1};
    *box_rows = box[1];
    return make_tmap(tm, dtype, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  const uint32_t bw = a->W < 128 ? a->W : 128;
  const uint32_t bh = (uint32_t)a->H < 128 / bw ? a->H : 128 / bw;
  uint32_t bn = 128 / (bw * bh);
  if (bn > (uint32_t)a->rows) bn = a->rows;
  const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->rows};
  const uint64_t strides[3] = {cb, cb * a->W, cb * a->W * a->H};
  const uint32_t box[4] = {(uint32_t)bke, bw, bh, bn};
  *box_rows = bw * bh * bn;
  return make_tmap(tm, dtype, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int igemm_tc(const sg_igemm_args* a, cudaStream_t stream) {
  using namespace tc;
  const bool tf32 = a->act_dtype == SG_F32;
  SG_REQUIRE(tf32 || a->act_dtype == SG_BF16 || a->act_dtype == SG_F16, "sg_igemm(tc): act_dtype %d", a->act_dtype);
  const int esz = tf32 ? 4 : 2, bke = 128 / esz;  // one k-block = one 128-byte swizzle span of channels
  if (tf32) {
    SG_REQUIRE(a->a_lo && a->w_lo, "sg_igemm(tc, fp32): the split-tf32 engine needs a_lo and w_lo (sg_split_tf32)");
    SG_REQUIRE(a->out_dtype == 0, "sg_igemm(tc, fp32): outputs are fp32");
    SG_REQUIRE(((reinterpret_cast<uintptr_t>(a->a_lo) | reinterpret_cast<uintptr_t>(a->w_lo)) & 15) == 0,
               "sg_igemm(tc): operands must be 16-byte aligned");
  } else {
    SG_REQUIRE(a->out_dtype == 0 || a->out_dtype == SG_BF16 || a->out_dtype == SG_F16, "sg_igemm(tc): out_dtype %d", a->out_dtype);
  }
  SG_REQUIRE(a->Cin % bke == 0, "sg_igemm(tc): Cin=%d %% %d != 0", a->Cin, bke);
  SG_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0,
             "sg_igemm(tc): operands must be 16-byte aligned");
  const int BN = (a->Cout % 128 == 0) ? 128 : 64;
  const int HW = a->H * a->W;
  IgemmGeom g;
  g.M = (int64_t)a->rows * HW;
  g.H = a->H; g.W = a->W; g.HW = HW;
  g.Cout = a->Cout; g.taps = a->taps; g.cblocks = a->Cin / bke;
  g.bke = bke;
  g.passes = tf32 ? 3 : 1;
  g.n_tiles = a->Cout / BN;
  g.tiles_per_sample = HW >= 128 ? HW / 128 : 0;
  g.samples_per_tile = HW >= 128 ? 0 : 128 / HW;
  g.idesc = tf32 ? make_idesc_tf32(128, BN, 0) : make_idesc(a->act_dtype, 128, BN, 0, 0);
  SG_REQUIRE(g.M < (1ll << 31), "sg_igemm(tc): M too large");
  SG_REQUIRE(a->Cout <= MAX_COUT, "sg_igemm(tc): Cout=%d > %d", a->Cout, MAX_COUT);

  CUtensorMap tmA, tmB, tmA_lo, tmB_lo;
  uint32_t box_rows = 0;
  int rc;
  // Cout = 64 3x3 layers on 64- / 32-pixel-wide images: the slab pipeline (see V2)
  const bool slab = !tf32 && BN == 64 && a->taps == 9 && (a->W == 64 || a->W == 32) && a->H % (256 / a->W) == 0 &&
                    HW % 256 == 0;
  if ((rc = make_a_map(&tmA, a, a->a, a->act_dtype, esz, bke, g.M, &box_rows, slab))) return rc;
  const uint64_t wdims[2] = {(uint64_t)a->Cin, (uint64_t)a->taps * a->Cout};
  const uint64_t wstrides[1] = {(uint64_t)a->Cin * esz};
  const uint32_t wbox[2] = {(uint32_t)bke, (uint32_t)BN};
  if ((rc = make_tmap(&tmB, a->act_dtype, 2, a->w, wdims, wstrides, wbox, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if (tf32) {
    if ((rc = make_a_map(&tmA_lo, a, a->a_lo, a->act_dtype, esz, bke, g.M, &box_rows))) return rc;
    if ((rc = make_tmap(&tmB_lo, a->act_dtype, 2, a->w_lo, wdims, wstrides, wbox, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  } else {
    tmA_lo = tmA;
    tmB_lo = tmB;
  }
  const bool resident = slab && a->Cin == tc::BK;  // the 64 -> 64 layers: nine 8 KB weight tiles stay in shared memory
  g.tx_bytes = slab ? box_rows * 128u + (resident ? 0u : 3u * (uint32_t)BN * 128u) : 2u * box_rows * 128u + (uint32_t)BN * 128u;

  IgemmEpi ep;
  ep.bias = a->bias; ep.residual = a->residual; ep.out_f32 = a->out_f32; ep.out_act = a->out_act;
  ep.partials = a->partials; ep.act = a->act; ep.act_dtype = a->out_dtype ? a->out_dtype : a->act_dtype;
  if (tf32) {  // fp32 engine: a single fp32 output (callers may pass it in either slot, like the SIMT engine)
    if (!ep.out_f32) ep.out_f32 = reinterpret_cast<float*>(a->out_act);
    ep.out_act = nullptr;
  }
  ep.P = sg_igemm_partials(SG_ENGINE_TC, a->H, a->W, a->Cout);
  if (tf32) return BN == 128 ? launch2<128, 1>(tmA, tmB, tmA_lo, tmB_lo, g, ep, stream) : launch2<64, 1>(tmA, tmB, tmA_lo, tmB_lo, g, ep, stream);
  if (resident) return launch2<64, 0, 2>(tmA, tmB, tmA_lo, tmB_lo, g, ep, stream);
  if (slab) return launch2<64, 0, 1>(tmA, tmB, tmA_lo, tmB_lo, g, ep, stream);
  return BN == 128 ? launch2<128, 0>(tmA, tmB, tmA_lo, tmB_lo, g, ep, stream) : launch2<64, 0>(tmA, tmB, tmA_lo, tmB_lo, g, ep, stream);
}

}  // namespace sg
