// Shared pieces of the tcgen05 attention kernels (attention_tc.cu: v1 / v8 / v11, attention_tc12.cu: v12).
#pragma once

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
  uint32_t idesc_s, idesc_o, idesc_ol;  // idesc_ol: v12, N = d + 16 (O | row sums)
  int act_dtype;
  float redo_log2;  // largest tolerated (tile max - reference max) * c before the tile is recomputed
  float l_max;      // v12: largest tolerated row sum of the fast pass
};


// packed fp32 arithmetic of sm_100 (two lanes per instruction: half the issue slots of the scalar forms), the
// three-input maximum, and the FMA-pipe exp2 used for a fraction of the softmax exponentials
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// exp2 of two scaled exponents on the FMA / ALU pipes (same polynomial as ex2_poly, two lanes per instruction)
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  un2(x2, x0, x1);
  x2 = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
  const uint64_t magic = pk2(12582912.0f, 12582912.0f);
  const uint64_t t2 = add2(x2, magic);
  const uint64_t f2 = sub2(x2, sub2(t2, magic));
  uint64_t p = fma2(f2, pk2(0.05517145f, 0.05517145f), pk2(0.24261084f, 0.24261084f));
  p = fma2(p, f2, pk2(0.69326097f, 0.69326097f));
  p = fma2(p, f2, pk2(0.99992812f, 0.99992812f));
  float t0, t1, q0, q1;
  un2(t2, t0, t1);
  un2(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// which of the 8 pairs of a 16-column chunk take the polynomial: POLY of 8, spread out so MUFU and FMA work interleave
template <int POLY>
__device__ __forceinline__ constexpr bool pair_is_poly(int pair) {
  return POLY == 0 ? false
         : POLY == 1 ? (pair & 7) == 5
         : POLY == 2 ? (pair & 3) == 3
         : POLY == 3 ? ((pair & 7) == 2 || (pair & 7) == 5 || (pair & 7) == 7)
                     : (pair & 1) == 1;
}

// barrier ids are immediates: a register id makes ptxas reserve all 16 named barriers for the CTA
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

}  // namespace tc
}  // namespace sg
