/*
 * sgb200.h -- C ABI of the B200-native CFG-DDPM sampling kernels (libsgb200.so).
 *
 * The reference (gibbona1/SpectrogramGenAI) has no FFI layer: its hot path is Python calling
 * torch.nn modules.  Each entry point below therefore cites the torch
 * call site(s) in /root/reference/src/diff_modules.py that it replaces.  INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch owns all memory), with two documented exceptions
 *     (the <= 9 KB weights of sg_conv_in and of the fused output conv are HOST pointers: they travel BY VALUE as kernel
 *     launch parameters); kernels never allocate, free or synchronise;
 *   - the library keeps no mutable device state: every launch carries its own operands, so calls on different streams,
 *     for different models (model and ema_model) or on different devices of one process are independent.  Host-side
 *     caches are read-only after first use and keyed by device (function attributes, SM count);
 *   - activations are channels-last: [rows, H, W, C] ("NHWC"), rows = batch rows
 *     (2n when the conditional and unconditional passes are batched).  H and W are powers
 *     of two;
 *   - "act" tensors are SG_BF16 / SG_F16 (tcgen05 engine: 16-bit operands, fp32 accumulation in TMEM) or SG_F32:
 *     with SG_ENGINE_TC the fp32-accurate tensor-core engine (split-TF32 operands x = hi + lo, three kind::tf32 MMAs per
 *     product into one fp32 TMEM accumulator), with SG_ENGINE_SIMT the CUDA-core comparator kernels;
 *   - every function returns 0 on success, non-zero on error; sg_last_error() returns a
 *     thread-local human-readable reason.  No exceptions, no CPU fallback: a non-sm_100
 *     device is an error (SG_ERR_ARCH);
 *   - `stream` is a cudaStream_t passed as void*; all launches are capturable in a CUDA graph.
 */
#ifndef SGB200_H
#define SGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 15

typedef enum { SG_F32 = 0, SG_BF16 = 1, SG_F16 = 2 } sg_dtype;
typedef enum { SG_ENGINE_SIMT = 0, SG_ENGINE_TC = 1 } sg_engine;
/* sg_igemm epilogue activation: GELU is applied after the bias and BEFORE the residual add (ff_self, :61);
 * RELU_POST after the residual add (the VQAE decoder's relu(conv(x) + x), :343-348). */
typedef enum { SG_ACT_NONE = 0, SG_ACT_GELU = 1, SG_ACT_RELU_POST = 2 } sg_act;

enum {
  SG_OK = 0,
  SG_ERR_ARG = 1,    /* bad shape / null pointer / unsupported combination */
  SG_ERR_ARCH = 2,   /* device is not sm_100 (B200)                        */
  SG_ERR_LAUNCH = 3, /* CUDA launch or driver error                        */
};

typedef void* sg_stream_t;

int sg_abi_version(void);
/* Programmatic dependent launch of the step kernels issued by the CALLING THREAD from now on: 0 = off (every launch fully
 * serialised), 1 = every launch, 2 (default) = only grids of at most 4 x #SM CTAs.  Thread-local launch policy, like the
 * current device; it affects launch overlap only, never results.  Measured: mode 2 lowers the latency of a CFG step by
 * 5 % at batch 1-8 and costs 0.6 % at batch 512, so the Python plan selects 2 for small batches and 0 for large ones. */
int sg_set_pdl(int mode);
const char* sg_last_error(void);
/* 0 iff `device` is a compute-capability 10.0 device (B200: the library ships an sm_100a image only). */
int sg_device_check(int device);
/* Make `device` current for this library's CUDA runtime on the calling thread (one rank = one GPU). */
int sg_set_device(int device);

/* ---- K5: sinusoidal timestep encoding + label embedding + the six emb_layer projections ----
 * replaces UNet.pos_encoding (:168-173), label_emb add (:214-215) and emb_layer of Down/Up
 * (:105-108,:112 / :126-129,:135).
 *   t        fp32 [rows] timesteps (the reference promotes long -> fp32 in the multiply at :171),
 *            or NULL to use (float)*step (device int32) for every row
 *   y        int64 [rows] class ids; a negative id means "unconditional row" (y=None, :427);
 *            NULL = all rows unconditional
 *   inv_freq fp32 [128]    1/10000^(2k/256), built by the caller exactly as :169 does
 *   label    fp32 [num_classes, 256] (may be NULL iff every y < 0)
 *   w_emb    fp32 [emb_total, 256], b_emb fp32 [emb_total]: emb_layer.1 of down1..3, up1..3 concatenated
 *   temb     fp32 [rows, 256] out (encoding + label row);  emb fp32 [rows, emb_total] out
 */
int sg_time_embed(const float* t, const int32_t* step, const int64_t* y, const float* inv_freq,
                  const float* label, int num_classes, const float* w_emb, const float* b_emb,
                  int emb_total, int rows, float* temb, float* emb, sg_stream_t stream);

/* ---- inc.double_conv.0: 3x3 conv, c_in in {1..4}, NCHW fp32 input -> raw fp32 NHWC [rows,S,S,64] ----
 * replaces nn.Conv2d(c_in, 64, 3, padding=1, bias=False) (:82 via :144).  Row r reads sample
 * r % n_src of x (the cond/uncond halves share x).  Also emits GroupNorm partial sums:
 * partials fp32 [rows, P, 2] (sum, sum of squares), P = sg_conv_in_partials(S).
 * `w` is a HOST pointer: the 64 x c_in x 9 weights (<= 9216 B) are passed by value as a kernel parameter, so every FMA
 * reads its weight as a constant-bank operand and each launch owns its copy (no global bank, no repack launch).
 */
int sg_conv_in_partials(int S);
int sg_conv_in(const float* x, int n_src, int c_in, int S, const float* w /* HOST [64,c_in,3,3] */, int rows,
               void* raw, int raw_dtype /* SG_F32 or SG_F16 (saturating) */, float* partials, sg_stream_t stream);

/* ---- K1: implicit-GEMM 3x3 (taps=9, pad 1) or 1x1 (taps=1; Linear) over NHWC activations ----
 * replaces nn.Conv2d(.,.,3,padding=1,bias=False) (:82,:85) and the Linear layers of
 * SelfAttention (in_proj/out_proj inside nn.MultiheadAttention :56,:69; ff_self :60,:62).
 * out[m, co] = sum_{tap, ci} a[pixel(m) + tap offset, ci] * w[tap, co, ci]; epilogue, in order:
 * + bias[co]; GELU(erf) if gelu; + residual[m, co].  M = rows*H*W.
 * SG_ENGINE_SIMT: fp32 CUDA-core kernel (act_dtype = SG_F32).  SG_ENGINE_TC: tcgen05/TMEM kernel
 * fed by TMA: act_dtype = SG_BF16 or SG_F16 (fp32 accumulate), or act_dtype = SG_F32 = the fp32-accurate engine on
 * split-TF32 operands: a / w hold the hi parts (a: the fp32 activation itself), a_lo / w_lo the lo parts (sg_split_tf32), and the kernel accumulates
 * a_lo w_hi + a_hi w_lo + a_hi w_hi in one fp32 TMEM tile (a single TF32 pass: 3e-4 per conv; split: the tensor core's
 * truncating fp32 accumulation remains, a uniform relative shrink of ~1e-8 K that the GroupNorm behind every conv removes).
 * Requires Cin % 64 == 0 (TC 16-bit) / % 32 (TC fp32) / % 16 (SIMT), Cout % 64 == 0.
 */
typedef struct {
  const void* a;         /* act  [rows, H, W, Cin]                              */
  const void* w;         /* act  [taps, Cout, Cin]  (packed by the caller)      */
  const float* bias;     /* fp32 [Cout] or NULL                                 */
  const float* residual; /* fp32 [M, Cout] or NULL                              */
  float* out_f32;        /* fp32 [M, Cout] or NULL                              */
  void* out_act;         /* act  [M, Cout] or NULL                              */
  float* partials;       /* fp32 [rows, P, 2] GroupNorm partial sums or NULL    */
  const void* a_lo;      /* fp32 engine on tensor cores: lo part of a (same shape), else NULL */
  const void* w_lo;      /* fp32 engine on tensor cores: lo part of w (same shape), else NULL */
  int32_t rows, H, W, Cin, Cout, taps;
  int32_t act;           /* sg_act: applied as described below                  */
  int32_t engine;        /* sg_engine                                           */
  int32_t act_dtype;     /* sg_dtype of a, w and (unless out_dtype says otherwise) out_act */
  int32_t out_dtype;     /* 0: out_act has act_dtype.  SG_F16 / SG_BF16: out_act is stored in that 16-bit type
                            whatever the operand type (tcgen05 engine only) -- used to keep the raw conv output that
                            feeds GroupNorm in fp16 (saturating, 11-bit mantissa) instead of fp32              */
} sg_igemm_args;
int sg_igemm_partials(int engine, int H, int W, int Cout); /* P for the given geometry */
int sg_igemm(const sg_igemm_args* args, sg_stream_t stream);

/* ---- K2: GroupNorm(1, C) finalize + affine (+GELU | +residual,GELU | +time-embedding) ----
 * replaces nn.GroupNorm(1,C) + nn.GELU (:83-84,:86), F.gelu(x + double_conv(x)) (:91) and
 * the broadcast "x + emb" of Down/Up (:113,:136).  mode: 0 = affine only, 1 = GELU(affine),
 * 2 = GELU(residual + affine) (residual fp32 [rows,HW,C] required).  emb (fp32, row stride
 * emb_stride, already offset to this layer's slice) is added last when non-NULL.
 * raw / partials hold raw_rows rows and output row r normalises raw row r % raw_rows (raw_rows == rows, or the
 * label-independent prefix of the UNet computed once for the conditional and unconditional halves, raw_rows == rows/2:
 * inc and down1's convolutions see no embedding, :187-188, :110-113).
 * raw is fp32 (raw_dtype SG_F32) or fp16 (SG_F16, written by sg_igemm with out_dtype = SG_F16); the statistics in
 * `partials` always come from the fp32 accumulators.
 * range_flag (device int32, may be NULL): GroupNorm is scale invariant, fp16 is not.  With an fp16 raw tensor the kernel
 * sets *range_flag = 1 when the row's mean square (known exactly from the fp32 statistics) lies outside [2^-20, 2^20],
 * i.e. when values could have saturated at 65504 or lost precision to fp16 subnormals; the caller then re-runs with
 * fp32 raw tensors.  Never cleared by the library.
 */
int sg_gn_apply(const void* raw, int raw_dtype, const float* partials, int P, const float* gamma, const float* beta,
                int rows, int raw_rows, int HW, int C, int mode, const float* residual, const float* emb, int emb_stride,
                float* out_f32, void* out_act, int act_dtype, int32_t* range_flag, sg_stream_t stream);

/* ---- K2 + K3b: GELU(GroupNorm(raw) + cat([skip, upsample2x(x)])) -> act, residual recomputed from its sources ----
 * The first DoubleConv of Up is residual (:88-91 with :132-134): its input -- the concatenation of the skip and the
 * bilinearly upsampled x -- is added back after the second GroupNorm.  Instead of keeping an fp32 copy of that
 * concatenation (written by sg_upsample_cat, read here), this variant of sg_gn_apply mode 2 recomputes it from x fp32
 * [rows,h,w,Cx] and skip fp32 [skip_rows,2h,2w,Cs] (row r reads skip row r % skip_rows).  raw is fp16
 * [rows,2h,2w,Cs+Cx] (sg_igemm out_dtype = SG_F16), out act of the same shape; Cx, Cs multiples of 8.
 */
int sg_gn_apply_vcat(const void* raw, const float* partials, int P, const float* gamma, const float* beta, int rows,
                     const float* x, const float* skip, int skip_rows, int h, int w, int Cx, int Cs, void* out_act,
                     int act_dtype, int32_t* range_flag /* as sg_gn_apply */, sg_stream_t stream);

/* ---- K3a: MaxPool2d(2) (:100).  in fp32 [rows,H,W,C] -> fp32 and/or act [rows,H/2,W/2,C] ---- */
int sg_maxpool2(const float* in, int rows, int H, int W, int C, float* out_f32, void* out_act, int act_dtype,
                sg_stream_t stream);

/* ---- K3b: Upsample(x2, bilinear, align_corners=True) + cat([skip, x], dim=1) (:120,:132-133) ----
 * x fp32 [rows,h,w,Cx], skip fp32 [skip_rows,2h,2w,Cs] -> [rows,2h,2w,Cs+Cx] (skip channels first); output row r
 * reads skip row r % skip_rows (skip_rows == rows, or the shared label-independent `inc` output, rows/2).
 */
int sg_upsample_cat(const float* x, const float* skip, int rows, int skip_rows, int h, int w, int Cx, int Cs,
                    float* out_f32, void* out_act, int act_dtype, sg_stream_t stream);

/* ---- LayerNorm over C (:57 self.ln, :59 ff_self.0).  in fp32 [M,C] -> act [M,C]; C in {64,128,256} ---- */
int sg_layernorm(const float* in, const float* gamma, const float* beta, int64_t M, int C, void* out_act,
                 int act_dtype, sg_stream_t stream);

/* ---- SelfAttention head, fused (tcgen05 engine, C = 64 or 128): qkv = LayerNorm(x) Win^T + bin ----
 * replaces self.ln (:57,:67) + the in_proj of nn.MultiheadAttention (:56,:69).
 * x fp32 [M,C] (the residual stream, tokens row-major) -> qkv act [M,3C].  w_in act [3C,C] (torch layout),
 * b_in fp32 [3C].  One pass over x, one over qkv; the LayerNorm of a token is register math of the thread that
 * owns its accumulator row.  C != 64 is SG_ERR_ARG: callers use sg_layernorm + sg_igemm there.
 */
int sg_ln_inproj(const float* x, const float* ln_g, const float* ln_b, const void* w_in, const float* b_in, int64_t M,
                 int C, void* qkv, int act_dtype, sg_stream_t stream);

/* ---- SelfAttention tail, fused (tcgen05 engine, C = 64 or 128) ----
 * replaces out_proj + residual (:69-70), ff_self = LayerNorm -> Linear -> GELU -> Linear (:58-63) and the second
 * residual (:71):   a = att Wo^T + bo + x;   out = GELU(LayerNorm(a) W1^T + b1) W2^T + b2 + a.
 * att act [M,C] (attention core output), x fp32 [M,C], wo/w1/w2 act [C,C] (torch layout: [out, in]),
 * biases and LayerNorm affine fp32 [C]; out fp32 [M,C] (may alias x: a tile is read completely before it is written).
 * Three chained tcgen05 GEMMs per 128-token tile, operands handed from epilogue to GEMM through shared memory.
 */
int sg_attn_tail(const void* att, const float* x, const void* wo, const float* bo, const float* ln_g, const float* ln_b,
                 const void* w1, const float* b1, const void* w2, const float* b2, int64_t M, int C, float* out,
                 int act_dtype, sg_stream_t stream);

/* sg_attn_tail with the model's output conv fused behind it (outc, :166,:195: 1x1 conv C -> c_out + bias on the last
 * SelfAttention block's output): eps fp32 NCHW [M/HW, c_out, HW] = outc_b + out . outc_w^T, computed from the fp32
 * `out` row each thread already holds, so the [M,C] fp32 tensor is neither written nor re-read (sa6 at n = 512: 2.1 GB).
 * outc_w fp32 [c_out, C], outc_b fp32 [c_out] are HOST pointers (1 KB, passed by value as kernel parameters: constant-bank
 * FMA operands owned by the launch), c_out in 1..4, C = 64, HW = tokens per sample (power of two >= 128).
 * out may be NULL (only eps is produced) or a fp32 [M,C] buffer that also receives the block output.
 */
int sg_attn_tail_outc(const void* att, const float* x, const void* wo, const float* bo, const float* ln_g,
                      const float* ln_b, const void* w1, const float* b1, const void* w2, const float* b2, int64_t M,
                      int C, float* out, const float* outc_w, const float* outc_b, int c_out, int HW, float* eps,
                      int act_dtype, sg_stream_t stream);

/* ---- K4: multi-head self-attention core, never materialising the L x L matrix ----
 * replaces the scaled-dot-product inside nn.MultiheadAttention (:56,:69): per row and head,
 * softmax(q k^T / sqrt(d)) v.  qkv act [rows*L, 3C] = in_proj output (q | k | v, head h at
 * columns [h*d,(h+1)*d) of each third); out [rows*L, C].  d = C/heads in {16,32,64}.
 * SG_ENGINE_SIMT: qkv is fp32, out has dtype act_dtype (fp32 or 16-bit).  SG_ENGINE_TC: qkv and out are
 * both act_dtype (SG_BF16 / SG_F16).
 */
int sg_attention(const void* qkv, void* out, int rows, int L, int C, int heads, int engine, int act_dtype,
                 sg_stream_t stream);
/* K4 of the fp32-accurate tensor-core engine: the same product with split-TF32 operands, for L >= 128 (power of two).
 * sg_attn_prep_tf32: qkv fp32 [rows*L, 3C] (the in_proj output) -> qk_hi / qk_lo fp32 [rows*L, 2C] (split of q | k) and
 *   vt_hi / vt_lo fp32 [rows*heads*d, L] = the split of V transposed per (batch row, head), so that V^T is a K-major
 *   B operand (kind::tf32 accepts MN-major operands only in a dedicated swizzle).
 * sg_attention_tf32: out fp32 [rows*L, C]; S = Q K^T and O = P V are each three kind::tf32 MMAs (lo hi + hi lo + hi hi)
 *   into one fp32 TMEM accumulator; the probabilities are split in the kernel. */
int sg_attn_prep_tf32(const float* qkv, float* qk_hi, float* qk_lo, float* vt_hi, float* vt_lo, int rows, int L, int C,
                      int heads, sg_stream_t stream);
int sg_attention_tf32(const float* qk_hi, const float* qk_lo, const float* vt_hi, const float* vt_lo, float* out, int rows,
                      int L, int C, int heads, sg_stream_t stream);
/* Operands of the fp32-accurate tensor-core engine.  kind::tf32 multiplies the top 19 bits of every fp32 container, so
 * an fp32 tensor x is used as two operands whose sum keeps ~22 mantissa bits:
 *   hi != NULL (weights, packed once):  hi = tf32(x) (round to nearest), lo = tf32(x - hi);
 *   hi == NULL (activations):           x ITSELF is the high operand (the tensor core reads trunc19(x)) and only
 *                                       lo = tf32(x - trunc19(x)) is written.
 * The activation form needs no pass of its own when the kernel that produces x writes lo next to it: sg_gn_apply (fp32 raw),
 * sg_maxpool2 and sg_upsample_cat do so when called with act_dtype = SG_F32 and out_act = the lo tensor.
 * n % 4 == 0, buffers 16-byte aligned. */
int sg_split_tf32(const float* x, float* hi, float* lo, int64_t n, sg_stream_t stream);

/* ---- outc: 1x1 conv 64 -> c_out with bias, NHWC fp32 in, NCHW fp32 out (:166,:195) ---- */
int sg_conv_out(const float* in, const float* w /*[c_out,64]*/, const float* b, int rows, int HW, int c_out,
                float* eps, sg_stream_t stream);

/* ---- K6: CFG lerp + posterior update, one kernel (:426-439) ----
 * x fp32 [n, E] in/out (E = c*S*S).  eps fp32 [2n, E]: rows [0,n) conditional, rows [n,2n)
 * unconditional; with cfg_scale <= 0 only rows [0,n) are read (the single-forward branch, :426).
 * coef fp32 [T,3] = (1/sqrt(alpha), (1-alpha)/sqrt(1-alpha_hat), sqrt(beta)) built by the caller
 * with the reference's own expressions.  i = *step (device int32) is the current timestep.
 * Noise z: `noise` != NULL -> injected fp32 [T-1, n, E], z = noise[T - i] (index 0 is x_T);
 * else Philox4x32-10(seed; sample_base + sample, i, element) + Box-Muller.  z = 0 at i == 1 (:434-435).
 * Arithmetic uses un-fused fp32 multiplies/adds in the reference's evaluation order, so for equal
 * eps inputs and injected noise the update is bit-identical to the CPU reference.
 */
int sg_cfg_update(float* x, const float* eps, int n, int E, float cfg_scale, const float* coef, int T,
                  const int32_t* step, const float* noise, uint64_t seed, int64_t sample_base, sg_stream_t stream);
/* *step -= 1 (one thread); the last node of a captured sampling step. */
int sg_step_advance(int32_t* step, sg_stream_t stream);
/* x_T from the same Philox stream (throughput mode): x[sample, e] = N(0,1)(seed; sample_base+sample, step_tag, e). */
int sg_philox_normal(float* x, int n, int E, uint64_t seed, int64_t sample_base, int step_tag, sg_stream_t stream);

/* ---- K8: (clamp(x,-1,1)+1)/2*255 -> truncating uint8 cast (:440-441) ---- */
int sg_to_uint8(const float* x, int64_t count, uint8_t* out, sg_stream_t stream);

/* ---- forward-only training helpers (no backward pass) ----
 * sg_noise_images: Diffusion.noise_images (:404-409).  x fp32 [n,E], t int64 [n] (index into alpha_hat fp32 [T]):
 *   x_t = sqrt(alpha_hat[t]) * x + sqrt(1 - alpha_hat[t]) * eps, un-fused products / sum like the reference, so equal
 *   eps gives bit-identical x_t.  eps_in fp32 [n,E], or NULL to draw N(0,1) from the Philox stream keyed by
 *   (seed, sample_base + sample); eps_out (may be NULL) receives the eps used -- the method returns it.
 * sg_ema_update: EMA.update_average (:37-40) in place: ma = ma * beta + one_minus_beta * cur (the reference passes the
 *   Python float 1 - beta; both scalars are applied in fp32).
 * sg_mse: nn.MSELoss() (:478) of two fp32 buffers -> *out (fp32 scalar); scratch = sg_mse_scratch_doubles() doubles;
 *   deterministic (fixed partition, double partial sums).
 */
int sg_noise_images(const float* x, const int64_t* t, const float* alpha_hat, int T, int n, int E, const float* eps_in,
                    uint64_t seed, int64_t sample_base, float* x_t, float* eps_out, sg_stream_t stream);
int sg_ema_update(float* ma, const float* cur, int64_t n, float beta, float one_minus_beta, sg_stream_t stream);
int sg_mse_scratch_doubles(void);
int sg_mse(const float* a, const float* b, int64_t n, double* scratch, float* out, sg_stream_t stream);

/* ---- weight repack, once per load_state_dict (UNet_conditional.load_state_dict / Diffusion.load :509-510) ----
 * w: a Conv2d / Linear weight of the reference state_dict, fp32 [Cout, Cin, taps] (taps = kh*kw: 9 for the 3x3 convs,
 * 1 for Linear / 1x1) -> out [taps, Cout, Cin] in out_dtype (SG_F32 / SG_BF16 / SG_F16, round to nearest even): the
 * `w` operand of sg_igemm.  Kernel-owned scratch does not exist: the only sizes a host must ask for are the GroupNorm
 * partial counts (sg_igemm_partials, sg_conv_in_partials); every other buffer has the shape its entry point states. */
int sg_pack_weights(const float* w, int Cout, int Cin, int taps, void* out, int out_dtype, sg_stream_t stream);

/* ---- un-clamped image cast of the denoise-trajectory dumps (DiffusionVAE.sample :672-675) ----
 * out = uint8((x + 1) / 2 * 255) with torch's cast semantics for values outside [0, 255]: truncate to int32, keep the
 * low 8 bits. */
int sg_to_uint8_wrap(const float* x, int64_t count, uint8_t* out, sg_stream_t stream);

/* ---- DiffusionVAE decode tail (:702-706) ----
 * sg_vq_quantize: [clamp(-1,1) +] VQEmbeddingEMA.forward in eval mode (:290-318) over `count` fp32 values taken in
 *   groups of 4 consecutive elements (the reference's reshape(-1, 4) of the NCHW latent); codebook fp32 [n_codes, 4];
 *   quantized = x + (q - x) exactly as :313 writes it; indices int32 [count/4] or NULL.
 * sg_dec_in_proj: Decoder.in_proj (:330): 1x1 conv 4 -> Cout + bias; z fp32 NCHW [n,4,S,S], w fp32 [Cout,4],
 *   out NHWC [n,S,S,Cout] as fp32 and/or act.
 * sg_tconv2_u8: Decoder.strided_t_conv_2 (:336) + image tail (:704-705).  t = output of the first ConvTranspose2d
 *   computed by sg_igemm as a Linear and left un-shuffled: [2 (a)][n*S*S][2 (b) * C] of dtype t_dtype (fp32 or 16-bit),
 *   i.e. pixel (2h+a, 2w+b) of the 2S x 2S map; w2 fp32 [C,1,2,2], b2 fp32 [1].  out_u8 uint8 [n,1,4S,4S] =
 *   ((y+1)/2*255) truncated to int32 and reduced mod 256 (the un-clamped cast of the reference), out_f32 = y (either
 *   may be NULL).
 */
int sg_vq_quantize(const float* x, int64_t count, const float* codebook, int n_codes, int clamp, float* quantized,
                   int32_t* indices, sg_stream_t stream);
int sg_dec_in_proj(const float* z, const float* w, const float* b, int n, int S, int Cout, float* out_f32, void* out_act,
                   int act_dtype, sg_stream_t stream);
int sg_tconv2_u8(const void* t, int t_dtype, int n, int S, int C, const float* w2, const float* b2, uint8_t* out_u8,
                 float* out_f32, sg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H */
