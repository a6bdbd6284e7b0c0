/*
 * b200diff.h -- C ABI of libb200diff.so: the B200 (sm_100a) kernels behind the diffusion hot path.
 *
 * The reference (xyfJASON/diffusion-models-pytorch) has no FFI of its own: its boundary is the Python
 * call surface `model(x, t, **kw)` / `DDPM.denoise(...)`.  Each entry point below therefore cites the
 * reference ATen call sequence it replaces (file:line under /root/reference).  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a CUDA device pointer owned by the caller (PyTorch); the library never allocates
 *     user-visible memory and never synchronises the stream;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - activations handed between kernels are NHWC ("pixel-major"): fp32 for the residual stream,
 *     bf16 for tensor-core operands; model inputs/outputs are the reference's NCHW fp32;
 *   - return value 0 = ok; otherwise a cudaError_t or >= 1000 for argument errors,
 *     with text from b200_last_error().
 */
#ifndef B200DIFF_H_
#define B200DIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200DIFF_VERSION 200
/* fixed-point scales of the GroupNorm statistics exchanged between kernels ([B][C][2] int64: sum, sum of squares) */
#define B200_STAT_Q1 1073741824.0 /* 2^30 */
#define B200_STAT_Q2 4194304.0    /* 2^22 */

int b200_version(void);
const char* b200_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
long long b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1: convolution as implicit GEMM on tcgen05/TMEM, operands fed by TMA.
 * Replaces nn.Conv2d 3x3/1x1 (models/unet.py:16,26,28,72,118; models/modules.py:64,72,83-86) plus the
 * adds fused around it: bias, time-embedding broadcast add (models/unet.py:41), residual / shortcut add
 * (models/unet.py:43, models/modules.py:100) and the nearest-2x of Upsample (models/modules.py:63).
 *
 * D[M = B*Ho*Wo pixels, N] = sum over K-blocks of A[M, 64] * W[N, 64]^T, where K runs over
 * (tap, 64-channel chunk) of source a0 and then over the 64-channel chunks of the optional source a1
 * (used to fold the 1x1 shortcut conv of a ResBlock into the GEMM of its second 3x3 conv).
 * A source is a bf16 tensor [B][planes][Hs][Ws][C]; the A rows of output pixel (n, ho, wo) for a tap
 * (dw, dh, plane) are read at (n, plane, ho + dh, wo + dw, :) with zero fill outside [0,Hs)x[0,Ws).
 * Stride-2 convs use 4 parity planes; plain convs use planes = 1.
 * phases = 4 runs the four 2x2-tap sub-convolutions that together equal "nearest-2x then 3x3"; phase p
 * writes output pixel (2*ho + (p>>1), 2*wo + (p&1)) and uses weight rows [p*w_rows_per_phase, ...).
 * --------------------------------------------------------------------------------------------- */
enum { B200_OUT_F32_NHWC = 0, B200_OUT_BF16_NHWC = 1, B200_OUT_F32_NCHW = 2, B200_OUT_BF16_NCHW = 3 };

typedef struct b200_conv_desc {
  const void* a0;      /* bf16 source 0 */
  int a0_C, a0_H, a0_W, a0_planes;
  const void* a1;      /* bf16 source 1 (single tap) or NULL */
  int a1_C, a1_H, a1_W, a1_planes;
  const void* w;       /* packed bf16 weights [w_rows][w_K], K contiguous: K = ntaps0*a0_C + a1_C */
  int w_rows, w_K, w_rows_per_phase;
  int B, Ho, Wo;       /* output-tile pixel grid */
  int phases;          /* 1 or 4 */
  int N;               /* output channels */
  int ntaps0;          /* taps of source 0 (1, 4 or 9) */
  int8_t taps0[4][9][4]; /* [phase][tap] = {dw, dh, plane, 0} */
  int8_t tap1[4];      /* {dw, dh, plane, 0} of source 1 */
  const float* bias;     /* [N] or NULL */
  const float* rowadd;   /* per-image additive row [B][rowadd_ld] (time embedding projection) or NULL */
  int rowadd_ld;
  const float* residual; /* fp32 NHWC [B][out_H][out_W][res_ld] or NULL */
  int res_ld;
  long long* stats;      /* optional [B][N][2] int64, pre-zeroed: per-(image, channel) sum (fixed point, 2^30) and sum of
                            squares (2^22) of the NHWC output (before any rounding to bf16) are accumulated here for the
                            GroupNorm that follows.  Integer atomics: the result does not depend on the arrival order
                            of the CTAs, so forwards are bitwise reproducible (B200_STAT_Q1 / B200_STAT_Q2 below) */
  void* out;
  int out_mode;        /* B200_OUT_* */
  int out_ld;          /* channel stride of NHWC outputs; for NCHW outputs the channel count */
  int out_H, out_W;    /* output image size in pixels */
  int osy, osx;        /* output pixel = (ho*osy + (phase>>1), wo*osx + (phase&1)) */
} b200_conv_desc;

int b200_conv2d_fwd(const b200_conv_desc* d, void* stream);

/* Convolution whose epilogue applies the GroupNorm (+ AdaGN scale/shift) (+ SiLU) that FOLLOWS it, i.e. conv1 -> norm2
 * of a ResBlock (models/unet.py:20-24,38-40; models/unet_categorial_adagn.py:36-41,58-60), writing only the bf16
 * NHWC operand of the next convolution: out_norm = SiLU(GN(conv(x) + bias + rowadd)).  Possible when one 256-pixel
 * tile holds whole images (Ho*Wo in {16, 64, 256}) and its 128 channels whole groups, so the statistics are complete
 * inside a CTA (tiles are chosen so that they hold whole images and are full: B must be a multiple of 64 / (Ho*Wo) for
 * 4x4 images), or when an image is 2 / 4 tiles and N == 128 (the tiles' CTAs are launched cooperatively and exchange
 * their sums through g->xstats / g->xcount).  N % 128 == 0; N / groups a power of two <= 32.
 * Two forms:
 *   d->out == NULL  (conv1 -> norm2): only out_norm is written; d->stats / d->residual must be NULL;
 *   d->out != NULL  (block output: conv2 -> the NEXT block's norm1, models/unet.py:30-43 across two ResBlocks, or the
 *                   last block -> the output head's norm, models/unet.py:116-117): x = conv + bias (+ d->residual) is
 *                   written to d->out as fp32 NHWC with its statistics in d->stats (multi-tile images: in g->xstats),
 *                   and GN(x) to out_norm; no rowadd / scale / shift.
 * The engine uses both for every eligible layer at inference (B200_FUSE_GN2=0 / B200_FUSE_GN1=0 restore two launches). */
typedef struct b200_gn_fuse_desc {
  const float* gamma;        /* [N] */
  const float* beta;         /* [N] */
  const float* scale;        /* optional per-image rows [B][ss_ld]: y = GN(x) * (1 + scale) + shift */
  const float* shift;
  void* out_norm;            /* bf16 NHWC [B][Ho][Wo][out_norm_ld] (first N columns) */
  int out_norm_ld;           /* channel stride of out_norm (and out_raw_bf16); 0 = N */
  void* out_raw_bf16;        /* optional: bf16 copy of the un-normalised x, same layout as out_norm (operand of the
                              * consumer's fused 1x1 shortcut when it concatenates a skip connection); with d->out == NULL
                              * (a block output nobody reads as fp32) no embedding row / scale / shift may be given */
  int ss_ld;
  int groups;
  int apply_silu;
  float eps;
  /* images of 512 / 1024 pixels (2 / 4 tiles per image; N == 128): zeroed workspaces through which the co-scheduled
   * CTAs of an image share their statistics -- xstats int64 [B][N][2] (afterwards it holds the statistics of the conv
   * output, like b200_conv_desc.stats), xcount int64 [B] arrival counters.  NULL for images of <= 256 pixels. */
  long long* xstats;
  long long* xcount;
} b200_gn_fuse_desc;
int b200_conv2d_gn_fwd(const b200_conv_desc* d, const b200_gn_fuse_desc* g, void* stream);

/* First convolution of the UNet (models/unet.py:72,123): NCHW fp32 image, tiny Cin (1..4), 3x3 s1 p1,
 * -> fp32 NHWC [B][H][W][Cout].  Weights are the reference's OIHW fp32 tensor as is. */
int b200_conv3x3_first(const float* x_nchw, const float* w_oihw, const float* bias, float* out_nhwc,
                       long long* stats /* optional [B][Cout][2] int64, pre-zeroed, as in b200_conv_desc.stats */,
                       int B, int Cin, int H, int W, int Cout, void* stream);

/* The same convolution on the tensor cores (inference): b200_first_split writes the image as bf16 NHWC [B][H][W][64]
 * pixels [hi(Cin) | lo(Cin) | hi(Cin) | 0 ...] (hi = bf16(x), lo = bf16(x - hi)); a 3x3 b200_conv2d_fwd /
 * b200_conv2d_gn_fwd over it with a0_C = 64 and weights packed by b200_pack_weights mode 4 (tap_ld = 64) evaluates
 * x*w as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation. */
int b200_first_split(const float* x_nchw, void* out_bf16_nhwc64, int B, int Cin, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3: GroupNorm (+ optional per-sample scale/shift = AdaGN) (+ optional SiLU), fp32 NHWC in, bf16 NHWC out.
 * Replaces nn.GroupNorm + nn.SiLU (models/unet.py:14-15,23-24,116-117; models/modules.py:82,105-123) and
 * the torch.cat of the skip connection (models/unet.py:145): two fp32 sources are normalised as one
 * concatenated tensor of C0 + C1 channels.  Statistics are exact two-pass fp32 over a slab staged in
 * shared memory.  raw_out (optional) receives the un-normalised concatenation rounded to bf16 (operand of
 * the 1x1 shortcut conv).  resample: 0 none, 1 = 2x2 average pool, 2 = nearest 2x, applied to the activated values
 * before the single rounding to bf16 (up/down ResBlocks, models/unet_categorial_adagn.py:52-57).
 * --------------------------------------------------------------------------------------------- */
int b200_groupnorm_silu_fwd(const float* x0, int C0, const float* x1, int C1, int B, int HW, int W, int groups,
                            const float* gamma, const float* beta, float eps, const float* scale,
                            const float* shift, int ss_ld, int apply_silu, int resample, void* out_bf16,
                            void* raw_out_bf16, void* stream);

/* Streaming variant of K3 for inputs whose per-(image, channel) statistics [B][C][2] = (sum, sum of squares), int64
 * fixed point with scales B200_STAT_Q1 / B200_STAT_Q2, were
 * accumulated by the producing kernel (b200_conv_desc.stats): one coalesced pass, 4 B read + 2 B written per
 * element.  Same semantics and arguments as b200_groupnorm_silu_fwd otherwise. */
int b200_groupnorm_apply_fwd(const void* x0, int x0_is_bf16 /* x0 is bf16 NHWC (statistics taken from the fp32
                             accumulators of its producer); with a second (fp32) source or a raw copy: C0 % 8 == 0,
                             C1 % 8 == 0, C <= 2048, no resampling */, int C0, const long long* stats0,
                             const float* x1, int C1, const long long* stats1, int B, int HW, int W, int groups,
                             const float* gamma,
                             const float* beta, float eps, const float* scale, const float* shift, int ss_ld,
                             int apply_silu, int resample, void* out_bf16, void* raw_out_bf16, void* stream);

/* fp32 NHWC -> bf16 NHWC, optionally split into the 4 parity planes [B][2*(h&1)+(w&1)][H/2][W/2][C] that
 * the stride-2 convolution (models/modules.py:72) reads through unit-stride TMA boxes. */
int b200_cast_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, int parity_split, void* stream);

/* 2x2 average pool / nearest 2x on fp32 NHWC (models/unet_categorial_adagn.py:26-28). */
int b200_avgpool2_f32(const float* x, float* out, int B, int H, int W, int C, void* stream);
int b200_upsample2_f32(const float* x, float* out, int B, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2: fused softmax(q k^T * scale) v for the low-resolution self-attention blocks
 * (models/modules.py:92-97).  T <= 256 keys, head dim d in {64,128,256}; S and O live in TMEM.
 * q, k: bf16 [B][T][ld_qk] at column offsets q_off + h*d / k_off + h*d; v transposed: bf16 [B][heads*d][T].
 * out: bf16 [B][T][ld_out], head h at columns [h*d, (h+1)*d).
 * --------------------------------------------------------------------------------------------- */
int b200_attention_fwd(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out,
                       int B, int T, int heads, int d, float scale, void* stream);

/* Training form of K2: additionally writes lse fp32 [B][heads][T] = log2(sum_j exp2(s_ij * scale * log2e)) per score
 * row (T <= 256, d in {64,128,256}); b200_attention_bwd recomputes softmax rows from it instead of storing them. */
int b200_attention_fwd_lse(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out,
                           int B, int T, int heads, int d, float scale, float* lse, void* stream);

/* Adjoint of K2 in ONE launch (autograd of models/modules.py:92-97: bmm, softmax, bmm), head dim 64, T <= 256:
 *     P = softmax(scale q k^T) (recomputed), dV = P^T dO, dP = dO V^T, dS = scale * P o (dP - rowsum(dO o O)),
 *     dQ = dS K, dK = dS^T Q.
 * qk / vt as in b200_attention_fwd with ld_qk = 2C, q_off = 0, k_off = C (C = heads*64); o, d_o: bf16 [B][T][C];
 * lse from b200_attention_fwd_lse; dqk: bf16 [B][T][ld_dqk] with [dQ | dK] in its first 2C columns; dv: bf16
 * [B][T][ld_dv] (both may point into one [B][T][3C] tensor: one weight-gradient GEMM for q, k, v).  One CTA per (image, head);
 * the [T][T] score tensors never leave the SM (the batched-GEMM form moved four of them through HBM). */
int b200_attention_bwd(const void* qk, const void* vt, const void* o, const void* d_o, const float* lse, void* dqk,
                       int ld_dqk, void* dv, int ld_dv, int B, int T, int heads, int d, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2b: the WHOLE self-attention block (models/modules.py:77-102, SelfAttentionBlock.forward) in one launch:
 *     out = x + proj(softmax(q k^T * scale) v),  q, k, v = 1x1 convs of GroupNorm(x)      (no SiLU in this norm)
 * for the CIFAR-10 UNet's 16x16 blocks: T = 256 tokens, C = 256 channels, ONE head, 32 groups.  One 2-CTA cluster
 * per image (tcgen05 cta_group::2); q, k, v, the scores and the attention output stay in shared memory / TMEM.
 *   x        fp32 NHWC [B][T][C] residual stream, x_stats its producer statistics [B][C][2] (int64 fixed point)
 *   w        bf16 [4C][C]: rows [0,C) = Wq, [C,2C) = Wk, [2C,3C) = Wv, [3C,4C) = Wproj, each [out][in] (K-major)
 *   bias     fp32 [4C] in the same order
 *   out      fp32 NHWC [B][T][C] (must not alias x), out_stats: statistics of out for the next GroupNorm (or NULL)
 *   dbg[6]   tests only: bf16 [B][256][256] dumps of xn, q, k (without its bias: a per-row shift of the scores, which
 *            softmax cancels), v^T, P (unnormalised), o; all NULL in production
 * Other shapes take the five-launch path (b200_groupnorm_apply_fwd, b200_conv2d_fwd x3, b200_attention_fwd).
 * --------------------------------------------------------------------------------------------- */
typedef struct b200_attn_block_desc {
  const float* x;
  const long long* x_stats;
  const float* gamma;
  const float* beta;
  const void* w;
  const float* bias;
  float* out;
  long long* out_stats;
  void* dbg[6];
  int B, T, C, heads, groups;
  float eps;
  float scale;
  int pad_;
} b200_attn_block_desc;
int b200_attn_block_fwd(const b200_attn_block_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Time embedding path (models/modules.py:40-57, models/unet.py:64-69,18-21).
 * b200_time_embed: t[rows] (int64) -> [sin(t f), cos(t f)] over the dim/2 frequencies `freqs` (cos first when
 *   cos_first, the ADM variant models/adm/nn.py:103-121) -> Linear(dim,E) -> SiLU -> Linear(E,E) (+ class
 *   embedding row y[b] when y != NULL) -> out fp32 [rows][E]; out_silu_bf16 (optional) = SiLU(out) as bf16, the
 *   operand of the per-ResBlock projections, which run on K1 as one 1x1 "conv" over all blocks.
 * --------------------------------------------------------------------------------------------- */
int b200_time_embed(const int64_t* t, int rows, const float* freqs, int dim, int E, int cos_first, const float* w1,
                    const float* b1, const float* w2, const float* b2, const int64_t* y, const float* class_embed,
                    float* out, void* out_silu_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4: fused sampler update (diffusions/ddpm.py:174-261, diffusions/ddim.py:57-86, CFG mix
 * diffusions/ddim.py:177-187 / ddpm.py:335-347).  One elementwise pass over NCHW fp32 tensors:
 *   [cond/uncond eps -> x0 -> clip -> eps] -> CFG mix -> x0 -> clip -> eps -> mean -> + sqrt(var)*noise.
 * coef points at one row of the per-step coefficient table in DEVICE memory (so a captured CUDA graph
 * can be replayed with a different row): see B200_SC_* indices.
 * --------------------------------------------------------------------------------------------- */
enum {
  B200_SC_SQRT_RECIP_AC = 0, /* (1/ac_t)^0.5 */
  B200_SC_SQRT_RECIPM1_AC,   /* (1/ac_t - 1)^0.5 */
  B200_SC_SQRT_AC,           /* ac_t^0.5 (pred_v) */
  B200_SC_SQRT_1M_AC,        /* (1-ac_t)^0.5 (pred_v) */
  B200_SC_X0_COEF,           /* DDPM: mean_coef1; DDIM: sqrt(ac_prev) */
  B200_SC_XT_COEF,           /* DDPM: mean_coef2; DDIM: 0 */
  B200_SC_EPS_COEF,          /* DDPM: 0; DDIM: sqrt(1-ac_prev-var) */
  B200_SC_VAR,               /* fixed variance (0 at t == 0) */
  B200_SC_MIN_LOGVAR,        /* learned_range */
  B200_SC_MAX_LOGVAR,        /* learned_range */
  B200_SC_ADD_NOISE,         /* 1.0 if t != 0 else 0.0 */
  B200_SC_COUNT = 12
};
enum { B200_OBJ_EPS = 0, B200_OBJ_X0 = 1, B200_OBJ_V = 2 };

typedef struct b200_sampler_desc {
  const float* model_out;    /* [B][Cm][HW], Cm = C or 2C (learned variance) */
  const float* model_out_uncond; /* CFG second branch or NULL */
  const float* xt;           /* [B][C][HW] */
  const float* noise;        /* [B][C][HW] or NULL (treated as 0) */
  const float* coef;         /* device pointer to B200_SC_COUNT floats */
  int B, C, Cm, HW;
  int objective;             /* B200_OBJ_*; for CFG the mix is always done in eps space */
  int clip;                  /* clip_denoised */
  int learned_range;         /* per-pixel variance from channels [C, 2C) of model_out */
  double guidance_scale;     /* s; the mix is (1-s)*eps_uncond + s*eps_cond with both scalars rounded to fp32 */
  float* sample;             /* outputs, each optional (NULL = skip) */
  float* mean;
  float* pred_x0;
  float* pred_eps;
  float* var_out;            /* only written when learned_range */
} b200_sampler_desc;

int b200_sampler_step(const b200_sampler_desc* d, void* stream);

/* Euler / Heun sampler steps in sigma space (diffusions/euler.py:50-66, diffusions/heun.py:56-107), sigma = sqrt((1-ac)/ac).
 * The predict coefficients are those of the evaluation timestep (t for the first-order step, t_prev for Heun's second
 * evaluation); second_order = 1 averages the derivative with d1 and restarts from x1 (both saved by the first step). */
typedef struct b200_ode_desc {
  const float* model_out;    /* [B][Cm][HW] */
  const float* x;            /* [B][C][HW]: x_t (first order) or the first-order sample (second order) */
  const float* d1;           /* second order: derivative of the first-order step */
  const float* x1;           /* second order: x_t of the first-order step */
  int B, C, Cm, HW;
  int objective, clip, second_order;
  float sqrt_recip_ac, sqrt_recipm1_ac, sqrt_ac, sqrt_1m_ac;
  float sigma_t, sigma_prev;
  float* sample; float* pred_x0; float* deriv;   /* each optional */
} b200_ode_desc;
int b200_ode_step(const b200_ode_desc* d, void* stream);

/* Output stage (scripts/sample_uncond.py:189-195, utils/misc.py image_norm_to_float, torchvision save_image):
 * fp32 NCHW samples -> uint8 NHWC pixels, out = trunc(clamp((clamp(x,-1,1)+1)/2*255 + 0.5, 0, 255)). */
int b200_to_uint8_hwc(const float* x, uint8_t* out, int B, int C, int HW, void* stream);

/* q(x_t | x_0) (diffusions/ddpm.py:152-172): xt = sqrt(ac[t_b]) x0 + sqrt(1-ac[t_b]) eps, per-sample t.
 * alphas_cumprod has total_steps entries; a t_b outside [0, total_steps) (where the reference raises an IndexError) is
 * never dereferenced: that sample's output is NaN. */
int b200_diffuse(const float* x0, const float* eps, const int64_t* t, const float* alphas_cumprod, float* xt,
                 int B, int CHW, int total_steps, void* stream);

/* =============================================================================================
 * Backward pass (training step: scripts/train_ddpm.py:171-192 -> loss.backward() of diffusions/ddpm.py:122-138).
 * The reference relies on autograd; each entry point below is the hand-written adjoint of a forward kernel.
 * ============================================================================================= */

/* Batched GEMM on tcgen05 with selectable operand majorness:
 *   D_g[M][N] (+)= alpha * sum_k A_g(m, k) * B_g(n, k),   g = (batch b, head h)
 * An operand is a bf16 matrix [batch][rows][ld]; its window starts at column col_base + h * col_head.
 *   mn_major = 0: rows index M (resp. N), columns index K   (the "K-major" form, e.g. q, k of softmax(q k^T))
 *   mn_major = 1: rows index K, columns index M (resp. N)   (contraction over the slow dimension, no transpose pass)
 * per_head_batch: the operand is stored [batch*heads][rows][ld] (scores / probabilities) instead of with the
 * heads side by side in its columns.  Used by the attention backward (models/modules.py:92-97 adjoint) and by
 * the backward of the per-block embedding projections (models/unet.py:18-21,41). */
typedef struct b200_gemm_operand {
  const void* ptr;
  int rows, ld;
  long long batch_stride;   /* elements; 0 = rows * ld */
  int col_base, col_head;
  int mn_major;
  int per_head_batch;
} b200_gemm_operand;

typedef struct b200_gemm_desc {
  b200_gemm_operand a, b;
  int M, N, K;
  int batch, heads;
  void* out;                /* [..][M][out_ld]: element (b, h, m, n) at b*out_batch_stride + h*out_head_stride + m*out_ld + n */
  int out_bf16;             /* 0: fp32, 1: bf16 */
  int out_ld;
  long long out_batch_stride, out_head_stride;
  int accumulate;           /* fp32 only: D += (atomic adds) instead of D = */
  int split_k;              /* > 1: split the contraction over this many CTAs (fp32 atomics; D must be pre-zeroed
                               or hold the value to accumulate into) */
  float alpha;              /* 0 is read as 1 */
} b200_gemm_desc;

int b200_gemm_batched(const b200_gemm_desc* d, void* stream);

/* Weight gradient of a convolution executed by b200_conv2d_fwd (adjoint of nn.Conv2d w.r.t. its weight):
 *   dw[co][ci][tap] += sum over (n, ho, wo) dy[n][ho][wo][co] * x[n][plane_tap][ho + dh_tap][wo + dw_tap][x_c0 + ci]
 * dy: bf16 NHWC [B][Ho][Wo][dy_C]; x: bf16 [B][x_planes][x_H][x_W][x_C] exactly as the forward read it (same tap
 * table).  dw is fp32 with arbitrary element strides, so the result lands directly in the reference's OIHW
 * .grad tensor (dw_co_stride = Cin_total*kh*kw, dw_ci_stride = kh*kw, dw_tap_stride = 1) -- and always ACCUMULATES
 * (split over pixel ranges with atomics), like autograd's .grad accumulation. */
typedef struct b200_wgrad_desc {
  const void* dy; int dy_C, dy_c0;   /* the Cout gradient channels start at channel dy_c0 of dy */
  const void* x; int x_C, x_H, x_W, x_planes, x_c0;
  int B, Ho, Wo;
  int Cout, Cin;
  int ntaps;
  int8_t taps[9][4];        /* {dw, dh, plane, 0} */
  float* dw;
  long long dw_co_stride, dw_ci_stride, dw_tap_stride;
  void* scratch;            /* optional fp32 workspace: when it holds splits * ntaps * Cout * Cin floats the pixel-range
                               splits store plain partial tiles there and a second kernel reduces them into dw
                               (deterministic, no atomics); otherwise the splits accumulate into dw with atomics */
  long long scratch_bytes;
} b200_wgrad_desc;

int b200_conv2d_wgrad(const b200_wgrad_desc* d, void* stream);

/* Training-mode K3: b200_groupnorm_apply_fwd plus dropout after the activation (nn.Dropout of models/unet.py:24).
 * The keep mask is a counter-based hash of (seed, element index): the backward regenerates it, nothing is stored.
 * seed = drop_seed + *drop_seed_dev (a device scalar, may be NULL): the per-forward base lives in device memory so that
 * a captured CUDA graph of the training step draws fresh masks on every replay. */
int b200_groupnorm_apply_train_fwd(const void* x0, int x0_is_bf16, int C0, const long long* stats0, const float* x1, int C1,
                                   const long long* stats1, int B, int HW, int W, int groups, const float* gamma,
                                   const float* beta, float eps, const float* scale, const float* shift, int ss_ld,
                                   int apply_silu, int resample, float drop_p, unsigned long long drop_seed,
                                   const unsigned long long* drop_seed_dev,
                                   void* out_bf16, void* raw_out_bf16, void* stream);
/* keep mask (1.0 / 0.0) of that dropout for elements [0, n): lets the parity tests hand the oracle the same mask */
int b200_dropout_mask(float* out, long long n, float p, unsigned long long seed, void* stream);

/* Adjoint of K3 (GroupNorm [+ AdaGN scale/shift] [+ SiLU] [+ dropout] [+ 2x resample]), two streaming passes.
 * g: bf16 gradient w.r.t. the K3 output [B][HW_out][C0+C1]; x0/x1/stats*: the forward's inputs and statistics.
 * Outputs: dx as fp32 split over the two sources (store or accumulate each; optional fp32 addend [B][HW][C] folded
 * in, e.g. the residual branch's gradient), or as one bf16 tensor (+ optional per-(image, channel) sums of dx in
 * dx_rowsum, accumulated: the time-embedding-row gradient of the producing conv; dx_colsum: its bias gradient);
 * dgamma/dbeta [C] accumulate; dscale/dshift [B][dss_ld] are written. sums: workspace [B][8][C] floats. */
typedef struct b200_gn_bwd_desc {
  const void* g;
  const float* x0; int C0; const long long* stats0;
  const float* x1; int C1; const long long* stats1;
  int B, HW, W, groups;
  const float* gamma; const float* beta; float eps;
  const float* scale; const float* shift; int ss_ld;
  int apply_silu, resample;
  float drop_p; unsigned long long drop_seed; const unsigned long long* drop_seed_dev;
  float* sums;
  float* dx0; int dx0_accumulate;
  float* dx1; int dx1_accumulate;
  const float* addend;
  void* dx_bf16;
  float* dx_rowsum; int dx_rowsum_ld;   /* [B][dx_rowsum_ld] (0 = C), accumulated */
  float* dx_colsum;                     /* [C], accumulated: sum of dx over images and pixels */
  float* dgamma; float* dbeta;
  float* dscale; float* dshift; int dss_ld;
} b200_gn_bwd_desc;
int b200_groupnorm_bwd(const b200_gn_bwd_desc* d, void* stream);

/* Gradient plumbing: fp32 [rows][C] -> bf16 (tensor-core operand) with colsum[c] += column sums (bias gradient);
 * NCHW fp32 with C <= 8 -> NHWC bf16 zero-padded to Cpad channels (first / last conv); column sums of a bf16 window;
 * out (+)= scale * {2x2 average (mode 1) | nearest 2x (mode 2)} of x (adjoints of each other up to the scale);
 * fp32 -> bf16 nearest 2x (training-mode Upsample). */
int b200_cast_bf16_colsum(const float* x, void* out_bf16, float* colsum, long long rows, int C, void* stream);
int b200_nchw_to_nhwc_pad_bf16(const float* x, void* out_bf16, float* colsum, int B, int C, int HW, int Cpad, void* stream);
int b200_colsum_bf16(const void* x, float* colsum, long long rows, int ld, int c0, int C, void* stream);
int b200_resample_f32(const float* x, float* out, int B, int H, int W, int C, int mode, float scale, int accumulate,
                      void* stream);
int b200_upsample2_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, void* stream);

/* Row softmax pieces of the attention backward (adjoint of models/modules.py:94-95):
 * P = softmax(scale * S) as bf16; dS = scale * P o (dP - rowsum(dP o P)) as bf16.  S, dP fp32 [rows][T]. */
int b200_softmax_rows(const float* S, void* P_bf16, long long rows, int T, float scale, void* stream);
int b200_softmax_bwd_rows(const void* P_bf16, const float* dP, void* dS_bf16, long long rows, int T, float scale, void* stream);

/* Adjoint of b200_time_embed, per-row part: given d_semb = dL/d SiLU(emb) (fp32 [rows][E]) and the forward's emb, it
 * accumulates db1, db2 and the class-embedding rows (atomics) and writes the bf16 operands pe [rows][dim],
 * hid / demb / dpre [rows][E] of the weight-gradient GEMMs dW2 = demb^T hid, dW1 = dpre^T pe (b200_gemm_batched). */
int b200_time_embed_bwd(const int64_t* t, int rows, const float* freqs, int dim, int E, int cos_first, const float* w1,
                        const float* b1, const float* w2, const float* emb, const float* d_semb, const int64_t* y,
                        void* pe_bf16, void* hid_bf16, void* demb_bf16, void* dpre_bf16, float* db1, float* db2,
                        float* dclass, void* stream);

/* F.mse_loss(a, b) (mean) into loss[0], and its gradient da = grad_scale[0] * 2 (a - b) / n (diffusions/ddpm.py:136-138);
 * grad_scale is a device scalar (the incoming dL/dloss) or NULL for 1. */
int b200_mse_loss(const float* a, const float* b, float* loss, long long n, void* stream);
int b200_mse_loss_grad(const float* a, const float* b, const float* grad_scale, float* da, long long n, void* stream);

/* Fused optimizer tail (clip_grad_norm_ + torch.optim.Adam/AdamW step + models/ema.py:44-52 EMA update;
 * scripts/train_ddpm.py:186-188).  `chunks` is a DEVICE array of n_chunks descriptors, each covering up to 65536
 * consecutive elements of one parameter tensor and of its gradient / Adam moments / EMA shadow (ema may be NULL).
 * gnorm_sq (device scalar) receives the squared global gradient norm when max_grad_norm > 0 or want_norm.
 * Semantics are torch's: clip coefficient min(1, max_norm / (norm + 1e-6)); Adam with bias correction, L2
 * weight decay added to the gradient (adamw = 0) or decoupled (adamw = 1); EMA: e -= (1 - decay) (e - p_new). */
typedef struct b200_optim_chunk {
  float* p; const float* g; float* m; float* v; float* ema;
  int n; int pad_;
} b200_optim_chunk;

/* Optional device-resident optimizer state: with it the step count, bias corrections, gradual EMA decay and lr are
 * advanced / read on the device (one extra 1-thread kernel), so a captured CUDA graph of the whole training step can
 * be replayed without host-computed scalars.  The host initialises step / ema_updates / lr / ema_decay_max /
 * ema_gradual and may rewrite lr between replays. */
typedef struct b200_optim_dev_state {
  int step, ema_updates;
  float lr, bc1, bc2_sqrt, ema_decay, ema_decay_max;
  int ema_gradual;
} b200_optim_dev_state;

typedef struct b200_optim_desc {
  const void* chunks; int n_chunks;
  float lr, beta1, beta2, eps, weight_decay;
  int adamw;
  int step;                 /* 1-based step count (bias correction) */
  float max_grad_norm;      /* <= 0: no clipping */
  int want_norm;
  float* gnorm_sq;
  float ema_decay;          /* < 0: no EMA update */
  void* dev_state;          /* b200_optim_dev_state in device memory, or NULL (then step / lr / ema_decay above are used) */
} b200_optim_desc;
int b200_optimizer_step(const b200_optim_desc* d, void* stream);

/* Multi-tensor weight pack (see b200_conv_desc.w): one launch re-creates the bf16 GEMM operands of every layer from
 * the fp32 OIHW parameters after they changed.  `table` is a DEVICE array of n_entries descriptors.
 *   mode 0: dst[row0 + co][col0 + tap*Ci + ci] = src[co][ci][tap]                    (forward operand, bf16, row stride ld)
 *   mode 1: dst[row0 + ci][col0 + (taps-1-tap)*Co + co] = src[co][ci][tap]           (data-gradient operand)
 *   mode 2: dst_f32[i] = src[i] + src2[i], i < Co*Ci*taps                            (summed bias of a fused shortcut)
 *   mode 3: dst[row0 + co][col0 + tap*3*Ci + {0, Ci, 2*Ci} + ci] = {hi, hi, lo}(src[co][ci][tap]), taps <= 9
 *           (FP32 mode, see b200_split_cast: weight side of the 3-term bf16 split product)
 *   mode 4: as mode 3 with a tap stride of tap_ld (>= 3*Ci) columns instead of 3*Ci; the columns in between are not
 *           written (zero-fill dst once): weight side of b200_first_split's 64-channel pixels */
typedef struct b200_pack_entry {
  const float* src; const float* src2; void* dst;
  int Co, Ci, taps, mode;
  int row0, col0, ld, tap_ld;
} b200_pack_entry;
int b200_pack_weights(const void* table, int n_entries, void* stream);

/* =============================================================================================
 * FP32 mode ("bf16x3", csrc/precise.cu): the reference runs this path in fp32 (models/unet.py:121-152,
 * models/modules.py:92-97).  Every GEMM operand x is split as hi = bf16(x), lo = bf16(x - hi) and a product a*w is
 * evaluated as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation; the three terms are laid out along the GEMM K
 * dimension, activations as [hi | lo | hi] (pattern 0) and weights as [hi | hi | lo] (pattern 1) per group of channels,
 * so b200_conv2d_fwd / b200_gemm_batched run unchanged on 3x the channels.  Gate: eps rel-L2 <= 1e-4 vs fp32.
 * ============================================================================================= */
typedef struct b200_split_desc {
  const float* in;          /* fp32 [rows][in_ld]; the window is columns [in_col0, in_col0 + C) */
  long long rows;
  int in_ld, in_col0, C;
  int group;                /* channels per split group (C for a plain tensor, head dim for per-head operands) */
  int pattern;              /* 0: [hi | lo | hi], 1: [hi | hi | lo] */
  int act;                  /* 1: exact SiLU before the split */
  void* out;                /* bf16 [rows][out_ld], group g at out_col0 + 3*g*group (+ group, + 2*group) */
  int out_ld, out_col0;
  int planes_rows;          /* > 0: plane layout instead: out[b][3][planes_rows][C], b = row / planes_rows (the K dimension
                               of an MN-major GEMM operand, e.g. V of softmax(QK^T) V) */
  int parity_H, parity_W;   /* > 0: rows are the NHWC pixels of [B][H][W]; output rows are ordered as the 4 parity planes
                               [B][2*(h&1)+(w&1)][H/2][W/2] (input of the stride-2 conv, cf. b200_cast_bf16) */
} b200_split_desc;
int b200_split_cast(const b200_split_desc* d, void* stream);

/* b200_groupnorm_apply_fwd writing the split operand: out [B][HW_out][3*(C0+C1)] (pattern 0), optional raw copy
 * [B][HW][3*(C0+C1)]; statistics evaluated in fp64, exact SiLU. */
int b200_groupnorm_apply_split_fwd(const float* x0, int C0, const long long* stats0, const float* x1, int C1,
                                   const long long* stats1, int B, int HW, int W, int groups, const float* gamma,
                                   const float* beta, float eps, const float* scale, const float* shift, int ss_ld,
                                   int apply_silu, int resample, void* out_split, void* raw_out_split, void* stream);

/* softmax over the T columns of fp32 scores S [rows][T] (times `scale`) -> split probabilities bf16 [rows][3T] =
 * [p_hi | p_lo | p_hi] (models/modules.py:95). */
int b200_softmax_rows_split(const float* S, void* P_split, long long rows, int T, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DIFF_H_ */
