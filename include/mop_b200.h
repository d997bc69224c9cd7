/*
 * mop_b200.h - C ABI of libmop_b200.so: the B200 (sm_100a) implementation of the
 * Mixture-of-Products attention hot path of Eran-BA/MoP.
 *
 * The reference has no FFI of its own (it is eager PyTorch); each entry point
 * below replaces the body of one reference `nn.Module.forward` (and its
 * autograd) between the input projections and the output projection:
 *
 *   mop_edgewise_fwd/bwd   EdgewiseMSA.forward      mop/models/attention_variants.py:500-562
 *                          EdgewiseGateHead.forward mop/models/attention_variants.py:311-331
 *                          (clone: experiments/cifar100_edgewise_gates.py:99-115,211-310)
 *   mop_sdpa_fwd/bwd       MSA.forward              mop/models/components.py:61-64
 *                          BaselineMSA.forward      mop/models/attention_variants.py:42-46
 *                          MultiheadSelfAttention   mop/models/whisper_mop.py:163-175
 *                          MultiheadCrossAttention  mop/models/whisper_mop.py:212-219
 *   mop_quartet_fwd/bwd    CausalSelfAttention      mop/models/quartet_attn_patch.py:88-121
 *
 * Conventions
 *   - plain C, POD structs, raw device pointers; no torch / C++ types cross the ABI.
 *   - the caller owns every buffer (inputs, outputs, saved statistics, workspace);
 *     the library never allocates or frees device memory and never synchronises.
 *   - all work is enqueued on the `cuda_stream` argument (a cudaStream_t).
 *   - return 0 on success, a negative MOP_E* code otherwise; the message is
 *     available from mop_last_error() (thread local).  No exceptions cross the ABI.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with MOP_ECUDA.
 *   - every struct starts with `struct_bytes = sizeof(struct)`; a mismatch is
 *     rejected with MOP_EABI.
 *   - `dtype` selects the storage type of the activation tensors (qkv / q,k,v /
 *     y / dy / gradients).  Parameters, statistics and partial parameter
 *     gradients are always fp32.  MOP_F32 = "fp32 mode" (fp32 CUDA-core math,
 *     <=1e-5 of the reference); MOP_BF16 = bf16 storage, tensor-core contractions
 *     with fp32 accumulation and fp32 statistics.
 */
#ifndef MOP_B200_H
#define MOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOP_ABI_VERSION 9

enum { MOP_OK = 0, MOP_EINVAL = -1, MOP_EABI = -2, MOP_EUNSUPPORTED = -3, MOP_ECUDA = -4, MOP_EWORKSPACE = -5 };
enum { MOP_F32 = 0, MOP_BF16 = 1 };
enum { MOP_GATE_DENSE = 0, MOP_GATE_LOWRANK = 1, MOP_GATE_CONST = 2 };   /* CONST: fixed scalar gates (MultiHopMSA) */
/* which implementation ran (written to `impl_used` by the *_fwd / *_bwd calls) */
enum { MOP_IMPL_SIMT = 1, MOP_IMPL_TCGEN05 = 2 };
/* `impl` request: AUTO picks tcgen05 when the shape/dtype is covered, else SIMT */
enum { MOP_IMPL_AUTO = 0 };

int mop_abi_version(void);
const char* mop_last_error(void);
/* number of SMs of the current device (cached), <0 on error */
int mop_device_sm_count(void);

/* ------------------------------------------------------------------------- */
/* Edgewise (Mixture-of-Products) attention core                              */
/* ------------------------------------------------------------------------- */
typedef struct MopEdgewiseParams {
  int32_t struct_bytes;
  int32_t dtype;      /* MOP_F32 | MOP_BF16: type of qkv, y, dy, dqkv */
  int32_t impl;       /* MOP_IMPL_AUTO | MOP_IMPL_SIMT | MOP_IMPL_TCGEN05 */
  int32_t impl_used;  /* out */
  int32_t B, H, N, dk;
  int32_t V;          /* number of views / score maps (>=2)            :362 */
  int32_t Vp;         /* 1: share_qkv (one projection + per-view scales) :459-464; V: one projection per view :466-470 */
  int32_t gate_mode;  /* MOP_GATE_DENSE | MOP_GATE_LOWRANK               :250,273 */
  int32_t gate_rank;  /* r (lowrank)                                     :277 */
  int32_t hidden;     /* dense head hidden channels (reference: 16)      :445 */
  int32_t use_k3;     /* dense head only: 3x3 stage present              :253 */
  float beta_not;     /*                                                 :546 */
  float eps;          /* 1e-6                                            :516 */
  /* activations */
  const void* qkv;    /* [B,N,Vp,3,H,dk] contiguous, `dtype`: output of the qkv Linear(s) */
  void* y;            /* [B,N,H,dk] contiguous, `dtype`: merge-heads layout :563 */
  /* parameters (fp32, device) */
  const float* q_scale; /* [V,H,dk] or NULL (all ones)                   :376-378 */
  const float* k_scale;
  const float* v_scale;
  const float* chain_value_logit; /* [1]                                 :451 */
  const float* row_w;   /* lowrank: [4r,C], C = 2V+2                     :277 */
  const float* row_b;   /* [4r] */
  const float* col_w;   /* [4r,C]                                        :278 */
  const float* col_b;   /* [4r] */
  const float* conv1_w; /* dense: [hidden,C]                             :251 */
  const float* conv1_b; /* [hidden] */
  const float* mid3_w;  /* [hidden,hidden,3,3] iff use_k3                :254 */
  const float* mid3_b;  /* [hidden] */
  const float* conv2_w; /* [4,hidden]                                    :255 */
  const float* conv2_b; /* [4] */
  float* row_stats;     /* [B,H,N,2] fp32, optional: per row of the mixed score map its (integer) base-2 reference exponent and
                           the sum of the bf16-rounded probabilities.  Written by mop_edgewise_fwd when non-NULL; the
                           tcgen05 backward for N != 64 needs it as an input */
  float* y_base;        /* [B,N,H,dk] fp32, optional, same life cycle as row_stats: A V_1 (the output without the chain
                           value term) in full precision, from which the backward forms rowsum(dA . A) */
  float* aux;           /* [mop_edgewise_aux_floats()] fp32, optional (fwd out, bwd in).  The N = 64 tcgen05 kernels hand their
                           small per-(b,h) vectors from the forward to the backward instead of recomputing them: per-view and
                           mixed-map softmax row statistics, row/column feature means, low-rank gate factors (17 KB per (b,h)).
                           mop_edgewise_aux_floats() == 0: not used by this configuration (may be NULL) */
  /* backward only */
  const void* dy;       /* [B,N,H,dk] `dtype` */
  void* dqkv;           /* [B,N,Vp,3,H,dk] `dtype`, fully overwritten */
  float* dscale_part;   /* [R,3,V,dk], R = mop_edgewise_partial_rows(): partial grads of q/k/v_scale; row i belongs to head i % H
                           (R = B*H: one row per (b,h) problem; the N = 64 tcgen05 backward sums over the problems of a CTA and
                           writes one row per CTA).  The caller sums rows with equal i % H.  NULL if scales are NULL */
  float* dhead_part;    /* [R, mop_edgewise_head_param_count()] partial grads of the gate head, packed in the order
                           lowrank: row_w,row_b,col_w,col_b ; dense: conv1_w,conv1_b,(mid3_w,mid3_b),conv2_w,conv2_b */
  float* dlogit_part;   /* [B*H] (always one value per problem) */
  /* scratch */
  void* workspace;
  size_t workspace_bytes;
  /* gate_mode == MOP_GATE_CONST (variant D, MultiHopMSA attention_variants.py:163-231): the four gates are the fixed scalars
     {and_, or_, not_, chain} instead of a gate head, the chain product is A_1 A_2^(hops-1) (views 0,1,1,..) and there is no
     reverse chain; dhead_part is not written */
  float const_gates[4];
  int32_t hops;         /* chain length >= 2 (CONST mode); ignored otherwise */
  /* S lens bank (attention_variants.py:427-442, :523-533): lens_n depthwise 3x3 convolutions of the V score maps with dilation
     (= zero padding) lens_dil[l]; channel l*V + v = conv_l(S_v) is appended to the gate-head features (C = 2V + 2 + V lens_n) */
  int32_t lens_n;       /* 0: off; <= 4 */
  int32_t lens_dil[4];
  const float* lens_w;  /* [lens_n, V, 3, 3] fp32 (the lens_bank.{l}.weight tensors stacked) */
  float* dlens_part;    /* backward: [B*H, lens_n * V * 9] partial grads */
} MopEdgewiseParams;

size_t mop_edgewise_head_param_count(const MopEdgewiseParams* p);
/* 1 if mop_edgewise_fwd with these params would run the kernel whose backward wants `row_stats` and `y_base`: allocate
 * them, pass them to the forward and hand them back to the backward; 0 otherwise (both may be NULL) */
int mop_edgewise_needs_row_stats(const MopEdgewiseParams* p);
/* rows R of dscale_part / dhead_part that mop_edgewise_bwd writes for these params (B*H, or the persistent grid) */
int mop_edgewise_partial_rows(const MopEdgewiseParams* p);
/* number of floats of `aux` the forward would write / the backward needs for these params (0: none) */
/* Sums of the gradient partials of one mop_edgewise_bwd call in one launch (deterministic): dscale [3,V,H,dk] from dscale_part
 * [R,3,V,dk] (row i belongs to head i % H), dhead [nhead] from dhead_part [R,nhead], dlogit [1] from dlogit_part [G].  dscale_part /
 * dhead_part may be NULL together with their outputs. */
int mop_edgewise_reduce_partials(const float* dscale_part, const float* dhead_part, const float* dlogit_part, int R, int H, int V, int dk,
                                 int nhead, int G, float* dscale, float* dhead, float* dlogit, void* cuda_stream);
size_t mop_edgewise_aux_floats(const MopEdgewiseParams* p);
/* bytes of workspace needed by mop_edgewise_fwd (backward=0) / mop_edgewise_bwd (backward=1) */
size_t mop_edgewise_workspace_bytes(const MopEdgewiseParams* p, int backward);
int mop_edgewise_fwd(MopEdgewiseParams* p, void* cuda_stream);
int mop_edgewise_bwd(MopEdgewiseParams* p, void* cuda_stream);

/* ------------------------------------------------------------------------- */
/* Plain / causal / biased / cross attention                                  */
/* ------------------------------------------------------------------------- */
typedef struct MopSdpaParams {
  int32_t struct_bytes;
  int32_t dtype, impl, impl_used;
  int32_t B, H, Nq, Nk, dk;
  int32_t causal;       /* j > i -> -inf            whisper_mop.py:165-167 */
  float scale;          /* 1/sqrt(dk) */
  /* q[b,n,h,:] = q + b*q_sb + n*q_sn + h*q_sh (element strides, last dim contiguous) */
  const void *q, *k, *v;
  int64_t q_sb, q_sn, q_sh, k_sb, k_sn, k_sh, v_sb, v_sn, v_sh;
  void* y;              /* [B,Nq,H,dk] contiguous */
  /* optional additive bias (fp32), element strides may be 0 for broadcast   whisper_mop.py:169-170,213-214 */
  const float* bias; int64_t bias_sb, bias_sh, bias_sq, bias_sk;
  /* optional keep-mask: entries == 0 are filled with -inf      attention_variants.py:43-44 */
  const float* zero_mask; int64_t zm_sb, zm_sh, zm_sq, zm_sk;
  float* lse;           /* [B,H,Nq] saved row log-sum-exp (fwd out, bwd in) */
  /* backward */
  const void* dy;       /* [B,Nq,H,dk] */
  void *dq, *dk_, *dv;  /* contiguous [B,Nq,H,dk], [B,Nk,H,dk], [B,Nk,H,dk] */
  void* workspace; size_t workspace_bytes;
  /* attention dropout on the probabilities (attention_variants.py:45, whisper_mop.py:173,217): p = 0 -> off.  The keep mask is
     a function of (dropout_seed, dropout_offset, b, h, i, j) regenerated by every kernel; pass the forward's pair to the backward */
  float dropout_p;
  uint64_t dropout_seed, dropout_offset;
} MopSdpaParams;

size_t mop_sdpa_workspace_bytes(const MopSdpaParams* p, int backward);
int mop_sdpa_fwd(MopSdpaParams* p, void* cuda_stream);
int mop_sdpa_bwd(MopSdpaParams* p, void* cuda_stream);

/* ------------------------------------------------------------------------- */
/* Quartet causal attention                                                   */
/* ------------------------------------------------------------------------- */
typedef struct MopQuartetParams {
  int32_t struct_bytes;
  int32_t dtype, impl, impl_used;
  int32_t B, H, T, dk;
  int32_t use_quartet;  /* 0: single z-scored map         quartet_attn_patch.py:108-110 */
  float scale;          /* 1/sqrt(dk) */
  float eps;            /* score_norm_eps (1e-5) */
  /* all of q,k,v,q2,k2: [B,T,H,dk] contiguous (the Linear outputs viewed as heads, :84-92) */
  const void *q, *k, *v, *q2, *k2;
  const float* mixture;        /* [1] :52 */
  const float* quartet_scale;  /* [1] :55 */
  const float* add_mask; int64_t am_sb, am_sh, am_sq, am_sk; /* optional additive mask :115-116 */
  void* y;              /* [B,T,H,dk] */
  float* stats;         /* [B,H,T,3]: sigma1, sigma2, lse  (fwd out, bwd in) */
  float* y_f32;         /* [B,T,H,dk] fp32, optional (fwd out, bwd in): the output before rounding to `dtype`.  When set, the
                           backward forms delta = dO . y from it; the mixture / quartet_scale gradients are heavily
                           cancelling sums whose accuracy is bounded by that of delta */
  /* backward */
  const void* dy;
  void *dq, *dk_, *dv, *dq2, *dk2;
  float* dscalar_part;  /* [B*H,2] partial grads of (mixture, quartet_scale) */
  void* workspace; size_t workspace_bytes;
  /* attention dropout (quartet_attn_patch.py:24,118-119; TransformerConfig.dropout defaults to 0.1): see MopSdpaParams */
  float dropout_p;
  uint64_t dropout_seed, dropout_offset;
  /* ABI v9, backward only, optional: the workspace the forward call of the same inputs ran with (left untouched since).  The
     tcgen05 backward then reads the forward's key preparation (centred keys, Gram tile images) from it instead of redoing it. */
  const void* fwd_workspace; size_t fwd_workspace_bytes;
} MopQuartetParams;

size_t mop_quartet_workspace_bytes(const MopQuartetParams* p, int backward);
int mop_quartet_fwd(MopQuartetParams* p, void* cuda_stream);
int mop_quartet_bwd(MopQuartetParams* p, void* cuda_stream);

/* ------------------------------------------------------------------------- */
/* Fused residual-add + DropPath scale + LayerNorm              */
/* ------------------------------------------------------------------------- */
/* The elementwise neighbours of the attention call in every reference block:
 *   x = x + dp(attn(ln1(x)));  x = x + dp(mlp(ln2(x)))     experiments/cifar100_edgewise_gates.py:371-374,
 *   DropPath (per-sample mask / keep)                       mop/models/components.py:14-27
 * forward : x_new = x + scale[row / rows_per_sample] * r (r optional), y = LayerNorm(x_new) * gamma + beta
 * backward: dx = LN'(dy) + dx_new (optional), dr = scale * dx, per-CTA partials of dgamma / dbeta */
typedef struct MopLnParams {
  int32_t struct_bytes;
  int32_t rows, D;          /* x is [rows, D] fp32 contiguous, D <= 1024 */
  int32_t r_dtype, y_dtype; /* MOP_F32 | MOP_BF16: type of r / dr and of y / dy */
  int32_t rows_per_sample;  /* tokens per sample (rows sharing one DropPath scale) */
  int32_t nparts;           /* rows of dgamma_part / dbeta_part (>= mop_ln_partial_rows(rows)) */
  float eps;
  const void* x;            /* residual stream in, fp32 */
  const void* r;            /* branch output to add, or NULL */
  const float* scale;       /* [rows / rows_per_sample] per-sample factor (mask / keep), or NULL (= 1) */
  const float* gamma;
  const float* beta;
  void* x_new;              /* fp32 [rows, D]: x + scale r (fwd out when r != NULL; bwd in: the normalised tensor) */
  void* y;                  /* [rows, D] y_dtype */
  float* mean;              /* [rows] fwd out, bwd in */
  float* rstd;              /* [rows] */
  const void* dy;           /* [rows, D] y_dtype */
  const void* dx_new;       /* optional fp32 [rows, D]: gradient reaching x_new through the residual stream */
  void* dx;                 /* fp32 [rows, D] */
  void* dr;                 /* optional [rows, D] r_dtype */
  float* dgamma_part;       /* [nparts, D] */
  float* dbeta_part;        /* [nparts, D] */
} MopLnParams;

/* number of per-CTA partial rows the backward writes for a [rows, D] problem on the current device */
int mop_ln_partial_rows(int rows);
/* recommended number of partial rows (= CTAs of mop_ln_bwd) for feature count D; mop_ln_bwd launches exactly `nparts` CTAs and
 * writes every partial row, so any count from 1 to 16 x the SM count is valid */
int mop_ln_partial_rows_d(int rows, int D);
int mop_ln_fwd(MopLnParams* p, void* cuda_stream);
int mop_ln_bwd(MopLnParams* p, void* cuda_stream);

/* ------------------------------------------------------------------------- */
/* Whisper-MoP 2D gate                                           */
/* ------------------------------------------------------------------------- */
/* MoP2D.forward mop/models/whisper_mop.py:91-124 is linear and bias-free: ViewsConv2D (1x1) -> Kernels2D (k x k, zero padding
 * k/2) -> FuseExcInh2D (1x1) -> mean over mel bins -> 1 + a_pos g_pos - a_neg g_neg.  The caller folds the weights into ONE
 * k x k filter He (a_pos H_pos - a_neg H_neg); the library computes, without forming any [B, V+K, T, F] map,
 *   R[b,t,w]   = sum of the mel bins that filter column w sees at row t      (fwd out, bwd in; [B, T, ks])
 *   gate[b,t]  = 1 + (1/F) sum_{u,w} He[u,w] R[b, t+u-k/2, w]                 ([B, T])
 * backward: dHe_part[i] (nparts = mop_mop2d_partial_rows() rows of ks*ks, summed by the caller); mel is data (no gradient). */
int mop_mop2d_partial_rows(void);
int mop_mop2d_fwd(const float* mel, const float* He, float* R, float* gate, int B, int T, int F, int ks, void* cuda_stream);
int mop_mop2d_bwd(const float* R, const float* dgate, float* dHe_part, int nparts, int B, int T, int F, int ks, void* cuda_stream);

/* ------------------------------------------------------------------------- */
/* ViT-MoP token gate                                            */
/* ------------------------------------------------------------------------- */
/* mop/models/vit_mop.py:95-114 with components.py:255-303: ViewsLinear (D -> V per token) -> Kernels3 (conv3x3 V -> 16, SiLU,
 * conv1x1 16 -> K) -> FuseExcInh (conv1x1 V+K -> hid, SiLU, conv1x1 hid -> 2 + bias, sigmoid) -> gate = 1 + a_pos G+ - a_neg G-
 * -> out = tok * gate.  One fused kernel per direction; a_pos / a_neg = softplus(alpha) are applied by the caller (one-element
 * device tensors).  Backward outputs are per-CTA partial rows, summed by the caller. */
typedef struct MopTokenGateParams {
  int32_t struct_bytes;
  int32_t dtype;            /* MOP_F32 | MOP_BF16: tokens in / out */
  int32_t B, T, D;          /* tokens [B, T, D], T = Gh * Gw <= 256, D % 8 == 0, D <= 2048 */
  int32_t Gh, Gw;
  int32_t V, K, hid;        /* views (<= 8), kernel maps (<= 8), FuseExcInh hidden width (<= 16) */
  int32_t nparts;           /* CTAs of the backward = rows of dnet_part (mop_token_gate_partial_rows) */
  const void* x;            /* [B, T, D] */
  const float* views_w;     /* [V, D]            ViewsLinear.proj.weight */
  const float* k3_w;        /* [16, V, 3, 3]     Kernels3.k[0].weight */
  const float* k1_w;        /* [K, 16]           Kernels3.k[2].weight */
  const float* f1_w;        /* [hid, V + K]      FuseExcInh.fuse[0].weight */
  const float* f2_w;        /* [2, hid]          FuseExcInh.fuse[2].weight */
  const float* f2_b;        /* [2]               FuseExcInh.fuse[2].bias */
  const float* a_pos;       /* [1] softplus(alpha_pos) */
  const float* a_neg;       /* [1] softplus(alpha_neg) */
  void* out;                /* [B, T, D] */
  float* views;             /* [B, T, V] fwd out, bwd in */
  float* gate;              /* [B, T]    fwd out, bwd in */
  const void* dout;         /* [B, T, D] */
  void* dx;                 /* [B, T, D] */
  float* dwv_part;          /* [nparts * mop_token_gate_wv_groups(D), V, D] */
  float* dnet_part;         /* [nparts, n + 2]: k3_w | k1_w | f1_w | f2_w | f2_b (n = mop_token_gate_net_params) | d a_pos | d a_neg */
} MopTokenGateParams;
int mop_token_gate_partial_rows(int B);
int mop_token_gate_wv_groups(int D);
int mop_token_gate_net_params(const MopTokenGateParams* p);
int mop_token_gate_fwd(MopTokenGateParams* p, void* cuda_stream);
int mop_token_gate_bwd(MopTokenGateParams* p, void* cuda_stream);

/* GPT-MoP 1-D token gate: MoPBlock.apply_mop mop/models/gpt_mop.py:102-123 (ViewsLinear1D :19-33, Kernels1D :36-49,
 * FuseExcInh1D :52-67).  Linear and bias-free: the caller folds Kernels1D, FuseExcInh1D and alpha into a 3-tap filter of the views,
 * w_eff [3, V]; gate[t] = 1 + sum_tau sum_v w_eff[tau+1][v] views[t+tau][v] (zero outside the sequence), out = x * gate. */
typedef struct MopTokenGate1dParams {
  int32_t struct_bytes;
  int32_t dtype;            /* MOP_F32 | MOP_BF16: tokens in / out */
  int32_t B, T, D, V;       /* x [B, T, D], D % 8 == 0, D <= 2048, V <= 8 */
  int32_t nparts;           /* CTAs of the backward (mop_token_gate1d_partial_rows) */
  const void* x;
  const float* views_w;     /* [V, D] */
  const float* w_eff;       /* [3, V] */
  void* out;                /* [B, T, D] */
  float* views;             /* [B, T, V] fwd out, bwd in */
  float* gate;              /* [B, T]    fwd out, bwd in */
  const void* dout;
  void* dx;
  float* dwv_part;          /* [nparts * mop_token_gate_wv_groups(D), V, D] */
  float* dweff_part;        /* [nparts, 3 V] */
} MopTokenGate1dParams;
int mop_token_gate1d_partial_rows(int B, int T);
int mop_token_gate1d_fwd(MopTokenGate1dParams* p, void* cuda_stream);
int mop_token_gate1d_bwd(MopTokenGate1dParams* p, void* cuda_stream);

/* The dropout factor every attention kernel applies to P[b,h,i,j] for (p, seed, offset): out[bh, i, j] = 0 or 1 / (1 - p)
 * (fp32, [BH, Nq, Nk]).  Test infrastructure: lets the CPU oracle be evaluated under the kernels' own mask. */
int mop_dropout_mask(float* out, int BH, int Nq, int Nk, float p, uint64_t seed, uint64_t offset, void* cuda_stream);

/* Bring-up check of the tcgen05 building blocks: D = (a_mn ? A^T : A) * (b_mn ? B : B^T) on 64x64 fp32
 * device matrices (rounded to bf16), accumulator at TMEM lane offset {0,16} / column offset; D2 = D + 1
 * after a tcgen05.st/ld round trip. */
int mop_selftest_umma(const float* A, const float* B, float* D, float* D2, int a_mn, int b_mn, int lane_off,
                      int col_off, void* cuda_stream);

/* Bring-up check of the M=128 building blocks (thread-per-row 32x32b TMEM access, chunk-major tiles with Ra / Rb
 * rows): D[256 x Nn] = A[Ma x K] * B^T (b_mn = 0, B is [Nn x K]) or A * B[b_k0 : b_k0 + K, :] (b_mn = 1, B is [Kb x Nn]). */
int mop_selftest_umma128(const float* A, const float* B, float* D, int Ma, int Nn, int K, int b_mn, int Ra, int Rb, int Kb,
                         int b_k0, void* cuda_stream);

/* Bring-up check of the TMA tile load: rows [row0, row0 + R) of (batch, head) of a bf16 [B][N][H][dk] tensor with element
 * strides (sb, sn, sh) -> the R x 64 shared-memory tile image in the 128-byte swizzled operand layout (16-byte chunk c of row r
 * stored at chunk c ^ (r & 7)), copied to `out` (R * 128 bytes). */
int mop_selftest_tma(const void* x, void* out, int B, int N, int H, int dk, int64_t sb, int64_t sn, int64_t sh, int R, int row0,
                     int head, int batch, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MOP_B200_H */
