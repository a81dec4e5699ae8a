/*
 * bifold_b200 — C ABI of the B200-native BiFold policy hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and a CUDA stream, never
 * allocates, never synchronises, never throws.  Return value: 0 on success, a negative BFK_E_* code
 * otherwise (text via bfk_last_error()).  All pointers are DEVICE pointers unless stated.  The caller
 * owns every buffer.
 *
 * The reference (Barbany/bifold) has no native code and no FFI on this path: it is eager PyTorch.
 * Each entry point therefore cites the reference *Python* call it replaces (path:line relative to the
 * reference root; `HF:` = transformers/models/siglip/modeling_siglip.py 5.5.0, which the reference
 * calls at bifold/models/siglip.py:26,57,102).  INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.
 */
#ifndef BIFOLD_B200_H_
#define BIFOLD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* bfk_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define BFK_API __attribute__((visibility("default")))
#else
#define BFK_API
#endif

enum { BFK_F32 = 0, BFK_BF16 = 1 };
enum { BFK_OK = 0, BFK_E_INVALID = -1, BFK_E_UNSUPPORTED = -2, BFK_E_CUDA = -3 };
enum { BFK_BACKEND_AUTO = 0, BFK_BACKEND_SIMT = 1, BFK_BACKEND_TCGEN05 = 2,
       BFK_BACKEND_TCGEN05_PERSISTENT = 3 /* attention only: force the persistent forward kernel (attention_v2.cu) */ };
enum {
  BFK_ACT_NONE = 0,
  BFK_ACT_GELU_TANH = 1,   /* HF:323-327 (gelu_pytorch_tanh)            */
  BFK_ACT_GELU_ERF = 2,    /* bifold/models/modules.py:23 (nn.GELU())   */
  BFK_ACT_DGELU_TANH = 3,  /* out = acc * gelu_tanh'(aux)  (backward)   */
  BFK_ACT_DGELU_ERF = 4,   /* out = acc * gelu_erf'(aux)   (backward)   */
  /* training pair that moves the derivative into the forward epilogue (tanh / exp are computed once):
   *   forward : v = acc + bias;  out = gelu(v);  aux = gelu'(v)          backward: out = acc * aux */
  BFK_ACT_GELU_TANH_GRAD = 5,
  BFK_ACT_GELU_ERF_GRAD = 6,
  BFK_ACT_MUL_AUX = 7
};

BFK_API int32_t bfk_version(void);
BFK_API const char* bfk_last_error(void);        /* thread-local, host string */
BFK_API int32_t bfk_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM with fused epilogue:  D = act(alpha * A·Bᵀ + bias) + residual      (fp32 accumulate)
 *   A element (m,k): trans_a ? A[k*lda + m] : A[m*lda + k]
 *   B element (n,k): trans_b ? B[k*ldb + n] : B[n*ldb + k]    (trans_b = 0 is the nn.Linear layout)
 * Replaces: nn.Linear / F.linear in HF:285-287,309,324-326; bifold/models/transformer.py:28,42;
 * bifold/models/modules.py:22-25; nn.Conv2d patch embedding HF:124-130,178 (after bfk_im2col) and
 * their autograd backward (dgrad: trans_b = 1; wgrad: trans_a = trans_b = 1).
 * in_dtype BF16 + backend AUTO/TCGEN05 runs on tcgen05 tensor cores (TMA-fed, TMEM accumulators);
 * in_dtype F32 runs the SIMT fp32 kernel ("fp32 parity mode").
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t trans_a;
  const void* B; int64_t ldb; int32_t trans_b;
  int32_t in_dtype;
  void* D; int64_t ldd; int32_t out_dtype;
  const float* bias;                 /* [N] fp32 or NULL */
  int32_t act;                       /* BFK_ACT_*  */
  void* aux; int64_t ld_aux;         /* GELU_*: optional store of the pre-activation (out_dtype); GELU_*_GRAD: store of
                                        gelu'(pre-activation) (out_dtype); DGELU_* / MUL_AUX: tensor to read (in_dtype) */
  const float* residual; int64_t ldr; int32_t res_row_mod; /* fp32; row = res_row_mod ? m % res_row_mod : m */
  float alpha;
  int32_t accumulate;                /* D += ... (fp32 output only) */
  int32_t backend;
  void* workspace; int64_t workspace_bytes; /* optional fp32 scratch enabling split-K (plain fp32-output GEMMs, e.g. wgrad) */
  int32_t split_k;                          /* 0 = auto when a workspace is given, 1 = off, n = force n splits */
} bfk_gemm_args;
BFK_API int32_t bfk_gemm(const bfk_gemm_args* a, bfk_stream_t s);

/* Column sums of a row-major matrix: out[n] = sum_m X[m*ld + n]  (bias gradients).  Deterministic
 * two-stage reduction; `partials` must hold bfk_colsum_workspace_floats(M,N) floats. */
BFK_API size_t bfk_colsum_workspace_floats(int32_t M, int32_t N);
BFK_API int32_t bfk_colsum(const void* X, int32_t dtype, int32_t M, int32_t N, int64_t ld, float* out,
                   float* partials, bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (biased variance, eps inside sqrt), fp32 statistics.
 * Replaces nn.LayerNorm at HF:334-336,348,357,518,618 (eps 1e-6) and bifold/models/modules.py:11
 * (eps 1e-5).  Backward fuses the residual-gradient add and can emit a second copy in `dx2_dtype`.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t M, D;
  const float* x; int64_t ldx;
  const float* gamma; const float* beta; float eps;
  void* y; int64_t ldy; int32_t y_dtype;
  float* mean; float* rstd;          /* [M] saved for backward (may be NULL) */
  /* optional second output (NULL = none), same dtype as y: y2 = dropout(y, p, seed) -- the LoRA-branch input
   * (peft LoRA lora_dropout on the q_proj / v_proj inputs); same element hash as bfk_dropout_copy / bfk_layernorm_bwd.dy2 */
  void* y2; int64_t ldy2; float y2_dropout_p; uint64_t y2_seed;
  /* optional fused LoRA down-projection (bf16 y only, NULL = none):  u[m, j] = sum_k drop(y)[m, k] * lora_a[j, k],
   * j < 16 (q factors rows 0..r-1, v factors rows 8..8+r-1, unused rows zero) -- peft lora_A of q_proj / v_proj applied to
   * the dropped LayerNorm output; u is normally the 16 extra K columns of the same row (u = y + D, ld_u = ldy) */
  const void* lora_a; void* lora_u; int64_t ld_u;
  /* optional DEVICE counter added to y2_seed when the kernel runs (NULL = 0): lets a captured CUDA graph draw a new
   * dropout mask on every replay */
  const uint64_t* seed_ptr;
} bfk_layernorm_fwd_args;
BFK_API int32_t bfk_layernorm_fwd(const bfk_layernorm_fwd_args* a, bfk_stream_t s);

typedef struct {
  int32_t M, D;
  const void* dy; int64_t ld_dy; int32_t dy_dtype;
  const float* x; int64_t ldx;
  const float* mean; const float* rstd; const float* gamma;
  const float* dres; int64_t ld_dres;        /* residual-branch gradient added to dx, or NULL */
  float* dx; int64_t ld_dx;
  void* dx2; int64_t ld_dx2; int32_t dx2_dtype; /* optional second copy of dx, or NULL */
  float* dgamma; float* dbeta;               /* [D] overwritten, or NULL (frozen LN) */
  float* partials;                           /* bfk_layernorm_bwd_workspace_floats(M,D) floats if dgamma */
  const void* dy2; int64_t ld_dy2;           /* optional second upstream gradient (dy_dtype) added to dy under the
                                                LoRA dropout mask: dy += keep(seed,m,k) ? dy2/(1-p) : 0 */
  float dy2_dropout_p; uint64_t dy2_seed;
  const uint64_t* seed_ptr;          /* optional device counter added to dy2_seed (see bfk_layernorm_fwd_args) */
} bfk_layernorm_bwd_args;
BFK_API size_t bfk_layernorm_bwd_workspace_floats(int32_t M, int32_t D);
BFK_API int32_t bfk_layernorm_bwd(const bfk_layernorm_bwd_args* a, bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head attention on a packed QKV matrix.
 *   qkv row (b*N+n): [ q(h,d) | k(h,d) | v(h,d) ],  out row: [ (h,d) ].
 *   row_mask (B,N) fp32 or NULL: rows with mask == 0 get EVERY score set to mask_fill before the
 *   softmax (the reference's query-row mask, bifold/models/transformer.py:33-35) => uniform
 *   attention over all keys, and no gradient to q/k from that row.
 * Replaces HF:275-311 (SDPA, no mask) and bifold/models/transformer.py:26-41.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, N, H, hd;
  const void* qkv; int64_t ld_qkv;
  void* out; int64_t ld_out;
  float* lse;                        /* [B,H,N] log-sum-exp of the scaled scores (for backward) */
  const float* row_mask;
  float scale; float mask_fill;
  int32_t dtype;                     /* BFK_F32 | BFK_BF16 */
  int32_t backend;
  /* backward only */
  const void* dout; int64_t ld_dout;
  void* dqkv; int64_t ld_dqkv;
  float* delta;                      /* [B,H,N] workspace: rowsum(dout*out) */
} bfk_attention_args;
BFK_API int32_t bfk_attention_fwd(const bfk_attention_args* a, bfk_stream_t s);
BFK_API int32_t bfk_attention_bwd(const bfk_attention_args* a, bfk_stream_t s);

/* Cross attention of the alternate fusion module `crossattention` (replaces the torch.nn.MultiheadAttention call of
 * bifold/models/fusion.py:117-124): on the packed projection [B*Nf, 3*H*hd] = [q | k | v] of all fusion tokens, the rows
 * [q0, q0+nq) of every sample attend to the rows [k0, k0+nk).  out / dq are written for the query rows, dk / dv for the key
 * rows (their k and v slots); everything else in out / dqkv is left untouched (the caller zero-fills).
 * key_bias (optional, fp32 [B, nk]) is ADDED to the scaled scores; problem (b, h) reads row (b*H + h) % B - the reference's
 * float attn_mask with its tiled batch axis (fusion.py:107-113). */
typedef struct {
  int32_t B, Nf, H, hd;
  int32_t q0, nq, k0, nk;
  int32_t dtype;                     /* BFK_F32 | BFK_BF16 (qkv, out, dout, dqkv) */
  const void* qkv; int64_t ld_qkv;
  float scale;                       /* head_dim^-0.5 */
  const float* key_bias;             /* or NULL */
  void* out; int64_t ld_out;         /* [B*Nf, H*hd] */
  float* lse;                        /* [B, H, nq] */
  const void* dout; int64_t ld_dout; /* backward */
  void* dqkv; int64_t ld_dqkv;
  float* delta;                      /* [B, H, nq] workspace */
} bfk_cross_attention_args;
BFK_API int32_t bfk_cross_attention_fwd(const bfk_cross_attention_args* a, bfk_stream_t s);
BFK_API int32_t bfk_cross_attention_bwd(const bfk_cross_attention_args* a, bfk_stream_t s);
/* weights[b, i, j] = mean over the heads of softmax(scores)[b, h, i, j] (fp32 [B, nq, nk]): the second return value of
 * MultiheadAttention(need_weights=True, average_attn_weights=True), fusion.py:117-124; needs the forward's lse. */
BFK_API int32_t bfk_cross_attention_weights(const bfk_cross_attention_args* a, float* weights, bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Embeddings and token assembly.
 * ---------------------------------------------------------------------------------------------- */
/* im2col of non-overlapping p x p patches: out[(b*g*g + py*g + px), c*p*p + i*p + j] =
 * rgb[b,c,py*p+i,px*p+j]  (matches conv weight (D,3,p,p).flatten(1)); HF:178-179. */
BFK_API int32_t bfk_im2col(const float* rgb, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                   void* out, int64_t ld_out, int32_t out_dtype, bfk_stream_t s);
/* out[b*L + t, :] = tok_emb[ids[b,t], :] + pos_emb[t, :]   (HF:217-224); ids int64. */
BFK_API int32_t bfk_text_embed(const int64_t* ids, int32_t B, int32_t L, int32_t D, int32_t vocab,
                       const float* tok_emb, const float* pos_emb, float* out, bfk_stream_t s);

/* LoRA down-projection on the LayerNorm output (peft LoRA; bifold/models/siglip.py:33-40):
 *   h[m, D + j] = sum_k drop(h[m,k]) * Acat[j,k],  j < 16  (rows of Acat beyond 2r are zero).
 * The 16 extra columns ride along the K dimension of the QKV GEMM whose weight is [W | scale*B].
 * dropout_p > 0 applies an inverted-dropout mask from a counter-based hash of (seed, m, k). */
BFK_API int32_t bfk_lora_down(void* h, int64_t ldh, int32_t dtype, int32_t M, int32_t D, const void* Acat,
                      float dropout_p, uint64_t seed, bfk_stream_t s);
/* backward of the above w.r.t. h:  dh[m,k] += mask(m,k) * sum_j dh[m, D+j] * Acat[j,k] */
BFK_API int32_t bfk_lora_down_bwd(void* dh, int64_t ldh, int32_t dtype, int32_t M, int32_t D, const void* Acat,
                          float dropout_p, uint64_t seed, bfk_stream_t s);
/* dst[m,k] = keep(seed,m,k) ? src[m,k]/(1-p) : 0 — the dropped LoRA input (same counter-based mask as above). */
BFK_API int32_t bfk_dropout_copy(const void* src, int64_t lds, void* dst, int64_t ldd, int32_t dtype, int32_t M, int32_t D,
                                 float dropout_p, uint64_t seed, bfk_stream_t s);
/* LoRA factor gradients of one tower layer on the tensor cores (bf16 operands), replacing the autograd of peft's
 * lora_A / lora_B Linear pair on q_proj and v_proj (rank r <= 8, packed layout: q factors in columns 0..r-1, v factors
 * in columns 8..8+r-1 of the 16-column blocks):
 *   dBq[d, j] = scale * sum_m dqkv[m, d]      * u[m, j]          dBv[d, j] = scale * sum_m dqkv[m, 2D + d] * u[m, 8 + j]
 *   dAq[j, c] =         sum_m du[m, j]        * x[m, c]          dAv[j, c] =         sum_m du[m, 8 + j]    * x[m, c]
 * Two split-K GEMMs leave fp32 partials in `workspace` (bfk_lora_grads_workspace_bytes(D)), one kernel reduces them in a
 * fixed order (deterministic) into the four fp32 outputs (overwritten). */
typedef struct {
  int32_t M, D, r;
  const void* dqkv; int64_t ld_dqkv;  /* [M, 3D] bf16 gradient of the packed q|k|v projections */
  const void* u;    int64_t ld_u;     /* [M, 16]  bf16 A drop(h) */
  const void* du;   int64_t ld_du;    /* [M, 16]  bf16 gradient w.r.t. u */
  const void* x;    int64_t ld_x;     /* [M, D]   bf16 drop(h), the LoRA branch input */
  float scale;                        /* lora_alpha / r */
  float* dBq; float* dBv;             /* [D, r] */
  float* dAq; float* dAv;             /* [r, D] */
  void* workspace; int64_t workspace_bytes;
} bfk_lora_grads_args;
BFK_API size_t bfk_lora_grads_workspace_bytes(int32_t D);
BFK_API int32_t bfk_lora_grads(const bfk_lora_grads_args* a, bfk_stream_t s);

/* Tall-skinny Gram product used for the LoRA weight gradients (deterministic split over rows):
 *   out[i, j] = alpha * sum_m P[m*ldp + i] * mask?(Q[m*ldq + j]),  i < ni (<=16), j < nj.
 * `partials`: bfk_skinny_gram_workspace_floats(M, ni, nj). */
BFK_API size_t bfk_skinny_gram_workspace_floats(int32_t M, int32_t ni, int32_t nj);
BFK_API int32_t bfk_skinny_gram(const void* P, int64_t ldp, int32_t ni, const void* Q, int64_t ldq, int32_t nj,
                        int32_t dtype, int32_t M, float alpha, float* out, int64_t ldo, int32_t transpose_out,
                        float dropout_p, uint64_t seed, float* partials, bfk_stream_t s);

/* Fusion input assembly (bifold/models/siglip.py:62-68,91-130 + type embedding add
 * bifold/models/fusion.py:46-56):  X[b, r, :] for r in [0, Nf),
 *   Nf = (L+1) + (T+1)*(Np+1);  text rows get type 0, context and image rows type 1;
 *   context rows also get ctx_pos.  mask_out (B,Nf) fp32 is the attention-mask vector (or NULL). */
typedef struct {
  int32_t B, L, Np, T, D;
  const float* txt;        /* [B, L, D]   text tower output                     */
  const float* img;        /* [B, Np, D]  vision tower output of the frame      */
  const float* ctx;        /* [B*T, Np, D] vision tower output of context (T>0) */
  const float* text_token; const float* image_token;   /* [D] */
  const float* type_emb;   /* [2, D] */
  const float* ctx_pos;    /* [T*(Np+1), D] (T>0) */
  const int64_t* ctx_mask; /* [B, T] int64 (T>0) */
  float* X;                /* [B, Nf, D] */
  float* mask_out;         /* [B, Nf] or NULL */
} bfk_assemble_args;
BFK_API int32_t bfk_assemble_fwd(const bfk_assemble_args* a, bfk_stream_t s);
/* backward: dX [B,Nf,D] -> dtxt/dimg/dctx (any may be NULL) and parameter grads (fp32, overwritten). */
typedef struct {
  int32_t B, L, Np, T, D;
  const float* dX;
  float* dtxt; float* dimg; float* dctx;
  float* d_text_token; float* d_image_token; float* d_type_emb; float* d_ctx_pos;
  float* partials;         /* workspace: 2*64*D floats */
} bfk_assemble_bwd_args;
BFK_API int32_t bfk_assemble_bwd(const bfk_assemble_bwd_args* a, bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Heatmap tail.  The reference ConvDecoder (bifold/models/modules.py:33-74) is five 1x1 convs with
 * x2 bilinear upsamples between them and NO non-linearity, so each head is exactly
 *   logit = Up16( w_eff . token + b_eff ),  w_eff = W5 W4 W3 W2 W1,  Up16 = four x2 bilinear stages.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t heads, D;                   /* channels: D -> D/2 -> D/2 -> D/4 -> D/4 -> 1 (int division) */
  const float* const* W;              /* DEVICE array [heads*5] of weight pointers (cout,cin) row-major */
  const float* const* b;              /* DEVICE array [heads*5] of bias pointers */
  float* w_eff; float* b_eff;         /* [heads, D], [heads] */
  float* chain;                       /* [heads, 4, D] saved row vectors v2..v5 (for backward) */
  /* backward */
  const float* d_w_eff; const float* d_b_eff;
  float* const* dW; float* const* db; /* DEVICE arrays [heads*5] of gradient pointers (overwritten) */
} bfk_decoder_collapse_args;
BFK_API int32_t bfk_decoder_collapse_fwd(const bfk_decoder_collapse_args* a, bfk_stream_t s);
BFK_API int32_t bfk_decoder_collapse_bwd(const bfk_decoder_collapse_args* a, bfk_stream_t s);

/* grid[b, h, n] = w_eff[h] . X[b*x_batch_stride + (row0 + n)*D ...] + b_eff[h]
 * (bifold/models/pickplace.py:49-57,71 token -> grid map + collapsed decoder). */
typedef struct {
  int32_t B, heads, Np, D;
  const float* X; int64_t x_batch_stride; int32_t row0;   /* fp32 fused tokens */
  const float* w_eff; const float* b_eff;
  float* grid;                       /* [B, heads, Np] */
  /* backward */
  const float* dgrid;                /* [B, heads, Np] */
  int32_t Nf;                        /* rows per sample of dX */
  float* dX;                         /* [B, Nf, D] fully written (zeros outside the Np token rows) */
  void* dX2; int32_t dX2_dtype;      /* optional copy */
  float* d_w_eff; float* d_b_eff;    /* [heads, D], [heads] overwritten */
  float* partials;                   /* bfk_head_project_bwd_workspace_floats(B, heads, Np, D) */
} bfk_head_project_args;
BFK_API int32_t bfk_head_project_fwd(const bfk_head_project_args* a, bfk_stream_t s);
BFK_API size_t bfk_head_project_bwd_workspace_floats(int32_t B, int32_t heads, int32_t Np, int32_t D);
BFK_API int32_t bfk_head_project_bwd(const bfk_head_project_args* a, bfk_stream_t s);

/* Fills the 1-D taps of the composite x16 upsample (four x2 bilinear stages, align_corners=False;
 * bifold/models/modules.py:68-69): for output index y in [0, 16g): first input index and 3 weights.
 * HOST pointers.  Returns BFK_E_UNSUPPORTED if some output needs more than 3 taps. */
BFK_API int32_t bfk_upsample16_taps(int32_t g, int32_t* tap_start, float* tap_w /* [16g,3] */);

/* heat[h][b,y,x] = sigmoid(Up16(grid[b,h])) ; optional fused BCE (nn.BCELoss semantics,
 * bifold/losses/__init__.py:61,80-83) and first-max argmax (bifold/models/utils.py:66).
 * Replaces bifold/models/pickplace.py:75-106 + modules.py:57-74. */
typedef struct {
  int32_t B, heads, g;               /* H = W = 16 g */
  const float* grid;                 /* [B, heads, g*g] */
  const int32_t* tap_start; const float* tap_w;  /* DEVICE copies of bfk_upsample16_taps */
  float* heat;                       /* [heads, B, H, W] */
  const float* const* target;        /* DEVICE array [heads] of (B,H,W) fp32 target pointers, or NULL */
  const float* target_mask;          /* (B,H,W) fp32 multiplied into targets of heads flagged below, or NULL */
  uint32_t mask_heads;               /* bit h set: target_h *= target_mask */
  float* loss;                       /* [heads+1]: mean BCE per head, then their sum (if target) */
  unsigned long long* argmax_key;    /* [heads, B] workspace+result: (float bits << 32) | ~index, or NULL */
  float* partials;                   /* bfk_heatmap_fwd_workspace_floats(B, heads, g) floats, 8-byte aligned (required with target;
                                      * with argmax_key only: optional, NULL = atomic keys) */
  int32_t* counters;                 /* optional DEVICE int32 [>= 1], zero before the FIRST call and left zero by every call: the
                                      * last CTA then also reduces the per-band partials (one launch instead of three).  Must not
                                      * be shared by calls that can run concurrently.  NULL = separate finalize launch. */
} bfk_heatmap_fwd_args;
BFK_API int32_t bfk_heatmap_bands(int32_t g);   /* bands per map (loss partials per map) */
BFK_API size_t bfk_heatmap_fwd_workspace_floats(int32_t B, int32_t heads, int32_t g);
BFK_API int32_t bfk_heatmap_fwd(const bfk_heatmap_fwd_args* a, bfk_stream_t s);

/* dgrid from either the fused BCE gradient  dlogit = g_h*(p - t)         (mode 0)
 * or a generic upstream gradient             dlogit = dheat * p*(1-p)     (mode 1). */
typedef struct {
  int32_t B, heads, g, mode;
  const float* heat;                 /* [heads, B, H, W] */
  const float* const* target; const float* target_mask; uint32_t mask_heads;   /* mode 0 */
  const float* gscale;               /* DEVICE [heads (+1)]: dLoss/dloss_h (+ dLoss/dtotal)  (mode 0) */
  float gscale_mul; int32_t gscale_has_total; /* dlogit = gscale_mul*(gscale[h] (+ gscale[heads]))*(p - t) */
  const float* const* dheat;         /* DEVICE array [heads] of (B,H,W) pointers (mode 1); NULL entry = zero */
  const int32_t* tap_start; const float* tap_w;
  float* dgrid;                      /* [B, heads, g*g] */
  float* partials;                   /* bfk_heatmap_bwd_workspace_floats(B, heads, g) floats */
  int32_t* counters;                 /* optional DEVICE int32 [>= 1 + B*heads], zero before the first call and left zero by every
                                      * call (same buffer as bfk_heatmap_fwd_args.counters): the last band of a map sums the
                                      * bands.  NULL = separate second-stage launch. */
} bfk_heatmap_bwd_args;
BFK_API size_t bfk_heatmap_bwd_workspace_floats(int32_t B, int32_t heads, int32_t g);
BFK_API int32_t bfk_heatmap_bwd(const bfk_heatmap_bwd_args* a, bfk_stream_t s);

/* Standalone nn.BCELoss on probabilities (used when the caller only has heatmaps):
 * loss[0] = mean( -[t*max(log p,-100) + (1-t)*max(log1p(-p),-100)] ). partials: >= 1024 floats. */
BFK_API int32_t bfk_bce_prob(const float* p, const float* t, const float* t_mask, int64_t n, float* loss,
                     float* partials, bfk_stream_t s);
/* dp = gscale[0] * (p - t) / max(p*(1-p), 1e-12)   (torch binary_cross_entropy_backward) */
BFK_API int32_t bfk_bce_prob_bwd(const float* p, const float* t, const float* t_mask, int64_t n,
                         const float* gscale, float* dp, bfk_stream_t s);

/* First-max argmax per map: maps [n_maps, hw] fp32 -> key[n_maps] as in bfk_heatmap_fwd.
 * Replaces the D2H + numpy loop of bifold/models/utils.py:60-77. */
BFK_API int32_t bfk_argmax2d(const float* maps, int32_t n_maps, int32_t hw, unsigned long long* key, bfk_stream_t s);
/* get_action tail on device (bifold/models/utils.py:80-83, bifold/models/__init__.py:108-141):
 * decode keys, snap pick heads to the nearest mask pixel (first in row-major order among ties),
 * confidence gate.  heat [heads,B,H,W]; mask (B,H,W) fp32 or NULL; pick_heads/place_heads: indices
 * of the (left,right) pick / place heads (n_arms = 1 or 2).  out_xy int32 [heads,B,2] ([x,y], -1 for
 * a dropped arm), out_conf fp32 [heads,B]. */
typedef struct {
  int32_t B, heads, H, W, n_arms;
  const float* heat; const unsigned long long* key; const float* mask;
  int32_t pick_heads[2]; int32_t place_heads[2];
  float threshold;
  int32_t* out_xy; float* out_conf;
} bfk_action_args;
BFK_API int32_t bfk_action(const bfk_action_args* a, bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Fused Adam / AdamW over a flat fp32 buffer (torch.optim.Adam/AdamW semantics,
 * bifold/optim/__init__.py:9,12): g' = grad_scale*g (+ wd*p if !decoupled); p *= 1-lr*wd if
 * decoupled; m,v update; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  Optionally writes a bf16
 * shadow of the updated parameters.  `step` is the 1-based step count AFTER increment.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t n;
  float* p; const float* g; float* m; float* v;
  void* shadow_bf16;                 /* or NULL */
  float lr, beta1, beta2, eps, weight_decay; int32_t decoupled;
  int32_t step; float grad_scale;
  /* optional DEVICE float[3] = {lr, 1 - beta1^step, sqrt(1 - beta2^step)} that overrides lr / step when the kernel runs
   * (written by bfk_adamw_advance): the step count and learning rate of a captured CUDA graph advance on the device */
  const float* dyn;
} bfk_adamw_args;
BFK_API int32_t bfk_adamw(const bfk_adamw_args* a, bfk_stream_t s);
/* Multi-tensor strided cast: every segment copies a contiguous fp32 [rows, cols] matrix into dst[r * ld_dst + c] =
 * scale * src[r * cols + c] (dst_dtype BFK_F32 / BFK_BF16).  One launch refreshes all operand-dtype copies of the trainable
 * weights after an optimizer step (fusion weights -> bf16; LoRA factors -> the [W | (alpha/r) B] / [A_q; A_v] packs).
 * `segs` is a DEVICE array; first_chunk = running sum of ceil(rows*cols / 1024) over the previous segments. */
typedef struct {
  const float* src; void* dst;
  int64_t ld_dst, first_chunk;
  int32_t rows, cols, dst_dtype;
  float scale;
} bfk_cast_segment;
BFK_API int32_t bfk_multi_cast(const bfk_cast_segment* segs, int32_t nseg, int64_t total_chunks, bfk_stream_t s);

/* ++*step_dev;  dyn = {*lr_dev, 1 - beta1^step, sqrt(1 - beta2^step)}  (one thread, double precision like the host path) */
BFK_API int32_t bfk_adamw_advance(int32_t* step_dev, const float* lr_dev, float beta1, float beta2, float* dyn, bfk_stream_t s);
/* fp32 -> bf16 cast (shadow weights), optional transpose of a [rows, cols] matrix. */
BFK_API int32_t bfk_cast_bf16(const float* src, void* dst, int64_t n, bfk_stream_t s);
BFK_API int32_t bfk_cast_f32_to(const float* src, void* dst, int32_t dst_dtype, int64_t rows, int64_t cols,
                        int64_t ld_src, int64_t ld_dst, bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * GPU-side input stage (the per-sample host loops of bifold/data/processor.py that dominate host->device bytes).
 * bfk_gauss_targets: Processor._make_gaussmap (bifold/data/processor.py:405-425), evaluated in fp64 and rounded to fp32:
 *   centers int32 [n_maps, n_points, 2] = [cx, cy] (already rounded like round()/np.round); a point with cx == INT32_MIN is
 *   padding.  One valid point: strategy "first", out = exp(-((x-cx)^2 + (y-cy)^2) / (2 sigma^2)).  Several: strategy "gmm",
 *   sum of Gaussians / (2 pi sigma^2) divided by the map maximum.  No valid point ("no action for this arm",
 *   processor.py:296): all zeros.  out fp32 [n_maps, H, W].
 * bfk_normalize_u8: Processor._process_rgb through the SigLIP image processor (bifold/data/processor.py:143-148) for frames
 *   already at model resolution: uint8 [n, H, W, 3] -> fp32 [n, 3, H, W] = (x / 255 - mean) / std.
 * ---------------------------------------------------------------------------------------------- */
BFK_API int32_t bfk_gauss_targets(const int32_t* centers, int32_t n_maps, int32_t n_points, int32_t H, int32_t W, float sigma,
                          float* out, bfk_stream_t s);
BFK_API int32_t bfk_normalize_u8(const uint8_t* frames, int64_t n_frames, int32_t H, int32_t W, float mean, float std_, float* out,
                         bfk_stream_t s);

/* ------------------------------------------------------------------------------------------------
 * Mask branch of PickPlaceConvDecoder (compute_mask=True) and the mask losses.
 * bfk_mask_gate_fwd: pick heatmaps = sigmoid(pick) * mask_heatmap (bifold/models/pickplace.py:78-101,111-122);
 *   heads whose bit is clear in `pick_heads` are copied.  bfk_mask_gate_bwd: the product rule, with the mask side of it
 *   dropped when detach_mask is set (pickplace.py:80-89).  B*H*W must be a multiple of 4.
 * bfk_dice_loss_*:  DiceLoss  (bifold/losses/__init__.py:126-136): sum_b 1 - (2 sum(p t) + 1) / (sum p + sum t + 1).
 * bfk_focal_loss_*: FocalLoss (bifold/losses/__init__.py:139-155): scale * sum alpha_t * bce(p, t) * (1 - p_t)^gamma with
 *   nn.BCELoss's -100 log clamp; the reference's `loss.mean(1).sum()` is scale = 1 / H.  alpha < 0 disables alpha_t.
 * bfk_add_inplace: dst += src (the mask head's gradient w.r.t. the fusion INPUT joins the fusion backward's).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, heads, H, W;
  uint32_t pick_heads;               /* bit h set: head h is gated by the mask */
  const float* heat_raw;             /* [heads, B, H, W] sigmoid outputs */
  const float* mask_p;               /* [B, H, W] mask heatmap */
  float* heat_out;                   /* [heads, B, H, W] */
} bfk_mask_gate_args;
BFK_API int32_t bfk_mask_gate_fwd(const bfk_mask_gate_args* a, bfk_stream_t s);
typedef struct {
  int32_t B, heads, H, W;
  uint32_t pick_heads; int32_t detach_mask;
  const float* dheat_out;            /* [heads (+1), B, H, W] upstream gradient; map `heads` = d mask_heatmap if has_dmask */
  int32_t has_dmask;
  const float* heat_raw; const float* mask_p;
  float* dheat_raw;                  /* [heads, B, H, W] gradient w.r.t. the sigmoid outputs */
  float* dmask;                      /* [B, H, W] total gradient w.r.t. the mask heatmap */
} bfk_mask_gate_bwd_args;
BFK_API int32_t bfk_mask_gate_bwd(const bfk_mask_gate_bwd_args* a, bfk_stream_t s);
BFK_API int32_t bfk_add_inplace(float* dst, const float* src, int64_t n, bfk_stream_t s);
/* p, t: [B, n] fp32.  sums [B, 3] = {sum p t, sum p, sum t} (saved for the backward), loss [1]. */
BFK_API size_t bfk_dice_loss_workspace_floats(int32_t B);
BFK_API int32_t bfk_dice_loss_fwd(const float* p, const float* t, int32_t B, int64_t n, float* sums, float* loss, float* partials,
                          bfk_stream_t s);
BFK_API int32_t bfk_dice_loss_bwd(const float* t, int32_t B, int64_t n, const float* sums, const float* g /* DEVICE [1] */, float* dp,
                          bfk_stream_t s);
/* p, t: n fp32 elements; partials: >= 1024 floats; g: DEVICE [1] upstream gradient. */
BFK_API int32_t bfk_focal_loss_fwd(const float* p, const float* t, int64_t n, float alpha, float gamma, float scale, float* loss,
                           float* partials, bfk_stream_t s);
BFK_API int32_t bfk_focal_loss_bwd(const float* p, const float* t, int64_t n, float alpha, float gamma, float scale, const float* g,
                           float* dp, bfk_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* BIFOLD_B200_H_ */
