/* spe_b200.h -- C ABI of libspe_b200.so: the sm_100a hot path of MingXiangL/SPE.
 *
 * The reference (pure Python/PyTorch) has no FFI layer; its boundary for this path is
 * the Python module API of models/{cait,transformer,attention,matcher,conditional_detr,
 * position_encoding}.py and util/box_ops.py.  The Python shells in spe_b200/models mirror that API
 * and call ONLY the entry points below (ctypes).  Each entry point cites the
 * reference call site(s) it replaces.
 *
 * Conventions
 *   - every pointer is a BORROWED device pointer (caller allocates inputs, outputs, workspaces);
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises the host;
 *   - return value 0 = ok, <0 = error; spe_last_error() returns a thread-local message;
 *   - bf16 tensors are raw uint16 storage; "f32" = float; shapes are row-major unless stated;
 *   - no CPU fallback exists: without a CUDA device every compute entry point returns an error.
 */
#ifndef SPE_B200_H
#define SPE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* spe_last_error(void);
int spe_version(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t spe_launch_count(void);
/* device-wide L1 / shared-memory split preference: 1 = prefer shared memory (no carve-out reconfiguration between the ~200 KB tcgen05
 * kernels and the small row-wise kernels), 0 = driver default */
int spe_set_cache_config(int mode);
/* Optional device-side timing of kernel families with CUDA events on the launching stream (bench.py roofline).
 * Families: 0 gemm (work = algorithmic flops), 1 talking-softmax fwd, 2 talking-softmax bwd, 3 softmax,
 * 4 layernorm (work = algorithmic bytes), 5 matcher LSAP (work = images), 6 other, 7 batched attention GEMMs (QK^T / PV and
 * gradients; work = algorithmic bytes: they are bound by the N^2 operand in HBM). */
int spe_prof_enable(int on);
int spe_prof_collect(double* ms, double* work, int64_t* launches);   /* arrays of spe_prof_family_count() */
int spe_prof_family_count(void);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM:  C[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * B[b][n][k] )
 * bf16 operands staged by TMA (128B swizzle) into shared memory, fp32 accumulation in TMEM.
 * Replaces every nn.Linear / torch.bmm / `@` on the path: cait.py:376,379,389,390 (qkv, QK^T, PV,
 * proj), timm Mlp fc1/fc2, cait.py:115-133 (class attention), transformer.py:280-287,368-423
 * (in/out projections, FFN), attention.py:345,372 (bmm), conditional_detr.py:103-110 (heads),
 * and their backward (dgrad: B operand MN-major = the same weight; wgrad: both MN-major).
 * ------------------------------------------------------------------------------------------- */
enum { SPE_MAJOR_K = 0, SPE_MAJOR_MN = 1 };
enum { SPE_DT_BF16 = 0, SPE_DT_F32 = 1, SPE_DT_F16 = 2 };   /* F16: plain / alpha-scaled output only (the talking-heads logits) */
enum { SPE_ACT_NONE = 0, SPE_ACT_RELU = 1, SPE_ACT_GELU = 2,
       SPE_ACT_RELU_GRAD = 3,   /* out = v * (aux_in > 0)           */
       SPE_ACT_GELU_GRAD = 4 }; /* out = v * gelu'(aux_in)          */

typedef struct {
    int M, N, K;
    int batch1, batch2;               /* batch index b = b1 * batch2 + b2 (e.g. image x head)      */
    const void* A; int a_major;       /* K-major: (m,k) at m*lda+k;  MN-major: (m,k) at k*lda+m    */
    int64_t lda, a_sb1, a_sb2;        /* strides in elements; batch stride 0 (batch>1) = broadcast */
    const void* B; int b_major;       /* K-major: (n,k) at n*ldb+k;  MN-major: (n,k) at k*ldb+n    */
    int64_t ldb, b_sb1, b_sb2;
    void* C; int c_dtype;             /* SPE_DT_BF16 / SPE_DT_F32                                  */
    int64_t ldc, c_sb1, c_sb2;
    float alpha;
    const float* bias;                /* [N] or NULL : v = alpha*acc + bias[n]                     */
    int act;                          /* SPE_ACT_*   : v = act(v)                                  */
    const void* aux_in;               /* bf16 [M,N] (ld = ld_aux), for *_GRAD activations          */
    void* aux_out;                    /* bf16 [M,N] (ld = ld_aux) or NULL: v after bias, BEFORE act */
    int64_t ld_aux;
    const float* gamma;               /* [N] or NULL : v = gamma[n]*v   (LayerScale, cait.py:414)  */
    const float* residual;            /* f32 [M,N] (ld = ldr) or NULL : v += residual[m][n].  residual == C (same pitch, f32, no gamma/act)
                                         means C += alpha A B (+bias): tiles are reduce-added in place (gradient accumulation) */
    int64_t ldr, r_sb1, r_sb2;
    int split, split_stride;          /* if split>0: dest column = (n/split)*split_stride + n%split */
} spe_gemm_args;

int spe_gemm(const spe_gemm_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-head attention forward (TMA -> tcgen05 QK^T -> softmax in registers -> tcgen05 PV; logits stay in TMEM):
 *   O[b,:,h] = softmax_keys( scale * (Q_h K_h^T [+ Q2_h K2_h^T]) + key_padding_mask ) V_h
 * Replaces nn.MultiheadAttention's core (transformer.py:280) and models/attention.py:345-378 -- incl. the conditional
 * cross-attention, whose per-head [content | position] concat (transformer.py:408-414) is the sum of two QK^T products.
 * Heads are packed along the feature dim: q [B,Lq,H*d], k [B,Lk,H*d], v [B,Lk,H*dv] (bf16; *_ld = token stride,
 * *_sb = image stride, in elements; head h starts at column h*d).  d, d2, dv: multiples of 16, <= 64.
 * mask u8 [B,Lk] (1 = padded key) or NULL.  out bf16 [B,Lq,H*dv].  Optional outputs: P bf16 [B,H,Lq,ldP] (normalised
 * probabilities, zero in the padding columns; what the backward GEMMs consume), lse f32 [B,H,Lq] = log2-domain
 * log-sum-exp of the scaled logits (p = 2^(scale*log2e*s - lse)).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int B, H, Lq, Lk;
    int d, d2, dv;                         /* per-head dims: QK segment 1, QK segment 2 (0 if q2 == NULL), V */
    const void* q; int64_t q_ld, q_sb;
    const void* k; int64_t k_ld, k_sb;
    const void* v; int64_t v_ld, v_sb;
    const void* q2; int64_t q2_ld, q2_sb;  /* optional second QK segment (NULL: none)                        */
    const void* k2; int64_t k2_ld, k2_sb;
    const uint8_t* mask;
    float scale;
    void* out; int64_t out_ld, out_sb;
    void* P; int64_t ldP;                  /* optional */
    float* lse;                            /* optional */
} spe_attention_args;

int spe_attention_fwd(const spe_attention_args* a, void* stream);

/* Attention backward GEMMs fused (attention.py:345-372 / cait.py:379-389 backward):  ONE pass over the N^2 tensors
 *   dV[b,:,h] = P[b,h]^T dO[b,:,h]      dK[b,:,h] = alpha dS[b,h]^T Q[b,:,h]      dQ[b,:,h] = alpha dS[b,h] K[b,:,h]
 * dS, P bf16 [B,H,Lq,ld] (zero in the columns >= Lk); q, k, dO, dq, dk, dv_out bf16 with heads packed in the feature dim
 * (*_ld token stride, *_sb image stride, elements).  dv_out == NULL: dQ and dK only (second QK segment of the conditional
 * cross-attention).  delta != NULL (f32 [B,H,Lq] = rowsum(dO o O), spe_attention_delta): the dS operand holds dP = dO V^T and
 * the softmax backward dS = P o (dP - delta) is applied to the staged tiles in shared memory (P required) -- no separate
 * softmax-backward pass over the N^2 tensors.  workspace: f32 [spe_attention_bwd_gemms_workspace(B,H,Lq,d)] (dQ
 * accumulation, zeroed here). */
typedef struct {
    int B, H, Lq, Lk, d, dv;
    const void* dS; const void* P; int64_t ld;
    const void* q; int64_t q_ld, q_sb;
    const void* k; int64_t k_ld, k_sb;
    const void* dO; int64_t do_ld, do_sb;
    float alpha;
    void* dq; int64_t dq_ld, dq_sb;
    void* dk; int64_t dk_ld, dk_sb;
    void* dv_out; int64_t dv_ld, dv_sb;
    float* workspace;
    const float* delta;
} spe_attention_bwd_args;

int64_t spe_attention_bwd_gemms_workspace(int B, int H, int Lq, int d);
int spe_attention_bwd_gemms(const spe_attention_bwd_args* a, void* stream);
/* Fused attention backward, recomputing flavour (standard attention): given q, k, v (+ q2, k2), dO and the forward's row statistics
 * lse (spe_attention_fwd) and delta (spe_attention_delta), produces dq, dk, dv (+ dq2, dk2) without reading or writing any N^2
 * tensor: per (128-key block, head, image) the logits and dP = dO V^T are recomputed into TMEM, P = 2^(scale*log2e*S - lse) and
 * dS = P o (dP - delta) go to shared memory as bf16 tiles, and dV += P^T dO, dK += dS^T Q, dQ += dS K run on them (tcgen05).
 * Layout conventions as spe_attention_fwd.  workspace: f32 [B * Lq * H * max(d, d2)]. */
typedef struct {
    int B, H, Lq, Lk, d, d2, dv;
    const void* q; int64_t q_ld, q_sb;
    const void* k; int64_t k_ld, k_sb;
    const void* v; int64_t v_ld, v_sb;
    const void* q2; int64_t q2_ld, q2_sb;
    const void* k2; int64_t k2_ld, k2_sb;
    const void* dO; int64_t do_ld, do_sb;
    const uint8_t* mask;
    float scale;
    const float* lse; const float* delta;
    void* dq; int64_t dq_ld, dq_sb;
    void* dk; int64_t dk_ld, dk_sb;
    void* dv_out; int64_t dv_ld, dv_sb;
    void* dq2; int64_t dq2_ld, dq2_sb;
    void* dk2; int64_t dk2_ld, dk2_sb;
    float* workspace;
} spe_attention_bwd2_args;

int spe_attention_bwd(const spe_attention_bwd2_args* a, void* stream);

/* delta[b,h,q] = sum_c dO[b,q,h*dv+c] O[b,q,h*dv+c]  (bf16 inputs with heads packed, f32 [B,H,Lq] out) */
int spe_attention_delta(const void* dO, const void* O, int B, int H, int Lq, int dv, int64_t do_ld, int64_t do_sb, int64_t o_ld,
                        int64_t o_sb, float* delta, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row-wise / elementwise kernels of the backbone + transformer
 * ------------------------------------------------------------------------------------------- */
/* LayerNorm over the last dim (cait.py:414-415 eps 1e-6; transformer.py:284,288,384,425,427 eps 1e-5).
 * x f32 [rows,D] -> y_bf16 and/or y_f32 (either may be NULL); saves mean/rstd [rows]. */
int spe_layernorm_fwd(const float* x, const float* w, const float* b, float eps, int64_t rows, int D,
                      void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream);
/* dx = (dres ? dres : 0) + LN'(dy) ; dy given as bf16 and/or f32 (summed); dw/db accumulated (+=) in f32 */
int spe_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* dres, const float* x,
                      const float* w, const float* mean, const float* rstd, int64_t rows, int D,
                      float* dx, float* dw, float* db, void* stream);

/* Talking-heads mix + softmax + mix (cait.py:381-386):
 *   L[g] = sum_h Wl[g,h] S[h] + bl[g];  P = softmax_j(L);  A[g] = sum_h Ww[g,h] P[h] + bw[g]
 * S f32 [B,H,Nq,ldS] (Nk valid columns) -> A bf16 same layout (ldA).  */
/* stats (may be NULL): f32 [B*Nq, H] receives log2-domain softmax normalisers (row max + log2 sum) for the backward */
int spe_talking_softmax_fwd(const float* S, void* A, const float* Wl, const float* bl, const float* Ww,
                            const float* bw, float* stats, int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldA, void* stream);
/* backward: dA bf16 -> dS bf16 (may alias dA); parameter grads accumulated (+=) */
int spe_talking_softmax_bwd(const float* S, const void* dA, void* dS, const float* Wl, const float* bl,
                            const float* Ww, const float* bw, const float* stats, int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldA,
                            float* dWl, float* dbl, float* dWw, float* dbw, float* workspace, int64_t workspace_floats,
                            void* stream);
int64_t spe_talking_softmax_bwd_workspace(int B, int H, int Nq, int Nk);
/* The same two kernels with the logits S kept in FP16 in HBM (spe_gemm with c_dtype = SPE_DT_F16): the head mix rounds S to fp16
 * anyway (single-pass fp16 mma.sync), so the forward is bit-identical, and the N^2 logit traffic (S write + 2 reads per layer) and
 * the saved-activation footprint of S halve.  Row-staged kernels only: query with spe_talking_s16_supported first. */
int spe_talking_s16_supported(int H, int Nk, int64_t ldS, int64_t ldA);
int spe_talking_softmax_fwd_s16(const void* S_f16, void* A, const float* Wl, const float* bl, const float* Ww,
                                const float* bw, float* stats, int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldA, void* stream);
int spe_talking_softmax_bwd_s16(const void* S_f16, const void* dA, void* dS, const float* Wl, const float* bl,
                                const float* Ww, const float* bw, const float* stats, int B, int H, int Nq, int Nk, int64_t ldS, int64_t ldA,
                                float* dWl, float* dbl, float* dWw, float* dbw, float* workspace, int64_t workspace_floats,
                                void* stream);

/* ---------------------------------------------------------------------------------------------
 * FUSED talking-heads attention (csrc/talking_fused.cu): the whole of Attention_talking_head.forward between the qkv and
 * the proj linears (cait.py:377-389) and its backward, with NO [B,H,N,N] tensor in HBM:
 *   O[b,:,g] = ( sum_h Ww[g,h] softmax_keys( sum_h' Wl[h,h'] scale Q_h' K_h'^T + bl[h] ) + bw[g] ) V_g
 * q / k / v: bf16 [B,N,H*dh] views (*_ld = token stride, *_sb = image stride, elements; head h at column h*dh), typically the
 * three thirds of the packed qkv projection.  Logits live in TMEM (tcgen05.mma, M = 64), the H x H head mixes run on
 * mma.sync from the tcgen05.ld fragments, the mixed probabilities feed tcgen05.mma P V from shared memory.
 * lse2 (f32 [B,H,N], out): log2-domain logsumexp of the mixed logits, the only thing the backward needs besides q, k, v.
 * Supported head geometry: spe_talking_fused_supported(H, dh) (H in {4, 8}, dh = 48 = every CaiT variant up to S); other
 * shapes take the unfused spe_gemm + spe_talking_softmax_* pipeline.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int B, H, N, dh;
    const void* q; int64_t q_ld, q_sb;
    const void* k; int64_t k_ld, k_sb;
    const void* v; int64_t v_ld, v_sb;
    const float *Wl, *bl, *Ww, *bw;       /* proj_l / proj_w weights [H,H] (out, in) and biases [H], f32 */
    float scale;                          /* qk scale (dh^-0.5) */
    void* out; int64_t out_ld, out_sb;    /* bf16 [B,N,H*dh] */
    float* lse2;                          /* f32 [B,H,N] */
    void* workspace; int64_t workspace_bytes;
} spe_talking_fused_args;
int spe_talking_fused_supported(int H, int dh);
int64_t spe_talking_fused_fwd_workspace(int B, int H, int N, int dh);
int spe_talking_fused_fwd(const spe_talking_fused_args* a, void* stream);
/* backward: dO bf16 [B,N,H*dh] -> dqkv bf16 [B,N,3*H*dh] (token stride dqkv_ld; dq | dk | dv thirds), dWl / dWw f32 [H,H]
 * ACCUMULATED (+=).  dbl is identically 0 (softmax is shift invariant) and dbw = sum_b colsum(dO) . colsum(V) is exact outside.
 * Recomputes the logits from q, k and lse2 in three launches (delta + dWw; dQ + dWl; dK + dV). */
typedef struct {
    int B, H, N, dh;
    const void* q; int64_t q_ld, q_sb;
    const void* k; int64_t k_ld, k_sb;
    const void* v; int64_t v_ld, v_sb;
    const void* dO; int64_t do_ld, do_sb;
    const float *Wl, *bl, *Ww, *bw;
    float scale;
    const float* lse2;                    /* f32 [B,H,N] from spe_talking_fused_fwd */
    void* dqkv; int64_t dqkv_ld;          /* bf16 [B*N, 3*H*dh], rows contiguous over (b, n) */
    float *dWl, *dWw;                     /* f32 [H,H], += */
    void* workspace; int64_t workspace_bytes;
} spe_talking_fused_bwd_args;
int64_t spe_talking_fused_bwd_workspace(int B, int H, int N, int dh);
int spe_talking_fused_bwd(const spe_talking_fused_bwd_args* a, void* stream);

/* Plain softmax over keys with optional key-padding mask (attention.py:363-371, nn.MultiheadAttention).
 * S f32 [B,H,Nq,ldS] -> P bf16 [B,H,Nq,ldP]; mask u8 [B,Nk] (1 = padded -> -inf) or NULL.
 * If pmean != NULL also writes the head-mean of P, f32 [B,Nq,Nk] (cait.py:658-667 cams). */
int spe_softmax_fwd(const float* S, void* P, const uint8_t* mask, int B, int H, int Nq, int Nk, int64_t ldS,
                    int64_t ldP, float* pmean, void* stream);
/* TSCAM_cait_two_branch.std_reweighting (cait.py:801-806, :827): CAMs from the per-head class-attention probabilities.
 * P bf16 [B,H,Lq,ldP];  out f32 [B,C,N]:  out[b,c,n] = sum_h P[b,h,q0+c,k0+n] * w[b,h,c],
 * w = (std_h - min_h std) / max_h(std - min_h std),  std_h = unbiased std over the N keys.  No gradient. */
int spe_cam_std_reweight(const void* P_bf16, int B, int H, int Lq, int64_t ldP, int q0, int C, int k0, int N, float* out,
                         void* stream);
/* dS = P * (dP - sum_j P dP), bf16 in / bf16 out (dS may alias dP) */
int spe_softmax_bwd(const void* P, const void* dP, void* dS, int B, int H, int Nq, int Nk, int64_t ldP, void* stream);

/* LayerScale branch backward (cait.py:414-415: x + gamma * y):
 * dy_bf16 = gamma * dout ; dgamma += sum_rows dout*y ; dbias += sum_rows dy  */
int spe_layerscale_bwd(const float* dout, const void* y_bf16, const float* gamma, int64_t rows, int D,
                       void* dy_bf16, float* dgamma, float* dbias, void* stream);
/* column sums of a bf16 [rows, N] (ld) matrix, accumulated into out f32 [N] (bias gradients) */
int spe_colsum_bf16(const void* x, int64_t rows, int N, int64_t ld, float* out, void* stream);
/* batched: x bf16 [batch][rows][N] (row pitch ld, batch pitch batch_stride) -> out f32 [batch][N] (accumulated) */
int spe_colsum_bf16_batched(const void* x, int batch, int64_t rows, int N, int64_t ld, int64_t batch_stride, float* out, void* stream);
/* y_bf16 = (a*x + b*y) with f32 inputs (y may be NULL); also optional f32 output */
int spe_axpby_cast(const float* x, const float* y, float a, float b, int64_t n, void* out_bf16, float* out_f32,
                   void* stream);
/* bf16 shadows of many fp32 tensors in ONE launch (the per-step refresh of every weight's GEMM operand after the optimizer
 * moved the fp32 masters): segs = device array of `count` {src f32*, dst bf16*, n} records, max_n = largest n. */
typedef struct spe_cast_seg { const float* src; void* dst; int64_t n; } spe_cast_seg;
int spe_cast_f32_to_bf16_multi(const spe_cast_seg* segs, int count, int64_t max_n, void* stream);
int spe_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream);
/* ReLU backward for an activation fused into a GEMM epilogue: out = dout * (h > 0), all bf16 (transformer.py:21-33 MLP) */
int spe_relu_bwd_bf16(const void* dout, const void* h, void* out, int64_t n, void* stream);
/* out[i] += x_bf16[i] (f32 accumulate) */
int spe_add_bf16_into_f32(const void* x, float* out, int64_t n, void* stream);

/* Patch embedding im2col (cait.py:527, Conv2d k=s=16): img f32 [B,3,H,W] -> bf16 [B*h*w, 3*p*p] */
int spe_im2col_patch(const float* img, int B, int H, int W, int p, void* out_bf16, void* stream);
/* col2im is not needed (images carry no gradient). */

/* Bicubic resize (align_corners=False, A=-0.75) of pos_embed (cait.py:588-613):
 * src f32 [sh*sw, D] token-major -> dst f32 [dh*dw, D]; bwd scatters dst-grad into src-grad (+=). */
int spe_bicubic_tokens_fwd(const float* src, int sh, int sw, int D, float* dst, int dh, int dw, void* stream);
int spe_bicubic_tokens_bwd(const float* ddst, int dh, int dw, int D, float* dsrc, int sh, int sw, void* stream);

/* PositionEmbeddingSine(normalize=True) (position_encoding.py:37-57): mask u8 [B,h,w] -> pos f32 [B,h*w,D] */
int spe_sine_pos_2d(const uint8_t* mask, int B, int h, int w, int D, float* pos, void* pos_bf16, void* stream);
/* gen_sineembed_for_position (transformer.py:35-49, /128 exponent): ref f32 [n,2] -> emb f32 [n,D];
 * bwd: dref[n,2] = sum_c demb * d emb/d ref */
int spe_query_sine_fwd(const float* ref, int64_t n, int D, float* emb, void* stream);
int spe_query_sine_bwd(const float* ref, const float* demb, int64_t n, int D, float* dref, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Matcher (models/matcher.py:41-87) -- cost matrix + batched rectangular LSAP on the GPU
 * ------------------------------------------------------------------------------------------- */
/* Block-diagonal cost only: for image b, cost[b][q][g], g < G_b = gt_off[b+1]-gt_off[b].
 * logits f32 [B,Q,C], boxes f32 [B,Q,4] cxcywh, gt_labels i32 [sumG], gt_boxes f32 [sumG,4],
 * gt_off i32 [B+1]; cost f32 [B,Q,ldc] (ldc >= max G_b).  fp32, reference op order, no FMA contraction.
 * gt_images > 0: the targets describe gt_images images and problem b uses image b % gt_images (the L decoder levels of
 * conditional_detr.py:444-452 stacked along the batch, B = L * gt_images, share one target set); 0: one image per problem. */
int spe_match_cost(const float* logits, const float* boxes, const int32_t* gt_labels, const float* gt_boxes,
                   const int32_t* gt_off, int B, int Q, int C, float w_class, float w_bbox, float w_giou,
                   float* cost, int64_t ldc, int gt_images, void* stream);
/* scipy.optimize.linear_sum_assignment (matcher.py:86) for B independent problems: cost f32 [B,nr,ldc],
 * problem b uses the first nc[b] columns (nc = NULL -> all ldc).  Output query_to_col i32 [B,nr]
 * (-1 = unassigned row; when nr <= nc[b] every row is assigned).  Identical indices to scipy incl. ties:
 * rows sorted ascending with their columns == scipy's (row_ind, col_ind).
 * workspace: spe_lsap_workspace_bytes(B, nr, max_nc) bytes. */
int64_t spe_lsap_workspace_bytes(int B, int nr, int max_nc);
int spe_lsap_batched(const float* cost, int B, int nr, int64_t ldc, const int32_t* nc, int max_nc,
                     int32_t* row_to_col, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Set criterion (models/conditional_detr.py:237-319, 468-494, 504-561)
 * ------------------------------------------------------------------------------------------- */
/* Focal classification loss + gradient + class_error + cardinality, from the dense assignment.
 *   logits f32 [B,Q,C]; row_to_gt i32 [B,Q] (-1 none); gt_labels i32[sumG]; gt_off i32[B+1];
 *   gt_scores f32[sumG] or NULL (SetCriterionRefine weights :523-529)
 *   inv_num_boxes: device f32 scalar (1/num_boxes, conditional_detr.py:436-440)
 * outputs: out[0]=loss_ce, out[1]=class_error, out[2]=cardinality_error (f32[3]);
 *          dlogits f32 [B,Q,C] = d loss_ce / d logits (may be NULL). */
int64_t spe_focal_loss_workspace_bytes(int B, int Q);
int spe_focal_loss(const float* logits, const int32_t* row_to_gt, const int32_t* gt_labels, const int32_t* gt_off,
                   const float* gt_scores, const float* inv_num_boxes, int B, int Q, int C, float alpha, float gamma,
                   float* out, float* dlogits, void* workspace, void* stream);
/* L1 + GIoU on matched pairs (diag only; conditional_detr.py:300-319 / :540-561).
 * out[0]=loss_bbox, out[1]=loss_giou; dboxes f32 [B,Q,4] receives w_l1*dL1 + w_giou*dGIoU?  No: two planes
 * dboxes_l1 and dboxes_giou (each [B,Q,4], zero on unmatched rows; may be NULL). */
int spe_box_loss(const float* boxes, const int32_t* row_to_gt, const float* gt_boxes, const int32_t* gt_off,
                 const float* gt_scores, const float* inv_num_boxes, int B, int Q,
                 float* out, float* dboxes_l1, float* dboxes_giou, void* stream);
/* Multi-label BCE-with-logits mean (conditional_detr.py:225-235): x f32 [n], y f32 [n] -> out[0]; dx = d/dx */
int spe_bce_logits(const float* x, const float* y, int64_t n, float* out, float* dx, void* stream);

/* Pairwise IoU / GIoU (util/box_ops.py:33-74): a f32 [N,4], b f32 [M,4] xyxy -> iou/giou/union f32 [N,M] */
int spe_box_iou_pairwise(const float* a, int N, const float* b, int M, float* iou, float* uni, float* giou, void* stream);

/* CAM -> pseudo ground-truth box: engine.get_pseudo_label (engine.py:310-352) = cams_deit.resize_cam (cams_deit.py:9-14)
 * + cams_deit.get_bboxes (:34-58), for npairs (image, class) pairs at once, bit-exact with the OpenCV path the reference calls.
 *   cams f32 [B,C,h,w]; pairs i32 [npairs,2] = (image, class) on the device; the map is resized to rows x cols (the reference passes
 *   dsize = (H_img, W_img), i.e. rows = W_img, cols = H_img), min-max normalised, quantised to u8, thresholded at u8 > thr_u8
 *   (= int(cam_thr * 255)), and the bounding box of the contour with the largest cv2.contourArea is returned as
 *   boxes_out f32 [npairs,4] = cxcywh / (norm_x, norm_y, norm_x, norm_y)  and (optional) xyxy_out i32 [npairs,4] = [x, y, x+w, y+h].
 *   workspace: spe_cam_boxes_workspace_bytes(npairs, rows, cols) bytes. */
int64_t spe_cam_boxes_workspace_bytes(int npairs, int rows, int cols);
int spe_cam_boxes(const float* cams, int B, int C, int h, int w, const int32_t* pairs, int npairs, int rows, int cols, int thr_u8,
                  float norm_x, float norm_y, float* boxes_out, int32_t* xyxy_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Multi-box variant (engine.get_pseudo_label_multi_boxes, engine.py:356-398 = cams_deit.get_multi_bboxes :61-97): every contour of the
 * thresholded map (outer borders at any nesting level and hole borders, cv2.RETR_TREE) whose cv2.contourArea is >= area_ratio * the
 * largest one, by decreasing area; at most max_boxes (<= 64) per pair.  boxes_out f32 [npairs,max_boxes,4], xyxy_out i32 (optional),
 * counts_out i32 [npairs] (>= 1: a map without contours yields [0,0,1,1]). */
int spe_cam_boxes_multi(const float* cams, int B, int C, int h, int w, const int32_t* pairs, int npairs, int rows, int cols, int thr_u8,
                        float norm_x, float norm_y, double area_ratio, int max_boxes, float* boxes_out, int32_t* xyxy_out,
                        int32_t* counts_out, void* workspace, int64_t workspace_bytes, void* stream);

/* GT jitter + repeat of SetCriterion.forward in training mode (models/conditional_detr.py:410-431), device side.
 *   in : boxes f32 [sumG,4] cxcywh, labels i32 [sumG], scores f32 [sumG] or NULL, offsets i32 [B+1] (CSR over images; sumG = offsets[B]
 *        is read on the device, cap_total >= sumG sizes the launch)
 *   out: ratio rows per GT box: the first min(ratio-1, #kept) of n_try candidates box * U(1-jitter, 1+jitter)^4 with IoU(candidate, box) >
 *        iou_thr, in candidate order, then the original box; labels / scores repeated; offsets_out = ratio * offsets; counts_out (may be
 *        NULL) = ratio * per-image counts.
 *   rng_state: device u64[2] {seed, launch counter} -- the kernel advances the counter, so graph replays draw fresh candidates;
 *   ticket: device u32 scratch, zero-initialised once. */
int spe_gt_jitter_repeat(const float* boxes, const int32_t* labels, const float* scores, const int32_t* offsets, int B, int cap_total,
                         int ratio, float jitter, int n_try, float iou_thr, uint64_t* rng_state, float* boxes_out, int32_t* labels_out,
                         float* scores_out, int32_t* offsets_out, int32_t* counts_out, uint32_t* ticket, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step on the flat fp32 buffers: clip_grad_norm_ (engine.py:163-164) + torch.optim.AdamW with the three
 * learning-rate groups of main.py:177-190.  No host synchronisation: step count, bias corrections, clip coefficient and the
 * learning rates are device scalars.
 * ------------------------------------------------------------------------------------------- */
/* out[0] = sum_i x[i]^2 (deterministic two-stage reduction); workspace: spe_sumsq_workspace_floats() floats */
int64_t spe_sumsq_workspace_floats(void);
int spe_sumsq_f32(const float* x, int64_t n, float* out, float* workspace, void* stream);
/* state f32[4] = {beta1^t, beta2^t, clip coefficient, t}: initialise to {1, 1, 1, 0}; tick advances t by one step and sets the clip
 * coefficient min(1, max_norm / (sqrt(grad_sumsq[0]) + 1e-6)) (1 when grad_sumsq is NULL or max_norm <= 0) */
int spe_adamw_tick(float* state, float beta1, float beta2, const float* grad_sumsq, float max_norm, void* stream);
#define SPE_ADAMW_MAX_SEGMENTS 16
typedef struct {
    float* p;                 /* parameters  f32[n]  (updated in place) */
    const float* g;           /* gradients   f32[n]  (multiplied by the clip coefficient on the fly) */
    float* m;                 /* exp_avg     f32[n] */
    float* v;                 /* exp_avg_sq  f32[n] */
    int64_t n;                /* multiple of 4 */
    int nseg;                 /* the buffer is nseg contiguous segments, each belonging to one parameter group */
    int64_t seg_end[SPE_ADAMW_MAX_SEGMENTS];   /* exclusive end offsets, increasing, multiples of 4, last == n */
    int32_t seg_group[SPE_ADAMW_MAX_SEGMENTS];
    const float* lr;          /* device f32[groups] */
    const float* wd;          /* device f32[groups] (decoupled weight decay) */
    float beta1, beta2, eps;
    const float* state;       /* device f32[4], see spe_adamw_tick */
    float* g_out;             /* optional: the clipped gradients (may alias g), what clip_grad_norm_ leaves in .grad */
    void* shadow_bf16;        /* optional: bf16 copy of the updated parameters, bf16[n] */
} spe_adamw_args;
int spe_adamw_flat(const spe_adamw_args* args, void* stream);
/* x *= clip coefficient (state[2]): clip_grad_norm_ without an optimizer step */
int spe_scale_by_clip_coef(float* x, int64_t n, const float* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPE_B200_H */
