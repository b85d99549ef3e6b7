/* libcrct_b200 — C ABI of the B200-native CRCT question-answering hot path.
 *
 * The reference (levymsn/CQA-CRCT) is pure Python and has no FFI; every arithmetic step below
 * `CRCT/backbone/encoder_decorator.py:73 forward()` is a torch op.  This header is therefore the
 * boundary a maintainer binds with `ctypes`: plain pointers and sizes, no
 * torch types.  Each entry point names the reference lines (relative to /root/reference/CRCT/)
 * whose arithmetic it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; this library never
 *    allocates, frees or synchronises); all work is enqueued on the given stream;
 *  - activations are bf16 row-major, statistics / parameters / gradients are fp32;
 *  - gradient outputs documented "+=" are ACCUMULATED (the caller zeroes the gradient arena once
 *    per step, like `optimizer.zero_grad()` at CRCT/train.py:214);
 *  - return value: 0 on success, a negative crct_status_t otherwise; `crct_last_error()` gives the
 *    thread-local message.  Nothing throws across the ABI.  There is no CPU fallback: a device
 *    that is not sm_100 is CRCT_ERR_ARCH.
 */
#ifndef CRCT_B200_H
#define CRCT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* crct_stream_t; /* cudaStream_t */

typedef enum {
    CRCT_OK = 0,
    CRCT_ERR_ARG = -1,   /* null pointer / misaligned / inconsistent sizes */
    CRCT_ERR_CUDA = -2,  /* a CUDA runtime or driver call failed */
    CRCT_ERR_ARCH = -3,  /* current device is not compute capability 10.x */
    CRCT_ERR_SHAPE = -4  /* shape outside what the kernels support */
} crct_status_t;

const char* crct_last_error(void);
int crct_version(void);
/* CRCT_OK iff the current CUDA device can run the sm_100a kernels. */
int crct_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * K1  tcgen05/TMEM GEMM fed by TMA:  D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
 * Replaces every nn.Linear on the transformer path and its autograd backward
 * (backbone/vilbert.py:373-375,388-390,420,446,463,502-504,551,577,594,637-646,732,739,1453).
 *   a_major = 0: A stored [M,K] (K contiguous, row stride lda)     a_major = 1: A stored [K,M] (M contiguous)
 *   b_major = 0: B stored [N,K] (K contiguous, row stride ldb)     b_major = 1: B stored [K,N] (N contiguous)
 *   forward  y = x W^T      : A = x  (0), B = W  (0)
 *   dgrad    dx = dy W      : A = dy (0), B = W  (1)         (GEMM-K = out features)
 *   wgrad    dW = dy^T x    : A = dy (1), B = x  (1)         (GEMM-K = rows), fp32 output
 * -------------------------------------------------------------------------------------------- */
typedef enum {
    CRCT_EPI_BIAS = 0,      /* D = acc + bias                                    (bias may be NULL) */
    CRCT_EPI_BIAS_GELU = 1, /* u = acc + bias; D = gelu_erf(u); D2 = gelu_erf'(u) (if D2)   vilbert.py:111-117,454-457 */
    CRCT_EPI_BIAS_RES = 2,  /* D = dropout(acc + bias) + aux                      vilbert.py:424-428,467-471,749-756 */
    CRCT_EPI_MUL = 3,       /* D = acc * aux   (aux = the D2 saved by BIAS_GELU)   backward of vilbert.py:456 */
    CRCT_EPI_F32 = 4,       /* D (fp32) = acc, or D += acc when accumulate != 0 (wgrad, split-K) */
    CRCT_EPI_BIAS_RES_F32 = 5 /* CRCT_EPI_BIAS_RES with fp32 D AND fp32 aux: the residual stream and the pre-LayerNorm sum z are
                               * never rounded to bf16 (the reference's autocast keeps them fp32 too: its LayerNorm runs in fp32) */
} crct_epilogue_t;

typedef struct {
    const void* A;      /* bf16 */
    const void* B;      /* bf16 */
    void* D;            /* bf16 [M,N] (fp32 for CRCT_EPI_F32), row stride ldd */
    void* D2;           /* bf16 [M,N] row stride ldd, CRCT_EPI_BIAS_GELU only, may be NULL */
    const float* bias;  /* fp32 [N] or NULL */
    const void* aux;    /* bf16 [M,N] row stride ldaux (residual / addend / saved activation derivative) or NULL */
    int32_t M, N, K;
    int32_t lda, ldb, ldd, ldaux; /* in elements */
    int32_t a_major, b_major;
    int32_t epilogue;   /* crct_epilogue_t */
    int32_t accumulate; /* CRCT_EPI_F32: 1 = atomically add into D */
    int32_t split_k;    /* >= 1; > 1 requires CRCT_EPI_F32 with accumulate */
    int32_t block_n;    /* 0 = choose; else 128, 192 (single-CTA tiles only) or 256 */
    float dropout_p;    /* CRCT_EPI_BIAS_RES: 0 = off */
    uint64_t seed;      /* dropout stream; element counter = m * N + n */
    int32_t max_ctas;   /* 0 = one persistent CTA per SM; else cap (tests) */
    int32_t cta_group;  /* 0 = choose; 1 = one CTA per 128 x BN tile; 2 = CTA pair (tcgen05 cta_group::2) per 256 x BN tile */
    int32_t dbg[7];     /* descriptor overrides for bring-up; must be 0 in production */
    const uint64_t* salt; /* optional DEVICE word XOR-ed into `seed` at run time (see crct_bump_salt), or NULL */
    const int32_t* a_rows_dev; /* optional DEVICE word: the number of valid ROWS of A as stored (var-len packing, see crct_row_map)
                                * — M for a_major = 0 (tiles past it are skipped, rows past it are not written), GEMM-K for the
                                * wgrad form a_major = b_major = 1 (rows past it contribute nothing).  The M / K fields are then
                                * the upper bounds the buffers were allocated for.  Not with cta_group = 2. */
    int32_t rows_hint;         /* with a_rows_dev: the caller's ESTIMATE of *a_rows_dev (0 = unknown).  Only steers the tile-shape
                                * choice (waves on 148 SMs); never read for correctness. */
    const int32_t* drop_rows;  /* optional DEVICE int32 [M]: the dropout element counter of output row m is drop_rows[m] * N + n
                                * instead of m * N + n (packed rows draw the masks of their padded positions: crct_row_map's src_row) */
} crct_gemm_t;

int crct_gemm_bf16(const crct_gemm_t* args, crct_stream_t stream);
/* Up to 8 independent weight-gradient problems (the wgrad form above: a_major = b_major = 1, CRCT_EPI_F32, accumulate = 1) in ONE
 * launch: one persistent tile list over all problems, per-problem split-K chosen so that the group is about two waves of
 * equal-cost tiles.  A layer's weight gradients (QKV, attention output, FFN in / out — backward of vilbert.py:388-390,420,446,463)
 * are not on the backward's critical path; issued together they share one pipeline fill / tail and fill the SMs that the small
 * (visual-stream) problems leave idle.  `split_k` of a problem > 0 overrides the choice; `a_rows_dev` / `rows_hint` as above. */
int crct_gemm_wgrad_grouped(const crct_gemm_t* problems, int count, crct_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4  row kernels (HBM-bound, one warp per row, fp32 statistics).  Row width H: multiple of 8, <= 1024.
 * -------------------------------------------------------------------------------------------- */
/* fp32 master weights -> bf16 operand arena (n multiple of 8).  Replaces autocast's per-op casts, train.py:172. */
int crct_cast_f32_to_bf16(const float* src, void* dst_bf16, size_t n, crct_stream_t stream);
/* out[i] = (1 - mask[i]) * -10000  (vilbert.py:1380-1396).  kind: 0 = bool/uint8, 1 = int64, 2 = fp32. */
int crct_additive_mask(const void* mask, int kind, float* out, int n, crct_stream_t stream);
/* Row counts on the device.  Kernels that take `rows_dev` (a DEVICE int32, or NULL) process min(rows, *rows_dev) rows;
 * `rows` then only sizes the grid.  This is how the var-len ("packed") layout runs inside a captured CUDA graph: the number
 * of valid token / region rows of a batch (crct_row_map) is never read by the host. */
/* y = LayerNorm(z) with eps = 1e-12 inside the sqrt (vilbert.py:281-294).  z bf16 (z_f32 = 0) or fp32 (z_f32 = 1)
 * [rows,H], y bf16 (the next GEMM's operand); y32 (or NULL): the same values in fp32 — the residual stream stays fp32 end to
 * end, the next CRCT_EPI_BIAS_RES_F32 epilogue adds this copy; mean/rstd fp32 [rows] or both NULL (inference). */
int crct_layernorm_fwd(const void* z, const float* gamma, const float* beta, void* y, float* y32, float* mean, float* rstd,
                       int rows, int H, int z_f32, const int32_t* rows_dev, crct_stream_t stream);
/* out[b,:] (fp32) = LayerNorm(z[row,:]) for row = row_index[b] (row_index NULL: row = b * row_step), z fp32: the
 * first-token / first-region states the heads read (vilbert.py:958,973,1599-1600) straight from the last pre-LayerNorm sum,
 * without the bf16 rounding of the sequence output. */
int crct_layernorm_rows_f32(const float* z, const float* gamma, const float* beta, const int32_t* row_index, long long row_step,
                            float* out, int B, int H, crct_stream_t stream);
/* LayerNorm backward.  dy may carry the dropout that FOLLOWED the LayerNorm in the forward (embeddings:
 * p_in, seed_in); dzm is dz with the dropout that PRECEDED the residual add re-applied (p_out, seed_out; element
 * counter row*H+col — the same stream the CRCT_EPI_BIAS_RES epilogue used), i.e. the gradient of the dense output. */
typedef struct {
    const void* dy;      /* bf16 [rows,H] */
    const void* z;       /* bf16 [rows,H] pre-LayerNorm input saved by the forward */
    const float* mean;   /* [rows] */
    const float* rstd;   /* [rows] */
    const float* gamma;  /* [H] */
    void* dz;            /* bf16 [rows,H] */
    void* dzm;           /* bf16 [rows,H] or NULL (ignored when p_out == 0: dzm == dz) */
    float* dgamma;       /* [H] += */
    float* dbeta;        /* [H] += */
    float* dbias;        /* [H] += column sums of dzm (dz when p_out == 0), or NULL */
    int32_t rows, H;
    float p_in;  uint64_t seed_in;
    float p_out; uint64_t seed_out;
    const uint64_t* salt; /* optional device word XOR-ed into both seeds */
    int32_t z_f32;        /* 1: z is fp32 (what CRCT_EPI_BIAS_RES_F32 / the embeddings with z_f32 wrote) */
    const int32_t* rows_dev;
    const int32_t* drop_rows; /* as in crct_gemm_t: row index used by both dropout counters */
} crct_ln_bwd_t;
int crct_layernorm_bwd(const crct_ln_bwd_t* args, crct_stream_t stream);
/* Split form: crct_layernorm_bwd with dgamma = dbeta = dbias = NULL computes only dz / dzm (the part the backward
 * chain waits for); this call then accumulates the three column sums from dy, z and the dzm (dz when p_out == 0)
 * that call wrote.  Same struct; gamma is not read. */
int crct_layernorm_bwd_params(const crct_ln_bwd_t* args, crct_stream_t stream);
/* out[n] += sum_rows x[row,n]   (bias gradients).  x bf16 [rows,N], row stride ld. */
int crct_colsum_bf16(const void* x, float* out, int rows, int N, int ld, const int32_t* rows_dev, crct_stream_t stream);
/* softmax over the RoI feature axis, fp32 in -> bf16 GEMM operand (vilbert.py:1476).  src_row (or NULL): output row r is
 * computed from input row src_row[r] (packed output from the padded [B*R,F] input). */
int crct_softmax_rows(const float* x, void* out_bf16, int rows, int F, const int32_t* src_row, const int32_t* rows_dev,
                      crct_stream_t stream);

/* K5  text embedding (vilbert.py:320-358): word + position (question/answer tokens only, counted from the first
 * such token) + plotqa type (-1 -> 0, none for type 0) + Linear(4->H)(box) (none, bias included, for an all-zero
 * box) -> LayerNorm -> dropout.  z/mean/rstd may be NULL for inference. */
typedef struct {
    const int64_t* ids;    /* [B,T] */
    const int64_t* types;  /* [B,T] in {-1,0..11} */
    const float* loc;      /* [B,T,4] */
    const float* word; const float* pos; const float* type;  /* fp32 tables [V,H] [P,H] [12,H] */
    const float* w_loc; const float* b_loc;                  /* [H,4] [H] */
    const float* gamma; const float* beta;
    void* y; void* z; float* mean; float* rstd;              /* bf16 [B*T,H] x2, fp32 [B*T] x2 */
    int32_t B, T, H, max_pos;
    float dropout_p; uint64_t seed;
    const uint64_t* salt;
    int32_t z_f32;            /* 1: z is written in fp32 */
    const int32_t* src_row;   /* packed layout: output row r is token src_row[r] = b*T + t (crct_row_map); NULL = all B*T rows */
    const int32_t* rows_dev;
    float* y32;               /* optional fp32 copy of y (residual stream), as in crct_layernorm_fwd */
} crct_embed_text_t;
int crct_embed_text_fwd(const crct_embed_text_t* args, crct_stream_t stream);
/* scatter of dz (after crct_layernorm_bwd) into the tables; all outputs += . */
typedef struct {
    const int64_t* ids; const int64_t* types; const float* loc;
    const void* dz;        /* bf16 [B*T,H] */
    float* g_word; float* g_pos; float* g_type; float* g_wloc; float* g_bloc;
    int32_t B, T, H;
    const int32_t* src_row; const int32_t* rows_dev;   /* as in crct_embed_text_t */
} crct_embed_text_bwd_t;
int crct_embed_text_bwd(const crct_embed_text_bwd_t* args, crct_stream_t stream);

/* K6  visual embedding tail (vilbert.py:1478-1496): g = softmax(feat) W_img^T + b_img comes from K1;
 * z = g + Linear(4->Hv)(box) + color_emb[class] -> LayerNorm -> dropout. */
typedef struct {
    const void* g;         /* bf16 [rows,H] */
    const float* box;      /* [rows,4] */
    const int64_t* cls;    /* [rows] */
    const float* w_loc; const float* b_loc; const float* color; const float* gamma; const float* beta;
    void* y; void* z; float* mean; float* rstd;
    int32_t rows, H;
    float dropout_p; uint64_t seed;
    const uint64_t* salt;
    int32_t z_f32;
    const int32_t* src_row;   /* packed layout: g / y / z row r belongs to region src_row[r] = b*R + i of box / cls */
    const int32_t* rows_dev;
    float* y32;
} crct_embed_vis_t;
int crct_embed_vis_fwd(const crct_embed_vis_t* args, crct_stream_t stream);
typedef struct {
    const void* dz; const float* box; const int64_t* cls;
    float* g_color; float* g_wloc;   /* += ; the two bias gradients are crct_colsum_bf16(dz) */
    int32_t rows, H;
    const int32_t* src_row; const int32_t* rows_dev;
} crct_embed_vis_bwd_t;
int crct_embed_vis_bwd(const crct_embed_vis_bwd_t* args, crct_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2/K3  fused attention: ctx = dropout(softmax(q k^T / sqrt(dh) + mask_add)) v, one CTA per (sample, head).
 * Replaces vilbert.py:397-412 (text, 16 x 48), :527-543 (visual, 16 x 64) and both directions of the
 * co-attention :684-723 (32 x 32; call once with q = text / k,v = visual and once crossed).
 * Element (b, i, h, d) of q lives at q[(b*Lq + i)*ldq + h*dh + d]; k, v likewise with Lk.  dh in {32,48,64}.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
    const void* q; const void* k; const void* v;  /* bf16 */
    int32_t ldq, ldk, ldv;
    const float* mask_add;  /* [B,Lk] additive key mask (0 / -10000) */
    void* out;              /* bf16 [B*Lq, nh*dh], row stride ldo */
    int32_t ldo;
    float* lse;             /* [B,nh,Lq] log-sum-exp of the masked scores, or NULL (inference) */
    int32_t B, nh, dh, Lq, Lk;
    float dropout_p;        /* on the probabilities; element counter ((b*nh+h)*Lq+i)*Lk+j */
    uint64_t seed;
    const uint64_t* salt;
    /* packed (var-len) rows: when cu_q / cu_k (DEVICE int32 [B+1], crct_row_map) are given, sample b's queries are rows
     * [cu_q[b], cu_q[b+1]) of q / out and its keys rows [cu_k[b], cu_k[b+1]) of k / v; Lq / Lk are then the maxima (lse and the
     * dropout counters keep the [B,nh,Lq(,Lk)] indexing).  With cu_k, mask_add may be NULL: every packed key is valid. */
    const int32_t* cu_q; const int32_t* cu_k;
} crct_attn_fwd_t;
int crct_attn_fwd(const crct_attn_fwd_t* args, crct_stream_t stream);

typedef struct {
    const void* q; const void* k; const void* v;
    int32_t ldq, ldk, ldv;
    const float* mask_add;
    const void* out; int32_t ldo;     /* forward output */
    const void* dout; int32_t lddo;   /* bf16 gradient of out */
    const float* lse;
    void* dq; void* dk; void* dv;     /* bf16, same indexing as q/k/v with strides lddq/lddk/lddv; overwritten */
    int32_t lddq, lddk, lddv;
    int32_t B, nh, dh, Lq, Lk;
    float dropout_p;
    uint64_t seed;
    const uint64_t* salt;
    const int32_t* cu_q; const int32_t* cu_k;   /* as in crct_attn_fwd_t */
} crct_attn_bwd_t;
int crct_attn_bwd(const crct_attn_bwd_t* args, crct_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K7/K8  heads in fp32 on CUDA cores (M = batch size): poolers vilbert.py:955-976, classifier :1052-1060,
 * regressor regressor.py:36-42, hybrid loss vilbert.py:1586-1657 + encoder_decorator.py:144-153.
 * -------------------------------------------------------------------------------------------- */
/* C[m*ldc+n] (+)= dact( act( sum_k A[m*sa_m+k*sa_k] * B[k*sb_k+n*sb_n] + bias[n] ) )
 *   act: 0 none, 1 ReLU, 2 LeakyReLU(0.01), 3 tanh;  dact: multiply by (dmask[m*ldm+n] > 0 ? 1 : slope) when dmask.
 *   forward y = x W^T + b : A = x (sa_m=ldx, sa_k=1), B = W (sb_k=1, sb_n=ldw)
 *   dgrad   dx = dy W     : A = dy,                    B = W (sb_k=ldw, sb_n=1), dmask = post-activation of x
 *   wgrad   dW += dy^T x  : A = dy (sa_m=1, sa_k=ldy), B = x (sb_k=ldx, sb_n=1), accumulate = 1 */
typedef struct {
    const float* A; int64_t sa_m, sa_k;
    const float* B; int64_t sb_k, sb_n;
    float* C; int64_t ldc;
    const float* bias;
    const float* dmask; int64_t ldm;
    int32_t M, N, K;
    int32_t act;
    float slope;
    int32_t accumulate;
} crct_linear_t;
int crct_linear_f32(const crct_linear_t* args, crct_stream_t stream);
/* up to 12 independent problems in one launch (blockIdx.z); a NULL `A` stands for all-ones, which makes a bias gradient
 * `db += colsum(dy)` the problem {A = NULL, M = 1, B = dy (sb_k = ldy, sb_n = 1), K = rows, accumulate = 1}. */
int crct_linear_f32_batched(const crct_linear_t* problems, int count, crct_stream_t stream);
/* out[b,:] = float(src_bf16[b*row_stride + :]) — hidden state of the first token / region (vilbert.py:958,973,1599-1600) */
int crct_gather_first(const void* src_bf16, long long row_stride, float* out, int B, int H, crct_stream_t stream);
/* dst_bf16[b*row_stride + :] = g[b,:]; all other rows of dst must be zero (caller memsets) */
int crct_scatter_first(const float* g, void* dst_bf16, long long row_stride, int B, int H, crct_stream_t stream);
/* out[n] += sum_m x[m*ld+n] */
int crct_colsum_f32(const float* x, float* out, int M, int N, long long ld, crct_stream_t stream);
/* pooled = dropout_p(pt * pv)  (vilbert.py:1055) and its backward through the two ReLU poolers */
int crct_pool_mul_fwd(const float* pt, const float* pv, float* out, int n, float p, uint64_t seed, const uint64_t* salt,
                      crct_stream_t stream);
int crct_pool_mul_bwd(const float* dpooled, const float* pt, const float* pv, float* dut, float* duv, int n, float p,
                      uint64_t seed, const uint64_t* salt, crct_stream_t stream);
/* Dropout streams are counter-based: keep(element) = f(seed ^ *salt, element).  `seed` identifies the call site and is a
 * launch constant; `salt` is one device word the training loop advances once per step with this kernel, so a step
 * captured in a CUDA graph draws fresh masks on every replay while forward and backward of one step still agree. */
int crct_bump_salt(uint64_t* salt, crct_stream_t stream);
/* ------------------------------------------------------------------------------------------------
 * Var-len ("packed") rows.  The reference pads to T tokens / R regions (CRCT/utils.py:152,178) and masks the padding
 * additively (vilbert.py:1380-1396); masked keys get probability exactly 0, so padded rows influence nothing
 * (SURVEY.md §2.3).  crct_row_map compacts the rows with mask != 0:
 *   cu[b] = valid rows before sample b (cu[B] = total: pass `cu + B` as `rows_dev` / `a_rows_dev`),
 *   src_row[r] = b*L + t of packed row r.   mask kinds as in crct_additive_mask; a sample without valid rows keeps row 0.
 * crct_group_map: candidate n of an evaluation batch shares the packed region rows of question group[n] (f3):
 *   cu[n], and src_row[r] = the question-level packed row that candidate-level packed row r copies.
 * crct_gather_rows: dst[r,:] = src[idx[r],:] (rows of row_bytes, 16-byte multiples).
 * crct_gather_rows_f32 / crct_scatter_rows_f32: the first-token rows the heads read (vilbert.py:958,973,1599-1600) at
 *   arbitrary row indices (row_index = cu for the packed layout) — fp32 <-> bf16.
 * crct_fill_zero: cudaMemsetAsync on the stream (gradient buffers that are only scattered into).
 * -------------------------------------------------------------------------------------------- */
int crct_row_map(const void* mask, int kind, int B, int L, int32_t* cu, int32_t* src_row, crct_stream_t stream);
int crct_group_map(const int32_t* src_cu, const int64_t* group, int N, int32_t* cu, int32_t* src_row, crct_stream_t stream);
int crct_gather_rows(const void* src, const int32_t* idx, void* dst, int rows, long long row_bytes, const int32_t* rows_dev,
                     crct_stream_t stream);
int crct_gather_rows_f32(const void* src_bf16, long long ld, const int32_t* row_index, float* out, int B, int H, crct_stream_t stream);
int crct_scatter_rows_f32(const float* g, void* dst_bf16, long long ld, const int32_t* row_index, int B, int H, crct_stream_t stream);
int crct_fill_zero(void* dst, size_t bytes, crct_stream_t stream);
/* Same step, and the new value is also written to `snapshot`: one word per forward pass, which that pass and ITS backward
 * pass both read — a second forward before the first one's backward (l1 = model(a); l2 = model(b); (l1 + l2).backward(),
 * allowed by the reference because autograd stores its masks) then still recomputes the masks of its own forward. */
int crct_bump_salt_to(uint64_t* salt, uint64_t* snapshot, crct_stream_t stream);
/* Hybrid loss, metrics and (when dlogits/dpre are given) the gradients of
 *   loss = nsp_coeff * CE(logits, labels; ignore -1) + reg_coeff * mean_B(reg_loss)
 * R[b] = (value, needs_regression, tolerance, scale).  labels NULL = inference (no CE).  Outputs are dense [B]
 * vectors with zeros on rows that need no regression, exactly the `reg` list of vilbert.py:1590-1648.
 * scalars = {loss, nsp_loss, mean reg loss, #(+-5 %), #(<= tol_margin)}. */
typedef struct {
    const float* logits;   /* [B,2] */
    const float* reg;      /* [B] tanh output */
    const int64_t* labels; /* [B] or NULL */
    const float* R;        /* [B,4] */
    float* reg_pred; float* reg_loss; float* reg_l1; float* reg_dist;  /* [B] each */
    float* scalars;        /* [5] */
    float* dlogits;        /* [B,2] or NULL */
    float* dpre;           /* [B] gradient w.r.t. the pre-tanh regressor output, or NULL */
    int32_t B;
    int32_t l1;               /* 1 = L1Loss (-L1 flag), 0 = SmoothL1Loss(beta=0.5)   vilbert.py:1525-1528 */
    int32_t zero_impossible;  /* 1 = loss kind 'L1_smooth' (training glue): zero loss where |target| > 1   vilbert.py:1639-1641 */
    int32_t unit_grads;       /* 1 = dlogits = d nsp_loss / d logits, dpre[b] = d reg_loss[b] / d pre[b] (autograd chains the rest);
                                 0 = gradients of the combined loss (coefficients and the 1/B of the mean folded in) */
    float tol_margin, nsp_coeff, reg_coeff;
} crct_loss_t;
int crct_hybrid_loss(const crct_loss_t* args, crct_stream_t stream);
/* out[b,j] = x[b,j] * s[b*s_stride], s_stride in {0,1}: chains the upstream gradient of nsp_loss / reg_loss[b] */
int crct_scale_rows(const float* x, const float* s, int s_stride, float* out, int B, int n, crct_stream_t stream);

/* f1  fused AdamW over the flat arena (torch.optim.AdamW semantics, CRCT/utils.py:228-249 param groups):
 * group_of_block64[i/64] in 0..3 selects (lr, weight_decay); also rewrites the bf16 operand copy. */
typedef struct {
    float* w; const float* g; float* m; float* v;
    void* w_bf16;                       /* or NULL */
    const uint8_t* group_of_block64;    /* or NULL (group 0) */
    size_t n;
    float lr[4], weight_decay[4];
    float beta1, beta2, eps;
    int32_t step;                       /* 1-based */
    float grad_scale;                   /* e.g. 1/world_size after a sum all-reduce */
    const float* dyn;                   /* optional DEVICE array {lr[0..3], 1-beta1^t, sqrt(1-beta2^t)} overriding lr/step (CUDA-graph replay) */
    size_t group_offset;                /* element i belongs to block (i + group_offset) / 64 of group_of_block64: lets a data-parallel
                                           rank update ITS SHARD of a gradient bucket (a range that starts inside a 64-element block) */
} crct_adamw_t;
int crct_adamw(const crct_adamw_t* args, crct_stream_t stream);

/* f3  evaluation: candidate expansion and per-question answer selection on the device.
 * crct_expand_blocks: dst[n] = src[group[n]] for n_blocks blocks of bytes_per_block bytes (multiple of 4; 16-byte vectors
 * when sizes and pointers allow).  Fans the per-QUESTION visual embedding [Q, R*Hv] (and additive mask [Q, R]) out to the
 * per-CANDIDATE rows the encoder works on — replaces the host-side replication of CRCT/fig_dataloader.py:690-693,697-703. */
int crct_expand_blocks(const void* src, const int64_t* group, void* dst, long long n_blocks, long long bytes_per_block,
                       crct_stream_t stream);
/* Per question q with candidates offsets[q] .. offsets[q+1]-1: prob = softmax(logits)[:,0] (CRCT/evaluation.py:254-258),
 * answer[q] = first argmax of prob within the question (evaluation.py:291) or forced[q] when `forced` is given (the
 * '_REGS' branch, evaluation.py:289); sel_* = the regression columns of that candidate (evaluation.py:293-295). */
typedef struct {
    const float* logits;      /* [N,2] */
    const float* reg_pred;    /* [N] regression[0] */
    const float* reg_dist;    /* [N] regression[4] (relative distance, the 5 % measure) */
    const float* reg_l1;      /* [N] regression[2] */
    const int64_t* offsets;   /* [Q+1] exclusive prefix sum of num_ans */
    const int64_t* forced;    /* [Q] or NULL */
    int64_t* answer;          /* [Q] index within the question */
    float* prob;              /* [N] or NULL */
    float* sel_pred; float* sel_dist; float* sel_l1;   /* [Q] each */
    int32_t Q;
} crct_select_t;
int crct_select_answers(const crct_select_t* args, crct_stream_t stream);
/* Correctness flags and the running 6x2 accuracy table of `reduce_total_acc` (CRCT/evaluation.py:306-313,494-525):
 * flags[q] = {nsp_right, reg_right, reg_t_right, correct(+-5 %), correct(tolerance)}; total[row][0] += hits,
 * total[row][1] += population, rows = {nsp, cls-on-regression, reg 5 %, reg tolerance, total 5 %, total tolerance}. */
typedef struct {
    const int64_t* answer; const int64_t* gt_id;   /* [Q] */
    const uint8_t* needs_reg;                      /* [Q] */
    const float* sel_dist; const float* sel_l1;    /* [Q] */
    const float* tolerance;                        /* [Q] batch['tolerance_margin'] */
    uint8_t* flags;                                /* [Q,5] or NULL */
    double* total;                                 /* [6,2], accumulated */
    int32_t Q;
} crct_score_t;
int crct_score_answers(const crct_score_t* args, crct_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-model entry points: the INFERENCE forward of the model behind
 * `encoder_decorator.py:73 forward(..., evaluation=True)` — embeddings, the 12 + 6 + 6 layer schedule of
 * vilbert.py:822-946, poolers, classifier, regressor and the loss / metric vectors of vilbert.py:1586-1657 — scheduled by the
 * library itself (csrc/model.cu) on top of the operator entry points above, for hosts without Python (or that want one call
 * per batch: ~300 kernel launches enqueued from C++).  Same kernels in the same order as `VisualDialogEncoder.forward` in
 * evaluation mode: the outputs agree BIT FOR BIT (tests/test_model_capi_gpu.py).  The training forward / backward schedule stays in
 * the reference's host language (cqa_crct_b200/encoder.py), where autograd, DDP and the optimizers hook in.
 *   - the handle owns only host-side tables (configuration, parameter pointers); parameters, inputs, outputs and the
 *     workspace are DEVICE memory owned by the caller; nothing is allocated, freed or synchronised;
 *   - parameters are bound BY NAME (the reference's state_dict keys without the `bert_pretrained.` prefix): the fp32 master
 *     tensor and its bf16 operand copy; q|k|v (and query1|key1|value1, query2|key2|value2) must be adjacent in memory in that
 *     order, weights and biases separately (the fused projections read them as one [3H, H] matrix) — checked at bind time;
 *   - rows are packed (var-len) as in the Python host: masks select the valid tokens / regions.
 * -------------------------------------------------------------------------------------------- */
#define CRCT_MAX_CONNECTIONS 16
typedef struct {
    int32_t hidden_size, num_hidden_layers, num_attention_heads, intermediate_size;              /* text stream (config/vilbert.json) */
    int32_t v_hidden_size, v_num_hidden_layers, v_num_attention_heads, v_intermediate_size, v_feature_size;
    int32_t bi_hidden_size, bi_num_attention_heads;
    int32_t max_position_embeddings;
    int32_t num_connections;                                                                       /* len(v_biattention_id) */
    int32_t v_biattention_id[CRCT_MAX_CONNECTIONS], t_biattention_id[CRCT_MAX_CONNECTIONS];
    int32_t l1;                     /* params['L1']: 1 = L1Loss, 0 = SmoothL1Loss(beta = 0.5)   vilbert.py:1525-1528 */
    float tol_margin;               /* params['tol_margin'] */
} crct_config_t;
typedef struct crct_model_s* crct_handle_t;
int crct_create(const crct_config_t* config, crct_handle_t* out);
int crct_destroy(crct_handle_t h);
/* names[i]: parameter name; w32[i]: fp32 master (device); w16[i]: bf16 copy (device; may be NULL for biases, LayerNorm,
 * embedding tables and the fp32 heads); numel[i]: elements.  May be called again after the caller moved its buffers. */
int crct_bind_params(crct_handle_t h, const char* const* names, const float* const* w32, const void* const* w16, const size_t* numel, int n);
/* B candidate sequences of T tokens; Bq visual rows of R regions (Bq == B, or the number of QUESTIONS with `group`). */
size_t crct_workspace_bytes(crct_handle_t h, int B, int Bq, int T, int R);
typedef struct {
    const int64_t* tokens;       /* [B,T]            batch['tokens'] */
    const int64_t* segments;     /* [B,T]            token_type_ids in {-1,0..11} */
    const float* loc;            /* [B,T,4]          txt_loc */
    const void* attention_mask;  /* [B,T]            valid tokens; kind below */
    const float* image_feat;     /* [Bq,R,F] */
    const float* image_loc;      /* [Bq,R,4] */
    const int64_t* image_target; /* [Bq,R]           class ids -> color_emb */
    const void* image_mask;      /* [Bq,R]           valid regions */
    const float* R4;             /* [B,4]            gt_reg[0] = (value, needs_reg, tolerance, scale) */
    const int64_t* group;        /* [B] question index of every candidate (f3), or NULL when Bq == B */
    int32_t B, Bq, T, R;
    int32_t attention_mask_kind, image_mask_kind;   /* 0 = bool / uint8, 1 = int64, 2 = fp32 (as crct_additive_mask) */
    float text_fill, region_fill;                   /* expected fraction of valid rows (0 = unknown): tile-shape choice only */
} crct_batch_t;
typedef struct {
    float* logits;      /* [B,2]  seq_relationship_score */
    float* reg_pred;    /* [B]    reg[0] */
    float* reg_loss;    /* [B]    reg[1] */
    float* reg_l1;      /* [B]    reg[2] */
    float* reg_dist;    /* [B]    reg[4] */
    float* scalars;     /* [5]    {loss, nsp (0 at inference), mean reg loss, #(+-5 %), #(<= tol_margin)} = reg[3] counts in [3], [4] */
} crct_out_t;
int crct_forward(crct_handle_t h, const crct_batch_t* batch, const crct_out_t* out, void* workspace, size_t workspace_bytes,
                 crct_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 CHECK MODE (csrc/check_f32.cu): the operators above with fp32 activation storage and plain fp32 CUDA-core
 * arithmetic (exact erf, IEEE division, no tensor cores).  Same argument structs and the same dropout streams as the
 * production entry points; every pointer documented as bf16 above is fp32 here.  `VisualDialogEncoder(params,
 * precision='fp32')` runs the whole forward/backward schedule through these: <= 1e-4 against the fp64 oracle, and an
 * on-device yardstick for the bf16 path at sizes the CPU oracle cannot reach.  dk/dv of crct_f32_attn_bwd are
 * overwritten (deterministic two-pass backward, no atomics).
 * -------------------------------------------------------------------------------------------- */
int crct_f32_gemm(const crct_gemm_t* args, crct_stream_t stream);
int crct_f32_layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                           int rows, int H, crct_stream_t stream);
int crct_f32_layernorm_bwd(const crct_ln_bwd_t* args, crct_stream_t stream);
int crct_f32_layernorm_bwd_params(const crct_ln_bwd_t* args, crct_stream_t stream);
int crct_f32_attn_fwd(const crct_attn_fwd_t* args, crct_stream_t stream);
int crct_f32_attn_bwd(const crct_attn_bwd_t* args, crct_stream_t stream);
int crct_f32_embed_text_fwd(const crct_embed_text_t* args, crct_stream_t stream);
int crct_f32_embed_text_bwd(const crct_embed_text_bwd_t* args, crct_stream_t stream);
int crct_f32_embed_vis_fwd(const crct_embed_vis_t* args, crct_stream_t stream);
int crct_f32_embed_vis_bwd(const crct_embed_vis_bwd_t* args, crct_stream_t stream);
int crct_f32_softmax_rows(const float* x, float* out, int rows, int F, crct_stream_t stream);
int crct_f32_gather_first(const float* src, long long row_stride, float* out, int B, int H, crct_stream_t stream);
int crct_f32_scatter_first(const float* g, float* dst, long long row_stride, int B, int H, crct_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CRCT_B200_H */
