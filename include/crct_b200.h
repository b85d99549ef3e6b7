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
    CRCT_EPI_BIAS_GELU = 1, /* u = acc + bias; D2 = u (if D2); D = gelu_erf(u)    vilbert.py:111-117,454-457 */
    CRCT_EPI_BIAS_RES = 2,  /* D = dropout(acc + bias) + aux                      vilbert.py:424-428,467-471,749-756 */
    CRCT_EPI_DGELU = 3,     /* D = acc * gelu'(aux)                               backward of vilbert.py:456 */
    CRCT_EPI_F32 = 4        /* D (fp32) = acc, or D += acc when accumulate != 0 (wgrad, split-K) */
} crct_epilogue_t;

typedef struct {
    const void* A;      /* bf16 */
    const void* B;      /* bf16 */
    void* D;            /* bf16 [M,N] (fp32 for CRCT_EPI_F32), row stride ldd */
    void* D2;           /* bf16 [M,N] row stride ldd, CRCT_EPI_BIAS_GELU only, may be NULL */
    const float* bias;  /* fp32 [N] or NULL */
    const void* aux;    /* bf16 [M,N] row stride ldaux (residual / addend / pre-activation) or NULL */
    int32_t M, N, K;
    int32_t lda, ldb, ldd, ldaux; /* in elements */
    int32_t a_major, b_major;
    int32_t epilogue;   /* crct_epilogue_t */
    int32_t accumulate; /* CRCT_EPI_F32: 1 = atomically add into D */
    int32_t split_k;    /* >= 1; > 1 requires CRCT_EPI_F32 with accumulate */
    int32_t block_n;    /* 0 = choose; else 128 or 256 */
    float dropout_p;    /* CRCT_EPI_BIAS_RES: 0 = off */
    uint64_t seed;      /* dropout stream; element counter = m * N + n */
    int32_t max_ctas;   /* 0 = one persistent CTA per SM; else cap (tests) */
    int32_t dbg[7];     /* descriptor overrides for bring-up; must be 0 in production */
} crct_gemm_t;

int crct_gemm_bf16(const crct_gemm_t* args, crct_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CRCT_B200_H */
