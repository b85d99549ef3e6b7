// f3 — evaluation-side kernels: candidate expansion on the device and per-question answer selection.
//
// The reference's evaluation batch replicates every question's visual tensors once per candidate answer on the HOST
// (CRCT/fig_dataloader.py:690-693 expand + pad, :697-703 cut_batch_padding) and picks the answer with a Python loop over
// questions that syncs the device several times per question (CRCT/evaluation.py:287-296).  Here the visual embedding is
// computed once per question and fanned out to its candidates on the device (crct_expand_blocks), and the selection,
// the correctness flags and the 6x2 accuracy table of `reduce_total_acc` (CRCT/evaluation.py:494-525) are two launches
// without a host read-back.
#include "common.cuh"

namespace {

// dst[n] = src[group[n]] for N blocks of `vec_per_block` 16-byte vectors (or `vec_per_block` 4-byte words when W = 4)
template <typename V>
__global__ void __launch_bounds__(256) expand_blocks_kernel(const V* __restrict__ src, const long long* __restrict__ group,
                                                            V* __restrict__ dst, long long total, int vec_per_block) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / vec_per_block;
        const int c = (int)(i - n * vec_per_block);
        dst[i] = __ldg(src + __ldg(group + n) * vec_per_block + c);
    }
}

struct SelectParams {
    const float* logits; const float* reg_pred; const float* reg_dist; const float* reg_l1;
    const long long* offsets; const long long* forced;
    long long* answer; float* prob; float* sel_pred; float* sel_dist; float* sel_l1;
    int Q;
};

// one warp per question: p0 = softmax(logits)[:, 0] for each of its candidates (evaluation.py:254-258), first index of
// the maximum (torch.argmax, evaluation.py:291) unless `forced` gives the index (the '_REGS' branch, :289), then the
// three regression columns of that candidate (:293-295)
__global__ void __launch_bounds__(128) select_answers_kernel(const SelectParams p) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= p.Q) return;
    const long long lo = p.offsets[q], hi = p.offsets[q + 1];
    float best = -1.f;
    long long arg = hi;                  // larger than any candidate index: loses every tie
    for (long long i = lo + lane; i < hi; i += 32) {
        const float z0 = p.logits[2 * i], z1 = p.logits[2 * i + 1];
        const float mx = fmaxf(z0, z1);
        const float e0 = expf(z0 - mx), e1 = expf(z1 - mx);
        const float p0 = e0 / (e0 + e1);
        if (p.prob) p.prob[i] = p0;
        if (p0 > best) { best = p0; arg = i; }      // strictly greater: the first maximum of this lane's stride wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const long long oa = __shfl_xor_sync(0xFFFFFFFFu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) {
        long long a = (hi > lo && arg < hi) ? arg - lo : 0;      // all-NaN scores: candidate 0
        if (p.forced) a = p.forced[q];
        p.answer[q] = a;
        const bool ok = hi > lo && a >= 0 && lo + a < hi;
        const long long i = lo + a;
        p.sel_pred[q] = ok ? p.reg_pred[i] : 0.f;
        p.sel_dist[q] = ok ? p.reg_dist[i] : 0.f;
        p.sel_l1[q] = ok ? p.reg_l1[i] : 0.f;
    }
}

struct ScoreParams {
    const long long* answer; const long long* gt_id; const uint8_t* needs_reg; const float* sel_dist; const float* sel_l1;
    const float* tol; uint8_t* flags; double* total; int Q;
};

// flags[q] = {nsp_right, reg_right, reg_t_right, correct (+-5 %), correct (tolerance)} (evaluation.py:306-313) and
// total[6][2] += this batch's counts, row order of reduce_total_acc (evaluation.py:498-517)
__global__ void __launch_bounds__(256) score_answers_kernel(const ScoreParams p) {
    __shared__ int red[7][256];
    int c[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int q = threadIdx.x; q < p.Q; q += 256) {
        const bool need = p.needs_reg[q] != 0;
        const bool nsp = p.answer[q] == p.gt_id[q];
        const bool r5 = (p.sel_dist[q] <= 0.05f) && need;
        const bool rt = (p.sel_l1[q] <= p.tol[q]) && need;
        const bool ok5 = nsp && (!need || r5), okt = nsp && (!need || rt);
        if (p.flags) {
            uint8_t* f = p.flags + 5 * (size_t)q;
            f[0] = nsp; f[1] = r5; f[2] = rt; f[3] = ok5; f[4] = okt;
        }
        c[0] += nsp; c[1] += nsp && need; c[2] += need; c[3] += r5; c[4] += rt; c[5] += ok5; c[6] += okt;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) red[k][threadIdx.x] = c[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int k = 0; k < 7; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double n = p.Q, need = red[2][0];
        double* t = p.total;
        t[0] += red[0][0]; t[1] += n;          // nsp acc
        t[2] += red[1][0]; t[3] += need;       // classification acc on the regression questions
        t[4] += red[3][0]; t[5] += need;       // regression within 5 %
        t[6] += red[4][0]; t[7] += need;       // regression within the tolerance margin
        t[8] += red[5][0]; t[9] += n;          // total (+-5 %)
        t[10] += red[6][0]; t[11] += n;        // total (tolerance)
    }
}

}  // namespace

extern "C" CRCT_API int crct_expand_blocks(const void* src, const int64_t* group_, void* dst, long long n_blocks,
                                           long long bytes_per_block, crct_stream_t s) {
    const long long* group = reinterpret_cast<const long long*>(group_);
    if (!src || !group || !dst || n_blocks < 0 || bytes_per_block <= 0 || (bytes_per_block & 3))
        CRCT_FAIL(CRCT_ERR_ARG, "crct_expand_blocks: null pointer, or a block size that is not a positive multiple of 4 bytes");
    if (n_blocks == 0) return CRCT_OK;
    const int sms = crct_num_sms();
    if (sms <= 0) return CRCT_ERR_CUDA;
    const bool wide = (bytes_per_block & 15) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    const long long per = bytes_per_block / (wide ? 16 : 4);
    if (per > 0x7FFFFFFFLL) CRCT_FAIL(CRCT_ERR_ARG, "crct_expand_blocks: block too large");
    const long long total = n_blocks * per;
    const int grid = (int)((total + 255) / 256 < (long long)sms * 8 ? (total + 255) / 256 : (long long)sms * 8);
    if (wide)
        expand_blocks_kernel<uint4><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const uint4*>(src), group,
                                                                   reinterpret_cast<uint4*>(dst), total, (int)per);
    else
        expand_blocks_kernel<uint32_t><<<grid, 256, 0, as_stream(s)>>>(reinterpret_cast<const uint32_t*>(src), group,
                                                                      reinterpret_cast<uint32_t*>(dst), total, (int)per);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_select_answers(const crct_select_t* a, crct_stream_t s) {
    if (!a || !a->logits || !a->reg_pred || !a->reg_dist || !a->reg_l1 || !a->offsets || !a->answer || !a->sel_pred ||
        !a->sel_dist || !a->sel_l1)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_select_answers: null pointer");
    if (a->Q < 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_select_answers: negative question count");
    if (a->Q == 0) return CRCT_OK;
    SelectParams p;
    p.logits = a->logits; p.reg_pred = a->reg_pred; p.reg_dist = a->reg_dist; p.reg_l1 = a->reg_l1;
    p.offsets = reinterpret_cast<const long long*>(a->offsets); p.forced = reinterpret_cast<const long long*>(a->forced);
    p.answer = reinterpret_cast<long long*>(a->answer); p.prob = a->prob;
    p.sel_pred = a->sel_pred; p.sel_dist = a->sel_dist; p.sel_l1 = a->sel_l1; p.Q = a->Q;
    select_answers_kernel<<<(a->Q + 3) / 4, 128, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_score_answers(const crct_score_t* a, crct_stream_t s) {
    if (!a || !a->answer || !a->gt_id || !a->needs_reg || !a->sel_dist || !a->sel_l1 || !a->tolerance || !a->total)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_score_answers: null pointer");
    if (a->Q < 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_score_answers: negative question count");
    if (a->Q == 0) return CRCT_OK;
    ScoreParams p;
    p.answer = reinterpret_cast<const long long*>(a->answer); p.gt_id = reinterpret_cast<const long long*>(a->gt_id);
    p.needs_reg = a->needs_reg; p.sel_dist = a->sel_dist; p.sel_l1 = a->sel_l1; p.tol = a->tolerance; p.flags = a->flags;
    p.total = a->total; p.Q = a->Q;
    score_answers_kernel<<<1, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}
