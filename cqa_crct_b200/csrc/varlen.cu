// Var-len ("packed") row layout: only the VALID token / region rows of a batch go through the encoder.
//
// The reference pads every sequence to T = 124 tokens and R = 44 regions (CRCT/utils.py:152,178) and masks the padding
// additively (backbone/vilbert.py:1380-1396: (1 - mask) * -10000 on the keys).  A masked key's probability is
// exp(s - 10000 - max) = 0 exactly in fp32, so padded rows never influence a valid row, the [:,0] rows the heads read, or
// any gradient (SURVEY.md §2.3 probe) — they are pure dead work: ~30 % of the text rows and ~45 % of the region rows of a
// PlotQA-shaped batch.  Here the rows where mask == 1 are compacted once, right at the embeddings:
//   cu[b]      = number of valid rows of samples 0 .. b-1   (cu[B] = total = the `rows_dev` word of every kernel)
//   src_row[r] = b * L + t  of packed row r
// and every GEMM / LayerNorm / attention / column-sum kernel runs on min(rows, *rows_dev) rows (device-side count: the
// captured CUDA graph is the same for every batch).  Attention needs no mask any more: sample b attends over the packed
// rows [cu[b], cu[b+1]) only.  A sample with NO valid row keeps its row 0 (the reference would attend uniformly over the
// padding there; the loader never produces it: [CLS] and the <IMG> region are always valid).
#include "common.cuh"

namespace {

constexpr int MAP_THREADS = 1024;
constexpr int MAX_B = 8192;

__device__ __forceinline__ bool mask_on(const void* mask, int kind, size_t i) {
    if (kind == 0) return reinterpret_cast<const uint8_t*>(mask)[i] != 0;
    if (kind == 1) return reinterpret_cast<const long long*>(mask)[i] != 0;
    return reinterpret_cast<const float*>(mask)[i] != 0.f;
}

// one CTA: per-sample counts (one warp per sample) -> exclusive scan -> compacted positions
__global__ void __launch_bounds__(MAP_THREADS) row_map_kernel(const void* __restrict__ mask, int kind, int B, int L,
                                                              int* __restrict__ cu, int* __restrict__ src_row) {
    __shared__ int len_s[MAX_B + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = MAP_THREADS / 32;
    for (int b = warp; b < B; b += nw) {
        int n = 0;
        for (int t = lane; t < L; t += 32) n += mask_on(mask, kind, (size_t)b * L + t) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) len_s[b] = n > 0 ? n : 1;
    }
    __syncthreads();
    if (warp == 0) {                                     // exclusive scan over B by one warp, 32 samples per round
        int carry = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
            const int b = b0 + lane;
            const int v = b < B ? len_s[b] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (b < B) len_s[b] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) len_s[B] = carry;
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= B; b += MAP_THREADS) cu[b] = len_s[b];
    for (int b = warp; b < B; b += nw) {
        int base = len_s[b];
        const int total = len_s[b + 1] - base;
        int written = 0;
        for (int t0 = 0; t0 < L; t0 += 32) {
            const int t = t0 + lane;
            const bool on = t < L && mask_on(mask, kind, (size_t)b * L + t);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) src_row[base + __popc(m & ((1u << lane) - 1u))] = b * L + t;
            base += __popc(m);
            written += __popc(m);
        }
        if (written == 0 && lane == 0 && total == 1) src_row[len_s[b]] = b * L;      // all-masked sample keeps its row 0
    }
}

// candidates of one question share the question's packed region rows (f3): per-candidate offsets + source rows
__global__ void __launch_bounds__(MAP_THREADS) group_map_kernel(const int* __restrict__ src_cu, const long long* __restrict__ group, int N,
                                                                int* __restrict__ cu, int* __restrict__ src_row) {
    __shared__ int len_s[MAX_B + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = MAP_THREADS / 32;
    for (int n = threadIdx.x; n < N; n += MAP_THREADS) {
        const int q = (int)group[n];
        len_s[n] = src_cu[q + 1] - src_cu[q];
    }
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int b0 = 0; b0 < N; b0 += 32) {
            const int b = b0 + lane;
            const int v = b < N ? len_s[b] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (b < N) len_s[b] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) len_s[N] = carry;
    }
    __syncthreads();
    for (int n = threadIdx.x; n <= N; n += MAP_THREADS) cu[n] = len_s[n];
    for (int n = warp; n < N; n += nw) {
        const int q = (int)group[n];
        const int s0 = src_cu[q], len = len_s[n + 1] - len_s[n], d0 = len_s[n];
        for (int i = lane; i < len; i += 32) src_row[d0 + i] = s0 + i;
    }
}

// dst[r, :] = src[idx[r], :] for r < *rows_dev; 16-byte vectors, one warp per row
__global__ void __launch_bounds__(256) gather_rows16_kernel(const uint4* __restrict__ src, const int* __restrict__ idx, uint4* __restrict__ dst,
                                                           int rows, int vec_per_row, const int* __restrict__ rows_dev) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    const int n = rows_dev != nullptr ? min(rows, __ldg(rows_dev)) : rows;
    if (row >= n) return;
    const size_t s = (size_t)__ldg(idx + row) * vec_per_row, d = (size_t)row * vec_per_row;
    for (int i = lane; i < vec_per_row; i += 32) dst[d + i] = __ldg(src + s + i);
}

// out[b, :] = float(src[row_index[b] * ld + :])  — first-token / first-region hidden state of every sample
__global__ void __launch_bounds__(256) gather_rows_f32_kernel(const bf16* __restrict__ src, long long ld, const int* __restrict__ row_index,
                                                             float* __restrict__ out, int B, int H) {
    const int b = blockIdx.x;
    const bf16* s = src + (size_t)__ldg(row_index + b) * ld;
    for (int i = threadIdx.x * 8; i < H; i += 256 * 8) {
        float f[8];
        load8_bf16(s + i, f);
        *reinterpret_cast<float4*>(out + (size_t)b * H + i) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(out + (size_t)b * H + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
}
__global__ void __launch_bounds__(256) scatter_rows_f32_kernel(const float* __restrict__ g, bf16* __restrict__ dst, long long ld,
                                                              const int* __restrict__ row_index, int B, int H) {
    const int b = blockIdx.x;
    bf16* d = dst + (size_t)__ldg(row_index + b) * ld;
    for (int i = threadIdx.x * 8; i < H; i += 256 * 8) {
        float f[8];
        load8_f32(g + (size_t)b * H + i, f);
        store8_bf16(d + i, f);
    }
}

}  // namespace

extern "C" CRCT_API int crct_row_map(const void* mask, int kind, int B, int L, int32_t* cu, int32_t* src_row, crct_stream_t s) {
    if (!mask || !cu || !src_row || kind < 0 || kind > 2) CRCT_FAIL(CRCT_ERR_ARG, "crct_row_map: bad argument");
    if (B <= 0 || L <= 0 || B > MAX_B) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_row_map: B=%d must be in [1, %d], L=%d > 0", B, MAX_B, L);
    row_map_kernel<<<1, MAP_THREADS, 0, as_stream(s)>>>(mask, kind, B, L, cu, src_row);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_group_map(const int32_t* src_cu, const int64_t* group, int N, int32_t* cu, int32_t* src_row, crct_stream_t s) {
    if (!src_cu || !group || !cu || !src_row) CRCT_FAIL(CRCT_ERR_ARG, "crct_group_map: null pointer");
    if (N <= 0 || N > MAX_B) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_group_map: N=%d must be in [1, %d]", N, MAX_B);
    group_map_kernel<<<1, MAP_THREADS, 0, as_stream(s)>>>(src_cu, reinterpret_cast<const long long*>(group), N, cu, src_row);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_gather_rows(const void* src, const int32_t* idx, void* dst, int rows, long long row_bytes,
                                         const int32_t* rows_dev, crct_stream_t s) {
    if (!src || !idx || !dst) CRCT_FAIL(CRCT_ERR_ARG, "crct_gather_rows: null pointer");
    if (row_bytes <= 0 || (row_bytes % 16) || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15))
        CRCT_FAIL(CRCT_ERR_ARG, "crct_gather_rows: rows must be 16-byte multiples at 16-byte aligned addresses");
    if (rows <= 0) return CRCT_OK;
    gather_rows16_kernel<<<(rows + 7) / 8, 256, 0, as_stream(s)>>>(reinterpret_cast<const uint4*>(src), idx, reinterpret_cast<uint4*>(dst), rows,
                                                                  (int)(row_bytes / 16), rows_dev);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_gather_rows_f32(const void* src_bf16, long long ld, const int32_t* row_index, float* out, int B, int H,
                                             crct_stream_t s) {
    if (!src_bf16 || !row_index || !out) CRCT_FAIL(CRCT_ERR_ARG, "crct_gather_rows_f32: null pointer");
    if (B <= 0 || H <= 0 || (H % 8) || (ld % 8)) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_gather_rows_f32: H and ld must be multiples of 8");
    gather_rows_f32_kernel<<<B, 256, 0, as_stream(s)>>>(reinterpret_cast<const bf16*>(src_bf16), ld, row_index, out, B, H);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_scatter_rows_f32(const float* g, void* dst_bf16, long long ld, const int32_t* row_index, int B, int H,
                                              crct_stream_t s) {
    if (!g || !dst_bf16 || !row_index) CRCT_FAIL(CRCT_ERR_ARG, "crct_scatter_rows_f32: null pointer");
    if (B <= 0 || H <= 0 || (H % 8) || (ld % 8)) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_scatter_rows_f32: H and ld must be multiples of 8");
    scatter_rows_f32_kernel<<<B, 256, 0, as_stream(s)>>>(g, reinterpret_cast<bf16*>(dst_bf16), ld, row_index, B, H);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_fill_zero(void* dst, size_t bytes, crct_stream_t s) {
    if (!dst) CRCT_FAIL(CRCT_ERR_ARG, "crct_fill_zero: null pointer");
    if (bytes == 0) return CRCT_OK;
    CRCT_CUDA(cudaMemsetAsync(dst, 0, bytes, as_stream(s)));
    return CRCT_OK;
}
