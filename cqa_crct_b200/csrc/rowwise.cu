// K4/K5/K6 — fused, vectorised (128-bit), warp-shuffle row kernels: LayerNorm forward/backward,
// text / visual embedding assembly, RoI-feature softmax, column sums, casts, masks.
// All are HBM-bound: one warp owns one row, the row lives in registers, statistics are fp32.
//
// reference: CRCT/backbone/vilbert.py:281-294 (BertLayerNorm, eps inside the sqrt),
//            :320-358 (BertEmbeddingLocation), :1474-1496 (BertImageEmbeddings), :1380-1396 (masks).
#include "common.cuh"

namespace {

constexpr int MAXC = 4;                 // 16-byte chunks per lane: rows up to 32*4*8 = 1024 columns
constexpr int ROW_THREADS = 256;        // 8 warps = 8 rows per CTA
constexpr float LN_EPS = 1e-12f;

// a row of H (= 8*nchunks) values distributed over a warp: lane l holds chunks l, l+32, ...
struct Row {
    float v[MAXC][8];
};

__device__ __forceinline__ void row_stats(const Row& r, int nchunks, int lane, int H, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s += r.v[c][j];
        }
    mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = r.v[c][j] - mean; q += d * d; }
        }
    rstd = rsqrtf(warp_sum(q) / (float)H + LN_EPS);
}

// Row counts may live on the device (var-len packing: the number of valid rows of a batch is only known there, and the
// captured step must not depend on it): `rows_dev` non-NULL overrides `rows`, which then only sized the grid.
__device__ __forceinline__ int dyn_rows(int rows, const int* __restrict__ rows_dev) {
    return rows_dev != nullptr ? min(rows, __ldg(rows_dev)) : rows;
}
// 8 consecutive values of a pre-LayerNorm row: fp32 (production: z is kept in fp32 between the GEMM epilogue and the
// LayerNorm — rounding it to bf16 was the largest single contributor to the end-to-end error) or bf16
template <bool F32>
__device__ __forceinline__ void load8_z(const void* base, size_t off, float (&f)[8]) {
    if constexpr (F32) load8_f32(reinterpret_cast<const float*>(base) + off, f);
    else load8_bf16(reinterpret_cast<const bf16*>(base) + off, f);
}
__device__ __forceinline__ void store8_f32(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// y = (z - mean) * rstd * gamma + beta, optional dropout on y; writes y (bf16)
// (and, when y32_row is given, the same values unrounded: the fp32 copy the next residual add reads)
__device__ __forceinline__ void ln_write(const Row& z, int nchunks, int lane, float mean, float rstd, const float* gamma,
                                         const float* beta, bf16* y_row, float* y32_row, uint64_t row_idx0, uint32_t thr, float scale,
                                         uint64_t seed) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            float g[8], b[8], o[8];
            load8_f32(gamma + ch * 8, g);
            load8_f32(beta + ch * 8, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (z.v[c][j] - mean) * rstd * g[j] + b[j];
            if (thr != 0u) dropout8(o, seed, row_idx0 + ch * 8, thr, scale);
            store8_bf16(y_row + ch * 8, o);
            if (y32_row != nullptr) store8_f32(y32_row + ch * 8, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float f[8];
        load8_f32(src + i * 8, f);
        store8_bf16(dst + i * 8, f);
    }
}

__global__ void __launch_bounds__(ROW_THREADS) additive_mask_kernel(const void* __restrict__ mask, int kind, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m;
    if (kind == 0) m = reinterpret_cast<const uint8_t*>(mask)[i] ? 1.f : 0.f;            // torch.bool
    else if (kind == 1) m = (float)reinterpret_cast<const long long*>(mask)[i];            // int64
    else m = reinterpret_cast<const float*>(mask)[i];                                      // fp32
    out[i] = (1.0f - m) * -10000.0f;                                                       // vilbert.py:1391,1396
}

// ------------------------------------------------------------------------------------------------
template <bool ZF32>
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_fwd_kernel(const void* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                     bf16* __restrict__ y, float* __restrict__ y32, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
                     const int* __restrict__ rows_dev) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (ROW_THREADS / 32) + warp;
    if (row >= dyn_rows(rows, rows_dev)) return;
    const int nchunks = H >> 3;
    Row r;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) load8_z<ZF32>(z, (size_t)row * H + (lane + 32 * c) * 8, r.v[c]);
    float mean, rstd;
    row_stats(r, nchunks, lane, H, mean, rstd);
    ln_write(r, nchunks, lane, mean, rstd, gamma, beta, y + (size_t)row * H, y32 != nullptr ? y32 + (size_t)row * H : nullptr, 0, 0u, 1.f, 0);
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// LayerNorm of selected rows, fp32 in and out: the first-token / first-region hidden states the heads read
// (vilbert.py:958,973,1599-1600) are taken from the LAST pre-LayerNorm sum (fp32) and never rounded to bf16 — the poolers and
// the regressor are fp32, and with SmoothL1 / L1 losses their gradient scales with the regression error itself.
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_rows_f32_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const int* __restrict__ row_index, long long row_step, float* __restrict__ out, int B, int H) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (ROW_THREADS / 32) + warp;
    if (b >= B) return;
    const size_t row = row_index != nullptr ? (size_t)__ldg(row_index + b) : (size_t)b * (size_t)row_step;
    const int nchunks = H >> 3;
    Row r;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) load8_f32(z + row * H + (lane + 32 * c) * 8, r.v[c]);
    float mean, rstd;
    row_stats(r, nchunks, lane, H, mean, rstd);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            float g[8], bb[8], o[8];
            load8_f32(gamma + ch * 8, g);
            load8_f32(beta + ch * 8, bb);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (r.v[c][j] - mean) * rstd * g[j] + bb[j];
            store8_f32(out + (size_t)b * H + ch * 8, o);
        }
    }
}

// dz = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)), dxh = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; dbias += sum_rows dzm (optional)
// dy may carry an input dropout mask (seed_in), dzm = dz * keep(seed_out) * scale (optional second output).
// One warp per row.  The three column sums are accumulated in per-warp shared-memory rows, so registers hold only row
// data: the current row stays PACKED (bf16, as loaded) and is unpacked again for the second pass, which leaves room
// to fetch the NEXT row before the current one is processed — the kernel is latency-bound (two CTAs per SM, a
// dependent load -> reduce -> store chain per row), so that prefetch is what keeps HBM busy.
template <int NCH>
struct PackedRow {
    uint4 dy[NCH], z[NCH];
    float mean, rstd;
};
template <int NCH>
__device__ __forceinline__ void load_packed_row(PackedRow<NCH>& r, const bf16* __restrict__ dy, const bf16* __restrict__ z,
                                                const float* __restrict__ mean, const float* __restrict__ rstd, int row, int H,
                                                int lane, int nchunks) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            r.dy[c] = __ldg(reinterpret_cast<const uint4*>(dy + (size_t)row * H + ch * 8));
            r.z[c] = __ldg(reinterpret_cast<const uint4*>(z + (size_t)row * H + ch * 8));
        }
    }
    r.mean = __ldg(mean + row);
    r.rstd = __ldg(rstd + row);
}
__device__ __forceinline__ void unpack8f(const uint4& v, float (&f)[8]) {
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <int NCH>
__global__ void __launch_bounds__(ROW_THREADS, 2)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, const float* __restrict__ gamma, bf16* __restrict__ dz,
                     bf16* __restrict__ dzm, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                     int rows, int H, uint32_t thr_in, float scale_in, uint64_t seed_in, uint32_t thr_out, float scale_out,
                     uint64_t seed_out, const unsigned long long* __restrict__ salt) {
    extern __shared__ __align__(16) float acc_s[];       // [3][8 warps][H]
    if (salt != nullptr) { const unsigned long long sv = __ldg(salt); seed_in ^= sv; seed_out ^= sv; }
    constexpr int NW = ROW_THREADS / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = H >> 3;
    float* my_g = acc_s + (size_t)(0 * NW + warp) * H;
    float* my_b = acc_s + (size_t)(1 * NW + warp) * H;
    float* my_d = acc_s + (size_t)(2 * NW + warp) * H;
    for (int i = lane; i < H; i += 32) { my_g[i] = 0.f; my_b[i] = 0.f; my_d[i] = 0.f; }
    __syncwarp();
    const float invH = 1.0f / (float)H;
    const int stride = gridDim.x * NW;
    int row = blockIdx.x * NW + warp;
    PackedRow<NCH> cur, nxt;
    if (row < rows) load_packed_row<NCH>(cur, dy, z, mean_in, rstd_in, row, H, lane, nchunks);
    for (; row < rows; row += stride) {
        if (row + stride < rows) load_packed_row<NCH>(nxt, dy, z, mean_in, rstd_in, row + stride, H, lane, nchunks);
        const float rstd = cur.rstd, nmr = -cur.mean * cur.rstd;     // xhat = z * rstd + nmr
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunks) {
                float d[8], x[8], gm[8], pg[8], pb[8];
                unpack8f(cur.dy[c], d);
                unpack8f(cur.z[c], x);
                load8_f32(gamma + ch * 8, gm);
                load8_f32(my_g + ch * 8, pg);
                load8_f32(my_b + ch * 8, pb);
                if (thr_in != 0u) dropout8(d, seed_in, (uint64_t)row * H + ch * 8, thr_in, scale_in);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = fmaf(x[j], rstd, nmr);
                    pg[j] = fmaf(d[j], xh, pg[j]);
                    pb[j] += d[j];
                    const float dxh = d[j] * gm[j];
                    s1 += dxh;
                    s2 = fmaf(dxh, xh, s2);
                }
                *reinterpret_cast<float4*>(my_g + ch * 8) = make_float4(pg[0], pg[1], pg[2], pg[3]);
                *reinterpret_cast<float4*>(my_g + ch * 8 + 4) = make_float4(pg[4], pg[5], pg[6], pg[7]);
                *reinterpret_cast<float4*>(my_b + ch * 8) = make_float4(pb[0], pb[1], pb[2], pb[3]);
                *reinterpret_cast<float4*>(my_b + ch * 8 + 4) = make_float4(pb[4], pb[5], pb[6], pb[7]);
            }
        }
        // dz = rstd * (dxh - m1 - xhat * m2) = dxh * rstd - (m1 rstd) - xhat * (m2 rstd)
        const float m1r = warp_sum(s1) * invH * rstd, m2r = warp_sum(s2) * invH * rstd;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunks) {
                float d[8], x[8], gm[8], o[8];
                unpack8f(cur.dy[c], d);
                unpack8f(cur.z[c], x);
                load8_f32(gamma + ch * 8, gm);
                if (thr_in != 0u) dropout8(d, seed_in, (uint64_t)row * H + ch * 8, thr_in, scale_in);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = fmaf(x[j], rstd, nmr);
                    o[j] = fmaf(-xh, m2r, fmaf(d[j] * gm[j], rstd, -m1r));
                }
                store8_bf16(dz + (size_t)row * H + ch * 8, o);
                if (dzm != nullptr) {
                    dropout8(o, seed_out, (uint64_t)row * H + ch * 8, thr_out, scale_out);
                    store8_bf16(dzm + (size_t)row * H + ch * 8, o);
                }
                if (dbias != nullptr) {
                    float pd[8];
                    load8_f32(my_d + ch * 8, pd);
                    *reinterpret_cast<float4*>(my_d + ch * 8) = make_float4(pd[0] + o[0], pd[1] + o[1], pd[2] + o[2], pd[3] + o[3]);
                    *reinterpret_cast<float4*>(my_d + ch * 8 + 4) = make_float4(pd[4] + o[4], pd[5] + o[5], pd[6] + o[6], pd[7] + o[7]);
                }
            }
        }
        cur = nxt;
    }
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += ROW_THREADS) {
        float sg = 0.f, sb = 0.f, sd = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            sg += acc_s[(size_t)(0 * NW + w) * H + col];
            sb += acc_s[(size_t)(1 * NW + w) * H + col];
            sd += acc_s[(size_t)(2 * NW + w) * H + col];
        }
        if (dgamma != nullptr) atomicAdd(dgamma + col, sg);
        if (dbeta != nullptr) atomicAdd(dbeta + col, sb);
        if (dbias != nullptr) atomicAdd(dbias + col, sd);
    }
}

// ---- LayerNorm backward, split form (what the training step uses): the input gradient is on the backward's critical
// path, the three column sums (dgamma, dbeta, dense-bias gradient) are weight gradients nothing waits for.
//   ln_bwd_dz_kernel     : one warp per row, no accumulators -> few registers, many rows in flight; 61 MB at H = 768.
//   ln_bwd_params_kernel : a lane owns 8 columns, accumulates the three sums in registers over a strided set of rows
//                          (4 rows in flight), CTA reduce through smem, one atomic per column per CTA; re-reads dy, z and
//                          dzm, which the dz kernel has just left in L2.
template <int NCH, bool ZF32>
__global__ void __launch_bounds__(ROW_THREADS)
ln_bwd_dz_kernel(const bf16* __restrict__ dy, const void* __restrict__ z, const float* __restrict__ mean_in,
                 const float* __restrict__ rstd_in, const float* __restrict__ gamma, bf16* __restrict__ dz, bf16* __restrict__ dzm,
                 int rows, int H, uint32_t thr_in, float scale_in, uint64_t seed_in, uint32_t thr_out, float scale_out,
                 uint64_t seed_out, const unsigned long long* __restrict__ salt, const int* __restrict__ rows_dev,
                 const int* __restrict__ drop_rows) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (ROW_THREADS / 32) + warp;
    if (row >= dyn_rows(rows, rows_dev)) return;
    if (salt != nullptr) { const unsigned long long sv = __ldg(salt); seed_in ^= sv; seed_out ^= sv; }
    const uint64_t drow = drop_rows != nullptr ? (uint64_t)__ldg(drop_rows + row) : (uint64_t)row;     // dropout counters: padded position
    const int nchunks = H >> 3;
    float g[NCH][8], x[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {                     // all loads of the row first
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            load8_bf16(dy + (size_t)row * H + ch * 8, g[c]);
            load8_z<ZF32>(z, (size_t)row * H + ch * 8, x[c]);
        }
    }
    const float rstd = __ldg(rstd_in + row), nmr = -__ldg(mean_in + row) * rstd;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            float gm[8];
            load8_f32(gamma + ch * 8, gm);
            if (thr_in != 0u) dropout8(g[c], seed_in, drow * H + ch * 8, thr_in, scale_in);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x[c][j] = fmaf(x[c][j], rstd, nmr);
                g[c][j] *= gm[j];
                s1 += g[c][j];
                s2 = fmaf(g[c][j], x[c][j], s2);
            }
        }
    }
    const float invH = 1.0f / (float)H;
    const float m1r = warp_sum(s1) * invH * rstd, m2r = warp_sum(s2) * invH * rstd;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(-x[c][j], m2r, fmaf(g[c][j], rstd, -m1r));
            store8_bf16(dz + (size_t)row * H + ch * 8, o);
            if (dzm != nullptr) {
                dropout8(o, seed_out, drow * H + ch * 8, thr_out, scale_out);
                store8_bf16(dzm + (size_t)row * H + ch * 8, o);
            }
        }
    }
}

template <bool ZF32>
__global__ void __launch_bounds__(ROW_THREADS)
ln_bwd_params_kernel(const bf16* __restrict__ dy, const void* __restrict__ z, const bf16* __restrict__ dzm,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dbias, int rows, int H, uint32_t thr_in, float scale_in,
                     uint64_t seed_in, const unsigned long long* __restrict__ salt, const int* __restrict__ rows_dev,
                     const int* __restrict__ drop_rows) {
    pdl_launch_dependents();
    pdl_wait();
    rows = dyn_rows(rows, rows_dev);
    constexpr int NW = ROW_THREADS / 32;
    __shared__ float red[NW][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 256 + lane * 8;
    if (salt != nullptr) seed_in ^= __ldg(salt);
    float ag[8], ab[8], ad[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[j] = 0.f; ab[j] = 0.f; ad[j] = 0.f; }
    if (col < H) {
        const int stride = gridDim.y * NW;
        constexpr int U = 4;
        for (int row0 = blockIdx.y * NW + warp; row0 < rows; row0 += U * stride) {
            uint4 pd[U], pz[U], pz2[U], pm[U];
            float mean[U], rstd[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                  // all loads of U rows first
                const int row = row0 + u * stride;
                if (row < rows) {
                    pd[u] = __ldg(reinterpret_cast<const uint4*>(dy + (size_t)row * H + col));
                    if constexpr (ZF32) {
                        pz[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(z) + (size_t)row * H + col));
                        pz2[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(z) + (size_t)row * H + col + 4));
                    } else {
                        pz[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(z) + (size_t)row * H + col));
                    }
                    if (dbias != nullptr) pm[u] = __ldg(reinterpret_cast<const uint4*>(dzm + (size_t)row * H + col));
                    mean[u] = __ldg(mean_in + row);
                    rstd[u] = __ldg(rstd_in + row);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = row0 + u * stride;
                if (row < rows) {
                    float d[8], x[8];
                    unpack8f(pd[u], d);
                    if constexpr (ZF32) {
                        x[0] = __uint_as_float(pz[u].x); x[1] = __uint_as_float(pz[u].y); x[2] = __uint_as_float(pz[u].z); x[3] = __uint_as_float(pz[u].w);
                        x[4] = __uint_as_float(pz2[u].x); x[5] = __uint_as_float(pz2[u].y); x[6] = __uint_as_float(pz2[u].z); x[7] = __uint_as_float(pz2[u].w);
                    } else {
                        unpack8f(pz[u], x);
                    }
                    if (thr_in != 0u)
                        dropout8(d, seed_in, (drop_rows != nullptr ? (uint64_t)__ldg(drop_rows + row) : (uint64_t)row) * H + col, thr_in, scale_in);
                    const float nmr = -mean[u] * rstd[u];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        ag[j] = fmaf(d[j], fmaf(x[j], rstd[u], nmr), ag[j]);
                        ab[j] += d[j];
                    }
                    if (dbias != nullptr) {
                        float m[8];
                        unpack8f(pm[u], m);
#pragma unroll
                        for (int j = 0; j < 8; ++j) ad[j] += m[j];
                    }
                }
            }
        }
    }
    float* outs[3] = {dgamma, dbeta, dbias};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (outs[k] == nullptr) continue;                  // uniform across the CTA
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = k == 0 ? ag[j] : (k == 1 ? ab[j] : ad[j]);
        __syncthreads();
        const int c = blockIdx.x * 256 + threadIdx.x;
        if (c < H) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) sum += red[w][threadIdx.x];
            atomicAdd(outs[k] + c, sum);
        }
    }
}

// out[n] += sum_rows x[row, n]
__global__ void __launch_bounds__(ROW_THREADS)
colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, int rows, int N, int ld, const int* __restrict__ rows_dev) {
    pdl_launch_dependents();
    pdl_wait();
    rows = dyn_rows(rows, rows_dev);
    __shared__ float red[ROW_THREADS / 32][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 256 + lane * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (col < N) {
        const int stride = gridDim.y * (ROW_THREADS / 32);
        int row = blockIdx.y * (ROW_THREADS / 32) + warp;
        for (; row + 3 * stride < rows; row += 4 * stride) {          // 4 independent 16-byte loads in flight
            float f0[8], f1[8], f2[8], f3[8];
            load8_bf16(x + (size_t)row * ld + col, f0);
            load8_bf16(x + (size_t)(row + stride) * ld + col, f1);
            load8_bf16(x + (size_t)(row + 2 * stride) * ld + col, f2);
            load8_bf16(x + (size_t)(row + 3 * stride) * ld + col, f3);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        }
        for (; row < rows; row += stride) {
            float f[8];
            load8_bf16(x + (size_t)row * ld + col, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = threadIdx.x;
    if (blockIdx.x * 256 + c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < ROW_THREADS / 32; ++w) s += red[w][c];
        atomicAdd(out + blockIdx.x * 256 + c, s);
    }
}

// softmax over the F RoI features of one region, fp32 in -> bf16 GEMM operand (vilbert.py:1476)
__global__ void __launch_bounds__(ROW_THREADS)
softmax_rows_kernel(const float* __restrict__ x, bf16* __restrict__ out, int rows, int F, const int* __restrict__ src_row,
                    const int* __restrict__ rows_dev) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (ROW_THREADS / 32) + warp;
    if (row >= dyn_rows(rows, rows_dev)) return;
    const int src = src_row != nullptr ? __ldg(src_row + row) : row;      // packed output row <- padded input row
    const int nchunks = F >> 3;
    Row r;
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) {
            load8_f32(x + (size_t)src * F + (lane + 32 * c) * 8, r.v[c]);
#pragma unroll
            for (int j = 0; j < 8; ++j) mx = fmaxf(mx, r.v[c][j]);
        }
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { r.v[c][j] = __expf(r.v[c][j] - mx); s += r.v[c][j]; }
        }
    const float inv = 1.0f / warp_sum(s);
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
        if (lane + 32 * c < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r.v[c][j] *= inv;
            store8_bf16(out + (size_t)row * F + (lane + 32 * c) * 8, r.v[c]);
        }
}

// ------------------------------------------------------------------------------------------------
// text embedding (vilbert.py:320-358): word + position(QA tokens only, counted from the first QA token)
// + plotqa type (-1 -> 0, zero for type 0) + Linear(4->H)(box) (zero incl. bias for all-zero boxes) -> LN -> dropout
struct TextEmbArgs {
    const long long* ids; const long long* types; const float* loc;
    const float* word; const float* pos; const float* type; const float* w_loc; const float* b_loc;
    const float* gamma; const float* beta;
    bf16* y; void* z; float* mean; float* rstd;
    int B, T, H;
    uint32_t thr; float scale; uint64_t seed; const unsigned long long* salt;
    int z_f32; const int* src_row; const int* rows_dev; float* y32;
};

__device__ __forceinline__ int first_qa_index(const long long* types_row, int T, int lane) {
    int first = T;
    for (int t = lane; t < T; t += 32) {
        const long long ty = types_row[t];
        if ((ty == -1 || ty == 1) && t < first) first = t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    return first;
}

// `row` = output row (packed when src_row is given), `src` = its token position b*T + t in the padded inputs
__global__ void __launch_bounds__(ROW_THREADS) embed_text_fwd_kernel(const TextEmbArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (ROW_THREADS / 32) + warp;
    if (row >= dyn_rows(a.B * a.T, a.rows_dev)) return;
    const int src = a.src_row != nullptr ? __ldg(a.src_row + row) : row;
    const int b = src / a.T, t = src % a.T;
    const int H = a.H, nchunks = H >> 3;
    const int first = first_qa_index(a.types + (size_t)b * a.T, a.T, lane);
    const long long ty = a.types[src];
    const bool qa = (ty == -1 || ty == 1);
    const long long id = a.ids[src];
    const float4 bx = *reinterpret_cast<const float4*>(a.loc + (size_t)src * 4);
    const bool loc_on = (fabsf(bx.x) + fabsf(bx.y) + fabsf(bx.z) + fabsf(bx.w)) != 0.f;
    const long long ty_idx = ty == -1 ? 0 : ty;
    Row r;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            load8_f32(a.word + (size_t)id * H + ch * 8, r.v[c]);
            float tmp[8];
            if (qa) {
                load8_f32(a.pos + (size_t)(t - first) * H + ch * 8, tmp);
#pragma unroll
                for (int j = 0; j < 8; ++j) r.v[c][j] += tmp[j];
            }
            if (ty != 0) {
                load8_f32(a.type + (size_t)ty_idx * H + ch * 8, tmp);
#pragma unroll
                for (int j = 0; j < 8; ++j) r.v[c][j] += tmp[j];
            }
            if (loc_on) {
                load8_f32(a.b_loc + ch * 8, tmp);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 w = *reinterpret_cast<const float4*>(a.w_loc + (size_t)(ch * 8 + j) * 4);
                    r.v[c][j] += tmp[j] + w.x * bx.x + w.y * bx.y + w.z * bx.z + w.w * bx.w;
                }
            }
            if (a.z) {
                if (a.z_f32) store8_f32(reinterpret_cast<float*>(a.z) + (size_t)row * H + ch * 8, r.v[c]);
                else store8_bf16(reinterpret_cast<bf16*>(a.z) + (size_t)row * H + ch * 8, r.v[c]);
            }
        }
    }
    // LayerNorm statistics are taken on the bf16-rounded z when z is materialised in bf16, so that the backward
    // (which re-reads z) sees exactly the normalised values of the forward
    if (a.z && !a.z_f32) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (lane + 32 * c < nchunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) r.v[c][j] = __bfloat162float(__float2bfloat16(r.v[c][j]));
            }
    }
    float mean, rstd;
    row_stats(r, nchunks, lane, H, mean, rstd);
    ln_write(r, nchunks, lane, mean, rstd, a.gamma, a.beta, a.y + (size_t)row * H, a.y32 != nullptr ? a.y32 + (size_t)row * H : nullptr,
             (uint64_t)src * H, a.thr, a.scale, a.salt ? (a.seed ^ __ldg(a.salt)) : a.seed);
    if (lane == 0 && a.mean) { a.mean[row] = mean; a.rstd[row] = rstd; }
}

// backward scatter of dz (bf16, already through the LayerNorm backward) into the embedding tables
__device__ __forceinline__ void add8_global(float* dst, const float (&d)[8]) {
    ptx::red_add_f32x4(dst, d[0], d[1], d[2], d[3]);
    ptx::red_add_f32x4(dst + 4, d[4], d[5], d[6], d[7]);
}

struct TextEmbBwdArgs {
    const long long* ids; const long long* types; const float* loc; const bf16* dz;
    float* g_word; float* g_pos; float* g_type; float* g_wloc; float* g_bloc;
    int B, T, H;
    const int* src_row; const int* rows_dev;
};

__global__ void __launch_bounds__(ROW_THREADS) embed_text_bwd_kernel(const TextEmbBwdArgs a) {
    extern __shared__ float red[];                        // [8 warps][H] reused 5x
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, nchunks = H >> 3;
    Row acc[5];                                           // d b_loc, d w_loc[:,0..3]
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k].v[c][j] = 0.f;
    const int rows = dyn_rows(a.B * a.T, a.rows_dev);
    for (int row = blockIdx.x * (ROW_THREADS / 32) + warp; row < rows; row += gridDim.x * (ROW_THREADS / 32)) {
        const int src = a.src_row != nullptr ? __ldg(a.src_row + row) : row;
        const int b = src / a.T, t = src % a.T;
        const int first = first_qa_index(a.types + (size_t)b * a.T, a.T, lane);
        const long long ty = a.types[src];
        const bool qa = (ty == -1 || ty == 1);
        const long long id = a.ids[src];
        const float4 bx = *reinterpret_cast<const float4*>(a.loc + (size_t)src * 4);
        const bool loc_on = (fabsf(bx.x) + fabsf(bx.y) + fabsf(bx.z) + fabsf(bx.w)) != 0.f;
        const long long ty_idx = ty == -1 ? 0 : ty;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunks) {
                float d[8];
                load8_bf16(a.dz + (size_t)row * H + ch * 8, d);
                add8_global(a.g_word + (size_t)id * H + ch * 8, d);         // 128-bit reductions: 4x fewer L2 atomic ops
                if (qa) add8_global(a.g_pos + (size_t)(t - first) * H + ch * 8, d);
                if (ty != 0) add8_global(a.g_type + (size_t)ty_idx * H + ch * 8, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (loc_on) {
                        acc[0].v[c][j] += d[j];
                        acc[1].v[c][j] += d[j] * bx.x; acc[2].v[c][j] += d[j] * bx.y;
                        acc[3].v[c][j] += d[j] * bx.z; acc[4].v[c][j] += d[j] * bx.w;
                    }
                }
            }
        }
    }
    for (int k = 0; k < 5; ++k) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) red[warp * H + ch * 8 + j] = acc[k].v[c][j];
            }
        }
        __syncthreads();
        for (int col = threadIdx.x; col < H; col += ROW_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < ROW_THREADS / 32; ++w) s += red[w * H + col];
            if (k == 0) atomicAdd(a.g_bloc + col, s);
            else atomicAdd(a.g_wloc + (size_t)col * 4 + (k - 1), s);
        }
    }
}

// visual embedding tail (vilbert.py:1478-1496): z = G + Linear(4->Hv)(box) + color_emb[class] ; LN ; dropout
struct VisEmbArgs {
    const bf16* g; const float* box; const long long* cls;
    const float* w_loc; const float* b_loc; const float* color; const float* gamma; const float* beta;
    bf16* y; void* z; float* mean; float* rstd;
    int rows, H;
    uint32_t thr; float scale; uint64_t seed; const unsigned long long* salt;
    int z_f32; const int* src_row; const int* rows_dev; float* y32;
};

__global__ void __launch_bounds__(ROW_THREADS) embed_vis_fwd_kernel(const VisEmbArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (ROW_THREADS / 32) + warp;
    if (row >= dyn_rows(a.rows, a.rows_dev)) return;
    const int src = a.src_row != nullptr ? __ldg(a.src_row + row) : row;     // g / y / z rows are packed, box / cls padded
    const int H = a.H, nchunks = H >> 3;
    const float4 bx = *reinterpret_cast<const float4*>(a.box + (size_t)src * 4);
    const long long cls = a.cls[src];
    Row r;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nchunks) {
            float tmp[8], bl[8];
            load8_bf16(a.g + (size_t)row * H + ch * 8, r.v[c]);
            load8_f32(a.color + (size_t)cls * H + ch * 8, tmp);
            load8_f32(a.b_loc + ch * 8, bl);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 w = *reinterpret_cast<const float4*>(a.w_loc + (size_t)(ch * 8 + j) * 4);
                r.v[c][j] += tmp[j] + bl[j] + w.x * bx.x + w.y * bx.y + w.z * bx.z + w.w * bx.w;
            }
            if (a.z) {
                if (a.z_f32) {
                    store8_f32(reinterpret_cast<float*>(a.z) + (size_t)row * H + ch * 8, r.v[c]);
                } else {
                    store8_bf16(reinterpret_cast<bf16*>(a.z) + (size_t)row * H + ch * 8, r.v[c]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) r.v[c][j] = __bfloat162float(__float2bfloat16(r.v[c][j]));
                }
            }
        }
    }
    float mean, rstd;
    row_stats(r, nchunks, lane, H, mean, rstd);
    ln_write(r, nchunks, lane, mean, rstd, a.gamma, a.beta, a.y + (size_t)row * H, a.y32 != nullptr ? a.y32 + (size_t)row * H : nullptr,
             (uint64_t)src * H, a.thr, a.scale, a.salt ? (a.seed ^ __ldg(a.salt)) : a.seed);
    if (lane == 0 && a.mean) { a.mean[row] = mean; a.rstd[row] = rstd; }
}

struct VisEmbBwdArgs {
    const bf16* dz; const float* box; const long long* cls;
    float* g_color; float* g_wloc;
    int rows, H;
    const int* src_row; const int* rows_dev;
};

__global__ void __launch_bounds__(ROW_THREADS) embed_vis_bwd_kernel(const VisEmbBwdArgs a) {
    extern __shared__ float red[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, nchunks = H >> 3;
    Row acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[k].v[c][j] = 0.f;
    const int rows = dyn_rows(a.rows, a.rows_dev);
    for (int row = blockIdx.x * (ROW_THREADS / 32) + warp; row < rows; row += gridDim.x * (ROW_THREADS / 32)) {
        const int src = a.src_row != nullptr ? __ldg(a.src_row + row) : row;
        const float4 bx = *reinterpret_cast<const float4*>(a.box + (size_t)src * 4);
        const long long cls = a.cls[src];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunks) {
                float d[8];
                load8_bf16(a.dz + (size_t)row * H + ch * 8, d);
                add8_global(a.g_color + (size_t)cls * H + ch * 8, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0].v[c][j] += d[j] * bx.x; acc[1].v[c][j] += d[j] * bx.y;
                    acc[2].v[c][j] += d[j] * bx.z; acc[3].v[c][j] += d[j] * bx.w;
                }
            }
        }
    }
    for (int k = 0; k < 4; ++k) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nchunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) red[warp * H + ch * 8 + j] = acc[k].v[c][j];
            }
        }
        __syncthreads();
        for (int col = threadIdx.x; col < H; col += ROW_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < ROW_THREADS / 32; ++w) s += red[w * H + col];
            atomicAdd(a.g_wloc + (size_t)col * 4 + k, s);
        }
    }
}

__global__ void bump_salt_kernel(unsigned long long* salt, unsigned long long* snapshot) {
    unsigned long long z = *salt + 0x9E3779B97F4A7C15ull;          // splitmix64 step
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    *salt = z;
    if (snapshot) *snapshot = z;
}

inline int check_row_width(int H, const char* what) {
    if (H <= 0 || (H % 8) || H > MAXC * 256) CRCT_FAIL(CRCT_ERR_SHAPE, "%s: row width %d must be a multiple of 8 and <= %d", what, H, MAXC * 256);
    return CRCT_OK;
}
inline int row_grid(int rows) { return (rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32); }

}  // namespace

extern "C" CRCT_API int crct_bump_salt(uint64_t* salt, crct_stream_t s) {
    if (!salt) CRCT_FAIL(CRCT_ERR_ARG, "crct_bump_salt: null pointer");
    bump_salt_kernel<<<1, 1, 0, as_stream(s)>>>(reinterpret_cast<unsigned long long*>(salt), nullptr);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_bump_salt_to(uint64_t* salt, uint64_t* snapshot, crct_stream_t s) {
    if (!salt || !snapshot) CRCT_FAIL(CRCT_ERR_ARG, "crct_bump_salt_to: null pointer");
    bump_salt_kernel<<<1, 1, 0, as_stream(s)>>>(reinterpret_cast<unsigned long long*>(salt), reinterpret_cast<unsigned long long*>(snapshot));
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_cast_f32_to_bf16(const float* src, void* dst, size_t n, crct_stream_t s) {
    if (!src || !dst) CRCT_FAIL(CRCT_ERR_ARG, "crct_cast_f32_to_bf16: null pointer");
    if (n % 8) CRCT_FAIL(CRCT_ERR_ARG, "crct_cast_f32_to_bf16: n must be a multiple of 8");
    if (n == 0) return CRCT_OK;
    const size_t n8 = n / 8;
    size_t blocks = (n8 + ROW_THREADS - 1) / ROW_THREADS;
    const size_t cap = (size_t)crct_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    cast_kernel<<<(unsigned)blocks, ROW_THREADS, 0, as_stream(s)>>>(src, reinterpret_cast<bf16*>(dst), n8);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_additive_mask(const void* mask, int kind, float* out, int n, crct_stream_t s) {
    if (!mask || !out || n <= 0 || kind < 0 || kind > 2) CRCT_FAIL(CRCT_ERR_ARG, "crct_additive_mask: bad argument");
    additive_mask_kernel<<<(n + ROW_THREADS - 1) / ROW_THREADS, ROW_THREADS, 0, as_stream(s)>>>(mask, kind, out, n);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_layernorm_fwd(const void* z, const float* gamma, const float* beta, void* y, float* y32, float* mean, float* rstd,
                                  int rows, int H, int z_f32, const int32_t* rows_dev, crct_stream_t s) {
    if (!z || !gamma || !beta || !y || (mean == nullptr) != (rstd == nullptr)) CRCT_FAIL(CRCT_ERR_ARG, "crct_layernorm_fwd: bad pointer");
    if (int rc = check_row_width(H, "crct_layernorm_fwd")) return rc;
    if (rows <= 0) return CRCT_OK;
    if (z_f32)
        CRCT_CUDA(crct_launch_pdl(layernorm_fwd_kernel<true>, dim3(row_grid(rows)), dim3(ROW_THREADS), 0, as_stream(s), z,
                                  gamma, beta, reinterpret_cast<bf16*>(y), y32, mean, rstd, rows, H, rows_dev));
    else
        CRCT_CUDA(crct_launch_pdl(layernorm_fwd_kernel<false>, dim3(row_grid(rows)), dim3(ROW_THREADS), 0, as_stream(s), z,
                                  gamma, beta, reinterpret_cast<bf16*>(y), y32, mean, rstd, rows, H, rows_dev));
    return CRCT_OK;
}

extern "C" CRCT_API int crct_layernorm_rows_f32(const float* z, const float* gamma, const float* beta, const int32_t* row_index,
                                                long long row_step, float* out, int B, int H, crct_stream_t s) {
    if (!z || !gamma || !beta || !out) CRCT_FAIL(CRCT_ERR_ARG, "crct_layernorm_rows_f32: null pointer");
    if (int rc = check_row_width(H, "crct_layernorm_rows_f32")) return rc;
    if (B <= 0) return CRCT_OK;
    layernorm_rows_f32_kernel<<<row_grid(B), ROW_THREADS, 0, as_stream(s)>>>(z, gamma, beta, row_index, row_step, out, B, H);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_layernorm_bwd(const crct_ln_bwd_t* a, crct_stream_t s) {
    if (!a || !a->dy || !a->z || !a->mean || !a->rstd || !a->gamma || !a->dz) CRCT_FAIL(CRCT_ERR_ARG, "crct_layernorm_bwd: null pointer");
    if (int rc = check_row_width(a->H, "crct_layernorm_bwd")) return rc;
    if (a->rows <= 0) return CRCT_OK;
    const bool dzm = a->dzm != nullptr && a->p_out > 0.f;
    const int nch = (a->H / 8 + 31) / 32;
    const float sc_in = a->p_in > 0.f ? 1.f / (1.f - a->p_in) : 1.f, sc_out = a->p_out > 0.f ? 1.f / (1.f - a->p_out) : 1.f;
    const bool sums = a->dgamma || a->dbeta || a->dbias;
    if (!sums || a->z_f32 || a->rows_dev || a->drop_rows) {              // input gradient only (crct_layernorm_bwd_params does the sums)
        auto launch = [&](auto kern) {
            crct_launch_pdl(kern, dim3(row_grid(a->rows)), dim3(ROW_THREADS), 0, as_stream(s),
                reinterpret_cast<const bf16*>(a->dy), a->z, a->mean, a->rstd, a->gamma,
                reinterpret_cast<bf16*>(a->dz), dzm ? reinterpret_cast<bf16*>(a->dzm) : nullptr, a->rows, a->H,
                crct_drop_threshold(a->p_in), sc_in, a->seed_in, crct_drop_threshold(a->p_out), sc_out, a->seed_out,
                reinterpret_cast<const unsigned long long*>(a->salt), a->rows_dev, a->drop_rows);
        };
        if (a->z_f32) {
            switch (nch) {
                case 1: launch(ln_bwd_dz_kernel<1, true>); break;
                case 2: launch(ln_bwd_dz_kernel<2, true>); break;
                case 3: launch(ln_bwd_dz_kernel<3, true>); break;
                default: launch(ln_bwd_dz_kernel<4, true>); break;
            }
        } else {
            switch (nch) {
                case 1: launch(ln_bwd_dz_kernel<1, false>); break;
                case 2: launch(ln_bwd_dz_kernel<2, false>); break;
                case 3: launch(ln_bwd_dz_kernel<3, false>); break;
                default: launch(ln_bwd_dz_kernel<4, false>); break;
            }
        }
        CRCT_LAUNCH_CHECK();
        if (sums) return crct_layernorm_bwd_params(a, s);      // fp32 z / device row counts: the split form serves both calls
        return CRCT_OK;
    }
    int grid = 2 * crct_num_sms();
    if (grid > row_grid(a->rows)) grid = row_grid(a->rows);
    const size_t smem = (size_t)3 * (ROW_THREADS / 32) * a->H * sizeof(float);
    auto launch = [&](auto kern) {
        static bool configured[MAXC + 1] = {};    // all instantiations share one pointer type: track per chunk count
        if (!configured[nch]) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (ROW_THREADS / 32) * MAXC * 256 * (int)sizeof(float));
            configured[nch] = true;
        }
        kern<<<grid, ROW_THREADS, smem, as_stream(s)>>>(
            reinterpret_cast<const bf16*>(a->dy), reinterpret_cast<const bf16*>(a->z), a->mean, a->rstd, a->gamma,
            reinterpret_cast<bf16*>(a->dz), dzm ? reinterpret_cast<bf16*>(a->dzm) : nullptr, a->dgamma, a->dbeta, a->dbias,
            a->rows, a->H, crct_drop_threshold(a->p_in), sc_in, a->seed_in, crct_drop_threshold(a->p_out), sc_out, a->seed_out,
            reinterpret_cast<const unsigned long long*>(a->salt));
    };
    switch (nch) {
        case 1: launch(layernorm_bwd_kernel<1>); break;
        case 2: launch(layernorm_bwd_kernel<2>); break;
        case 3: launch(layernorm_bwd_kernel<3>); break;
        default: launch(layernorm_bwd_kernel<4>); break;
    }
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_layernorm_bwd_params(const crct_ln_bwd_t* a, crct_stream_t s) {
    if (!a || !a->dy || !a->z || !a->mean || !a->rstd) CRCT_FAIL(CRCT_ERR_ARG, "crct_layernorm_bwd_params: null pointer");
    if (int rc = check_row_width(a->H, "crct_layernorm_bwd_params")) return rc;
    if (a->rows <= 0 || (!a->dgamma && !a->dbeta && !a->dbias)) return CRCT_OK;
    const void* dzm = (a->dzm != nullptr && a->p_out > 0.f) ? a->dzm : a->dz;
    if (a->dbias && !dzm) CRCT_FAIL(CRCT_ERR_ARG, "crct_layernorm_bwd_params: dbias needs dzm (or dz)");
    const int gx = (a->H + 255) / 256;
    int gy = (2 * crct_num_sms() + gx - 1) / gx;
    const int max_gy = (a->rows + 63) / 64;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    auto launch = [&](auto kern) {
        crct_launch_pdl(kern, dim3(gx, gy), dim3(ROW_THREADS), 0, as_stream(s),
            reinterpret_cast<const bf16*>(a->dy), a->z, reinterpret_cast<const bf16*>(dzm), a->mean, a->rstd,
            a->dgamma, a->dbeta, a->dbias, a->rows, a->H, crct_drop_threshold(a->p_in), a->p_in > 0.f ? 1.f / (1.f - a->p_in) : 1.f, a->seed_in,
            reinterpret_cast<const unsigned long long*>(a->salt), a->rows_dev, a->drop_rows);
    };
    if (a->z_f32) launch(ln_bwd_params_kernel<true>);
    else launch(ln_bwd_params_kernel<false>);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_colsum_bf16(const void* x, float* out, int rows, int N, int ld, const int32_t* rows_dev, crct_stream_t s) {
    if (!x || !out) CRCT_FAIL(CRCT_ERR_ARG, "crct_colsum_bf16: null pointer");
    if ((N % 8) || (ld % 8)) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_colsum_bf16: N and ld must be multiples of 8");
    if (rows <= 0 || N <= 0) return CRCT_OK;
    const int gx = (N + 255) / 256;
    int gy = (2 * crct_num_sms() + gx - 1) / gx;
    const int max_gy = (rows + 63) / 64;
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    crct_launch_pdl(colsum_kernel, dim3(gx, gy), dim3(ROW_THREADS), 0, as_stream(s), reinterpret_cast<const bf16*>(x), out, rows, N, ld, rows_dev);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_softmax_rows(const float* x, void* out, int rows, int F, const int32_t* src_row, const int32_t* rows_dev,
                                          crct_stream_t s) {
    if (!x || !out) CRCT_FAIL(CRCT_ERR_ARG, "crct_softmax_rows: null pointer");
    if (int rc = check_row_width(F, "crct_softmax_rows")) return rc;
    if (rows <= 0) return CRCT_OK;
    softmax_rows_kernel<<<row_grid(rows), ROW_THREADS, 0, as_stream(s)>>>(x, reinterpret_cast<bf16*>(out), rows, F, src_row, rows_dev);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_embed_text_fwd(const crct_embed_text_t* a, crct_stream_t s) {
    if (!a || !a->ids || !a->types || !a->loc || !a->word || !a->pos || !a->type || !a->w_loc || !a->b_loc || !a->gamma ||
        !a->beta || !a->y)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_embed_text_fwd: null pointer");
    if (int rc = check_row_width(a->H, "crct_embed_text_fwd")) return rc;
    if (a->T > a->max_pos) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_embed_text_fwd: T=%d exceeds the position table (%d)", a->T, a->max_pos);
    TextEmbArgs k;
    k.ids = reinterpret_cast<const long long*>(a->ids); k.types = reinterpret_cast<const long long*>(a->types); k.loc = a->loc;
    k.word = a->word; k.pos = a->pos; k.type = a->type; k.w_loc = a->w_loc; k.b_loc = a->b_loc; k.gamma = a->gamma; k.beta = a->beta;
    k.y = reinterpret_cast<bf16*>(a->y); k.z = a->z; k.mean = a->mean; k.rstd = a->rstd;
    k.B = a->B; k.T = a->T; k.H = a->H;
    k.z_f32 = a->z_f32; k.src_row = a->src_row; k.rows_dev = a->rows_dev; k.y32 = a->y32;
    k.thr = crct_drop_threshold(a->dropout_p); k.scale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f; k.seed = a->seed;
    k.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    if (a->B * a->T <= 0) return CRCT_OK;
    embed_text_fwd_kernel<<<row_grid(a->B * a->T), ROW_THREADS, 0, as_stream(s)>>>(k);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_embed_text_bwd(const crct_embed_text_bwd_t* a, crct_stream_t s) {
    if (!a || !a->ids || !a->types || !a->loc || !a->dz || !a->g_word || !a->g_pos || !a->g_type || !a->g_wloc || !a->g_bloc)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_embed_text_bwd: null pointer");
    if (int rc = check_row_width(a->H, "crct_embed_text_bwd")) return rc;
    TextEmbBwdArgs k;
    k.ids = reinterpret_cast<const long long*>(a->ids); k.types = reinterpret_cast<const long long*>(a->types); k.loc = a->loc;
    k.dz = reinterpret_cast<const bf16*>(a->dz);
    k.g_word = a->g_word; k.g_pos = a->g_pos; k.g_type = a->g_type; k.g_wloc = a->g_wloc; k.g_bloc = a->g_bloc;
    k.B = a->B; k.T = a->T; k.H = a->H;
    k.src_row = a->src_row; k.rows_dev = a->rows_dev;
    const int rows = a->B * a->T;
    if (rows <= 0) return CRCT_OK;
    int grid = crct_num_sms();
    if (grid > row_grid(rows)) grid = row_grid(rows);
    embed_text_bwd_kernel<<<grid, ROW_THREADS, (size_t)(ROW_THREADS / 32) * a->H * sizeof(float), as_stream(s)>>>(k);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_embed_vis_fwd(const crct_embed_vis_t* a, crct_stream_t s) {
    if (!a || !a->g || !a->box || !a->cls || !a->w_loc || !a->b_loc || !a->color || !a->gamma || !a->beta || !a->y)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_embed_vis_fwd: null pointer");
    if (int rc = check_row_width(a->H, "crct_embed_vis_fwd")) return rc;
    VisEmbArgs k;
    k.g = reinterpret_cast<const bf16*>(a->g); k.box = a->box; k.cls = reinterpret_cast<const long long*>(a->cls);
    k.w_loc = a->w_loc; k.b_loc = a->b_loc; k.color = a->color; k.gamma = a->gamma; k.beta = a->beta;
    k.y = reinterpret_cast<bf16*>(a->y); k.z = a->z; k.mean = a->mean; k.rstd = a->rstd;
    k.rows = a->rows; k.H = a->H;
    k.z_f32 = a->z_f32; k.src_row = a->src_row; k.rows_dev = a->rows_dev; k.y32 = a->y32;
    k.thr = crct_drop_threshold(a->dropout_p); k.scale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f; k.seed = a->seed;
    k.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    if (a->rows <= 0) return CRCT_OK;
    embed_vis_fwd_kernel<<<row_grid(a->rows), ROW_THREADS, 0, as_stream(s)>>>(k);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_embed_vis_bwd(const crct_embed_vis_bwd_t* a, crct_stream_t s) {
    if (!a || !a->dz || !a->box || !a->cls || !a->g_color || !a->g_wloc) CRCT_FAIL(CRCT_ERR_ARG, "crct_embed_vis_bwd: null pointer");
    if (int rc = check_row_width(a->H, "crct_embed_vis_bwd")) return rc;
    VisEmbBwdArgs k;
    k.dz = reinterpret_cast<const bf16*>(a->dz); k.box = a->box; k.cls = reinterpret_cast<const long long*>(a->cls);
    k.g_color = a->g_color; k.g_wloc = a->g_wloc; k.rows = a->rows; k.H = a->H;
    k.src_row = a->src_row; k.rows_dev = a->rows_dev;
    if (a->rows <= 0) return CRCT_OK;
    int grid = crct_num_sms();
    if (grid > row_grid(a->rows)) grid = row_grid(a->rows);
    embed_vis_bwd_kernel<<<grid, ROW_THREADS, (size_t)(ROW_THREADS / 32) * a->H * sizeof(float), as_stream(s)>>>(k);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}
