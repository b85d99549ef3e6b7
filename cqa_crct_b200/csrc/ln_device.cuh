// LayerNorm row arithmetic shared by the row kernels (rowwise.cu) and the GEMM that applies the LayerNorm following its
// residual epilogue in the same launch (gemm_tcgen05.cu): ONE definition, so both forms produce the same bits.
// reference: CRCT/backbone/vilbert.py:281-294 (BertLayerNorm, eps inside the sqrt).
#pragma once
#include "common.cuh"

namespace {

constexpr int MAXC = 4;                 // 16-byte chunks per lane: rows up to 32*4*8 = 1024 columns
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


}  // namespace
