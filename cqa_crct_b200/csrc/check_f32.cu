// fp32 CHECK MODE — the same operators as the production kernels with fp32 activation storage and plain fp32 CUDA-core
// arithmetic (no tensor cores, exact erf, IEEE division), written for obviousness rather than speed.
//
// Purpose (north star: "tighter in an fp32 check mode"): `VisualDialogEncoder(params, precision='fp32')` runs the SAME
// host schedule — forward, hand-derived backward, dropout streams, gradient arena — through these kernels, so that
//   (1) the schedule and the backward derivation are pinned to the fp64 oracle at <= 1e-4 on the device, where the bf16
//       path can only be held to its own rounding floor (DESIGN.md "Numerical floor"), and
//   (2) the bf16 production path can be compared with an fp32 run ON THE GPU at sizes the CPU oracle cannot reach in
//       seconds, with dropout ON: both modes draw identical masks (same counter-based streams, same element counters).
// Every entry point takes the argument struct of its production counterpart; only the element type of the activation
// pointers differs (fp32 where the production kernel takes bf16).  Reference lines as cited on the production kernels.
#include "common.cuh"

namespace {

__device__ __forceinline__ bool keep_at(uint64_t seed, uint64_t idx, uint32_t thr) { return thr == 0u || crct_keep(seed, idx, thr); }

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }   // vilbert.py:111-117
__device__ __forceinline__ float gelu_exact_grad(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// ------------------------------------------------------------------------------------------------ GEMM
struct GemmP {
    const float* A; const float* B; float* D; float* D2; const float* bias; const float* aux;
    int M, N, K; long long sam, sak, sbn, sbk; int ldd, ldaux, epi, accumulate;
    uint32_t thr; float dscale; uint64_t seed; const unsigned long long* salt;
};
constexpr int GT = 64, GK = 16;

// D[m,n] = epilogue(sum_k A(m,k) B(n,k)); 64x64 tile per CTA, 4x4 outputs per thread, k accumulated in order
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmP p) {
    __shared__ float As[GK][GT + 1], Bs[GK][GT + 1];
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += GK) {
        for (int e = threadIdx.x; e < GT * GK; e += 256) {
            // the faster-varying thread index follows the contiguous axis of the operand
            int mm, kk;
            if (p.sak == 1) { kk = e % GK; mm = e / GK; } else { mm = e % GT; kk = e / GT; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < p.M && k < p.K) ? p.A[m * p.sam + k * p.sak] : 0.f;
            int nn, kb;
            if (p.sbk == 1) { kb = e % GK; nn = e / GK; } else { nn = e % GT; kb = e / GT; }
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < p.N && k2 < p.K) ? p.B[n * p.sbn + k2 * p.sbk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const uint64_t seed = p.salt ? (p.seed ^ *p.salt) : p.seed;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m >= p.M || n >= p.N) continue;
            float f = acc[i][j];
            const size_t o = (size_t)m * p.ldd + n;
            if (p.epi == CRCT_EPI_F32) { p.D[o] = p.accumulate ? p.D[o] + f : f; continue; }
            if (p.epi != CRCT_EPI_MUL && p.bias) f += p.bias[n];
            if (p.epi == CRCT_EPI_BIAS_GELU) {
                if (p.D2) p.D2[o] = gelu_exact_grad(f);
                f = gelu_exact(f);
            } else if (p.epi == CRCT_EPI_BIAS_RES) {
                if (p.thr != 0u) f = keep_at(seed, (uint64_t)m * (uint64_t)p.N + (uint64_t)n, p.thr) ? f * p.dscale : 0.f;
                if (p.aux) f += p.aux[(size_t)m * p.ldaux + n];
            } else if (p.epi == CRCT_EPI_MUL) {
                if (p.aux) f *= p.aux[(size_t)m * p.ldaux + n];
            }
            p.D[o] = f;
        }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
constexpr int LN_MAXC = 32;        // columns per lane: rows up to 1024 wide
constexpr float LN_EPS_F = 1e-12f;

__global__ void __launch_bounds__(256) ln_fwd_f32_kernel(const float* __restrict__ z, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ y, float* mean_out,
                                                         float* rstd_out, int rows, int H) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* zr = z + (size_t)row * H;
    float s = 0.f;
    for (int c = lane; c < H; c += 32) s += zr[c];
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
    for (int c = lane; c < H; c += 32) { const float d = zr[c] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + LN_EPS_F);
    for (int c = lane; c < H; c += 32) y[(size_t)row * H + c] = (zr[c] - mean) * rstd * gamma[c] + beta[c];
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

struct LnBwdP {
    const float* dy; const float* z; const float* mean; const float* rstd; const float* gamma; float* dz; float* dzm;
    float* dgamma; float* dbeta; float* dbias; int rows, H;
    uint32_t thr_in; float scale_in; uint64_t seed_in; uint32_t thr_out; float scale_out; uint64_t seed_out;
    const unsigned long long* salt;
};

__global__ void __launch_bounds__(256) ln_bwd_dz_f32_kernel(const LnBwdP p) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= p.rows) return;
    const unsigned long long sv = p.salt ? *p.salt : 0ull;
    const uint64_t seed_in = p.seed_in ^ sv, seed_out = p.seed_out ^ sv;
    const int H = p.H;
    const float mean = p.mean[row], rstd = p.rstd[row];
    float dxh[LN_MAXC], xh[LN_MAXC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
        const int c = lane + 32 * k;
        if (c < H) {
            float d = p.dy[(size_t)row * H + c];
            if (p.thr_in != 0u) d = keep_at(seed_in, (uint64_t)row * H + c, p.thr_in) ? d * p.scale_in : 0.f;
            xh[k] = (p.z[(size_t)row * H + c] - mean) * rstd;
            dxh[k] = d * p.gamma[c];
            s1 += dxh[k];
            s2 += dxh[k] * xh[k];
        }
    }
    const float m1 = warp_sum(s1) / (float)H, m2 = warp_sum(s2) / (float)H;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
        const int c = lane + 32 * k;
        if (c < H) {
            const float o = rstd * (dxh[k] - m1 - xh[k] * m2);
            p.dz[(size_t)row * H + c] = o;
            if (p.dzm) p.dzm[(size_t)row * H + c] = keep_at(seed_out, (uint64_t)row * H + c, p.thr_out) ? o * p.scale_out : 0.f;
        }
    }
}

// dgamma[c] += sum_rows dy*xhat, dbeta[c] += sum_rows dy, dbias[c] += sum_rows (dzm or dz); thread per column, rows split over blockIdx.y
__global__ void __launch_bounds__(256) ln_bwd_params_f32_kernel(const LnBwdP p) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= p.H) return;
    const uint64_t seed_in = p.seed_in ^ (p.salt ? *p.salt : 0ull);
    const float* dsrc = (p.dzm && p.thr_out != 0u) ? p.dzm : p.dz;
    float ag = 0.f, ab = 0.f, ad = 0.f;
    for (int row = blockIdx.y; row < p.rows; row += gridDim.y) {
        float d = p.dy[(size_t)row * p.H + c];
        if (p.thr_in != 0u) d = keep_at(seed_in, (uint64_t)row * p.H + c, p.thr_in) ? d * p.scale_in : 0.f;
        ag += d * ((p.z[(size_t)row * p.H + c] - p.mean[row]) * p.rstd[row]);
        ab += d;
        if (p.dbias) ad += dsrc[(size_t)row * p.H + c];
    }
    if (p.dgamma) atomicAdd(p.dgamma + c, ag);
    if (p.dbeta) atomicAdd(p.dbeta + c, ab);
    if (p.dbias) atomicAdd(p.dbias + c, ad);
}

// ------------------------------------------------------------------------------------------------ attention
struct AttnP {
    const float* q; const float* k; const float* v; int ldq, ldk, ldv; const float* mask_add;
    float* out; const float* out_c; int ldo; const float* dout; int lddo; float* lse; const float* lse_c;
    float* dq; float* dk; float* dv; int lddq, lddk, lddv;
    int B, nh, dh, Lq, Lk; float scale; uint32_t thr; float dscale; uint64_t seed; const unsigned long long* salt;
};
constexpr int ATT_MAXK = 512, ATT_MAXD = 64;

// warp per query row: scores over all keys in shared memory, softmax, dropout, P V
__global__ void __launch_bounds__(256) attn_fwd_f32_kernel(const AttnP p) {
    __shared__ float sc[8][ATT_MAXK];
    __shared__ float qv[8][ATT_MAXD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / p.nh, h = bh % p.nh;
    const int i = blockIdx.y * 8 + warp;
    if (i >= p.Lq) return;
    const uint64_t seed = p.salt ? (p.seed ^ *p.salt) : p.seed;
    const float* qi = p.q + ((size_t)b * p.Lq + i) * p.ldq + h * p.dh;
    for (int d = lane; d < p.dh; d += 32) qv[warp][d] = qi[d];
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < p.Lk; j += 32) {
        const float* kj = p.k + ((size_t)b * p.Lk + j) * p.ldk + h * p.dh;
        float s = 0.f;
        for (int d = 0; d < p.dh; ++d) s = fmaf(qv[warp][d], kj[d], s);
        s = s * p.scale + p.mask_add[(size_t)b * p.Lk + j];                 // vilbert.py:401-403
        sc[warp][j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < p.Lk; j += 32) { const float e = expf(sc[warp][j] - mx); sc[warp][j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < p.Lk; j += 32) {
        float pr = sc[warp][j] * inv;                                         // softmax, :405
        if (p.thr != 0u) pr = keep_at(seed, (uint64_t)(((uint32_t)bh * (uint32_t)p.Lq + (uint32_t)i) * (uint32_t)p.Lk + (uint32_t)j), p.thr) ? pr * p.dscale : 0.f;   // :407
        sc[warp][j] = pr;
    }
    __syncwarp();
    for (int d = lane; d < p.dh; d += 32) {
        float o = 0.f;
        for (int j = 0; j < p.Lk; ++j) o = fmaf(sc[warp][j], p.v[((size_t)b * p.Lk + j) * p.ldv + h * p.dh + d], o);
        p.out[((size_t)b * p.Lq + i) * p.ldo + h * p.dh + d] = o;
    }
    if (lane == 0 && p.lse) p.lse[(size_t)bh * p.Lq + i] = mx + logf(sum);
}

// pass Q: warp per query row -> dq
__global__ void __launch_bounds__(256) attn_bwd_q_f32_kernel(const AttnP p) {
    __shared__ float ds[8][ATT_MAXK];
    __shared__ float qv[8][ATT_MAXD], gv[8][ATT_MAXD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / p.nh, h = bh % p.nh;
    const int i = blockIdx.y * 8 + warp;
    if (i >= p.Lq) return;
    const uint64_t seed = p.salt ? (p.seed ^ *p.salt) : p.seed;
    const size_t qo = ((size_t)b * p.Lq + i);
    float Di = 0.f;
    for (int d = lane; d < p.dh; d += 32) {
        qv[warp][d] = p.q[qo * p.ldq + h * p.dh + d];
        gv[warp][d] = p.dout[qo * p.lddo + h * p.dh + d];
        Di += gv[warp][d] * p.out_c[qo * p.ldo + h * p.dh + d];
    }
    Di = warp_sum(Di);                                  // sum_j P_ij dP_ij = dO_i . O_i
    __syncwarp();
    const float lse = p.lse_c[(size_t)bh * p.Lq + i];
    for (int j = lane; j < p.Lk; j += 32) {
        const float* kj = p.k + ((size_t)b * p.Lk + j) * p.ldk + h * p.dh;
        const float* vj = p.v + ((size_t)b * p.Lk + j) * p.ldv + h * p.dh;
        float s = 0.f, dpd = 0.f;
        for (int d = 0; d < p.dh; ++d) { s = fmaf(qv[warp][d], kj[d], s); dpd = fmaf(gv[warp][d], vj[d], dpd); }
        const float pr = expf(s * p.scale + p.mask_add[(size_t)b * p.Lk + j] - lse);
        const bool kp = keep_at(seed, (uint64_t)(((uint32_t)bh * (uint32_t)p.Lq + (uint32_t)i) * (uint32_t)p.Lk + (uint32_t)j), p.thr);
        const float dP = kp ? dpd * p.dscale : 0.f;
        ds[warp][j] = pr * (dP - Di);
    }
    __syncwarp();
    for (int d = lane; d < p.dh; d += 32) {
        float o = 0.f;
        for (int j = 0; j < p.Lk; ++j) o = fmaf(ds[warp][j], p.k[((size_t)b * p.Lk + j) * p.ldk + h * p.dh + d], o);
        p.dq[qo * p.lddq + h * p.dh + d] = o * p.scale;
    }
}

// pass K: warp per key row -> dk, dv (lanes split the queries, then reduce)
__global__ void __launch_bounds__(256) attn_bwd_k_f32_kernel(const AttnP p) {
    __shared__ float kv[8][ATT_MAXD], vv[8][ATT_MAXD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / p.nh, h = bh % p.nh;
    const int j = blockIdx.y * 8 + warp;
    if (j >= p.Lk) return;
    const uint64_t seed = p.salt ? (p.seed ^ *p.salt) : p.seed;
    const size_t ko = ((size_t)b * p.Lk + j);
    for (int d = lane; d < p.dh; d += 32) { kv[warp][d] = p.k[ko * p.ldk + h * p.dh + d]; vv[warp][d] = p.v[ko * p.ldv + h * p.dh + d]; }
    __syncwarp();
    const float mk = p.mask_add[(size_t)b * p.Lk + j];
    float dk[ATT_MAXD], dv[ATT_MAXD];
#pragma unroll
    for (int d = 0; d < ATT_MAXD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int i = lane; i < p.Lq; i += 32) {
        const size_t qo = ((size_t)b * p.Lq + i);
        const float* qi = p.q + qo * p.ldq + h * p.dh;
        const float* gi = p.dout + qo * p.lddo + h * p.dh;
        const float* oi = p.out_c + qo * p.ldo + h * p.dh;
        float s = 0.f, dpd = 0.f, Di = 0.f;
        for (int d = 0; d < p.dh; ++d) { s = fmaf(qi[d], kv[warp][d], s); dpd = fmaf(gi[d], vv[warp][d], dpd); Di = fmaf(gi[d], oi[d], Di); }
        const float pr = expf(s * p.scale + mk - p.lse_c[(size_t)bh * p.Lq + i]);
        const bool kp = keep_at(seed, (uint64_t)(((uint32_t)bh * (uint32_t)p.Lq + (uint32_t)i) * (uint32_t)p.Lk + (uint32_t)j), p.thr);
        const float m = kp ? p.dscale : 0.f;
        const float dS = pr * (dpd * m - Di), pd = pr * m;
#pragma unroll
        for (int d = 0; d < ATT_MAXD; ++d)
            if (d < p.dh) { dk[d] = fmaf(dS, qi[d], dk[d]); dv[d] = fmaf(pd, gi[d], dv[d]); }
    }
#pragma unroll
    for (int d = 0; d < ATT_MAXD; ++d) {
        if (d < p.dh) {
            const float a = warp_sum(dk[d]), c = warp_sum(dv[d]);
            if (lane == 0) { p.dk[ko * p.lddk + h * p.dh + d] = a * p.scale; p.dv[ko * p.lddv + h * p.dh + d] = c; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ embeddings
struct TextP {
    const long long* ids; const long long* types; const float* loc; const float* word; const float* pos; const float* type;
    const float* w_loc; const float* b_loc; const float* gamma; const float* beta; float* y; float* z; float* mean; float* rstd;
    const float* dz; float* g_word; float* g_pos; float* g_type; float* g_wloc; float* g_bloc;
    int B, T, H, max_pos; uint32_t thr; float scale; uint64_t seed; const unsigned long long* salt;
};

__device__ __forceinline__ int first_qa(const long long* types_row, int T, int lane) {       // vilbert.py:331-341
    int first = T;
    for (int t = lane; t < T; t += 32) { const long long ty = types_row[t]; if ((ty == -1 || ty == 1) && t < first) first = t; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    return first;
}

// shared tail: LayerNorm of the lane-distributed row + dropout, vilbert.py:355-357 / 1493-1495
__device__ __forceinline__ void ln_tail(const float (&v)[LN_MAXC], int H, int lane, int row, const float* gamma, const float* beta, float* y,
                                        float* mean_out, float* rstd_out, uint32_t thr, float scale, uint64_t seed) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) if (lane + 32 * k < H) s += v[k];
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) if (lane + 32 * k < H) { const float d = v[k] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + LN_EPS_F);
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
        const int c = lane + 32 * k;
        if (c < H) {
            float o = (v[k] - mean) * rstd * gamma[c] + beta[c];
            if (thr != 0u) o = keep_at(seed, (uint64_t)row * H + c, thr) ? o * scale : 0.f;
            y[(size_t)row * H + c] = o;
        }
    }
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

__global__ void __launch_bounds__(256) embed_text_fwd_f32_kernel(const TextP a) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.B * a.T) return;
    const int b = row / a.T, t = row % a.T, H = a.H;
    const int first = first_qa(a.types + (size_t)b * a.T, a.T, lane);
    const long long ty = a.types[row], id = a.ids[row];
    const bool qa = (ty == -1 || ty == 1);
    const float* bx = a.loc + (size_t)row * 4;
    const bool loc_on = (fabsf(bx[0]) + fabsf(bx[1]) + fabsf(bx[2]) + fabsf(bx[3])) != 0.f;
    const long long ty_idx = ty == -1 ? 0 : ty;
    float v[LN_MAXC];
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
        const int c = lane + 32 * k;
        v[k] = 0.f;
        if (c < H) {
            float x = a.word[(size_t)id * H + c];
            if (qa) x += a.pos[(size_t)(t - first) * H + c];
            if (ty != 0) x += a.type[(size_t)ty_idx * H + c];
            if (loc_on) x += a.b_loc[c] + a.w_loc[c * 4] * bx[0] + a.w_loc[c * 4 + 1] * bx[1] + a.w_loc[c * 4 + 2] * bx[2] + a.w_loc[c * 4 + 3] * bx[3];
            v[k] = x;
            if (a.z) a.z[(size_t)row * H + c] = x;
        }
    }
    ln_tail(v, H, lane, row, a.gamma, a.beta, a.y, a.mean, a.rstd, a.thr, a.scale, a.salt ? (a.seed ^ *a.salt) : a.seed);
}

__global__ void __launch_bounds__(256) embed_text_bwd_f32_kernel(const TextP a) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.B * a.T) return;
    const int b = row / a.T, t = row % a.T, H = a.H;
    const int first = first_qa(a.types + (size_t)b * a.T, a.T, lane);
    const long long ty = a.types[row], id = a.ids[row];
    const bool qa = (ty == -1 || ty == 1);
    const float* bx = a.loc + (size_t)row * 4;
    const bool loc_on = (fabsf(bx[0]) + fabsf(bx[1]) + fabsf(bx[2]) + fabsf(bx[3])) != 0.f;
    const long long ty_idx = ty == -1 ? 0 : ty;
    for (int c = lane; c < H; c += 32) {
        const float d = a.dz[(size_t)row * H + c];
        atomicAdd(a.g_word + (size_t)id * H + c, d);
        if (qa) atomicAdd(a.g_pos + (size_t)(t - first) * H + c, d);
        if (ty != 0) atomicAdd(a.g_type + (size_t)ty_idx * H + c, d);
        if (loc_on) {
            atomicAdd(a.g_bloc + c, d);
            for (int k = 0; k < 4; ++k) atomicAdd(a.g_wloc + (size_t)c * 4 + k, d * bx[k]);
        }
    }
}

struct VisP {
    const float* g; const float* box; const long long* cls; const float* w_loc; const float* b_loc; const float* color;
    const float* gamma; const float* beta; float* y; float* z; float* mean; float* rstd;
    const float* dz; float* g_color; float* g_wloc;
    int rows, H; uint32_t thr; float scale; uint64_t seed; const unsigned long long* salt;
};

__global__ void __launch_bounds__(256) embed_vis_fwd_f32_kernel(const VisP a) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.rows) return;
    const int H = a.H;
    const float* bx = a.box + (size_t)row * 4;
    const long long cls = a.cls[row];
    float v[LN_MAXC];
#pragma unroll
    for (int k = 0; k < LN_MAXC; ++k) {
        const int c = lane + 32 * k;
        v[k] = 0.f;
        if (c < H) {
            const float x = a.g[(size_t)row * H + c] + a.color[(size_t)cls * H + c] + a.b_loc[c] + a.w_loc[c * 4] * bx[0] +
                            a.w_loc[c * 4 + 1] * bx[1] + a.w_loc[c * 4 + 2] * bx[2] + a.w_loc[c * 4 + 3] * bx[3];
            v[k] = x;
            if (a.z) a.z[(size_t)row * H + c] = x;
        }
    }
    ln_tail(v, H, lane, row, a.gamma, a.beta, a.y, a.mean, a.rstd, a.thr, a.scale, a.salt ? (a.seed ^ *a.salt) : a.seed);
}

__global__ void __launch_bounds__(256) embed_vis_bwd_f32_kernel(const VisP a) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= a.rows) return;
    const float* bx = a.box + (size_t)row * 4;
    const long long cls = a.cls[row];
    for (int c = lane; c < a.H; c += 32) {
        const float d = a.dz[(size_t)row * a.H + c];
        atomicAdd(a.g_color + (size_t)cls * a.H + c, d);
        for (int k = 0; k < 4; ++k) atomicAdd(a.g_wloc + (size_t)c * 4 + k, d * bx[k]);
    }
}

// ------------------------------------------------------------------------------------------------ small ones
__global__ void __launch_bounds__(256) softmax_rows_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int F) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * F;
    float mx = -INFINITY;
    for (int c = lane; c < F; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < F; c += 32) s += expf(xr[c] - mx);
    s = warp_sum(s);
    for (int c = lane; c < F; c += 32) out[(size_t)row * F + c] = expf(xr[c] - mx) / s;
}

__global__ void first_rows_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long row_stride, int B, int H, int scatter) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, c = i % H;
    if (scatter) dst[(long long)b * row_stride + c] = src[i];
    else dst[i] = src[(long long)b * row_stride + c];
}

inline int check_h(int H, const char* what) {
    if (H <= 0 || H > LN_MAXC * 32) CRCT_FAIL(CRCT_ERR_SHAPE, "%s: row width %d must be in 1..%d", what, H, LN_MAXC * 32);
    return CRCT_OK;
}
inline void drop_consts(float p, uint32_t& thr, float& scale) { thr = crct_drop_threshold(p); scale = p > 0.f ? 1.f / (1.f - p) : 1.f; }

}  // namespace

extern "C" CRCT_API int crct_f32_gemm(const crct_gemm_t* a, crct_stream_t s) {
    if (!a || !a->A || !a->B || !a->D) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_gemm: null pointer");
    if (a->M <= 0 || a->N <= 0 || a->K <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_f32_gemm: empty problem");
    if (a->epilogue < CRCT_EPI_BIAS || a->epilogue > CRCT_EPI_F32) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_gemm: unknown epilogue %d", a->epilogue);
    GemmP p;
    p.A = reinterpret_cast<const float*>(a->A); p.B = reinterpret_cast<const float*>(a->B); p.D = reinterpret_cast<float*>(a->D);
    p.D2 = reinterpret_cast<float*>(a->D2); p.bias = a->bias; p.aux = reinterpret_cast<const float*>(a->aux);
    p.M = a->M; p.N = a->N; p.K = a->K;
    const int lda = a->lda ? a->lda : (a->a_major ? a->M : a->K), ldb = a->ldb ? a->ldb : (a->b_major ? a->N : a->K);
    p.sam = a->a_major ? 1 : lda; p.sak = a->a_major ? lda : 1;
    p.sbn = a->b_major ? 1 : ldb; p.sbk = a->b_major ? ldb : 1;
    p.ldd = a->ldd ? a->ldd : a->N; p.ldaux = a->ldaux ? a->ldaux : a->N;
    p.epi = a->epilogue; p.accumulate = a->accumulate;
    drop_consts(a->epilogue == CRCT_EPI_BIAS_RES ? a->dropout_p : 0.f, p.thr, p.dscale);
    p.seed = a->seed; p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    gemm_f32_kernel<<<dim3((a->N + GT - 1) / GT, (a->M + GT - 1) / GT), 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                               int rows, int H, crct_stream_t s) {
    if (!z || !gamma || !beta || !y || (!mean != !rstd)) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_layernorm_fwd: null pointer");
    if (rows <= 0) return CRCT_OK;
    ln_fwd_f32_kernel<<<(rows + 7) / 8, 256, 0, as_stream(s)>>>(z, gamma, beta, y, mean, rstd, rows, H);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_layernorm_bwd(const crct_ln_bwd_t* a, crct_stream_t s) {
    if (!a || !a->dy || !a->z || !a->mean || !a->rstd || !a->gamma || !a->dz) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_layernorm_bwd: null pointer");
    if (int rc = check_h(a->H, "crct_f32_layernorm_bwd")) return rc;
    if (a->rows <= 0) return CRCT_OK;
    LnBwdP p;
    p.dy = reinterpret_cast<const float*>(a->dy); p.z = reinterpret_cast<const float*>(a->z); p.mean = a->mean; p.rstd = a->rstd;
    p.gamma = a->gamma; p.dz = reinterpret_cast<float*>(a->dz);
    p.dzm = a->p_out > 0.f ? reinterpret_cast<float*>(a->dzm) : nullptr;
    p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dbias = a->dbias; p.rows = a->rows; p.H = a->H;
    drop_consts(a->p_in, p.thr_in, p.scale_in);
    drop_consts(a->p_out, p.thr_out, p.scale_out);
    if (a->p_out > 0.f && !a->dzm) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_layernorm_bwd: p_out > 0 needs dzm");
    p.seed_in = a->seed_in; p.seed_out = a->seed_out; p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    ln_bwd_dz_f32_kernel<<<(a->rows + 7) / 8, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    if (a->dgamma || a->dbeta || a->dbias) {
        ln_bwd_params_f32_kernel<<<dim3((a->H + 255) / 256, 64), 256, 0, as_stream(s)>>>(p);
        CRCT_LAUNCH_CHECK();
    }
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_layernorm_bwd_params(const crct_ln_bwd_t* a, crct_stream_t s) {
    if (!a || !a->dy || !a->z || !a->mean || !a->rstd || !a->dz) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_layernorm_bwd_params: null pointer");
    if (a->rows <= 0 || !(a->dgamma || a->dbeta || a->dbias)) return CRCT_OK;
    LnBwdP p;
    memset(&p, 0, sizeof(p));
    p.dy = reinterpret_cast<const float*>(a->dy); p.z = reinterpret_cast<const float*>(a->z); p.mean = a->mean; p.rstd = a->rstd;
    p.dz = reinterpret_cast<float*>(a->dz); p.dzm = a->p_out > 0.f ? reinterpret_cast<float*>(a->dzm) : nullptr;
    p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dbias = a->dbias; p.rows = a->rows; p.H = a->H;
    drop_consts(a->p_in, p.thr_in, p.scale_in);
    drop_consts(a->p_out, p.thr_out, p.scale_out);
    p.seed_in = a->seed_in; p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    ln_bwd_params_f32_kernel<<<dim3((a->H + 255) / 256, 64), 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

static int fill_attn(AttnP& p, const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, const float* mask, int B, int nh, int dh,
                     int Lq, int Lk, float dropout_p, uint64_t seed, const uint64_t* salt, const char* what) {
    if (!q || !k || !v || !mask) CRCT_FAIL(CRCT_ERR_ARG, "%s: null pointer", what);
    if (B <= 0 || nh <= 0 || Lq <= 0 || Lk <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "%s: empty problem", what);
    if (dh <= 0 || dh > ATT_MAXD || Lk > ATT_MAXK) CRCT_FAIL(CRCT_ERR_SHAPE, "%s: head dim %d (max %d) / key length %d (max %d)", what, dh, ATT_MAXD, Lk, ATT_MAXK);
    if ((double)B * nh * Lq * Lk >= 4294967296.0) CRCT_FAIL(CRCT_ERR_SHAPE, "%s: B*nh*Lq*Lk must stay below 2^32", what);
    memset(&p, 0, sizeof(p));
    p.q = reinterpret_cast<const float*>(q); p.k = reinterpret_cast<const float*>(k); p.v = reinterpret_cast<const float*>(v);
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.mask_add = mask; p.B = B; p.nh = nh; p.dh = dh; p.Lq = Lq; p.Lk = Lk;
    p.scale = 1.0f / sqrtf((float)dh);
    drop_consts(dropout_p, p.thr, p.dscale);
    p.seed = seed; p.salt = reinterpret_cast<const unsigned long long*>(salt);
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_attn_fwd(const crct_attn_fwd_t* a, crct_stream_t s) {
    if (!a || !a->out) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_attn_fwd: null pointer");
    AttnP p;
    if (int rc = fill_attn(p, a->q, a->k, a->v, a->ldq, a->ldk, a->ldv, a->mask_add, a->B, a->nh, a->dh, a->Lq, a->Lk, a->dropout_p, a->seed, a->salt,
                           "crct_f32_attn_fwd")) return rc;
    p.out = reinterpret_cast<float*>(a->out); p.ldo = a->ldo; p.lse = a->lse;
    attn_fwd_f32_kernel<<<dim3(a->B * a->nh, (a->Lq + 7) / 8), 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_attn_bwd(const crct_attn_bwd_t* a, crct_stream_t s) {
    if (!a || !a->out || !a->dout || !a->lse || !a->dq || !a->dk || !a->dv) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_attn_bwd: null pointer");
    AttnP p;
    if (int rc = fill_attn(p, a->q, a->k, a->v, a->ldq, a->ldk, a->ldv, a->mask_add, a->B, a->nh, a->dh, a->Lq, a->Lk, a->dropout_p, a->seed, a->salt,
                           "crct_f32_attn_bwd")) return rc;
    p.out_c = reinterpret_cast<const float*>(a->out); p.ldo = a->ldo; p.dout = reinterpret_cast<const float*>(a->dout); p.lddo = a->lddo;
    p.lse_c = a->lse;
    p.dq = reinterpret_cast<float*>(a->dq); p.dk = reinterpret_cast<float*>(a->dk); p.dv = reinterpret_cast<float*>(a->dv);
    p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
    attn_bwd_q_f32_kernel<<<dim3(a->B * a->nh, (a->Lq + 7) / 8), 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    attn_bwd_k_f32_kernel<<<dim3(a->B * a->nh, (a->Lk + 7) / 8), 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_embed_text_fwd(const crct_embed_text_t* a, crct_stream_t s) {
    if (!a || !a->ids || !a->types || !a->loc || !a->word || !a->pos || !a->type || !a->w_loc || !a->b_loc || !a->gamma || !a->beta || !a->y)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_embed_text_fwd: null pointer");
    if (int rc = check_h(a->H, "crct_f32_embed_text_fwd")) return rc;
    TextP p;
    memset(&p, 0, sizeof(p));
    p.ids = reinterpret_cast<const long long*>(a->ids); p.types = reinterpret_cast<const long long*>(a->types); p.loc = a->loc;
    p.word = a->word; p.pos = a->pos; p.type = a->type; p.w_loc = a->w_loc; p.b_loc = a->b_loc; p.gamma = a->gamma; p.beta = a->beta;
    p.y = reinterpret_cast<float*>(a->y); p.z = reinterpret_cast<float*>(a->z); p.mean = a->mean; p.rstd = a->rstd;
    p.B = a->B; p.T = a->T; p.H = a->H; p.max_pos = a->max_pos;
    drop_consts(a->dropout_p, p.thr, p.scale);
    p.seed = a->seed; p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    embed_text_fwd_f32_kernel<<<(a->B * a->T + 7) / 8, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_embed_text_bwd(const crct_embed_text_bwd_t* a, crct_stream_t s) {
    if (!a || !a->ids || !a->types || !a->loc || !a->dz || !a->g_word || !a->g_pos || !a->g_type || !a->g_wloc || !a->g_bloc)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_embed_text_bwd: null pointer");
    TextP p;
    memset(&p, 0, sizeof(p));
    p.ids = reinterpret_cast<const long long*>(a->ids); p.types = reinterpret_cast<const long long*>(a->types); p.loc = a->loc;
    p.dz = reinterpret_cast<const float*>(a->dz); p.g_word = a->g_word; p.g_pos = a->g_pos; p.g_type = a->g_type; p.g_wloc = a->g_wloc;
    p.g_bloc = a->g_bloc; p.B = a->B; p.T = a->T; p.H = a->H;
    embed_text_bwd_f32_kernel<<<(a->B * a->T + 7) / 8, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_embed_vis_fwd(const crct_embed_vis_t* a, crct_stream_t s) {
    if (!a || !a->g || !a->box || !a->cls || !a->w_loc || !a->b_loc || !a->color || !a->gamma || !a->beta || !a->y)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_embed_vis_fwd: null pointer");
    if (int rc = check_h(a->H, "crct_f32_embed_vis_fwd")) return rc;
    VisP p;
    memset(&p, 0, sizeof(p));
    p.g = reinterpret_cast<const float*>(a->g); p.box = a->box; p.cls = reinterpret_cast<const long long*>(a->cls);
    p.w_loc = a->w_loc; p.b_loc = a->b_loc; p.color = a->color; p.gamma = a->gamma; p.beta = a->beta;
    p.y = reinterpret_cast<float*>(a->y); p.z = reinterpret_cast<float*>(a->z); p.mean = a->mean; p.rstd = a->rstd;
    p.rows = a->rows; p.H = a->H;
    drop_consts(a->dropout_p, p.thr, p.scale);
    p.seed = a->seed; p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    embed_vis_fwd_f32_kernel<<<(a->rows + 7) / 8, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_embed_vis_bwd(const crct_embed_vis_bwd_t* a, crct_stream_t s) {
    if (!a || !a->dz || !a->box || !a->cls || !a->g_color || !a->g_wloc) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_embed_vis_bwd: null pointer");
    VisP p;
    memset(&p, 0, sizeof(p));
    p.dz = reinterpret_cast<const float*>(a->dz); p.box = a->box; p.cls = reinterpret_cast<const long long*>(a->cls);
    p.g_color = a->g_color; p.g_wloc = a->g_wloc; p.rows = a->rows; p.H = a->H;
    embed_vis_bwd_f32_kernel<<<(a->rows + 7) / 8, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_softmax_rows(const float* x, float* out, int rows, int F, crct_stream_t s) {
    if (!x || !out || F <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_softmax_rows: bad argument");
    if (rows <= 0) return CRCT_OK;
    softmax_rows_f32_kernel<<<(rows + 7) / 8, 256, 0, as_stream(s)>>>(x, out, rows, F);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_gather_first(const float* src, long long row_stride, float* out, int B, int H, crct_stream_t s) {
    if (!src || !out || B <= 0 || H <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_gather_first: bad argument");
    first_rows_f32_kernel<<<(B * H + 255) / 256, 256, 0, as_stream(s)>>>(src, out, row_stride, B, H, 0);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_f32_scatter_first(const float* g, float* dst, long long row_stride, int B, int H, crct_stream_t s) {
    if (!g || !dst || B <= 0 || H <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_f32_scatter_first: bad argument");
    first_rows_f32_kernel<<<(B * H + 255) / 256, 256, 0, as_stream(s)>>>(g, dst, row_stride, B, H, 1);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}
