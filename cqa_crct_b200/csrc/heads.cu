// K7/K8 — the small-M tail of the model in fp32 on CUDA cores: poolers, classifier, regressor MLPs,
// the hybrid loss (cross-entropy + L1 / SmoothL1) and their backward.  M = batch size (80 / 512), N <= 1024,
// and two outputs are 1 and 2 columns wide, so none of this maps to UMMA tiles; it is 0.03 % of the FLOPs
// and is launch/latency bound.  Weights are read in fp32 straight from the master parameters, so the
// class logits and the regression value do not see bf16 weight rounding.
//
// reference: CRCT/backbone/vilbert.py:955-976 (poolers), :1052-1060 (classifier), CRCT/backbone/regressor.py:36-42,
//            vilbert.py:1586-1657 (losses, metrics), CRCT/backbone/encoder_decorator.py:144-153 (loss combine).
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>

namespace {

constexpr int BM = 64, BN = 64, BK = 64, LIN_THREADS = 256;

struct LinParams {
    const float* A; long long sa_m, sa_k;     // A(m,k) = A[m*sa_m + k*sa_k]
    const float* B; long long sb_k, sb_n;     // B(k,n) = B[k*sb_k + n*sb_n]
    float* C; long long ldc;                  // C[m*ldc + n]
    const float* bias;                        // [N] or null
    const float* dmask; long long ldm;        // multiply result by (dmask[m*ldm+n] > 0 ? 1 : slope), or null
    int M, N, K;
    int act;                                  // 0 none, 1 relu, 2 leaky(0.01), 3 tanh
    float slope;
    int accumulate;
    int ksplit;                               // CTAs of the cluster that share one output tile along K (1, 2, 4, 8)
};

// 64x64 output tile per CLUSTER, 4x4 per thread.  These problems are weight-streaming (M = batch, every weight used
// M times) and far too small to fill the GPU along M and N alone, so a serial K loop is a chain of exposed DRAM
// latencies (measured: 23-118 us per launch, 0.9 ms per step).  The K range is therefore split across the CTAs of a
// thread-block cluster (grid.z = cluster size S <= 8): every CTA loads its K chunk (at most two 64-deep slabs, all
// loads in flight at once), the partial tiles meet through distributed shared memory, and CTA r finishes rows
// [64 r / S, 64 (r+1) / S) of the tile (bias, activation, mask, store) — deterministic, no workspace, no atomics.
// (barrier.cluster only waits for non-exited threads, so groups of a cluster whose tile is out of range just leave.)
constexpr int MAX_BATCH = 12;
struct LinBatch {
    LinParams p[MAX_BATCH];
    int gy;            // M tiles per problem: blockIdx.y = problem * gy + m tile
};

// blockIdx.y also selects one of several INDEPENDENT small problems (e.g. wgrad + dgrad + bias-grad of one head layer,
// or the same layer of the two regressor pipes): one launch instead of up to 12 latency-bound ones.
__global__ void __launch_bounds__(LIN_THREADS) linear_f32_kernel(const LinBatch batch) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const LinParams& p = batch.p[blockIdx.y / batch.gy];
    // a problem with a short K uses ksplit < cluster size: the cluster then covers (cluster size / ksplit) N tiles
    const int S = p.ksplit, groups = (int)gridDim.z / S, grp = (int)blockIdx.z / S, r = (int)blockIdx.z % S;
    const int m0 = (int)(blockIdx.y % batch.gy) * BM, n0 = ((int)blockIdx.x * groups + grp) * BN;
    if (m0 >= p.M || n0 >= p.N) return;                 // a K-split group shares (m0, n0): it leaves together
    __shared__ float smem[2 * BK * (BM + 4)];
    float (*As)[BM + 4] = reinterpret_cast<float (*)[BM + 4]>(smem);
    float (*Bs)[BN + 4] = reinterpret_cast<float (*)[BN + 4]>(smem + BK * (BM + 4));
    float (*Ps)[BN + 1] = reinterpret_cast<float (*)[BN + 1]>(smem);       // partial tile, reuses the slabs after the K loop
    constexpr int PER = (BM * BK) / LIN_THREADS;        // 16 elements of A and of B per thread per slab
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kchunk = ((p.K + S - 1) / S + BK - 1) / BK * BK;
    const int k_lo = r * kchunk, k_hi = min(p.K, k_lo + kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool a_kfast = p.sa_k == 1, b_nfast = p.sb_n == 1;
    float ra[PER], rb[PER];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = tid + i * LIN_THREADS;
            const int kk = a_kfast ? (e % BK) : (e / BM), m = a_kfast ? (e / BK) : (e % BM);
            ra[i] = (m0 + m < p.M && k0 + kk < k_hi) ? (p.A ? p.A[(long long)(m0 + m) * p.sa_m + (long long)(k0 + kk) * p.sa_k] : 1.f) : 0.f;
            const int kb = b_nfast ? (e / BN) : (e % BK), n = b_nfast ? (e % BN) : (e / BK);
            rb[i] = (n0 + n < p.N && k0 + kb < k_hi) ? p.B[(long long)(k0 + kb) * p.sb_k + (long long)(n0 + n) * p.sb_n] : 0.f;
        }
    };
    if (k_lo < k_hi) fetch(k_lo);
    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = tid + i * LIN_THREADS;
            const int kk = a_kfast ? (e % BK) : (e / BM), m = a_kfast ? (e / BK) : (e % BM);
            As[kk][m] = ra[i];
            const int kb = b_nfast ? (e / BN) : (e % BK), n = b_nfast ? (e % BN) : (e / BK);
            Bs[kb][n] = rb[i];
        }
        __syncthreads();
        if (k0 + BK < k_hi) fetch(k0 + BK);
        // M = batch size (80) is 1.25 tiles: in the second m tile only the warps whose rows exist do the arithmetic
        // (a warp owns rows [m0 + 8 * warp, + 8)); the others only take part in the loads and barriers
        if (m0 + (threadIdx.x >> 5) * 8 < p.M) {
#pragma unroll 8
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Ps[ty * 4 + i][tx * 4 + j] = acc[i][j];
    cluster.sync();                                      // every CTA's partial tile is complete and visible
    const int rows = (BM + S - 1) / S;
    const int r_lo = r * rows, r_hi = min(BM, r_lo + rows);
    for (int e = tid; e < (r_hi - r_lo) * BN; e += LIN_THREADS) {
        const int ml = r_lo + e / BN, nl = e % BN;
        const int m = m0 + ml, n = n0 + nl;
        if (m >= p.M || n >= p.N) continue;
        float v = 0.f;
        for (int s2 = 0; s2 < S; ++s2) v += cluster.map_shared_rank(smem, grp * S + s2)[ml * (BN + 1) + nl];
        if (p.bias) v += p.bias[n];
        if (p.act == 1) v = fmaxf(v, 0.f);
        else if (p.act == 2) v = v > 0.f ? v : 0.01f * v;
        else if (p.act == 3) v = tanhf(v);
        if (p.dmask) v *= (p.dmask[(long long)m * p.ldm + n] > 0.f ? 1.f : p.slope);
        float* c = p.C + (long long)m * p.ldc + n;
        if (p.accumulate) *c += v; else *c = v;
    }
    cluster.sync();                                      // nobody leaves while its partial tile is still being read
}

// out[b, :] = float(src[b * row_stride + :])  (first token / region of every sample)
__global__ void gather_first_kernel(const bf16* __restrict__ src, long long row_stride, float* __restrict__ out, int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, c = i % H;
    out[i] = __bfloat162float(src[(long long)b * row_stride + c]);
}
// dst[b * row_stride + :] = bf16(g[b, :]); every other row of dst must already be zero
__global__ void scatter_first_kernel(const float* __restrict__ g, bf16* __restrict__ dst, long long row_stride, int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, c = i % H;
    dst[(long long)b * row_stride + c] = __float2bfloat16(g[i]);
}
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N, long long ld) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += x[(long long)m * ld + n];
    out[n] += s;
}
// pooled = dropout(pt * pv)   (vilbert.py:1055)
__global__ void pool_mul_fwd_kernel(const float* __restrict__ pt, const float* __restrict__ pv, float* __restrict__ out, int n,
                                    uint32_t thr, float scale, uint64_t seed, const unsigned long long* __restrict__ salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (thr != 0u && salt) seed ^= __ldg(salt);
    float v = pt[i] * pv[i];
    if (thr != 0u) v = crct_keep(seed, (uint64_t)i, thr) ? v * scale : 0.f;
    out[i] = v;
}
// dut = dpooled * keep * pv * [pt > 0] ; duv = dpooled * keep * pt * [pv > 0]   (ReLU poolers, vilbert.py:959-960,974-975)
__global__ void pool_mul_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ pt, const float* __restrict__ pv,
                                    float* __restrict__ dut, float* __restrict__ duv, int n, uint32_t thr, float scale, uint64_t seed,
                                    const unsigned long long* __restrict__ salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (thr != 0u && salt) seed ^= __ldg(salt);
    float d = dpooled[i];
    if (thr != 0u) d = crct_keep(seed, (uint64_t)i, thr) ? d * scale : 0.f;
    const float a = pt[i], b = pv[i];
    dut[i] = a > 0.f ? d * b : 0.f;
    duv[i] = b > 0.f ? d * a : 0.f;
}

struct LossParams {
    const float* logits; const float* reg; const long long* labels; const float* R;
    float* reg_pred; float* reg_loss; float* reg_l1; float* reg_dist;
    float* scalars;                    // [0] total loss, [1] nsp loss, [2] mean reg loss, [3] #within 5 %, [4] #within tol
    float* dlogits; float* dpre;       // gradients (training) or null
    int B;
    int l1; int smooth_kind; int unit_grads; float tol_margin; float nsp_coeff, reg_coeff;
};

// one CTA; vilbert.py:1586-1657 evaluated densely over all B rows (rows with R[:,1] != 1 are masked)
__global__ void __launch_bounds__(256) loss_kernel(const LossParams p) {
    __shared__ float red[5][256];
    float s_nsp = 0.f, s_valid = 0.f, s_reg = 0.f, s_r5 = 0.f, s_rt = 0.f;
    for (int b = threadIdx.x; b < p.B; b += 256) {
        const float val = p.R[b * 4 + 0], needs = p.R[b * 4 + 1], scale = p.R[b * 4 + 3];
        const bool need = needs == 1.0f;
        const float pred = p.reg[b];
        float out_pred = 0.f, out_loss = 0.f, out_l1 = 0.f, out_dist = 0.f, dl = 0.f;
        if (need) {
            const float tgt = val / scale;                       // vilbert.py:1617
            const float diff = pred - tgt;
            const float l1v = fabsf(diff);
            float lossv;
            if (p.l1) { lossv = l1v; dl = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f); }
            else {                                               // SmoothL1Loss(beta = 0.5), vilbert.py:1528
                const float beta = 0.5f;
                if (l1v < beta) { lossv = 0.5f * diff * diff / beta; dl = diff / beta; }
                else { lossv = l1v - 0.5f * beta; dl = diff > 0.f ? 1.f : -1.f; }
            }
            float dist = l1v / fabsf(tgt);                       // vilbert.py:1632-1636
            if (tgt == 0.f) dist = 1.f;
            const bool both0 = (pred == 0.f) && (tgt == 0.f);
            if (both0) dist = 0.f;
            if (dist <= 0.05f || both0) s_r5 += 1.f;
            if (l1v <= p.tol_margin) s_rt += 1.f;
            bool live = true;
            if (p.smooth_kind && fabsf(tgt) > 1.f) live = false;  // kind != 'L1' zeroes impossible targets, vilbert.py:1639-1641
            out_pred = pred * scale; out_l1 = l1v; out_dist = dist;
            if (live) out_loss = lossv; else dl = 0.f;
        }
        p.reg_pred[b] = out_pred; p.reg_loss[b] = out_loss; p.reg_l1[b] = out_l1; p.reg_dist[b] = out_dist;
        s_reg += out_loss;
        if (p.dpre) p.dpre[b] = (p.unit_grads ? dl : p.reg_coeff * dl / (float)p.B) * (1.f - pred * pred);   // through tanh (regressor.py:33)
        if (p.labels) {
            const long long lab = p.labels[b];
            const float z0 = p.logits[b * 2], z1 = p.logits[b * 2 + 1];
            const float mx = fmaxf(z0, z1);
            const float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
            if (lab != -1) { s_nsp += lse - (lab == 0 ? z0 : z1); s_valid += 1.f; }
        }
    }
    red[0][threadIdx.x] = s_nsp; red[1][threadIdx.x] = s_valid; red[2][threadIdx.x] = s_reg; red[3][threadIdx.x] = s_r5; red[4][threadIdx.x] = s_rt;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
#pragma unroll
            for (int k = 0; k < 5; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
        __syncthreads();
    }
    const float nvalid = fmaxf(red[1][0], 1.f);
    const float nsp = red[0][0] / nvalid, mreg = red[2][0] / (float)p.B;
    if (threadIdx.x == 0) {
        p.scalars[0] = p.labels ? p.nsp_coeff * nsp + p.reg_coeff * mreg : 0.f;   // encoder_decorator.py:144-153
        p.scalars[1] = nsp; p.scalars[2] = mreg; p.scalars[3] = red[3][0]; p.scalars[4] = red[4][0];
    }
    if (p.labels && p.dlogits) {
        for (int b = threadIdx.x; b < p.B; b += 256) {
            const long long lab = p.labels[b];
            const float z0 = p.logits[b * 2], z1 = p.logits[b * 2 + 1];
            const float mx = fmaxf(z0, z1);
            const float e0 = expf(z0 - mx), e1 = expf(z1 - mx), inv = 1.f / (e0 + e1);
            const float w = lab != -1 ? (p.unit_grads ? 1.f : p.nsp_coeff) / nvalid : 0.f;
            p.dlogits[b * 2] = w * (e0 * inv - (lab == 0 ? 1.f : 0.f));
            p.dlogits[b * 2 + 1] = w * (e1 * inv - (lab == 1 ? 1.f : 0.f));
        }
    }
}

// fused multi-tensor AdamW over the flat parameter arena (f1: CRCT/utils.py:228-249 + torch.optim.AdamW semantics),
// refreshing the bf16 operand copy in the same pass.  Every 64-element block of the arena carries a group id that
// selects (lr, weight_decay): language vs vision learning rate x decay vs no-decay (bias / LayerNorm).
struct AdamParams {
    float* w; const float* g; float* m; float* v; bf16* w16; const uint8_t* group; size_t n;
    float lr[4], wd[4];
    float beta1, beta2, eps, bc1, bc2_sqrt, gscale;
    const float* dyn;          // device {lr[4], bc1, bc2_sqrt} overriding the by-value fields (graph replay)
    size_t group_offset;
};
__device__ __forceinline__ void adamw_one(const AdamParams& p, float lr, float wd, float bc1, float bc2_sqrt, float g, float& w, float& m, float& v) {
    const float gr = g * p.gscale;
    w *= 1.f - lr * wd;
    m = p.beta1 * m + (1.f - p.beta1) * gr;
    v = p.beta2 * v + (1.f - p.beta2) * gr * gr;
    w -= lr / bc1 * m / (sqrtf(v) / bc2_sqrt + p.eps);
}
__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams p) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (size_t)gridDim.x * blockDim.x) {
        const int grp = p.group ? (p.group[(i + p.group_offset) >> 6] & 3) : 0;
        const float lr = p.dyn ? __ldg(p.dyn + grp) : p.lr[grp], wd = p.wd[grp];
        const float bc1 = p.dyn ? __ldg(p.dyn + 4) : p.bc1, bc2_sqrt = p.dyn ? __ldg(p.dyn + 5) : p.bc2_sqrt;
        float wi = p.w[i], mi = p.m[i], vi = p.v[i];
        adamw_one(p, lr, wd, bc1, bc2_sqrt, p.g[i], wi, mi, vi);
        p.m[i] = mi; p.v[i] = vi;
        p.w[i] = wi;
        if (p.w16) p.w16[i] = __float2bfloat16(wi);
    }
}
// 128-bit form: 4 consecutive elements per thread (same 64-element block, hence the same (lr, weight decay) group); the same
// per-element arithmetic as the scalar kernel, bit for bit.  HBM-bound: 30 B per parameter.
__global__ void __launch_bounds__(256) adamw_vec4_kernel(const AdamParams p) {
    const size_t n4 = p.n >> 2;
    const float bc1 = p.dyn ? __ldg(p.dyn + 4) : p.bc1, bc2_sqrt = p.dyn ? __ldg(p.dyn + 5) : p.bc2_sqrt;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const size_t i = q << 2;
        const int grp = p.group ? (p.group[(i + p.group_offset) >> 6] & 3) : 0;
        const float lr = p.dyn ? __ldg(p.dyn + grp) : p.lr[grp], wd = p.wd[grp];
        float4 w = *reinterpret_cast<const float4*>(p.w + i), m = *reinterpret_cast<const float4*>(p.m + i), v = *reinterpret_cast<const float4*>(p.v + i);
        const float4 g = __ldcs(reinterpret_cast<const float4*>(p.g + i));      // read once per step: streaming
        adamw_one(p, lr, wd, bc1, bc2_sqrt, g.x, w.x, m.x, v.x);
        adamw_one(p, lr, wd, bc1, bc2_sqrt, g.y, w.y, m.y, v.y);
        adamw_one(p, lr, wd, bc1, bc2_sqrt, g.z, w.z, m.z, v.z);
        adamw_one(p, lr, wd, bc1, bc2_sqrt, g.w, w.w, m.w, v.w);
        *reinterpret_cast<float4*>(p.m + i) = m;
        *reinterpret_cast<float4*>(p.v + i) = v;
        *reinterpret_cast<float4*>(p.w + i) = w;
        if (p.w16) {
            uint2 o;
            o.x = pack_bf16x2(w.x, w.y); o.y = pack_bf16x2(w.z, w.w);
            *reinterpret_cast<uint2*>(p.w16 + i) = o;
        }
    }
}

}  // namespace

static int fill_lin(LinParams& p, const crct_linear_t* a) {
    if (!a->B || !a->C) CRCT_FAIL(CRCT_ERR_ARG, "crct_linear_f32: null pointer");
    if (a->act < 0 || a->act > 3) CRCT_FAIL(CRCT_ERR_ARG, "crct_linear_f32: unknown activation %d", a->act);
    p.A = a->A; p.sa_m = a->sa_m; p.sa_k = a->sa_k; p.B = a->B; p.sb_k = a->sb_k; p.sb_n = a->sb_n;
    p.C = a->C; p.ldc = a->ldc; p.bias = a->bias; p.dmask = a->dmask; p.ldm = a->ldm;
    p.M = a->M; p.N = a->N; p.K = a->K; p.act = a->act; p.slope = a->slope; p.accumulate = a->accumulate;
    return CRCT_OK;
}

extern "C" CRCT_API int crct_linear_f32_batched(const crct_linear_t* problems, int count, crct_stream_t s) {
    if (!problems || count < 1 || count > MAX_BATCH) CRCT_FAIL(CRCT_ERR_ARG, "crct_linear_f32_batched: 1..%d problems", MAX_BATCH);
    LinBatch b;
    memset(&b, 0, sizeof(b));
    int gy = 0, n = 0, S = 1;
    // one 64-deep slab per CTA where the cluster size allows (K <= 512): the per-launch time of these latency chains fell from
    // 15-17 us to 11 us with one slab instead of two (tools/section_times.py, heads 0.60 -> 0.53 ms per step)
    constexpr int KSPLIT_SLABS = 1;
    auto ksplit = [](int K) {                            // at most KSPLIT_SLABS 64-deep slabs per CTA, up to 8 CTAs
        int sp = 1;
        while (sp < 8 && sp * KSPLIT_SLABS * BK < K) sp *= 2;
        return sp;
    };
    for (int i = 0; i < count; ++i) {
        if (problems[i].M <= 0 || problems[i].N <= 0 || problems[i].K <= 0) continue;
        if (int rc = fill_lin(b.p[n], &problems[i])) return rc;
        b.p[n].ksplit = ksplit(problems[i].K);
        S = max(S, b.p[n].ksplit);
        gy = max(gy, (problems[i].M + BM - 1) / BM);
        ++n;
    }
    if (n == 0) return CRCT_OK;
    int gx = 0;
    for (int i = 0; i < n; ++i) {
        const int groups = S / b.p[i].ksplit, tiles = (b.p[i].N + BN - 1) / BN;
        gx = max(gx, (tiles + groups - 1) / groups);
    }
    b.gy = gy;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(gx, gy * n, S);
    cfg.blockDim = dim3(LIN_THREADS, 1, 1);
    cfg.stream = as_stream(s);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = S;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    CRCT_CUDA(cudaLaunchKernelEx(&cfg, linear_f32_kernel, b));
    return CRCT_OK;
}

extern "C" CRCT_API int crct_linear_f32(const crct_linear_t* a, crct_stream_t s) {
    if (!a || !a->A) CRCT_FAIL(CRCT_ERR_ARG, "crct_linear_f32: null pointer");
    return crct_linear_f32_batched(a, 1, s);
}

extern "C" CRCT_API int crct_gather_first(const void* src, long long row_stride, float* out, int B, int H, crct_stream_t s) {
    if (!src || !out || B <= 0 || H <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_gather_first: bad argument");
    gather_first_kernel<<<(B * H + 255) / 256, 256, 0, as_stream(s)>>>(reinterpret_cast<const bf16*>(src), row_stride, out, B, H);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_scatter_first(const float* g, void* dst, long long row_stride, int B, int H, crct_stream_t s) {
    if (!g || !dst || B <= 0 || H <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_scatter_first: bad argument");
    scatter_first_kernel<<<(B * H + 255) / 256, 256, 0, as_stream(s)>>>(g, reinterpret_cast<bf16*>(dst), row_stride, B, H);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_colsum_f32(const float* x, float* out, int M, int N, long long ld, crct_stream_t s) {
    if (!x || !out) CRCT_FAIL(CRCT_ERR_ARG, "crct_colsum_f32: null pointer");
    if (M <= 0 || N <= 0) return CRCT_OK;
    colsum_f32_kernel<<<(N + 127) / 128, 128, 0, as_stream(s)>>>(x, out, M, N, ld);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_pool_mul_fwd(const float* pt, const float* pv, float* out, int n, float p, uint64_t seed, const uint64_t* salt,
                                          crct_stream_t s) {
    if (!pt || !pv || !out || n <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_pool_mul_fwd: bad argument");
    pool_mul_fwd_kernel<<<(n + 255) / 256, 256, 0, as_stream(s)>>>(pt, pv, out, n, crct_drop_threshold(p), p > 0.f ? 1.f / (1.f - p) : 1.f, seed,
                                                                 reinterpret_cast<const unsigned long long*>(salt));
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_pool_mul_bwd(const float* dpooled, const float* pt, const float* pv, float* dut, float* duv, int n, float p,
                                 uint64_t seed, const uint64_t* salt, crct_stream_t s) {
    if (!dpooled || !pt || !pv || !dut || !duv || n <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_pool_mul_bwd: bad argument");
    pool_mul_bwd_kernel<<<(n + 255) / 256, 256, 0, as_stream(s)>>>(dpooled, pt, pv, dut, duv, n, crct_drop_threshold(p),
                                                                 p > 0.f ? 1.f / (1.f - p) : 1.f, seed,
                                                                 reinterpret_cast<const unsigned long long*>(salt));
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_hybrid_loss(const crct_loss_t* a, crct_stream_t s) {
    if (!a || !a->logits || !a->reg || !a->R || !a->reg_pred || !a->reg_loss || !a->reg_l1 || !a->reg_dist || !a->scalars)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_hybrid_loss: null pointer");
    if (a->B <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_hybrid_loss: empty batch");
    LossParams p;
    p.logits = a->logits; p.reg = a->reg; p.labels = reinterpret_cast<const long long*>(a->labels); p.R = a->R;
    p.reg_pred = a->reg_pred; p.reg_loss = a->reg_loss; p.reg_l1 = a->reg_l1; p.reg_dist = a->reg_dist; p.scalars = a->scalars;
    p.dlogits = a->dlogits; p.dpre = a->dpre; p.B = a->B; p.l1 = a->l1; p.smooth_kind = a->zero_impossible; p.unit_grads = a->unit_grads;
    p.tol_margin = a->tol_margin; p.nsp_coeff = a->nsp_coeff; p.reg_coeff = a->reg_coeff;
    loss_kernel<<<1, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

extern "C" CRCT_API int crct_adamw(const crct_adamw_t* a, crct_stream_t s) {
    if (!a || !a->w || !a->g || !a->m || !a->v) CRCT_FAIL(CRCT_ERR_ARG, "crct_adamw: null pointer");
    if (a->n == 0) return CRCT_OK;
    if (a->step < 1 && !a->dyn) CRCT_FAIL(CRCT_ERR_ARG, "crct_adamw: step counts from 1");
    AdamParams p;
    p.w = a->w; p.g = a->g; p.m = a->m; p.v = a->v; p.w16 = reinterpret_cast<bf16*>(a->w_bf16); p.group = a->group_of_block64; p.n = a->n;
    for (int i = 0; i < 4; ++i) { p.lr[i] = a->lr[i]; p.wd[i] = a->weight_decay[i]; }
    p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps;
    p.bc1 = 1.f - powf(a->beta1, (float)a->step);
    p.bc2_sqrt = sqrtf(1.f - powf(a->beta2, (float)a->step));
    p.gscale = a->grad_scale;
    p.dyn = a->dyn;
    p.group_offset = a->group_offset;
    const size_t cap = (size_t)crct_num_sms() * 8;
    const bool vec4 = (a->n % 4 == 0) && (a->group_offset % 4 == 0) &&
                      !(((uintptr_t)a->w | (uintptr_t)a->g | (uintptr_t)a->m | (uintptr_t)a->v) & 15) && !((uintptr_t)a->w_bf16 & 7);
    size_t blocks = ((vec4 ? a->n / 4 : a->n) + 255) / 256;
    if (blocks > cap) blocks = cap;
    if (vec4) adamw_vec4_kernel<<<(unsigned)blocks, 256, 0, as_stream(s)>>>(p);
    else adamw_kernel<<<(unsigned)blocks, 256, 0, as_stream(s)>>>(p);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}

namespace {
// out[b,j] = x[b,j] * s[b]  (s_stride = 1)  or  x[b,j] * s[0]  (s_stride = 0): upstream gradient of the loss outputs
__global__ void scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ s, int s_stride, float* __restrict__ out, int B, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n) return;
    out[i] = x[i] * s[(i / n) * s_stride];
}
}  // namespace

extern "C" CRCT_API int crct_scale_rows(const float* x, const float* s, int s_stride, float* out, int B, int n, crct_stream_t st) {
    if (!x || !s || !out || B <= 0 || n <= 0 || (s_stride != 0 && s_stride != 1)) CRCT_FAIL(CRCT_ERR_ARG, "crct_scale_rows: bad argument");
    scale_rows_kernel<<<(B * n + 255) / 256, 256, 0, as_stream(st)>>>(x, s, s_stride, out, B, n);
    CRCT_LAUNCH_CHECK();
    return CRCT_OK;
}
