// K2/K3 — fused attention forward / backward for the three CRCT attention shapes
// (text self 16x48, visual self 16x64, co-attention 32x32 with Lq != Lk).
//
//   ctx = dropout(softmax(q k^T / sqrt(d) + additive_mask)) v        (one CTA per (sample, head))
//
// reference: CRCT/backbone/vilbert.py:397-412 (text), :527-543 (visual), :684-723 (both co-attention
// directions; the same kernel is called with (q=text,k/v=visual) and (q=visual,k/v=text)).
// Scores, probabilities and their gradients never leave registers: the reference materialises a
// [B,h,Lq,Lk] fp32 tensor four times per attention (78.7 MB per text layer at B=80).  The kernels are
// HBM-bound (FLOPs are 2.2 % of the model), so the contractions use warp-level mma.sync m16n8k16 bf16
// on smem-resident heads; the roofline that bounds them is bytes(q,k,v,ctx) / HBM bandwidth.
//
// Backward = two register-resident passes over the recomputed probabilities (no atomics, no cross-warp
// reduction): pass A owns 16 keys per warp -> dK, dV;  pass B owns 16 queries per warp -> dQ.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int ATT_THREADS = 128;
constexpr int NWARPS = ATT_THREADS / 32;
constexpr int KB = 64;                      // keys (pass B / forward) or queries (pass A) per inner block
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four transposed 8x8 b16 matrices: B fragments for two adjacent n-tiles from a row-major [k][n] tile
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

template <int DH>
struct Smem {
    static constexpr int LD = DH + 8;       // +16 B per row: conflict-free fragment reads, 16 B aligned rows
};

// A fragment (16 rows x 16 k) of a row-major smem tile X[row][k]
template <int LD>
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const bf16* X, int r0, int k0, int lane) {
    const int m = lane >> 3, rr = lane & 7;           // one ldmatrix.x4: (rows r0.., k0..) (rows r0+8.., k0..) (rows r0.., k0+8..) (rows r0+8.., k0+8..)
    ldmatrix_x4(a, ptx::smem_u32(X + (r0 + rr + (m & 1) * 8) * LD + k0 + (m >> 1) * 8));
}
// B fragments of n-tiles n0 and n0+8 for one 16-wide k-step, B[k][n] = Y[n][k] (k contiguous): r[0],r[1] -> n0 ; r[2],r[3] -> n0+8
template <int LD>
__device__ __forceinline__ void load_b_frag_x2(uint32_t (&r)[4], const bf16* Y, int n0, int k0, int lane) {
    const int m = lane >> 3, rr = lane & 7;
    ldmatrix_x4(r, ptx::smem_u32(Y + (n0 + rr + (m >> 1) * 8) * LD + k0 + (m & 1) * 8));
}
// B fragment (16 k x 8 n) where B[k][n] = Y[n][k], Y row-major in smem (k contiguous)
template <int LD>
__device__ __forceinline__ void load_b_frag(uint32_t& b0, uint32_t& b1, const bf16* Y, int n0, int k0, int g, int t) {
    b0 = *reinterpret_cast<const uint32_t*>(Y + (n0 + g) * LD + k0 + 2 * t);
    b1 = *reinterpret_cast<const uint32_t*>(Y + (n0 + g) * LD + k0 + 8 + 2 * t);
}
// B fragments for n-tiles n0 and n0+8 where B[k][n] = Y[k][n], Y row-major in smem (n contiguous)
template <int LD>
__device__ __forceinline__ void load_b_frag_trans(uint32_t (&r)[4], const bf16* Y, int k0, int n0, int lane) {
    const int m = lane >> 3, rr = lane & 7;
    const bf16* p = Y + (k0 + rr + (m & 1) * 8) * LD + n0 + (m >> 1) * 8;
    ldmatrix_x4_trans(r, ptx::smem_u32(p));
}

// cooperative load of `rows` rows (DH bf16 each, row stride ld in global) into smem [rows_pad][LD], zero padded
template <int DH>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int ld, int rows, int rows_pad) {
    constexpr int LD = Smem<DH>::LD;
    constexpr int CH = DH / 8;
    // cp.async (LDGSTS): fire-and-forget 16-byte copies, all in flight at once; rows beyond `rows` are zero-filled
    for (int i = threadIdx.x; i < rows_pad * CH; i += ATT_THREADS) {
        const int r = i / CH, c = i % CH;
        const bf16* g = src + (size_t)(r < rows ? r : 0) * ld + c * 8;
        const uint32_t nbytes = r < rows ? 16u : 0u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(ptx::smem_u32(dst + r * LD + c * 8)), "l"(g), "r"(nbytes) : "memory");
    }
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct FwdParams {
    const bf16* q; const bf16* k; const bf16* v;
    int ldq, ldk, ldv;
    const float* mask_add;
    bf16* out; int ldo;
    float* lse;
    int B, nh, Lq, Lk;
    float scale;
    uint32_t thr; float dscale; uint64_t seed; const unsigned long long* salt;
    const int* cu_q; const int* cu_k;       // packed rows: sample b owns rows [cu[b], cu[b+1]) (Lq / Lk are then the maxima)
};

template <int DH>
// occupancy is what hides the load-then-compute structure of a CTA: 5 CTAs per SM where registers allow (dh <= 48)
__global__ void __launch_bounds__(ATT_THREADS, DH <= 48 ? 5 : 4) attn_fwd_kernel(const FwdParams p) {
    constexpr int LD = Smem<DH>::LD;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    pdl_launch_dependents();
    pdl_wait();
    const int LQP_S = (p.Lq + 15) & ~15, LKP_S = (p.Lk + KB - 1) & ~(KB - 1);       // shared-memory layout: the maxima
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* Ks = Qs + LQP_S * LD;
    bf16* Vs = Ks + LKP_S * LD;
    float* mask_s = reinterpret_cast<float*>(Vs + LKP_S * LD);

    const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;
    // this sample's rows: padded layout b * L, or the packed range [cu[b], cu[b+1])
    size_t qrow0 = (size_t)b * p.Lq, krow0 = (size_t)b * p.Lk;
    int Lq = p.Lq, Lk = p.Lk;
    if (p.cu_q != nullptr) { const int c0 = __ldg(p.cu_q + b); qrow0 = c0; Lq = min(p.Lq, __ldg(p.cu_q + b + 1) - c0); }
    if (p.cu_k != nullptr) { const int c0 = __ldg(p.cu_k + b); krow0 = c0; Lk = min(p.Lk, __ldg(p.cu_k + b + 1) - c0); }
    const int LQP = (Lq + 15) & ~15, LKP = (Lk + KB - 1) & ~(KB - 1);
    load_tile<DH>(Qs, p.q + qrow0 * p.ldq + h * DH, p.ldq, Lq, LQP);
    load_tile<DH>(Ks, p.k + krow0 * p.ldk + h * DH, p.ldk, Lk, LKP);
    load_tile<DH>(Vs, p.v + krow0 * p.ldv + h * DH, p.ldv, Lk, LKP);
    // scores are kept in the log2 domain: s2 = s * scale * log2(e) + mask * log2(e), p = 2^(s2 - m2)  (one MUFU.EX2 per element)
    for (int i = threadIdx.x; i < LKP; i += ATT_THREADS)
        mask_s[i] = i < Lk ? (p.mask_add != nullptr ? p.mask_add[(size_t)b * p.Lk + i] * LOG2E : 0.f) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float scale2 = p.scale * LOG2E;
    const CrctDrop32 drop = crct_drop32((p.thr != 0u && p.salt) ? (p.seed ^ __ldg(p.salt)) : p.seed, p.thr);
    for (int q0 = warp * 16; q0 < LQP; q0 += NWARPS * 16) {
        uint32_t aq[DH / 16][4];
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) load_a_frag<LD>(aq[kk], Qs, q0, kk * 16, lane);
        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
        float o[DH / 8][4];
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }

        for (int kb = 0; kb < LKP; kb += KB) {
            float s[KB / 8][4];
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
#pragma unroll
                for (int j = 0; j < KB / 8; j += 2) {
                    uint32_t r[4];
                    load_b_frag_x2<LD>(r, Ks, kb + j * 8, kk * 16, lane);
                    mma16816(s[j], aq[kk], r[0], r[1]);
                    mma16816(s[j + 1], aq[kk], r[2], r[3]);
                }
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) {
                const float2 mk = *reinterpret_cast<const float2*>(&mask_s[kb + j * 8 + 2 * t]);
                s[j][0] = fmaf(s[j][0], scale2, mk.x); s[j][1] = fmaf(s[j][1], scale2, mk.y);
                s[j][2] = fmaf(s[j][2], scale2, mk.x); s[j][3] = fmaf(s[j][3], scale2, mk.y);
                mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
            }
            float corr[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
                const float m_new = fmaxf(m_run[r], mx[r]);
                corr[r] = fast_exp2(m_run[r] - m_new);
                m_run[r] = m_new;
                l_run[r] *= corr[r];
            }
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) { o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1]; }
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) {
                s[j][0] = fast_exp2(s[j][0] - m_run[0]); s[j][1] = fast_exp2(s[j][1] - m_run[0]);
                s[j][2] = fast_exp2(s[j][2] - m_run[1]); s[j][3] = fast_exp2(s[j][3] - m_run[1]);
                l_run[0] += s[j][0] + s[j][1];
                l_run[1] += s[j][2] + s[j][3];
            }
            if (p.thr != 0u) {                 // dropout on the probabilities (vilbert.py:407); l_run stays undropped
                const uint32_t i0b = (blockIdx.x * p.Lq + q0 + g) * p.Lk + kb + 2 * t, i1b = i0b + 8u * p.Lk;
#pragma unroll
                for (int j = 0; j < KB / 8; ++j) {
                    bool k0, k1, k2, k3;
                    crct_keep2_32(drop, i0b + j * 8, k0, k1);
                    crct_keep2_32(drop, i1b + j * 8, k2, k3);
                    s[j][0] = k0 ? s[j][0] * p.dscale : 0.f; s[j][1] = k1 ? s[j][1] * p.dscale : 0.f;
                    s[j][2] = k2 ? s[j][2] * p.dscale : 0.f; s[j][3] = k3 ? s[j][3] * p.dscale : 0.f;
                }
            }
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk) {
                uint32_t ap[4];
                ap[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
                ap[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
                ap[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
                ap[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
                for (int nt = 0; nt < DH / 16; ++nt) {
                    uint32_t r[4];
                    load_b_frag_trans<LD>(r, Vs, kb + kk * 16, nt * 16, lane);
                    mma16816(o[2 * nt], ap, r[0], r[1]);
                    mma16816(o[2 * nt + 1], ap, r[2], r[3]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
            l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
        }
        const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
        const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) {
            const int col = h * DH + j * 8 + 2 * t;
            if (r0 < Lq) *reinterpret_cast<uint32_t*>(p.out + (qrow0 + r0) * p.ldo + col) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
            if (r1 < Lq) *reinterpret_cast<uint32_t*>(p.out + (qrow0 + r1) * p.ldo + col) = pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
        }
        if (p.lse != nullptr && t == 0) {
            if (r0 < Lq) p.lse[(size_t)blockIdx.x * p.Lq + r0] = (m_run[0] + __log2f(l_run[0])) * LN2;
            if (r1 < Lq) p.lse[(size_t)blockIdx.x * p.Lq + r1] = (m_run[1] + __log2f(l_run[1])) * LN2;
        }
    }
}

struct BwdParams {
    const bf16* q; const bf16* k; const bf16* v;
    int ldq, ldk, ldv;
    const float* mask_add;
    const bf16* out; int ldo;
    const bf16* dout; int lddo;
    const float* lse;
    bf16* dq; bf16* dk; bf16* dv;
    int lddq, lddk, lddv;
    int B, nh, Lq, Lk;
    float scale;
    uint32_t thr; float dscale; uint64_t seed; const unsigned long long* salt;
    const int* cu_q; const int* cu_k;
};

template <int DH, int PASS>     // PASS 0: this warp owns 16 keys -> dK, dV ; PASS 1: this warp owns 16 queries -> dQ
// PASS 0: dK, dV (warp owns 16 keys).  PASS 1: dQ (warp owns 16 queries; recomputes S and dP).  PASS 2: both in one kernel —
// pass A also parks dS^T (bf16, exactly the operand of its dK contraction) in shared memory and, after one barrier, a third
// phase contracts it with K: 5 contractions and one operand load instead of 7 and two.  Needs LKP x (LQP + 8) x 2 more
// bytes of shared memory (35 KB at 124 x 124), so 2 CTAs per SM; shapes whose dS^T does not fit use passes 0 + 1.
// 3 CTAs per SM for the split passes and for the fused kernel at dh = 32 (co-attention: 49 KB of shared memory per CTA);
// <= 168 registers; pass A of dh = 64 would spill 170 B and stays at 2
__global__ void __launch_bounds__(ATT_THREADS, ((PASS == 2 && DH != 32) || (DH == 64 && PASS == 0)) ? 2 : 3) attn_bwd_kernel(const BwdParams p) {
    constexpr int LD = Smem<DH>::LD;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    pdl_launch_dependents();
    pdl_wait();
    const int LQP_S = (p.Lq + KB - 1) & ~(KB - 1), LKP_S = (p.Lk + KB - 1) & ~(KB - 1);      // shared-memory layout: the maxima
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* dOs = Qs + LQP_S * LD;
    bf16* Ks = dOs + LQP_S * LD;
    bf16* Vs = Ks + LKP_S * LD;
    float* mask_s = reinterpret_cast<float*>(Vs + LKP_S * LD);
    float* lse_s = mask_s + LKP_S;
    float* D_s = lse_s + LQP_S;
    const int LDS = LQP_S + 8;                                     // dS^T[key][query] row stride: conflict-free 32-bit stores
    bf16* dSs = reinterpret_cast<bf16*>(D_s + LQP_S);              // PASS 2 only

    const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;
    size_t qrow0 = (size_t)b * p.Lq, krow0 = (size_t)b * p.Lk;     // padded layout, or the packed range [cu[b], cu[b+1])
    int Lq = p.Lq, Lk = p.Lk;
    if (p.cu_q != nullptr) { const int c0 = __ldg(p.cu_q + b); qrow0 = c0; Lq = min(p.Lq, __ldg(p.cu_q + b + 1) - c0); }
    if (p.cu_k != nullptr) { const int c0 = __ldg(p.cu_k + b); krow0 = c0; Lk = min(p.Lk, __ldg(p.cu_k + b + 1) - c0); }
    const int LQP = (Lq + KB - 1) & ~(KB - 1), LKP = (Lk + KB - 1) & ~(KB - 1);
    load_tile<DH>(Qs, p.q + qrow0 * p.ldq + h * DH, p.ldq, Lq, LQP);
    load_tile<DH>(dOs, p.dout + qrow0 * p.lddo + h * DH, p.lddo, Lq, LQP);
    load_tile<DH>(Ks, p.k + krow0 * p.ldk + h * DH, p.ldk, Lk, LKP);
    load_tile<DH>(Vs, p.v + krow0 * p.ldv + h * DH, p.ldv, Lk, LKP);
    for (int i = threadIdx.x; i < LKP; i += ATT_THREADS)
        mask_s[i] = i < Lk ? (p.mask_add != nullptr ? p.mask_add[(size_t)b * p.Lk + i] * LOG2E : 0.f) : -INFINITY;
    // D[q] = sum_d dO[q,d] * O[q,d]; padded queries get lse = +inf so their probabilities vanish
    for (int i = threadIdx.x; i < LQP; i += ATT_THREADS) {
        float d = 0.f, l = INFINITY;
        if (i < Lq) {
            const bf16* o = p.out + (qrow0 + i) * p.ldo + h * DH;
            const bf16* dd = p.dout + (qrow0 + i) * p.lddo + h * DH;
#pragma unroll
            for (int c = 0; c < DH / 8; ++c) {
                float fo[8], fd[8];
                load8_bf16(o + c * 8, fo);
                load8_bf16(dd + c * 8, fd);
#pragma unroll
                for (int j = 0; j < 8; ++j) d += fo[j] * fd[j];
            }
            l = p.lse[(size_t)blockIdx.x * p.Lq + i] * LOG2E;
        }
        D_s[i] = d;
        lse_s[i] = l;
    }
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint32_t drop_base = blockIdx.x * (uint32_t)p.Lq * (uint32_t)p.Lk;
    const CrctDrop32 drop = crct_drop32((p.thr != 0u && p.salt) ? (p.seed ^ __ldg(p.salt)) : p.seed, p.thr);
    const float scale2 = p.scale * LOG2E;

    // ---------------- pass A: this warp owns 16 keys -> dK, dV ----------------
    if constexpr (PASS == 0 || PASS == 2)
    for (int k0 = warp * 16; k0 < LKP && k0 < ((Lk + 15) & ~15); k0 += NWARPS * 16) {
        uint32_t ak[DH / 16][4], av[DH / 16][4];
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
            load_a_frag<LD>(ak[kk], Ks, k0, kk * 16, lane);
            load_a_frag<LD>(av[kk], Vs, k0, kk * 16, lane);
        }
        const float mrow[2] = {mask_s[k0 + g], mask_s[k0 + g + 8]};
        float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) { dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f; dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f; }
        for (int qb = 0; qb < LQP; qb += KB) {
            float st[KB / 8][4], dpt[KB / 8][4];
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) { st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f; dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
#pragma unroll
                for (int j = 0; j < KB / 8; j += 2) {
                    uint32_t r[4];
                    load_b_frag_x2<LD>(r, Qs, qb + j * 8, kk * 16, lane);
                    mma16816(st[j], ak[kk], r[0], r[1]);             // S^T = K Q^T
                    mma16816(st[j + 1], ak[kk], r[2], r[3]);
                    load_b_frag_x2<LD>(r, dOs, qb + j * 8, kk * 16, lane);
                    mma16816(dpt[j], av[kk], r[0], r[1]);            // dP^T = V dO^T
                    mma16816(dpt[j + 1], av[kk], r[2], r[3]);
                }
            // P^T, dS^T in place: st <- P^T (dropped, for dV), dpt <- dS^T * scale (for dK)
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) {
                const int qi0 = qb + j * 8 + 2 * t;
                const float2 ls = *reinterpret_cast<const float2*>(&lse_s[qi0]);
                const float2 dd = *reinterpret_cast<const float2*>(&D_s[qi0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int r = c >> 1;
                    const float pr = fast_exp2(fmaf(st[j][c], scale2, mrow[r]) - ((c & 1) ? ls.y : ls.x));
                    float fac = 1.f;
                    if (p.thr != 0u) fac = crct_keep32(drop, drop_base + (uint32_t)(qi0 + (c & 1)) * p.Lk + (k0 + g + 8 * r)) ? p.dscale : 0.f;
                    st[j][c] = pr * fac;
                    dpt[j][c] = pr * (dpt[j][c] * fac - ((c & 1) ? dd.y : dd.x)) * p.scale;
                }
                if constexpr (PASS == 2) {                           // rows = keys k0+g / k0+g+8, columns = queries qi0, qi0+1
                    *reinterpret_cast<uint32_t*>(dSs + (k0 + g) * LDS + qi0) = pack_bf16x2(dpt[j][0], dpt[j][1]);
                    *reinterpret_cast<uint32_t*>(dSs + (k0 + g + 8) * LDS + qi0) = pack_bf16x2(dpt[j][2], dpt[j][3]);
                }
            }
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk) {
                uint32_t ap[4], ads[4];
                ap[0] = pack_bf16x2(st[2 * kk][0], st[2 * kk][1]);       ap[1] = pack_bf16x2(st[2 * kk][2], st[2 * kk][3]);
                ap[2] = pack_bf16x2(st[2 * kk + 1][0], st[2 * kk + 1][1]); ap[3] = pack_bf16x2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
                ads[0] = pack_bf16x2(dpt[2 * kk][0], dpt[2 * kk][1]);       ads[1] = pack_bf16x2(dpt[2 * kk][2], dpt[2 * kk][3]);
                ads[2] = pack_bf16x2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]); ads[3] = pack_bf16x2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
                for (int nt = 0; nt < DH / 16; ++nt) {
                    uint32_t r[4];
                    load_b_frag_trans<LD>(r, dOs, qb + kk * 16, nt * 16, lane);   // dV += P^T dO
                    mma16816(dv[2 * nt], ap, r[0], r[1]);
                    mma16816(dv[2 * nt + 1], ap, r[2], r[3]);
                    load_b_frag_trans<LD>(r, Qs, qb + kk * 16, nt * 16, lane);    // dK += dS^T Q
                    mma16816(dk[2 * nt], ads, r[0], r[1]);
                    mma16816(dk[2 * nt + 1], ads, r[2], r[3]);
                }
            }
        }
        const int r0 = k0 + g, r1 = k0 + g + 8;
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) {
            const int col = h * DH + j * 8 + 2 * t;
            if (r0 < Lk) {
                *reinterpret_cast<uint32_t*>(p.dk + (krow0 + r0) * p.lddk + col) = pack_bf16x2(dk[j][0], dk[j][1]);
                *reinterpret_cast<uint32_t*>(p.dv + (krow0 + r0) * p.lddv + col) = pack_bf16x2(dv[j][0], dv[j][1]);
            }
            if (r1 < Lk) {
                *reinterpret_cast<uint32_t*>(p.dk + (krow0 + r1) * p.lddk + col) = pack_bf16x2(dk[j][2], dk[j][3]);
                *reinterpret_cast<uint32_t*>(p.dv + (krow0 + r1) * p.lddv + col) = pack_bf16x2(dv[j][2], dv[j][3]);
            }
        }
    }

    // ---------------- phase C (fused variant): dQ = dS K from the parked dS^T; this warp owns 16 queries ----------------
    if constexpr (PASS == 2) {
        __syncthreads();
        const int LK16 = (Lk + 15) & ~15;
        for (int q0 = warp * 16; q0 < ((Lq + 15) & ~15); q0 += NWARPS * 16) {
            float dq[DH / 8][4];
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }
            for (int k0 = 0; k0 < LK16; k0 += 16) {
                // A fragment dS[q0 .. q0+15][k0 .. k0+15] out of dS^T[key][query]: four transposed 8x8 blocks
                // (queries +0 / +8 within keys +0, then within keys +8)
                uint32_t ads[4];
                const int m = lane >> 3, rr = lane & 7;
                ldmatrix_x4_trans(ads, ptx::smem_u32(dSs + (k0 + rr + (m >> 1) * 8) * LDS + q0 + (m & 1) * 8));
#pragma unroll
                for (int nt = 0; nt < DH / 16; ++nt) {
                    uint32_t r[4];
                    load_b_frag_trans<LD>(r, Ks, k0, nt * 16, lane);              // dQ += dS K
                    mma16816(dq[2 * nt], ads, r[0], r[1]);
                    mma16816(dq[2 * nt + 1], ads, r[2], r[3]);
                }
            }
            const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) {
                const int col = h * DH + j * 8 + 2 * t;
                if (r0 < Lq) *reinterpret_cast<uint32_t*>(p.dq + (qrow0 + r0) * p.lddq + col) = pack_bf16x2(dq[j][0], dq[j][1]);
                if (r1 < Lq) *reinterpret_cast<uint32_t*>(p.dq + (qrow0 + r1) * p.lddq + col) = pack_bf16x2(dq[j][2], dq[j][3]);
            }
        }
    }

    // ---------------- pass B: this warp owns 16 queries -> dQ ----------------
    if constexpr (PASS == 1)
    for (int q0 = warp * 16; q0 < ((Lq + 15) & ~15); q0 += NWARPS * 16) {
        uint32_t aq[DH / 16][4], ado[DH / 16][4];
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
            load_a_frag<LD>(aq[kk], Qs, q0, kk * 16, lane);
            load_a_frag<LD>(ado[kk], dOs, q0, kk * 16, lane);
        }
        const float lrow[2] = {lse_s[q0 + g], lse_s[q0 + g + 8]};
        const float drow[2] = {D_s[q0 + g], D_s[q0 + g + 8]};
        float dq[DH / 8][4];
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }
        for (int kb = 0; kb < LKP; kb += KB) {
            float s[KB / 8][4], dp[KB / 8][4];
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk)
#pragma unroll
                for (int j = 0; j < KB / 8; j += 2) {
                    uint32_t r[4];
                    load_b_frag_x2<LD>(r, Ks, kb + j * 8, kk * 16, lane);
                    mma16816(s[j], aq[kk], r[0], r[1]);              // S = Q K^T
                    mma16816(s[j + 1], aq[kk], r[2], r[3]);
                    load_b_frag_x2<LD>(r, Vs, kb + j * 8, kk * 16, lane);
                    mma16816(dp[j], ado[kk], r[0], r[1]);            // dP = dO V^T
                    mma16816(dp[j + 1], ado[kk], r[2], r[3]);
                }
#pragma unroll
            for (int j = 0; j < KB / 8; ++j) {
                const int key0 = kb + j * 8 + 2 * t;
                const float2 mk = *reinterpret_cast<const float2*>(&mask_s[key0]);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    bool k0 = true, k1 = true;
                    if (p.thr != 0u) crct_keep2_32(drop, drop_base + (uint32_t)(q0 + g + 8 * r) * p.Lk + key0, k0, k1);
                    const float pr0 = fast_exp2(fmaf(s[j][2 * r], scale2, mk.x) - lrow[r]);
                    const float pr1 = fast_exp2(fmaf(s[j][2 * r + 1], scale2, mk.y) - lrow[r]);
                    dp[j][2 * r] = pr0 * ((k0 ? dp[j][2 * r] * p.dscale : 0.f) - drow[r]) * p.scale;
                    dp[j][2 * r + 1] = pr1 * ((k1 ? dp[j][2 * r + 1] * p.dscale : 0.f) - drow[r]) * p.scale;
                }
            }
#pragma unroll
            for (int kk = 0; kk < KB / 16; ++kk) {
                uint32_t ads[4];
                ads[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);       ads[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
                ads[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]); ads[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
                for (int nt = 0; nt < DH / 16; ++nt) {
                    uint32_t r[4];
                    load_b_frag_trans<LD>(r, Ks, kb + kk * 16, nt * 16, lane);    // dQ += dS K
                    mma16816(dq[2 * nt], ads, r[0], r[1]);
                    mma16816(dq[2 * nt + 1], ads, r[2], r[3]);
                }
            }
        }
        const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) {
            const int col = h * DH + j * 8 + 2 * t;
            if (r0 < Lq) *reinterpret_cast<uint32_t*>(p.dq + (qrow0 + r0) * p.lddq + col) = pack_bf16x2(dq[j][0], dq[j][1]);
            if (r1 < Lq) *reinterpret_cast<uint32_t*>(p.dq + (qrow0 + r1) * p.lddq + col) = pack_bf16x2(dq[j][2], dq[j][3]);
        }
    }
}

template <int DH>
int launch_fwd(const FwdParams& p, cudaStream_t st) {
    constexpr int LD = Smem<DH>::LD;
    const int LQP = (p.Lq + 15) & ~15, LKP = (p.Lk + KB - 1) & ~(KB - 1);
    const size_t smem = (size_t)(LQP + 2 * LKP) * LD * 2 + (size_t)LKP * 4;
    if (smem > 200 * 1024) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_fwd: Lq=%d Lk=%d do not fit in shared memory", p.Lq, p.Lk);
    static size_t configured = 0;
    if (smem > configured) {
        CRCT_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    CRCT_CUDA(crct_launch_pdl(attn_fwd_kernel<DH>, dim3(p.B * p.nh), dim3(ATT_THREADS), smem, st, p));
    return CRCT_OK;
}

template <int DH>
int launch_bwd(const BwdParams& p, cudaStream_t st) {
    constexpr int LD = Smem<DH>::LD;
    const int LQP = (p.Lq + KB - 1) & ~(KB - 1), LKP = (p.Lk + KB - 1) & ~(KB - 1);
    const size_t smem = (size_t)(2 * LQP + 2 * LKP) * LD * 2 + (size_t)(LKP + 2 * LQP) * 4;
    if (smem > 200 * 1024) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_bwd: Lq=%d Lk=%d do not fit in shared memory", p.Lq, p.Lk);
    // fused single kernel when dS^T fits next to the operand tiles with two CTAs per SM
    const size_t smem_fused = smem + (size_t)LKP * (LQP + 8) * 2;
    static const bool split_only = getenv("CRCT_ATTN_BWD_SPLIT") != nullptr;       // A/B switch
    if (smem_fused <= 110 * 1024 && !split_only) {
        static size_t configured_fused = 0;
        if (smem_fused > configured_fused) {
            CRCT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused));
            configured_fused = smem_fused;
        }
        CRCT_CUDA(crct_launch_pdl(attn_bwd_kernel<DH, 2>, dim3(p.B * p.nh), dim3(ATT_THREADS), smem_fused, st, p));      // dK, dV, dQ
        return CRCT_OK;
    }
    static size_t configured = 0;
    if (smem > configured) {
        CRCT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CRCT_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    CRCT_CUDA(crct_launch_pdl(attn_bwd_kernel<DH, 0>, dim3(p.B * p.nh), dim3(ATT_THREADS), smem, st, p));      // dK, dV
    CRCT_CUDA(crct_launch_pdl(attn_bwd_kernel<DH, 1>, dim3(p.B * p.nh), dim3(ATT_THREADS), smem, st, p));      // dQ
    return CRCT_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" CRCT_API int crct_attn_fwd(const crct_attn_fwd_t* a, crct_stream_t s) {
    if (!a || !a->q || !a->k || !a->v || (!a->mask_add && !a->cu_k) || !a->out) CRCT_FAIL(CRCT_ERR_ARG, "crct_attn_fwd: null pointer");
    if (a->B <= 0 || a->nh <= 0 || a->Lq <= 0 || a->Lk <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_fwd: empty problem");
    if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8) || !aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->out))
        CRCT_FAIL(CRCT_ERR_ARG, "crct_attn_fwd: operands must be 16-byte aligned with leading dimensions multiple of 8");
    if ((double)a->B * a->nh * a->Lq * a->Lk >= 4294967296.0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_fwd: B*nh*Lq*Lk must stay below 2^32");
    if (crct_attn_tc_eligible(a->dh, a->Lq, a->Lk, a->q, a->k, a->v, a->ldq, a->ldk, a->ldv, 0)) return crct_attn_fwd_tc(a, s);   // tcgen05 / TMEM / TMA
    FwdParams p;
    p.q = reinterpret_cast<const bf16*>(a->q); p.k = reinterpret_cast<const bf16*>(a->k); p.v = reinterpret_cast<const bf16*>(a->v);
    p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.mask_add = a->mask_add;
    p.out = reinterpret_cast<bf16*>(a->out); p.ldo = a->ldo; p.lse = a->lse;
    p.B = a->B; p.nh = a->nh; p.Lq = a->Lq; p.Lk = a->Lk;
    p.scale = 1.0f / sqrtf((float)a->dh);
    p.thr = crct_drop_threshold(a->dropout_p); p.dscale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f; p.seed = a->seed;
    p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    p.cu_q = a->cu_q; p.cu_k = a->cu_k;
    switch (a->dh) {
        case 32: return launch_fwd<32>(p, as_stream(s));
        case 48: return launch_fwd<48>(p, as_stream(s));
        case 64: return launch_fwd<64>(p, as_stream(s));
    }
    CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_fwd: head dim %d not in {32,48,64}", a->dh);
}

extern "C" CRCT_API int crct_attn_bwd(const crct_attn_bwd_t* a, crct_stream_t s) {
    if (!a || !a->q || !a->k || !a->v || (!a->mask_add && !a->cu_k) || !a->out || !a->dout || !a->lse || !a->dq || !a->dk || !a->dv)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_attn_bwd: null pointer");
    if (a->B <= 0 || a->nh <= 0 || a->Lq <= 0 || a->Lk <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_bwd: empty problem");
    if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8) || (a->lddo % 8) || (a->lddq % 2) || (a->lddk % 2) || (a->lddv % 2) ||
        !aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->out) || !aligned16(a->dout))
        CRCT_FAIL(CRCT_ERR_ARG, "crct_attn_bwd: operands must be 16-byte aligned with leading dimensions multiple of 8");
    if ((double)a->B * a->nh * a->Lq * a->Lk >= 4294967296.0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_bwd: B*nh*Lq*Lk must stay below 2^32");
    if (crct_attn_tc_eligible(a->dh, a->Lq, a->Lk, a->q, a->k, a->v, a->ldq, a->ldk, a->ldv, 1) && !(a->lddq % 8) && !(a->lddk % 8) && !(a->lddv % 8))
        return crct_attn_bwd_tc(a, s);                  // tcgen05 / TMEM / TMA
    BwdParams p;
    p.q = reinterpret_cast<const bf16*>(a->q); p.k = reinterpret_cast<const bf16*>(a->k); p.v = reinterpret_cast<const bf16*>(a->v);
    p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.mask_add = a->mask_add;
    p.out = reinterpret_cast<const bf16*>(a->out); p.ldo = a->ldo;
    p.dout = reinterpret_cast<const bf16*>(a->dout); p.lddo = a->lddo; p.lse = a->lse;
    p.dq = reinterpret_cast<bf16*>(a->dq); p.dk = reinterpret_cast<bf16*>(a->dk); p.dv = reinterpret_cast<bf16*>(a->dv);
    p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
    p.B = a->B; p.nh = a->nh; p.Lq = a->Lq; p.Lk = a->Lk;
    p.scale = 1.0f / sqrtf((float)a->dh);
    p.thr = crct_drop_threshold(a->dropout_p); p.dscale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f; p.seed = a->seed;
    p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    p.cu_q = a->cu_q; p.cu_k = a->cu_k;
    switch (a->dh) {
        case 32: return launch_bwd<32>(p, as_stream(s));
        case 48: return launch_bwd<48>(p, as_stream(s));
        case 64: return launch_bwd<64>(p, as_stream(s));
    }
    CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_bwd: head dim %d not in {32,48,64}", a->dh);
}
