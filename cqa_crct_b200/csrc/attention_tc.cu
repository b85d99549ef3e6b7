// K2/K3 on the 5th-generation tensor cores — fused attention forward / backward with tcgen05.mma, TMEM accumulators and
// TMA operand loads, for every CRCT attention whose sequences fit ONE 128-row tile (text self 16 x 48 at T <= 124, visual
// self 16 x 64 at R <= 44, both co-attention directions 32 x 32): one CTA per (sample, head).
//
//   forward    S = Q K^T  -> TMEM        (tcgen05.mma  M = 128 queries, N = key tile, K = dh; Q, K by TMA, 128B swizzle)
//              P = dropout(softmax(S / sqrt(dh) [+ mask]))   one thread per query row: tcgen05.ld, exp2, bf16 P -> smem
//              O = P V    -> TMEM        (A = P from shared memory K-major, B = V as loaded: MN-major)
//   backward   S = Q K^T, dP = dO V^T -> TMEM;  P, dS (bf16) -> smem;
//              dV = P^T dO, dK = dS^T Q (A = the same P / dS tiles read MN-major), dQ = dS K (A = dS read K-major)
//
// reference: CRCT/backbone/vilbert.py:397-412 (text), :527-543 (visual), :684-723 (co-attention, both directions).
// Q / K / V / dO tiles are [rows x 64 columns] boxes of the packed projections (head h = columns [h*dh, h*dh + 64): the
// columns past dh belong to the next head and are never multiplied — the MMAs run K = dh).  Rows past a sample's length
// (packed layout: the next sample's rows; padded layout: masked rows) are excluded by INDEX (probability exactly 0, operand
// rows cleared in shared memory where a 0 x garbage product could otherwise produce NaN).  Dropout counters, log-sum-exp
// and masks are laid out exactly as in the mma.sync kernels of attention.cu, which remain the path for longer sequences
// (stress shape T = 248) and the A/B reference (CRCT_ATTN_LEGACY=1).
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

int crct_make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer);

namespace {

constexpr int TC_THREADS = 256;                 // two warps per 32-lane TMEM quarter: each takes half of the score columns
constexpr int QT = 128;                       // query tile = TMEM lanes
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// UMMA shared-memory descriptor, SWIZZLE_128B (see gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, majors, N >> 3, M >> 4
__device__ __forceinline__ constexpr uint32_t idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
// K-major operand tile [rows][64 bf16] (one 128-byte swizzle atom wide): descriptor of the 16-element k-step `ks`
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int ks) { return smem_desc(tile + (uint32_t)ks * 32u, 16u, 1024u); }
// MN-major operand: tile [k rows][64 (M or N) elements]; 64-wide atoms `atom_bytes` apart; k-step = 16 rows
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int ks, uint32_t atom_bytes) {
    return smem_desc(tile + (uint32_t)ks * 2048u, atom_bytes, 1024u);
}
__device__ __forceinline__ void reg_fence16(uint32_t (&v)[16]) {      // the asynchronously loaded values exist from here on
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
// byte offset of the 16-byte chunk holding columns [8*chunk, 8*chunk + 8) of row r in a [rows][64] 128B-swizzled tile
__device__ __forceinline__ uint32_t swz(int r, int chunk) { return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4); }

// 16 fp32 accumulator values -> 16 bf16 = two 16-byte stores
__device__ __forceinline__ void store16_bf16(bf16* d, const uint32_t (&v)[16]) {
    uint4 a, b;
    a.x = pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1]));   a.y = pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3]));
    a.z = pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5]));   a.w = pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7]));
    b.x = pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9]));   b.y = pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11]));
    b.z = pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13])); b.w = pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15]));
    *reinterpret_cast<uint4*>(d) = a;
    *reinterpret_cast<uint4*>(d + 8) = b;
}

// ---- per-score scalar work on 8 consecutive columns (one 16-byte chunk of the P / dS tiles).  CHECK: the chunk crosses the
// sample's last key (columns >= Lk give probability 0); MASK: an additive key mask is present (padded layout).
template <bool CHECK, bool MASK>
__device__ __forceinline__ void fwd_chunk8(const uint32_t* v, int c, int Lk, float scale2, const float* mask_s, float m, float& l,
                                           const CrctDrop32& drop, uint32_t ctr, float dscale, uint32_t (&pk)[4]) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        // explicit roundings (no fused contraction): the masked and the mask-free form give the same bits where the mask is 0,
        // so the packed layout reproduces the padded one exactly
        float s0 = __fmul_rn(__uint_as_float(v[j]), scale2), s1 = __fmul_rn(__uint_as_float(v[j + 1]), scale2);
        if constexpr (MASK) { s0 = __fadd_rn(s0, mask_s[c + j]); s1 = __fadd_rn(s1, mask_s[c + j + 1]); }
        float p0 = fast_exp2(__fsub_rn(s0, m)), p1 = fast_exp2(__fsub_rn(s1, m));
        if constexpr (CHECK) {
            if (c + j >= Lk) p0 = 0.f;
            if (c + j + 1 >= Lk) p1 = 0.f;
        }
        l += p0 + p1;
        if (drop.thr != 0u) {
            bool k0, k1;
            crct_keep2_32(drop, ctr + (uint32_t)j, k0, k1);
            p0 = k0 ? p0 * dscale : 0.f;
            p1 = k1 ? p1 * dscale : 0.f;
        }
        pk[j >> 1] = pack_bf16x2(p0, p1);
    }
}
template <bool CHECK, bool MASK>
__device__ __forceinline__ void bwd_chunk8(const uint32_t* vs, const uint32_t* vd, int c, int Lk, float scale2, const float* mask_s, float lse2,
                                           float Dr, float scale, const CrctDrop32& drop, uint32_t ctr, float dscale, uint32_t* pr_out,
                                           uint32_t (&ds)[4]) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        float s0 = __fmul_rn(__uint_as_float(vs[j]), scale2), s1 = __fmul_rn(__uint_as_float(vs[j + 1]), scale2);
        if constexpr (MASK) { s0 = __fadd_rn(s0, mask_s[c + j]); s1 = __fadd_rn(s1, mask_s[c + j + 1]); }
        float pr0 = fast_exp2(__fsub_rn(s0, lse2)), pr1 = fast_exp2(__fsub_rn(s1, lse2));
        float g0 = __uint_as_float(vd[j]), g1 = __uint_as_float(vd[j + 1]);
        if constexpr (CHECK) {                         // past the last key: probability 0, and dP there may be garbage (select, not multiply)
            if (c + j >= Lk) { pr0 = 0.f; g0 = 0.f; }
            if (c + j + 1 >= Lk) { pr1 = 0.f; g1 = 0.f; }
        }
        float f0 = 1.f, f1 = 1.f;
        if (drop.thr != 0u) {
            bool k0, k1;
            crct_keep2_32(drop, ctr + (uint32_t)j, k0, k1);
            f0 = k0 ? dscale : 0.f;
            f1 = k1 ? dscale : 0.f;
        }
        // dS = P (dP * keep - D) * scale
        ds[j >> 1] = pack_bf16x2(pr0 * (g0 * f0 - Dr) * scale, pr1 * (g1 * f1 - Dr) * scale);
        pr_out[j >> 1] = pack_bf16x2(pr0 * f0, pr1 * f1);
    }
}

struct TcParams {
    const float* mask_add;
    const bf16* out;  int ldo;                // forward: written; backward: read (D = sum dO * O)
    bf16* out_w;
    const bf16* dout; int lddo;
    float* lse;
    bf16* dq; bf16* dk; bf16* dv;
    int lddq, lddk, lddv;
    int B, nh, Lq, Lk;
    float scale;
    uint32_t thr; float dscale; uint64_t seed; const unsigned long long* salt;
    const int* cu_q; const int* cu_k;
};

// ------------------------------------------------------------------------------------------------ forward
// 256 threads: thread t works on TMEM lane (query row) t & 127 and on column half t >> 7 of the score tile — two warps per
// 32-lane group (a warp may only touch the TMEM lanes of its own quarter, warp & 3), twice the warps per CTA to hide the
// tcgen05.ld / MUFU / shared-memory latencies of the per-score scalar work.
// shared memory (1024-byte aligned): Q [128][128 B] | K [LKT][128 B] | V [LKT][128 B] | barriers | mask [LKT] | red [2][2][128] f32
// P [128][LKT] bf16 (K-major, 64-key atoms of 16 KB) aliases Q (+ K): both are dead once S is in TMEM.
template <int DH, int LKT>
__global__ void __launch_bounds__(TC_THREADS) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - raw);
    constexpr uint32_t Q_BYTES = QT * 128, KV_BYTES = LKT * 128, OFF_BAR = Q_BYTES + 2 * KV_BYTES;
    constexpr int CH = LKT / 2;                                       // score columns per thread half
    const uint32_t sQ = base, sK = base + Q_BYTES, sV = sK + KV_BYTES;
    const uint32_t bar_tma = base + OFF_BAR, bar_mma = bar_tma + 8, tmem_slot = bar_tma + 16;
    float* mask_s = reinterpret_cast<float*>(g + OFF_BAR + 32);
    float* red_m = mask_s + LKT;                                      // [2][128] partial row maxima
    float* red_l = red_m + 2 * QT;                                    // [2][128] partial row sums
    uint8_t* gP = g;                                                  // P tile (generic pointer), aliases Q / K
    uint8_t* gV = g + Q_BYTES + KV_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (QT - 1), hsel = tid >> 7;
    pdl_launch_dependents();
    if (tid == 32) {                                   // warp 1: descriptors + barriers; warp 0 (converged) allocates tensor memory
        ptx::tma_prefetch_desc(&tmQ); ptx::tma_prefetch_desc(&tmK); ptx::tma_prefetch_desc(&tmV);
        ptx::mbar_init(bar_tma, 1);
        ptx::mbar_init(bar_mma, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, LKT);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(g + OFF_BAR + 16);
    pdl_wait();

    const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;
    int qrow0 = b * p.Lq, krow0 = b * p.Lk, Lq = p.Lq, Lk = p.Lk;
    if (p.cu_q != nullptr) { qrow0 = __ldg(p.cu_q + b); Lq = min(p.Lq, __ldg(p.cu_q + b + 1) - qrow0); }
    if (p.cu_k != nullptr) { krow0 = __ldg(p.cu_k + b); Lk = min(p.Lk, __ldg(p.cu_k + b + 1) - krow0); }
    const int Lk16 = (Lk + 15) & ~15;

    if (tid == 0) {
        ptx::mbar_arrive_expect_tx(bar_tma, Q_BYTES + 2 * KV_BYTES);
        ptx::tma_load_2d(sQ, &tmQ, bar_tma, h * DH, qrow0);
        ptx::tma_load_2d(sK, &tmK, bar_tma, h * DH, krow0);
        ptx::tma_load_2d(sV, &tmV, bar_tma, h * DH, krow0);
    }
    if (tid < LKT) mask_s[tid] = (p.mask_add != nullptr && tid < Lk) ? p.mask_add[(size_t)b * p.Lk + tid] * LOG2E : 0.f;
    ptx::mbar_wait(bar_tma, 0);
    if (tid == 0) {                                                   // S = Q K^T : DH / 16 instructions, 128 x LKT x 16 each
        ptx::tc_fence_after();
        constexpr uint32_t id = idesc(QT, LKT, false, false);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) ptx::tc_mma_bf16(tmem, desc_kmajor(sQ, ks), desc_kmajor(sK, ks), id, ks > 0 ? 1u : 0u);
        ptx::tc_commit(bar_mma);
    }
    // V rows in [Lk, Lk16) meet P columns that are written as 0: clear them, so that no stale NaN pattern is multiplied
    if (tid < QT && tid >= Lk && tid < Lk16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(gV + (size_t)tid * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();                                                  // mask_s visible
    ptx::mbar_wait(bar_mma, 0);
    __syncwarp();
    ptx::tc_fence_after();

    // ---- softmax of row `row` (TMEM lane), scores in the log2 domain: s2 = s * scale * log2(e) + mask * log2(e)
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float scale2 = p.scale * LOG2E;
    const bool live = row < Lq;
    const int cb = hsel * CH;                                         // this thread's columns: [cb, cb + CH)
    float m = -INFINITY;
    for (int c0 = cb; c0 < cb + CH && c0 < Lk; c0 += 32) {            // warp-uniform bounds
        uint32_t v[32];
        ptx::tc_ld_32x32(trow + (uint32_t)c0, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j < Lk) m = fmaxf(m, __fadd_rn(__fmul_rn(__uint_as_float(v[j]), scale2), mask_s[c0 + j]));
    }
    red_m[hsel * QT + row] = m;
    __syncthreads();
    m = fmaxf(red_m[row], red_m[QT + row]);
    const CrctDrop32 drop = crct_drop32((p.thr != 0u && p.salt) ? (p.seed ^ __ldg(p.salt)) : p.seed, p.thr);
    const uint32_t ctr0 = ((uint32_t)blockIdx.x * (uint32_t)p.Lq + (uint32_t)row) * (uint32_t)p.Lk;
    float l = 0.f;
    const bool masked = p.mask_add != nullptr;
    for (int c0 = cb; c0 < cb + CH && c0 < Lk16; c0 += 32) {
        uint32_t v[32];
        ptx::tc_ld_32x32(trow + (uint32_t)c0, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {                              // 8 columns = one 16-byte chunk of the P tile
            const int c = c0 + ch * 8;
            if (c >= Lk16) break;
            uint32_t pk[4] = {0u, 0u, 0u, 0u};
            if (live) {                                               // rows past the sample's last query write zeros
                const uint32_t ctr = ctr0 + (uint32_t)c;
                if (c + 8 <= Lk) {
                    if (masked) fwd_chunk8<false, true>(v + ch * 8, c, Lk, scale2, mask_s, m, l, drop, ctr, p.dscale, pk);
                    else fwd_chunk8<false, false>(v + ch * 8, c, Lk, scale2, mask_s, m, l, drop, ctr, p.dscale, pk);
                } else {
                    if (masked) fwd_chunk8<true, true>(v + ch * 8, c, Lk, scale2, mask_s, m, l, drop, ctr, p.dscale, pk);
                    else fwd_chunk8<true, false>(v + ch * 8, c, Lk, scale2, mask_s, m, l, drop, ctr, p.dscale, pk);
                }
            }
            *reinterpret_cast<uint4*>(gP + (size_t)(c >> 6) * (QT * 128) + swz(row, (c & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    red_l[hsel * QT + row] = l;
    ptx::fence_proxy_async();                                         // P / cleared V rows (generic stores) -> tensor-core reads
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {                                                   // O = P V : Lk16 / 16 instructions, 128 x DH x 16 each
        ptx::tc_fence_after();
        constexpr uint32_t id = idesc(QT, DH, false, true);
        for (int ks = 0; ks < Lk16 / 16; ++ks)
            ptx::tc_mma_bf16(tmem, desc_kmajor(sQ + (uint32_t)(ks >> 2) * (QT * 128), ks & 3), desc_mnmajor(sV, ks, LKT * 128), id, ks > 0 ? 1u : 0u);
        ptx::tc_commit(bar_mma);
    }
    l = red_l[row] + red_l[QT + row];
    if (live && hsel == 0 && p.lse != nullptr) p.lse[(size_t)blockIdx.x * p.Lq + row] = (m + __log2f(l)) * LN2;
    ptx::mbar_wait(bar_mma, 1);
    __syncwarp();
    ptx::tc_fence_after();
    {
        // tcgen05.ld is warp-collective (.sync.aligned): every lane executes the same loads, only the stores are predicated;
        // the two warps of a lane group share the DH / 16 column chunks
        const float inv = live ? 1.f / l : 0.f;
        bf16* dst = p.out_w + (size_t)(qrow0 + row) * p.ldo + h * DH;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
            if (((c0 >> 4) & 1) != hsel) continue;                    // warp-uniform
            uint32_t v[16];
            ptx::tc_ld_32x16(trow + (uint32_t)c0, v);
            ptx::tc_wait_ld();
            if (live) {
                uint4 o0, o1;
                o0.x = pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);   o0.y = pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
                o0.z = pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);   o0.w = pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
                o1.x = pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);   o1.y = pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
                o1.z = pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv); o1.w = pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
                *reinterpret_cast<uint4*>(dst + c0) = o0;
                *reinterpret_cast<uint4*>(dst + c0 + 8) = o1;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); ptx::tmem_dealloc(tmem, LKT); }
}

template <int DH, int LKT>
int launch_fwd_tc(const crct_attn_fwd_t* a, const TcParams& p, cudaStream_t st) {
    CUtensorMap tmQ, tmK, tmV;
    const uint64_t W = (uint64_t)a->nh * a->dh;
    if (int rc = crct_make_tmap_bf16_2d(&tmQ, a->q, W, (uint64_t)a->B * a->Lq, a->ldq, 64, QT)) return rc;
    if (int rc = crct_make_tmap_bf16_2d(&tmK, a->k, W, (uint64_t)a->B * a->Lk, a->ldk, 64, LKT)) return rc;
    if (int rc = crct_make_tmap_bf16_2d(&tmV, a->v, W, (uint64_t)a->B * a->Lk, a->ldv, 64, LKT)) return rc;
    constexpr int SMEM = QT * 128 + 2 * LKT * 128 + 32 + LKT * 4 + 4 * QT * 4 + 1024;
    static bool configured = false;
    if (!configured) {
        CRCT_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<DH, LKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    CRCT_CUDA(crct_launch_pdl(attn_fwd_tc_kernel<DH, LKT>, dim3(a->B * a->nh), dim3(TC_THREADS), SMEM, st, tmQ, tmK, tmV, p));
    return CRCT_OK;
}

// ------------------------------------------------------------------------------------------------ backward
// shared memory: Q [128][128 B] | dO [128][128 B] | K [LKT][128 B] | V [LKT][128 B] | X [128 q][128 keys] bf16 = two 64-key atoms of
// 16 KB | barriers | mask [128] | D partials [2][128].  X holds dS first (dK = dS^T Q and dQ = dS K read it MN-major / K-major), then
// — once those MMAs have retired — P (dV = P^T dO), which the threads keep packed in registers meanwhile: one pass over the
// scores, one 32 KB tile, 96 KB per CTA = two CTAs per SM.  TMEM (256 columns): S [0,128) and dP [128,256) first, then
// dK [0,64), dQ [64,128), dV [128,192).  Thread t: TMEM lane (query / key row) t & 127, column half t >> 7 (see the forward).
template <int DH, int LKT, int QBOX>
__global__ void __launch_bounds__(TC_THREADS) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                                                                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                                                                 const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* g = smem_raw + (base - raw);
    constexpr uint32_t Q_BYTES = QT * 128, KV_BYTES = LKT * 128, X_BYTES = 2 * QT * 128;
    constexpr uint32_t OFF_DO = Q_BYTES, OFF_K = 2 * Q_BYTES, OFF_V = OFF_K + KV_BYTES, OFF_X = OFF_V + KV_BYTES, OFF_BAR = OFF_X + X_BYTES;
    constexpr int CH = LKT / 2;
    const uint32_t sQ = base, sdO = base + OFF_DO, sK = base + OFF_K, sV = base + OFF_V, sX = base + OFF_X;
    const uint32_t bar_tma = base + OFF_BAR, bar_mma = bar_tma + 8, tmem_slot = bar_tma + 16;
    float* mask_s = reinterpret_cast<float*>(g + OFF_BAR + 32);       // [128]
    float* red_d = mask_s + QT;                                       // [2][128] partial D
    uint8_t* gX = g + OFF_X;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (QT - 1), hsel = tid >> 7;
    pdl_launch_dependents();
    if (tid == 32) {
        ptx::tma_prefetch_desc(&tmQ); ptx::tma_prefetch_desc(&tmdO); ptx::tma_prefetch_desc(&tmK); ptx::tma_prefetch_desc(&tmV);
        ptx::mbar_init(bar_tma, 1);
        ptx::mbar_init(bar_mma, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(g + OFF_BAR + 16);
    pdl_wait();

    const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;
    int qrow0 = b * p.Lq, krow0 = b * p.Lk, Lq = p.Lq, Lk = p.Lk;
    if (p.cu_q != nullptr) { qrow0 = __ldg(p.cu_q + b); Lq = min(p.Lq, __ldg(p.cu_q + b + 1) - qrow0); }
    if (p.cu_k != nullptr) { krow0 = __ldg(p.cu_k + b); Lk = min(p.Lk, __ldg(p.cu_k + b + 1) - krow0); }
    const int Lk16 = (Lk + 15) & ~15, Lq16 = (Lq + 15) & ~15;

    if (tid == 0) {
        ptx::mbar_arrive_expect_tx(bar_tma, 2 * QBOX * 128 + 2 * KV_BYTES);
        ptx::tma_load_2d(sQ, &tmQ, bar_tma, h * DH, qrow0);
        ptx::tma_load_2d(sdO, &tmdO, bar_tma, h * DH, qrow0);
        ptx::tma_load_2d(sK, &tmK, bar_tma, h * DH, krow0);
        ptx::tma_load_2d(sV, &tmV, bar_tma, h * DH, krow0);
    }
    // this thread's query row: log-sum-exp (log2 domain) and its half of D = sum_d dO * O (global loads in flight next to the TMA)
    const bool live = row < Lq;
    float lse2 = 0.f;
    {
        float dpart = 0.f;
        if (live) {
            lse2 = p.lse[(size_t)blockIdx.x * p.Lq + row] * LOG2E;
            const bf16* o = p.out + (size_t)(qrow0 + row) * p.ldo + h * DH;
            const bf16* dd = p.dout + (size_t)(qrow0 + row) * p.lddo + h * DH;
#pragma unroll
            for (int c = 0; c < DH / 8; ++c) {
                if ((c & 1) != hsel) continue;
                float fo[8], fd[8];
                load8_bf16(o + c * 8, fo);
                load8_bf16(dd + c * 8, fd);
#pragma unroll
                for (int j = 0; j < 8; ++j) dpart = fmaf(fo[j], fd[j], dpart);
            }
        }
        red_d[hsel * QT + row] = dpart;
    }
    if (tid < QT) mask_s[tid] = (p.mask_add != nullptr && tid < Lk) ? p.mask_add[(size_t)b * p.Lk + tid] * LOG2E : 0.f;
    // the key half [64,128) of X is read by the M = 128 MMAs even when every key lives in the first atom: keep it finite
    if (Lk16 <= 64 && tid < QT) {
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(gX + QT * 128 + (size_t)tid * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::mbar_wait(bar_tma, 0);
    if (tid == 0) {                                   // S = Q K^T -> [0,LKT) ; dP = dO V^T -> [128,128+LKT)
        ptx::tc_fence_after();
        constexpr uint32_t id = idesc(QT, LKT, false, false);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) ptx::tc_mma_bf16(tmem, desc_kmajor(sQ, ks), desc_kmajor(sK, ks), id, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) ptx::tc_mma_bf16(tmem + 128u, desc_kmajor(sdO, ks), desc_kmajor(sV, ks), id, ks > 0 ? 1u : 0u);
        ptx::tc_commit(bar_mma);
    }
    __syncthreads();                                  // mask_s, red_d visible
    const float Dr = red_d[row] + red_d[QT + row];
    ptx::mbar_wait(bar_mma, 0);
    __syncwarp();
    ptx::tc_fence_after();
    // operand rows that meet zero rows / columns of P and dS must be finite: clear Q, dO rows [Lq, Lq16) and K rows [Lk, Lk16)
    // (S and dP, which read them, are complete)
    if (tid < QT && tid >= Lq && tid < Lq16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            *reinterpret_cast<uint4*>(g + (size_t)tid * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(g + OFF_DO + (size_t)tid * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (tid < QT && tid >= Lk && tid < Lk16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(g + OFF_K + (size_t)tid * 128 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }

    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float scale2 = p.scale * LOG2E;
    const CrctDrop32 drop = crct_drop32((p.thr != 0u && p.salt) ? (p.seed ^ __ldg(p.salt)) : p.seed, p.thr);
    const uint32_t ctr0 = ((uint32_t)blockIdx.x * (uint32_t)p.Lq + (uint32_t)row) * (uint32_t)p.Lk;
    const int cb = hsel * CH;                         // this thread's score columns: [cb, cb + CH)
    const bool masked = p.mask_add != nullptr;
    uint32_t preg[CH / 2];                            // this row's (dropped) probabilities, packed bf16, until X is free again
    uint32_t vsb[2][16], vdb[2][16];                  // double-buffered TMEM loads: chunk i + 1 is in flight during the math of chunk i
    if (cb < Lk16) {
        ptx::tc_ld_32x16(trow + (uint32_t)cb, vsb[0]);
        ptx::tc_ld_32x16(trow + 128u + (uint32_t)cb, vdb[0]);
    }
#pragma unroll
    for (int i = 0; i < CH; i += 16) {
        const int c0 = cb + i;
        if (c0 < Lk16) {                              // warp-uniform
            uint32_t (&vs)[16] = vsb[(i >> 4) & 1];
            uint32_t (&vd)[16] = vdb[(i >> 4) & 1];
            ptx::tc_wait_ld();
            reg_fence16(vs);
            reg_fence16(vd);
            if (i + 16 < CH && c0 + 16 < Lk16) {
                ptx::tc_ld_32x16(trow + (uint32_t)(c0 + 16), vsb[((i >> 4) + 1) & 1]);
                ptx::tc_ld_32x16(trow + 128u + (uint32_t)(c0 + 16), vdb[((i >> 4) + 1) & 1]);
            }
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int c = c0 + ch * 8;
                uint32_t ds[4] = {0u, 0u, 0u, 0u};
                uint32_t* pr = &preg[(i + ch * 8) >> 1];
                pr[0] = pr[1] = pr[2] = pr[3] = 0u;
                if (live) {                           // rows past the sample's last query stay zero
                    const uint32_t ctr = ctr0 + (uint32_t)c;
                    if (c + 8 <= Lk) {
                        if (masked) bwd_chunk8<false, true>(vs + ch * 8, vd + ch * 8, c, Lk, scale2, mask_s, lse2, Dr, p.scale, drop, ctr, p.dscale, pr, ds);
                        else bwd_chunk8<false, false>(vs + ch * 8, vd + ch * 8, c, Lk, scale2, mask_s, lse2, Dr, p.scale, drop, ctr, p.dscale, pr, ds);
                    } else {
                        if (masked) bwd_chunk8<true, true>(vs + ch * 8, vd + ch * 8, c, Lk, scale2, mask_s, lse2, Dr, p.scale, drop, ctr, p.dscale, pr, ds);
                        else bwd_chunk8<true, false>(vs + ch * 8, vd + ch * 8, c, Lk, scale2, mask_s, lse2, Dr, p.scale, drop, ctr, p.dscale, pr, ds);
                    }
                }
                *reinterpret_cast<uint4*>(gX + (size_t)(c >> 6) * (QT * 128) + swz(row, (c & 63) >> 3)) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
            }
        }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        ptx::tc_fence_after();
        // dK [keys x DH] = dS^T Q : A = X read MN-major (M = keys, two 64-key atoms), B = Q as loaded (MN-major), K = queries
        constexpr uint32_t id_t = idesc(QT, DH, true, true);
        for (int ks = 0; ks < Lq16 / 16; ++ks)
            ptx::tc_mma_bf16(tmem, desc_mnmajor(sX, ks, QT * 128), desc_mnmajor(sQ, ks, QT * 128), id_t, ks > 0 ? 1u : 0u);
        // dQ [queries x DH] = dS K : A = X read K-major, B = K as loaded (MN-major), K = keys
        constexpr uint32_t id_q = idesc(QT, DH, false, true);
        for (int ks = 0; ks < Lk16 / 16; ++ks)
            ptx::tc_mma_bf16(tmem + 64u, desc_kmajor(sX + (uint32_t)(ks >> 2) * (QT * 128), ks & 3), desc_mnmajor(sK, ks, LKT * 128), id_q, ks > 0 ? 1u : 0u);
        ptx::tc_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, 1);                       // dK, dQ retired: X may be overwritten with P
    __syncwarp();
    ptx::tc_fence_after();
#pragma unroll
    for (int i = 0; i < CH; i += 8) {
        const int c = cb + i;
        if (c < Lk16)
            *reinterpret_cast<uint4*>(gX + (size_t)(c >> 6) * (QT * 128) + swz(row, (c & 63) >> 3)) =
                make_uint4(preg[i / 2], preg[i / 2 + 1], preg[i / 2 + 2], preg[i / 2 + 3]);
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {                                   // dV [keys x DH] = P^T dO
        ptx::tc_fence_after();
        constexpr uint32_t id_t = idesc(QT, DH, true, true);
        for (int ks = 0; ks < Lq16 / 16; ++ks)
            ptx::tc_mma_bf16(tmem + 128u, desc_mnmajor(sX, ks, QT * 128), desc_mnmajor(sdO, ks, QT * 128), id_t, ks > 0 ? 1u : 0u);
        ptx::tc_commit(bar_mma);
    }
    // meanwhile: dK (lanes = keys) and dQ (lanes = queries) leave TMEM; the two warps of a lane group share the column chunks
    const int col = h * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 16) {
        if (((c0 >> 4) & 1) != hsel) continue;        // warp-uniform
        uint32_t vk[16], vq[16];
        ptx::tc_ld_32x16(trow + (uint32_t)c0, vk);
        ptx::tc_ld_32x16(trow + 64u + (uint32_t)c0, vq);
        ptx::tc_wait_ld();
        if (row < Lk) store16_bf16(p.dk + (size_t)(krow0 + row) * p.lddk + col + c0, vk);
        if (live) store16_bf16(p.dq + (size_t)(qrow0 + row) * p.lddq + col + c0, vq);
    }
    ptx::mbar_wait(bar_mma, 0);                       // third completion of the barrier: parity 0 again
    __syncwarp();
    ptx::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 16) {
        if (((c0 >> 4) & 1) != hsel) continue;
        uint32_t vv[16];
        ptx::tc_ld_32x16(trow + 128u + (uint32_t)c0, vv);
        ptx::tc_wait_ld();
        if (row < Lk) store16_bf16(p.dv + (size_t)(krow0 + row) * p.lddv + col + c0, vv);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); ptx::tmem_dealloc(tmem, 256); }
}

template <int DH, int LKT, int QBOX>
int launch_bwd_tc(const crct_attn_bwd_t* a, const TcParams& p, cudaStream_t st) {
    CUtensorMap tmQ, tmdO, tmK, tmV;
    const uint64_t W = (uint64_t)a->nh * a->dh;
    if (int rc = crct_make_tmap_bf16_2d(&tmQ, a->q, W, (uint64_t)a->B * a->Lq, a->ldq, 64, QBOX)) return rc;
    if (int rc = crct_make_tmap_bf16_2d(&tmdO, a->dout, W, (uint64_t)a->B * a->Lq, a->lddo, 64, QBOX)) return rc;
    if (int rc = crct_make_tmap_bf16_2d(&tmK, a->k, W, (uint64_t)a->B * a->Lk, a->ldk, 64, LKT)) return rc;
    if (int rc = crct_make_tmap_bf16_2d(&tmV, a->v, W, (uint64_t)a->B * a->Lk, a->ldv, 64, LKT)) return rc;
    constexpr int SMEM = 2 * QT * 128 + 2 * LKT * 128 + 2 * QT * 128 + 32 + 3 * QT * 4 + 1024;
    static bool configured = false;
    if (!configured) {
        CRCT_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<DH, LKT, QBOX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    CRCT_CUDA(crct_launch_pdl(attn_bwd_tc_kernel<DH, LKT, QBOX>, dim3(a->B * a->nh), dim3(TC_THREADS), SMEM, st, tmQ, tmdO, tmK, tmV, p));
    return CRCT_OK;
}

template <int DH>
int dispatch_bwd_tc(const crct_attn_bwd_t* a, const TcParams& p, cudaStream_t st) {
    const bool wk = a->Lk > 64, wq = a->Lq > 64;
    if (wk) return wq ? launch_bwd_tc<DH, 128, 128>(a, p, st) : launch_bwd_tc<DH, 128, 64>(a, p, st);
    return wq ? launch_bwd_tc<DH, 64, 128>(a, p, st) : launch_bwd_tc<DH, 64, 64>(a, p, st);
}

}  // namespace

// Which calls take the tcgen05 path.  Eligibility: single-tile sequences, TMA-compatible operands.  Policy (measured on B200,
// B = 80 packed rows, dropout on, cold L2 — tools/attn_ab.py, profiles/r02_attention_ab.txt): the per-(sample, head) problems are
// tiny (3 + 8 MMA instructions), so a CTA's life is dominated by fixed latencies (TMA round trip, TMEM allocation, three
// MMA -> mbarrier -> tcgen05.ld hand-offs) and by the scalar softmax work; the tensor-core kernels win where 128 query rows
// keep all TMEM lanes busy (text self-attention forward 35 vs 44 us, text -> visual co-attention forward 35 vs 36 us) and lose
// where they do not (<= 44 visual queries: 23 vs 17 us, 48 vs 30 us) and in the backward (83 vs 75 us; 77 vs 54 us), whose
// mma.sync form keeps every contraction in registers.  IN THE STEP (three streams, captured graph) the choices 0 / 1 / 5 are
// indistinguishable (13.92 / 13.90 / 14.00 ms, +-0.05 run to run) and 15 costs 0.8 ms, so the default is 5: tcgen05 for the
// text self-attention in both directions and for the text -> visual co-attention forward.  CRCT_ATTN_TC_POLICY overrides
// (bit 0: forward with > 64 queries, bit 1: other forwards, bit 2: self-attention backward with > 64 queries, bit 3: other
// backwards; 15 = tcgen05 everywhere it is eligible), CRCT_ATTN_LEGACY=1 / CRCT_ATTN_LEGACY_NOW=1 (read per call) = mma.sync only.
bool crct_attn_tc_eligible(int dh, int Lq, int Lk, const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int backward) {
    static const bool legacy = getenv("CRCT_ATTN_LEGACY") != nullptr;
    if (legacy || getenv("CRCT_ATTN_LEGACY_NOW") != nullptr) return false;
    if (dh != 32 && dh != 48 && dh != 64) return false;
    if (Lq > 128 || Lk > 128) return false;
    if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) != 0 || (ldq % 8) || (ldk % 8) ||
        (ldv % 8))
        return false;
    const char* e = getenv("CRCT_ATTN_TC_POLICY");
    const int policy = e ? atoi(e) : 5;
    const int bit = backward ? ((Lq > 64 && Lk > 64) ? 4 : 8) : (Lq > 64 ? 1 : 2);
    return (policy & bit) != 0;
}

int crct_attn_fwd_tc(const crct_attn_fwd_t* a, crct_stream_t s) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.mask_add = a->mask_add;
    p.out_w = reinterpret_cast<bf16*>(a->out); p.ldo = a->ldo; p.lse = a->lse;
    p.B = a->B; p.nh = a->nh; p.Lq = a->Lq; p.Lk = a->Lk;
    p.scale = 1.0f / sqrtf((float)a->dh);
    p.thr = crct_drop_threshold(a->dropout_p); p.dscale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f; p.seed = a->seed;
    p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    p.cu_q = a->cu_q; p.cu_k = a->cu_k;
    cudaStream_t st = as_stream(s);
    const bool wide = a->Lk > 64;
    switch (a->dh) {
        case 32: return wide ? launch_fwd_tc<32, 128>(a, p, st) : launch_fwd_tc<32, 64>(a, p, st);
        case 48: return wide ? launch_fwd_tc<48, 128>(a, p, st) : launch_fwd_tc<48, 64>(a, p, st);
        case 64: return wide ? launch_fwd_tc<64, 128>(a, p, st) : launch_fwd_tc<64, 64>(a, p, st);
    }
    CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_fwd: head dim %d not in {32,48,64}", a->dh);
}

int crct_attn_bwd_tc(const crct_attn_bwd_t* a, crct_stream_t s) {
    if ((reinterpret_cast<uintptr_t>(a->dout) & 15) || (a->lddo % 8)) CRCT_FAIL(CRCT_ERR_ARG, "crct_attn_bwd: dout must be 16-byte aligned, lddo a multiple of 8");
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.mask_add = a->mask_add;
    p.out = reinterpret_cast<const bf16*>(a->out); p.ldo = a->ldo;
    p.dout = reinterpret_cast<const bf16*>(a->dout); p.lddo = a->lddo;
    p.lse = const_cast<float*>(a->lse);
    p.dq = reinterpret_cast<bf16*>(a->dq); p.dk = reinterpret_cast<bf16*>(a->dk); p.dv = reinterpret_cast<bf16*>(a->dv);
    p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
    p.B = a->B; p.nh = a->nh; p.Lq = a->Lq; p.Lk = a->Lk;
    p.scale = 1.0f / sqrtf((float)a->dh);
    p.thr = crct_drop_threshold(a->dropout_p); p.dscale = a->dropout_p > 0.f ? 1.f / (1.f - a->dropout_p) : 1.f; p.seed = a->seed;
    p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    p.cu_q = a->cu_q; p.cu_k = a->cu_k;
    cudaStream_t st = as_stream(s);
    switch (a->dh) {
        case 32: return dispatch_bwd_tc<32>(a, p, st);
        case 48: return dispatch_bwd_tc<48>(a, p, st);
        case 64: return dispatch_bwd_tc<64>(a, p, st);
    }
    CRCT_FAIL(CRCT_ERR_SHAPE, "crct_attn_bwd: head dim %d not in {32,48,64}", a->dh);
}
