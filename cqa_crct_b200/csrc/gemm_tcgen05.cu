// K1 — persistent, warp-specialised tcgen05/TMEM GEMM fed by TMA (sm_100a).
//
//   D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ),  bf16 operands, fp32 accumulation in tensor memory.
//
// Replaces the cuBLAS calls behind every nn.Linear of the reference transformer and their autograd
// backward (reference: CRCT/backbone/vilbert.py:373-375,388-390,420,446,463,502-504,551,577,594,
// 637-646,732,739,1453).  One CTA per SM, 12 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx-count)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers)
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-11  epilogue       (tcgen05.ld -> registers -> fused bias/GELU/dropout/residual -> global)
// Tile 128 x BN x 64 (BN = 128 or 256), 4-6 smem stages, double-buffered accumulators so the epilogue
// of tile i overlaps the MMAs of tile i+1.  K-major and MN-major operands are both native (UMMA
// descriptor major bits + transposed TMA boxes), so dgrad / wgrad need no transposed copies.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 16;       // 4 per SMSP: the fused epilogues are issue/latency-bound with fewer
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
constexpr int EPI_COLS = 16;            // accumulator columns per epilogue step

template <int BN>
struct Cfg {
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BN * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = (BN == 128) ? 256 : 512;      // two accumulator stages of BN columns; allocations are powers of two
    static constexpr int BAR_BYTES = 8 * (2 * STAGES + 4) + 16;
    static constexpr int BIAS_BYTES = 2 * 256 * 4;              // per accumulator stage: the tile's bias slice
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
};

struct KParams {
    int M, N, K;
    int num_n_tiles, num_tiles, split_k, kb_total, kb_per_split;
    void* D;
    void* D2;
    const float* bias;
    const bf16* aux;
    int ldd, ldaux;
    int wide;          // 1: D/D2/aux rows can be moved 32 bytes at a time (alignment, N % 16 == 0)
    int accumulate;
    uint32_t drop_thr;
    float drop_scale;
    uint64_t seed;
    const unsigned long long* salt;
    uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;   // bytes
    const int* a_rows_dev;   // device-side count of A's stored rows (var-len packing): M, or GEMM-K for the wgrad form
    int tile_m;              // rows of one output tile (128, or 256 for the CTA pair)
    int k_tail;              // wgrad with a_rows_dev: valid rows of the last k-block (0 = whole), filled in by the kernel
    const int* drop_rows;    // dropout counters use drop_rows[row] instead of row (packed rows keep their padded positions' masks)
};

// Problem size as the kernel sees it: the launch constants, or — with a_rows_dev — recomputed from the device-side row
// count.  Every role (producer, MMA issuer, epilogue) derives the same tile list from it.
template <bool A_MN, bool B_MN>
__device__ __forceinline__ KParams resolve_dynamic(const KParams& p) {
    KParams q = p;
    if (p.a_rows_dev != nullptr) {
        const int r = __ldg(p.a_rows_dev);
        if constexpr (A_MN && B_MN) {                       // wgrad: the reduction runs over the stored rows
            const int K = max(1, min(r, p.K));
            q.K = K;
            q.kb_total = (K + BLOCK_K - 1) / BLOCK_K;
            q.kb_per_split = (q.kb_total + p.split_k - 1) / p.split_k;
            q.k_tail = K % BLOCK_K;
        } else {
            const int M = max(1, min(r, p.M));
            q.M = M;
            q.num_tiles = ((M + p.tile_m - 1) / p.tile_m) * p.num_n_tiles * p.split_k;
        }
    }
    return q;
}

// wgrad over a device-side row count: rows [k_tail, 64) of the LAST k-block lie past the valid rows and hold whatever the
// buffers held before (possibly NaN bit patterns).  Both operands are MN-major there — one 128-byte shared-memory line
// per k index inside every 64-wide atom — so the lines of the invalid k are simply cleared (the swizzle permutes 16-byte
// chunks WITHIN a line) before the MMAs of that block are issued.  Called by all 32 lanes of the MMA warp.
__device__ __forceinline__ void zero_k_tail(uint8_t* stage_ptr, int atoms, int k_tail, int lane) {
    for (int a = 0; a < atoms; ++a) {
        uint8_t* atom = stage_ptr + (size_t)a * (BLOCK_K * 128);
        for (int off = k_tail * 128 + lane * 16; off < BLOCK_K * 128; off += 32 * 16)
            *reinterpret_cast<uint4*>(atom + off) = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
    __syncwarp();
}

// UMMA shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

// UMMA instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
template <int BN, bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

constexpr bool epi_uses_aux(int epi) { return epi == CRCT_EPI_BIAS_RES || epi == CRCT_EPI_MUL; }      // bf16 aux, fetched a tile ahead
constexpr bool epi_is_res(int epi) { return epi == CRCT_EPI_BIAS_RES || epi == CRCT_EPI_BIAS_RES_F32; }

// ---------------------------------------------------------------------------------------------
// Epilogue.  A lane owns one accumulator row; per step it handles EPI_COLS = 16 consecutive columns = 32 bytes of
// bf16 = one DRAM sector, moved with ONE 256-bit global access (sm_100: ld/st.global.v8.b32), so every warp
// instruction reads or writes 32 full sectors and nothing is staged through shared memory.  The four column-quarter
// warps of a lane group fill each row's 128-byte lines between them.  Operands that are not 32-byte aligned / a
// multiple of 16 columns take the 2 x 128-bit path (`wide` = 0).
// ---------------------------------------------------------------------------------------------
struct AuxRegs {
    uint32_t v[EPI_COLS / 2];       // 16 bf16 of one row
};

__device__ __forceinline__ void ld_row32(const bf16* src, bool wide, int cols_left, uint32_t (&v)[8]) {
    if (wide) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(src));
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (h * 8 < cols_left) t = __ldg(reinterpret_cast<const uint4*>(src + h * 8));
            v[h * 4] = t.x; v[h * 4 + 1] = t.y; v[h * 4 + 2] = t.z; v[h * 4 + 3] = t.w;
        }
    }
}
__device__ __forceinline__ void st_row32(bf16* dst, bool wide, int cols_left, const uint32_t (&v)[8]) {
    if (wide) {
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h * 8 < cols_left)
                *reinterpret_cast<uint4*>(dst + h * 8) = make_uint4(v[h * 4], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
    }
}

template <int EPI>
__device__ __forceinline__ void prefetch_aux(const KParams& p, int row, int col, AuxRegs& r) {
    if constexpr (epi_uses_aux(EPI)) {
#pragma unroll
        for (int i = 0; i < EPI_COLS / 2; ++i) r.v[i] = 0u;
        if (p.aux != nullptr && row < p.M && col < p.N) ld_row32(p.aux + (size_t)row * p.ldaux + col, p.wide != 0, p.N - col, r.v);
    }
}

// one accumulator row x EPI_COLS columns: bias (smem copy) / GELU (+ GELU') / dropout + residual / multiply -> global
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const KParams& p, int row, int col0, const uint32_t (&v)[EPI_COLS],
                                               const float* bias_s, const AuxRegs& aux, uint64_t seed, int drow) {
    if (row >= p.M) return;
    if constexpr (EPI == CRCT_EPI_F32) {
#pragma unroll
        for (int g = 0; g < EPI_COLS / 8; ++g) {
            const int col = col0 + g * 8;
            if (col >= p.N) break;
            float* d = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col;
            const float f0 = __uint_as_float(v[g * 8]), f1 = __uint_as_float(v[g * 8 + 1]), f2 = __uint_as_float(v[g * 8 + 2]),
                        f3 = __uint_as_float(v[g * 8 + 3]), f4 = __uint_as_float(v[g * 8 + 4]), f5 = __uint_as_float(v[g * 8 + 5]),
                        f6 = __uint_as_float(v[g * 8 + 6]), f7 = __uint_as_float(v[g * 8 + 7]);
            if (p.accumulate) {
                ptx::red_add_f32x4(d, f0, f1, f2, f3);
                ptx::red_add_f32x4(d + 4, f4, f5, f6, f7);
            } else {
                *reinterpret_cast<float4*>(d) = make_float4(f0, f1, f2, f3);
                *reinterpret_cast<float4*>(d + 4) = make_float4(f4, f5, f6, f7);
            }
        }
    } else {
        uint32_t o[EPI_COLS / 2], o2[EPI_COLS / 2];
#pragma unroll
        for (int g = 0; g < EPI_COLS / 8; ++g) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
            if constexpr (EPI != CRCT_EPI_MUL) {
                if (p.bias != nullptr) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + g * 8);         // smem broadcast
                    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + g * 8 + 4);
                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
            }
            if constexpr (EPI == CRCT_EPI_BIAS_GELU) {
                if (p.D2 != nullptr) {                      // training: also emit gelu'(u) for the backward (shares rcp/ex2)
                    float d[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float hq, e;
                        gelu_parts(f[j], hq, e);
                        d[j] = gelu_grad_from_parts(f[j], hq, e);
                        f[j] = gelu_from_parts(f[j], hq);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) o2[g * 4 + j] = pack_bf16x2(d[2 * j], d[2 * j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = gelu_f(f[j]);
                }
            }
            if constexpr (epi_is_res(EPI)) {
                if (p.drop_thr != 0u)
                    dropout8(f, seed, (uint64_t)drow * (uint64_t)p.N + (uint64_t)(col0 + g * 8), p.drop_thr, p.drop_scale);
            }
            if constexpr (epi_uses_aux(EPI)) {
                if (p.aux != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a = unpack_bf16x2(aux.v[g * 4 + j]);
                        if constexpr (EPI == CRCT_EPI_MUL) {
                            f[2 * j] *= a.x; f[2 * j + 1] *= a.y;
                        } else {
                            f[2 * j] += a.x; f[2 * j + 1] += a.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[g * 4 + j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
        }
        const size_t off = (size_t)row * p.ldd + col0;
        if constexpr (EPI == CRCT_EPI_BIAS_GELU) {
            if (p.D2 != nullptr) st_row32(reinterpret_cast<bf16*>(p.D2) + off, p.wide != 0, p.N - col0, o2);
        }
        st_row32(reinterpret_cast<bf16*>(p.D) + off, p.wide != 0, p.N - col0, o);
    }
}

// ---- CRCT_EPI_BIAS_RES_F32: z (fp32) = dropout(acc + bias) + aux (fp32).  The residual stream stays fp32 end to end: the
// LayerNorm writes its output twice (bf16 = the next GEMM's operand, fp32 = the next residual), and this epilogue adds the
// fp32 copy.  16 fp32 columns per lane and step = 64 bytes.  Like the bf16 aux it is fetched ONE TILE AHEAD (BN/4 registers per
// lane: 48 at BN = 192): with the first version's one-STEP-ahead loads every step waited a DRAM round trip (ncu: 27 us for the
// K = 768 out-projection against 15 us with the bias epilogue — three exposed ~1.5 us latencies per tile).
__device__ __forceinline__ void aux32_load(const KParams& p, int row, int col, float (&f)[EPI_COLS]) {
#pragma unroll
    for (int i = 0; i < EPI_COLS; ++i) f[i] = 0.f;
    if (row < p.M) {
        const float* src = reinterpret_cast<const float*>(p.aux) + (size_t)row * p.ldaux + col;
#pragma unroll
        for (int h = 0; h < EPI_COLS / 8; ++h) {
            if (col + h * 8 < p.N) {
                uint32_t v[8];
                asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(src + h * 8));
#pragma unroll
                for (int j = 0; j < 8; ++j) f[h * 8 + j] = __uint_as_float(v[j]);
            }
        }
    }
}

__device__ __forceinline__ void epilogue_chunk_res32(const KParams& p, int row, int col0, const uint32_t (&v)[EPI_COLS],
                                                     const float* bias_s, const float (&ax)[EPI_COLS], uint64_t seed, int drow) {
    if (row >= p.M) return;
#pragma unroll
    for (int g = 0; g < EPI_COLS / 8; ++g) {
        if (col0 + g * 8 >= p.N) break;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
        if (p.bias != nullptr) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + g * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + g * 8 + 4);
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        if (p.drop_thr != 0u) dropout8(f, seed, (uint64_t)drow * (uint64_t)p.N + (uint64_t)(col0 + g * 8), p.drop_thr, p.drop_scale);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += ax[g * 8 + j];
        float* d = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col0 + g * 8;
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(d), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])),
                        "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])) : "memory");
    }
}

__device__ __forceinline__ void reg_fence16(uint32_t (&v)[16]) {
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}

template <int BN>
struct EpiGeom {
    static constexpr int STEPS = (BN / 4) / EPI_COLS;
};
// what a lane keeps of the aux operand across tiles: its row's BN/4 bf16 columns of the next tile, or (fp32 aux) step 0 only
template <int BN, int EPI>
struct AuxTile {
    AuxRegs r[EpiGeom<BN>::STEPS];
};
template <int BN>
struct AuxTile<BN, CRCT_EPI_BIAS_RES_F32> {
    // BN <= 192: the lane's whole row slice of the tile (BN/4 fp32 columns = 48 registers), fetched a tile ahead.
    // BN = 256 (64 registers: spills under the 96-register cap of a 640-thread CTA): step 0 a tile ahead, the rest a step ahead.
    static constexpr bool TILE_AHEAD = BN <= 192;
    static constexpr int HELD = TILE_AHEAD ? EpiGeom<BN>::STEPS : 1;
    float f[HELD][EPI_COLS];
};

// aux (residual / multiplier) is fetched ONE TILE AHEAD: a lane keeps its row's BN/4 columns of the current tile in
// registers (STEPS x 32 B); as soon as a step's 32 bytes are consumed, the same registers receive the next tile's
// bytes, so every load has a whole tile time to arrive even when the epilogue is the critical path.
template <int BN, int EPI>
__device__ __forceinline__ void aux_load_tile(const KParams& p, int m0, int n0, int warp, int lane, AuxTile<BN, EPI>& aux) {
    const int row = m0 + (warp & 3) * 32 + lane;
    const int cbase = n0 + ((warp - EPI_WARP0) >> 2) * (BN / 4);
    if constexpr (EPI == CRCT_EPI_BIAS_RES_F32) {
#pragma unroll
        for (int c = 0; c < AuxTile<BN, EPI>::HELD; ++c) aux32_load(p, row, cbase + c * EPI_COLS, aux.f[c]);
    } else {
#pragma unroll
        for (int c = 0; c < EpiGeom<BN>::STEPS; ++c) prefetch_aux<EPI>(p, row, cbase + c * EPI_COLS, aux.r[c]);
    }
}

// Stage the tile's bias slice in smem (named barrier 1 among the epilogue warps); runs before the accumulator is ready.
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_prepare(const KParams& p, int n0, int warp, int lane, float* bias_s) {
    if constexpr (EPI != CRCT_EPI_F32 && EPI != CRCT_EPI_MUL) {
        if (p.bias != nullptr) {
            const int t = (warp - EPI_WARP0) * 32 + lane;   // 0..511
            if (t < BN) bias_s[t] = (n0 + t < p.N) ? p.bias[n0 + t] : 0.f;
        }
        asm volatile("bar.sync 1, %0;" :: "n"(NUM_EPI_WARPS * 32) : "memory");
    }
}

// Epilogue of one tile for one warp (16 epilogue warps: 4 TMEM lane groups x 4 column quarters); per 16-column step:
// tcgen05.ld (next step's load in flight) -> fused math -> one 256-bit store per output -> aux refill for the next
// tile (m0n, n0n; has_next = 0 on the CTA's last tile).
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const KParams& p, uint32_t tmem_acc, int m0, int n0, int warp, int lane,
                                              const float* bias_s, AuxTile<BN, EPI>& aux, bool has_next, int m0n, int n0n) {
    constexpr int STEPS = EpiGeom<BN>::STEPS;
    const int lane_grp = warp & 3;                          // TMEM lanes [32*lane_grp, +32) are this warp's
    const int col_q = (warp - EPI_WARP0) >> 2;              // column quarter
    uint64_t seed = p.seed;
    if constexpr (epi_is_res(EPI)) {
        if (p.drop_thr != 0u && p.salt != nullptr) seed ^= __ldg(p.salt);
    }
    const int row = m0 + lane_grp * 32 + lane;
    const int rown = m0n + lane_grp * 32 + lane;
    int drow = row;
    if constexpr (epi_is_res(EPI)) {
        if (p.drop_thr != 0u && p.drop_rows != nullptr && row < p.M) drow = __ldg(p.drop_rows + row);
    }
    const int cbase = col_q * (BN / 4);
    const uint32_t taddr = tmem_acc + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)cbase;
    uint32_t v[2][EPI_COLS];
    ptx::tc_ld_32x16(taddr, v[0]);
    if constexpr (EPI == CRCT_EPI_BIAS_RES_F32) {
        if constexpr (AuxTile<BN, EPI>::TILE_AHEAD) {
#pragma unroll
            for (int c = 0; c < STEPS; ++c) {
                const int cc = cbase + c * EPI_COLS;
                ptx::tc_wait_ld();
                reg_fence16(v[c & 1]);
                if (c + 1 < STEPS) ptx::tc_ld_32x16(taddr + (uint32_t)((c + 1) * EPI_COLS), v[(c + 1) & 1]);
                if (n0 + cc < p.N) epilogue_chunk_res32(p, row, n0 + cc, v[c & 1], bias_s + cc, aux.f[c], seed, drow);
                if (has_next) aux32_load(p, rown, n0n + cc, aux.f[c]);                         // the same registers: next tile's step c
            }
        } else {
            float ax[2][EPI_COLS];
#pragma unroll
            for (int i = 0; i < EPI_COLS; ++i) ax[0][i] = aux.f[0][i];
#pragma unroll
            for (int c = 0; c < STEPS; ++c) {
                const int cc = cbase + c * EPI_COLS;
                if (c + 1 < STEPS) aux32_load(p, row, n0 + cc + EPI_COLS, ax[(c + 1) & 1]);    // one step ahead
                ptx::tc_wait_ld();
                reg_fence16(v[c & 1]);
                if (c + 1 < STEPS) ptx::tc_ld_32x16(taddr + (uint32_t)((c + 1) * EPI_COLS), v[(c + 1) & 1]);
                if (n0 + cc < p.N) epilogue_chunk_res32(p, row, n0 + cc, v[c & 1], bias_s + cc, ax[c & 1], seed, drow);
            }
            if (has_next) aux32_load(p, rown, n0n + cbase, aux.f[0]);                          // step 0 of the next tile
        }
    } else {
#pragma unroll
        for (int c = 0; c < STEPS; ++c) {
            const int cc = cbase + c * EPI_COLS;
            ptx::tc_wait_ld();
            reg_fence16(v[c & 1]);              // the loaded values exist from here on (tcgen05.ld is asynchronous)
            if (c + 1 < STEPS) ptx::tc_ld_32x16(taddr + (uint32_t)((c + 1) * EPI_COLS), v[(c + 1) & 1]);
            if (n0 + cc < p.N) epilogue_chunk<EPI>(p, row, n0 + cc, v[c & 1], bias_s + cc, aux.r[c], seed, drow);     // warp-uniform
            if constexpr (epi_uses_aux(EPI)) {
                if (has_next) prefetch_aux<EPI>(p, rown, n0n + cc, aux.r[c]);
            }
        }
    }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p_launch) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* gbase = smem_raw + (base - raw_addr);
    const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
    auto smem_a = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
    auto smem_b = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES + C::A_BYTES; };
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * C::STAGES + i); };
    auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * C::STAGES + 2 + i); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();    // the next kernel of the stream may be scheduled behind this grid (see common.cuh)

    if (warp == 0 && lane == 0) {
        ptx::tma_prefetch_desc(&tmA);
        ptx::tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar(i), 1);
            ptx::mbar_init(tempty_bar(i), NUM_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                 // everything above touched only this CTA's shared / tensor memory
    const KParams p = resolve_dynamic<A_MN, B_MN>(p_launch);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int ks = tile % p.split_k;
                const int mn = tile / p.split_k;
                const int m0 = (mn / p.num_n_tiles) * BLOCK_M;
                const int n0 = (mn % p.num_n_tiles) * BN;
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);      // kb0 >= kb1: an empty split (device-side K), skipped by every role
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const int k0 = kb * BLOCK_K;
                    if constexpr (!A_MN) {
                        ptx::tma_load_2d(smem_a(stage), &tmA, full_bar(stage), k0, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; ++j)
                            ptx::tma_load_2d(smem_a(stage) + j * (BLOCK_K * 128), &tmA, full_bar(stage), m0 + j * 64, k0);
                    }
                    if constexpr (!B_MN) {
                        ptx::tma_load_2d(smem_b(stage), &tmB, full_bar(stage), k0, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            ptx::tma_load_2d(smem_b(stage) + j * (BLOCK_K * 128), &tmB, full_bar(stage), n0 + j * 64, k0);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // one lane issues; with a partial last k-block (k_tail, device-side K of a wgrad) the whole warp walks the loop so
        // that all 32 lanes can clear the invalid k lines of that block before its MMAs
        const bool tail = (A_MN && B_MN) && p.k_tail != 0;
        if (lane == 0 || tail) {
            constexpr uint32_t idesc = make_idesc<BN, A_MN, B_MN>();
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int ks = tile % p.split_k;
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                if (kb0 >= kb1) continue;
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                ++it;
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    if constexpr (A_MN && B_MN) {
                        if (tail && kb == p.kb_total - 1) {
                            zero_k_tail(gbase + (size_t)stage * C::STAGE_BYTES, BLOCK_M / 64, p.k_tail, lane);
                            zero_k_tail(gbase + (size_t)stage * C::STAGE_BYTES + C::A_BYTES, BN / 64, p.k_tail, lane);
                        }
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t adesc = make_smem_desc(smem_a(stage) + k * p.a_kstep, p.a_lbo, p.a_sbo);
                            const uint64_t bdesc = make_smem_desc(smem_b(stage) + k * p.b_kstep, p.b_lbo, p.b_sbo);
                            ptx::tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        ptx::tc_commit(empty_bar(stage));          // smem slot reusable once these MMAs retire
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                if (lane == 0) ptx::tc_commit(tfull_bar(acc));      // accumulator complete -> epilogue
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue =====================
        float* bias_s = reinterpret_cast<float*>(gbase + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);
        int it = 0;
        auto origin = [&](int tile, int& m0, int& n0) {
            const int mn = tile / p.split_k;
            m0 = (mn / p.num_n_tiles) * BLOCK_M;
            n0 = (mn % p.num_n_tiles) * BN;
        };
        AuxTile<BN, EPI> aux;
        if (blockIdx.x < p.num_tiles) {
            int m0, n0;
            origin(blockIdx.x, m0, n0);
            aux_load_tile<BN, EPI>(p, m0, n0, warp, lane, aux);
        }
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            {
                const int kb0 = (tile % p.split_k) * p.kb_per_split;
                if (kb0 >= min(p.kb_total, kb0 + p.kb_per_split)) continue;          // empty split (device-side K)
            }
            int m0, n0, m0n = 0, n0n = 0;
            origin(tile, m0, n0);
            const bool has_next = tile + (int)gridDim.x < p.num_tiles;
            if (has_next) origin(tile + (int)gridDim.x, m0n, n0n);
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            ++it;
            epilogue_prepare<BN, EPI>(p, n0, warp, lane, bias_s + acc * 256);
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            epilogue_tile<BN, EPI>(p, tmem_base + (uint32_t)(acc * BN), m0, n0, warp, lane, bias_s + acc * 256, aux, has_next, m0n, n0n);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Grouped weight-gradient launch: up to MAXG independent problems dW_g[out,in] += dy_g^T x_g (both operands MN-major, fp32
// split-K accumulation) through ONE persistent tile list.  A training step has 121 such GEMMs — four per text / visual layer
// (QKV, attention output, FFN in, FFN out), none of them on the backward's critical path; launched one by one each pays its
// own pipeline fill and tail (~5 us of a 14-33 us kernel) and its own wave quantisation (the 1900-row visual problems fill a
// fraction of the SMs).  The host defers a layer's weight gradients and issues them together: same tile kernel, same
// per-problem device-side K (`a_rows_dev`) and k-tail clearing, per-problem tensor maps indexed out of the kernel parameters.
// ---------------------------------------------------------------------------------------------
constexpr int MAXG = 8;
struct GroupParams {
    int count;
    int tile_start[MAXG + 1];            // prefix sums of tiles per problem
    KParams p[MAXG];
};
struct GroupMaps {
    CUtensorMap a[MAXG];
    CUtensorMap b[MAXG];
};

struct GroupTile {
    int g, m0, n0, kb0, kb1, kb_total, k_tail;
};
// tile index -> problem, tile origin and k-block range (device-side K resolved here: every role derives the same values)
template <int BN>
__device__ __forceinline__ GroupTile group_tile(const GroupParams& gp, int tile) {
    GroupTile t;
    int g = 0;
    while (g + 1 < gp.count && tile >= gp.tile_start[g + 1]) ++g;
    const KParams& P = gp.p[g];
    const int local = tile - gp.tile_start[g];
    int K = P.K;
    if (P.a_rows_dev != nullptr) K = max(1, min(__ldg(P.a_rows_dev), P.K));
    t.g = g;
    t.kb_total = (K + BLOCK_K - 1) / BLOCK_K;
    t.k_tail = P.a_rows_dev != nullptr ? K % BLOCK_K : 0;
    const int kbps = (t.kb_total + P.split_k - 1) / P.split_k;
    const int ks = local % P.split_k, mn = local / P.split_k;
    t.m0 = (mn / P.num_n_tiles) * BLOCK_M;
    t.n0 = (mn % P.num_n_tiles) * BN;
    t.kb0 = ks * kbps;
    t.kb1 = min(t.kb_total, t.kb0 + kbps);
    return t;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_wgrad_grouped_kernel(const __grid_constant__ GroupMaps tm, const __grid_constant__ GroupParams gp) {
    using C = Cfg<BN>;
    constexpr int EPI = CRCT_EPI_F32;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw_addr);
    const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
    auto smem_a = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
    auto smem_b = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES + C::A_BYTES; };
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * C::STAGES + i); };
    auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * C::STAGES + 2 + i); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar(i), 1);
            ptx::mbar_init(tempty_bar(i), NUM_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();
    const int total = gp.tile_start[gp.count];

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const GroupTile t = group_tile<BN>(gp, tile);
                for (int kb = t.kb0; kb < t.kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const int k0 = kb * BLOCK_K;
#pragma unroll
                    for (int j = 0; j < BLOCK_M / 64; ++j)
                        ptx::tma_load_2d(smem_a(stage) + j * (BLOCK_K * 128), &tm.a[t.g], full_bar(stage), t.m0 + j * 64, k0);
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        ptx::tma_load_2d(smem_b(stage) + j * (BLOCK_K * 128), &tm.b[t.g], full_bar(stage), t.n0 + j * 64, k0);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the loop: any problem may have a partial last k-block) =====================
        constexpr uint32_t idesc = make_idesc<BN, true, true>();
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const GroupTile t = group_tile<BN>(gp, tile);
            if (t.kb0 >= t.kb1) continue;
            const KParams& P = gp.p[t.g];
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            ++it;
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = t.kb0; kb < t.kb1; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                if (t.k_tail != 0 && kb == t.kb_total - 1) {
                    zero_k_tail(gbase + (size_t)stage * C::STAGE_BYTES, BLOCK_M / 64, t.k_tail, lane);
                    zero_k_tail(gbase + (size_t)stage * C::STAGE_BYTES + C::A_BYTES, BN / 64, t.k_tail, lane);
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t adesc = make_smem_desc(smem_a(stage) + k * P.a_kstep, P.a_lbo, P.a_sbo);
                        const uint64_t bdesc = make_smem_desc(smem_b(stage) + k * P.b_kstep, P.b_lbo, P.b_sbo);
                        ptx::tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::tc_commit(empty_bar(stage));
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            }
            if (lane == 0) ptx::tc_commit(tfull_bar(acc));
            __syncwarp();
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue (fp32 red.add into the gradient arena) =====================
        int it = 0;
        AuxTile<BN, EPI> aux;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const GroupTile t = group_tile<BN>(gp, tile);
            if (t.kb0 >= t.kb1) continue;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            ++it;
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            epilogue_tile<BN, EPI>(gp.p[t.g], tmem_base + (uint32_t)(acc * BN), t.m0, t.n0, warp, lane, nullptr, aux, false, 0, 0);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.  Each CTA stages its own
// 128 rows of A and only HALF of the B tile (BN/2 rows); one tcgen05.mma.cta_group::2 (M = 256) issued by the leader
// CTA drives both SMs' tensor cores and reads B from both shared memories.  Per SM and per k-block that is 32 KB of TMA
// writes and 8 KB of operand reads per MMA instead of 48 KB / 12 KB — the single-CTA kernel is shared-memory-bandwidth
// bound (96 B/clk of MMA reads + 96 B/clk of TMA writes against a 128 B/clk port), this one is not.
//   full[s]    (leader)  : 1 arrival (leader's expect_tx of both CTAs' bytes); both CTAs' TMA complete_tx land here
//   empty[s]   (each CTA): 1 arrival = multicast tcgen05.commit after the pair's MMAs on stage s retire
//   tfull[i]   (each CTA): 1 arrival = multicast commit after the last k-block of a tile
//   tempty[i]  (leader)  : 2 x 8 arrivals = every epilogue warp of both CTAs (remote mbarrier.arrive)
// ---------------------------------------------------------------------------------------------
template <int BN>
struct Cfg2 {
    static constexpr int STAGES = (BN == 256) ? 6 : 8;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 8 * (2 * STAGES + 4) + 16;
    static constexpr int BIAS_BYTES = 2 * 256 * 4;              // per accumulator stage: the tile's bias slice
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
};

template <int BN, bool A_MN, bool B_MN>
__device__ __forceinline__ constexpr uint32_t make_idesc_2cta() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
    using C = Cfg2<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - raw_addr);
    const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
    auto smem_a = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES; };
    auto smem_b = [&](int s) { return base + (uint32_t)s * C::STAGE_BYTES + C::A_BYTES; };
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * C::STAGES + i); };
    auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * C::STAGES + 2 + i); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();    // see the single-CTA kernel
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::tma_prefetch_desc(&tmA);
        ptx::tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(tfull_bar(i), 1);
            ptx::mbar_init(tempty_bar(i), 2 * NUM_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
        ptx::tmem_relinquish_2cta();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();                                   // barriers of BOTH CTAs initialised before any remote arrive
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                 // everything above touched only this CTA's shared / tensor memory

    if (warp == 0) {
        // ===================== TMA producer (both CTAs: own A rows, own half of B) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
                const int ks = tile % p.split_k;
                const int mn = tile / p.split_k;
                const int m0 = (mn / p.num_n_tiles) * (2 * BLOCK_M) + (int)rank * BLOCK_M;
                const int n0 = (mn % p.num_n_tiles) * BN + (int)rank * (BN / 2);
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                    const int k0 = kb * BLOCK_K;
                    if constexpr (!A_MN) {
                        ptx::tma_load_2d_2cta(smem_a(stage), &tmA, full_bar(stage), k0, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; ++j)
                            ptx::tma_load_2d_2cta(smem_a(stage) + j * (BLOCK_K * 128), &tmA, full_bar(stage), m0 + j * 64, k0);
                    }
                    if constexpr (!B_MN) {
                        ptx::tma_load_2d_2cta(smem_b(stage), &tmB, full_bar(stage), k0, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 128; ++j)
                            ptx::tma_load_2d_2cta(smem_b(stage) + j * (BLOCK_K * 128), &tmB, full_bar(stage), n0 + j * 64, k0);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only; one instruction drives both SMs) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_2cta<BN, A_MN, B_MN>();
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
                const int ks = tile % p.split_k;
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t adesc = make_smem_desc(smem_a(stage) + k * p.a_kstep, p.a_lbo, p.a_sbo);
                        const uint64_t bdesc = make_smem_desc(smem_b(stage) + k * p.b_kstep, p.b_lbo, p.b_sbo);
                        ptx::tc_mma_bf16_2cta(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    ptx::tc_commit_2cta(empty_bar(stage), 0x3);       // both CTAs' producers may refill stage
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
                ptx::tc_commit_2cta(tfull_bar(acc), 0x3);             // both CTAs' epilogues may drain their half
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue (each CTA drains its own 128 accumulator rows) =====================
        float* bias_s = reinterpret_cast<float*>(gbase + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);
        int it = 0;
        auto origin = [&](int tile, int& m0, int& n0) {
            const int mn = tile / p.split_k;
            m0 = (mn / p.num_n_tiles) * (2 * BLOCK_M) + (int)rank * BLOCK_M;
            n0 = (mn % p.num_n_tiles) * BN;
        };
        AuxTile<BN, EPI> aux;
        if (cluster_id < p.num_tiles) {
            int m0, n0;
            origin(cluster_id, m0, n0);
            aux_load_tile<BN, EPI>(p, m0, n0, warp, lane, aux);
        }
        for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
            int m0, n0, m0n = 0, n0n = 0;
            origin(tile, m0, n0);
            const bool has_next = tile + num_clusters < p.num_tiles;
            if (has_next) origin(tile + num_clusters, m0n, n0n);
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            epilogue_prepare<BN, EPI>(p, n0, warp, lane, bias_s + acc * 256);
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            epilogue_tile<BN, EPI>(p, tmem_base + (uint32_t)(acc * BN), m0, n0, warp, lane, bias_s + acc * 256, aux, has_next, m0n, n0n);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(tempty_bar(acc), 0);
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();                                   // the peer's smem / TMEM stay alive until the leader is done
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
        if (q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// 2-D bf16 tensor, `inner` contiguous elements per row, `outer` rows of stride ld elements
int make_tmap(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) CRCT_FAIL(CRCT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CRCT_FAIL(CRCT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u",
                                     (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    return CRCT_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int grid, cudaStream_t st) {
    auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, EPI>;
    static bool configured = false;       // per instantiation
    if (!configured) {
        CRCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
        configured = true;
    }
    CRCT_CUDA(crct_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), Cfg<BN>::SMEM_BYTES, st, tmA, tmB, p));
    return CRCT_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int grid, cudaStream_t st) {
    auto kern = gemm_tcgen05_2cta_kernel<BN, A_MN, B_MN, EPI>;
    static bool configured = false;       // per instantiation
    if (!configured) {
        CRCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::SMEM_BYTES));
        configured = true;
    }
    CRCT_CUDA(crct_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), Cfg2<BN>::SMEM_BYTES, st, tmA, tmB, p));     // __cluster_dims__(2,1,1): grid is even
    return CRCT_OK;
}

template <int BN>
int dispatch2(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int grid, cudaStream_t st) {
    if (!a_mn && !b_mn) {
        if (epi == CRCT_EPI_BIAS) return launch2<BN, false, false, CRCT_EPI_BIAS>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_GELU) return launch2<BN, false, false, CRCT_EPI_BIAS_GELU>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_RES) return launch2<BN, false, false, CRCT_EPI_BIAS_RES>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_RES_F32) return launch2<BN, false, false, CRCT_EPI_BIAS_RES_F32>(tmA, tmB, p, grid, st);
    } else if (!a_mn && b_mn) {
        if (epi == CRCT_EPI_BIAS) return launch2<BN, false, true, CRCT_EPI_BIAS>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_RES) return launch2<BN, false, true, CRCT_EPI_BIAS_RES>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_MUL) return launch2<BN, false, true, CRCT_EPI_MUL>(tmA, tmB, p, grid, st);
    } else if (a_mn && b_mn) {
        if (epi == CRCT_EPI_F32) return launch2<BN, true, true, CRCT_EPI_F32>(tmA, tmB, p, grid, st);
    }
    CRCT_FAIL(CRCT_ERR_ARG, "unsupported GEMM variant a_major=%d b_major=%d epilogue=%d", a_mn, b_mn, epi);
}

template <int BN>
int dispatch(int a_mn, int b_mn, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& p, int grid, cudaStream_t st) {
    if (!a_mn && !b_mn) {
        if (epi == CRCT_EPI_BIAS) return launch<BN, false, false, CRCT_EPI_BIAS>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_GELU) return launch<BN, false, false, CRCT_EPI_BIAS_GELU>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_RES) return launch<BN, false, false, CRCT_EPI_BIAS_RES>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_RES_F32) return launch<BN, false, false, CRCT_EPI_BIAS_RES_F32>(tmA, tmB, p, grid, st);
    } else if (!a_mn && b_mn) {
        if (epi == CRCT_EPI_BIAS) return launch<BN, false, true, CRCT_EPI_BIAS>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_BIAS_RES) return launch<BN, false, true, CRCT_EPI_BIAS_RES>(tmA, tmB, p, grid, st);
        if (epi == CRCT_EPI_MUL) return launch<BN, false, true, CRCT_EPI_MUL>(tmA, tmB, p, grid, st);
    } else if (a_mn && b_mn) {
        if (epi == CRCT_EPI_F32) return launch<BN, true, true, CRCT_EPI_F32>(tmA, tmB, p, grid, st);
    }
    CRCT_FAIL(CRCT_ERR_ARG, "unsupported GEMM variant a_major=%d b_major=%d epilogue=%d", a_mn, b_mn, epi);
}

}  // namespace

// 2-D bf16 tensor map with 128-byte swizzle for the other TMA users of the library (attention_tc.cu)
int crct_make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
    return make_tmap(tm, ptr, inner, outer, ld, box_inner, box_outer);
}

static inline bool max_ctas_unset(const crct_gemm_t* a) { return a->max_ctas <= 0; }

// auto policy for cta_group == 0
static bool crct_gemm_auto_pair(const crct_gemm_t* a) {
    // measured on B200 (profiles/r01_gemm_shapes.log): the CTA pair wins ~4 % once there are >= 8 tile columns to
    // share; narrower outputs quantise worse on 256-row tiles
    const char* e = getenv("CRCT_GEMM_PAIR_POLICY");            // tuning aid (tools/ab_policy.py)
    const int policy = e ? atoi(e) : 0;
    if (a->epilogue == CRCT_EPI_F32) return (policy & 1) != 0;
    if (policy & 2) return a->M >= 1024 && a->N >= 768;
    if (policy & 4) return false;
    return a->N >= 2048 && a->M >= 1024;
}

extern "C" CRCT_API int crct_gemm_bf16(const crct_gemm_t* a, crct_stream_t stream) {
    if (!a || !a->A || !a->B || !a->D) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: null pointer");
    if (a->M <= 0 || a->N <= 0 || a->K <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_gemm_bf16: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
    if (a->N % 8) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_gemm_bf16: N=%d must be a multiple of 8", a->N);
    if ((a->lda % 8) || (a->ldb % 8) || (a->ldd % 8) || (a->aux && (a->ldaux % 8)))
        CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: leading dimensions must be multiples of 8 elements");
    if (((uintptr_t)a->A | (uintptr_t)a->B | (uintptr_t)a->D | (uintptr_t)a->D2 | (uintptr_t)a->aux | (uintptr_t)a->bias) & 15)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: pointers must be 16-byte aligned");
    if (a->epilogue < CRCT_EPI_BIAS || a->epilogue > CRCT_EPI_BIAS_RES_F32) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: unknown epilogue %d", a->epilogue);
    if (a->a_rows_dev && a->cta_group == 2) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: a_rows_dev is not supported by the CTA-pair kernel");
    if (a->a_rows_dev && a->a_major && !a->b_major) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: a_rows_dev needs a_major = 0 (rows = M) or the wgrad form a_major = b_major = 1 (rows = K)");
    if ((a->epilogue == CRCT_EPI_MUL) && !a->aux) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: MUL epilogue needs aux");
    if (a->epilogue == CRCT_EPI_BIAS_RES_F32 && (!a->aux || ((uintptr_t)a->aux & 31) || ((uintptr_t)a->D & 31)))
        CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: RES_F32 epilogue needs a 32-byte aligned fp32 aux and D");
    const bool f32 = a->epilogue == CRCT_EPI_F32;
    int split_k = a->split_k;
    const int kb_total = (a->K + BLOCK_K - 1) / BLOCK_K;
    const int sms = a->max_ctas > 0 ? a->max_ctas : crct_num_sms();
    if (sms <= 0) return CRCT_ERR_CUDA;

    const bool pair = a->cta_group == 2 || (a->cta_group == 0 && !a->a_rows_dev && a->block_n != 192 && crct_gemm_auto_pair(a));
    const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
    int bn = a->block_n;
    if (bn == 0) {
        // rows the kernel is EXPECTED to run (packed layout: the host never reads *a_rows_dev; the caller's estimate of it only
        // picks the tile shape): a 1.09-wave problem on 128x256 tiles (6900 x 768 -> 162 tiles on 148 SMs) costs two full waves
        const int m_eff = (a->a_rows_dev && !a->a_major && a->rows_hint > 0 && a->rows_hint < a->M) ? a->rows_hint : a->M;
        auto cost = [&](int b, int penalty_pct) {
            long tiles = (long)((m_eff + tile_m - 1) / tile_m) * ((a->N + b - 1) / b);
            const long slots = pair ? sms / 2 : sms;
            return ((tiles + slots - 1) / slots) * (long)b * penalty_pct;
        };
        // 128x256 tiles move 1.33x fewer operand bytes per FLOP through L2/smem than 128x128 (measured faster on every
        // CRCT shape, wgrad included: profiles/r01_wgrad_shapes.log): a narrower tile must win by its penalty to be chosen
        bn = 256;
        long best = cost(256, 100);
        if (!pair && a->N > 128 && cost(192, 106) < best) { bn = 192; best = cost(192, 106); }
        if (a->N <= 128 || cost(128, 125) < best) bn = 128;
        if (f32 && a->accumulate && a->N >= 256) bn = 256;       // split-K refills the SMs: wave count is not the issue (measured)
    }
    if (bn != 128 && bn != 192 && bn != 256) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: block_n must be 0, 128, 192 or 256");
    if (bn == 192 && pair) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: block_n 192 is a single-CTA tile (cta_group 1)");
    const int num_m_tiles = (a->M + tile_m - 1) / tile_m;
    const int num_n_tiles = (a->N + bn - 1) / bn;
    const int slots = pair ? sms / 2 : sms;          // concurrently resident tiles
    if (pair && slots < 1) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: cta_group 2 needs at least 2 CTAs");
    if (split_k <= 0) {
        split_k = 1;
        if (f32 && a->accumulate) {
            const int mn = num_m_tiles * num_n_tiles;
            split_k = slots / mn;
            if (split_k > kb_total / 4) split_k = kb_total / 4;
            if (split_k < 1) split_k = 1;
        }
    }
    if (split_k > 1 && !(f32 && a->accumulate)) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_bf16: split_k > 1 needs CRCT_EPI_F32 with accumulate");
    if (split_k > kb_total) split_k = kb_total;
    const int kb_per_split = (kb_total + split_k - 1) / split_k;
    split_k = (kb_total + kb_per_split - 1) / kb_per_split;      // no empty split

    KParams p;
    memset(&p, 0, sizeof(p));
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.num_n_tiles = num_n_tiles;
    p.split_k = split_k;
    p.num_tiles = num_m_tiles * num_n_tiles * split_k;
    p.kb_total = kb_total;
    p.kb_per_split = kb_per_split;
    p.D = a->D; p.D2 = a->D2; p.bias = a->bias; p.aux = reinterpret_cast<const bf16*>(a->aux);
    p.ldd = a->ldd; p.ldaux = a->ldaux;
    {
        auto ok32 = [](const void* q, int ld) { return q == nullptr || ((reinterpret_cast<uintptr_t>(q) & 31u) == 0 && ld % 16 == 0); };
        const bool d32 = a->epilogue == CRCT_EPI_BIAS_RES_F32;        // fp32 D and aux: their own 32-byte path
        p.wide = (!f32 && !d32 && a->N % 16 == 0 && ok32(a->D, a->ldd) && ok32(a->D2, a->ldd) && ok32(a->aux, a->ldaux)) ? 1 : 0;
    }
    p.accumulate = a->accumulate;
    p.drop_thr = crct_drop_threshold(a->dropout_p);
    p.drop_scale = a->dropout_p > 0.f ? 1.0f / (1.0f - a->dropout_p) : 1.0f;
    p.seed = a->seed;
    p.salt = reinterpret_cast<const unsigned long long*>(a->salt);
    p.a_rows_dev = a->a_rows_dev;
    p.drop_rows = a->drop_rows;
    p.tile_m = tile_m;
    p.k_tail = 0;
    // K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO); LBO unused; +32 B per UMMA_K step.
    // MN-major, SWIZZLE_128B: one TMA box = 64 (MN) x BLOCK_K (K) -> K-groups of 8 rows 1024 B apart (SBO),
    //                         64-wide MN atoms BLOCK_K*128 B apart (LBO); +16 rows * 128 B per UMMA_K step.
    p.a_lbo = a->a_major ? BLOCK_K * 128 : 16;  p.a_sbo = 1024;  p.a_kstep = a->a_major ? UMMA_K * 128 : UMMA_K * 2;
    p.b_lbo = a->b_major ? BLOCK_K * 128 : 16;  p.b_sbo = 1024;  p.b_kstep = a->b_major ? UMMA_K * 128 : UMMA_K * 2;
    if (a->dbg[0]) {   // bring-up overrides: {enable, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep}
        p.a_lbo = a->dbg[1]; p.a_sbo = a->dbg[2]; p.a_kstep = a->dbg[3];
        p.b_lbo = a->dbg[4]; p.b_sbo = a->dbg[5]; p.b_kstep = a->dbg[6];
    }

    CUtensorMap tmA, tmB;
    int rc;
    if (!a->a_major) rc = make_tmap(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BLOCK_K, BLOCK_M);
    else             rc = make_tmap(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BLOCK_K);
    if (rc) return rc;
    if (!a->b_major) rc = make_tmap(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb, BLOCK_K, (uint32_t)(pair ? bn / 2 : bn));
    else             rc = make_tmap(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BLOCK_K);
    if (rc) return rc;

    cudaStream_t st = as_stream(stream);
    if (pair) {
        const int grid2 = 2 * (p.num_tiles < slots ? p.num_tiles : slots);
        if (bn == 256) return dispatch2<256>(a->a_major, a->b_major, a->epilogue, tmA, tmB, p, grid2, st);
        return dispatch2<128>(a->a_major, a->b_major, a->epilogue, tmA, tmB, p, grid2, st);
    }
    int grid = p.num_tiles < sms ? p.num_tiles : sms;
    {
        // balanced persistent grid: a 1.46-wave problem (216 tiles) takes two tile-times on 148 CTAs and on 108 CTAs alike — with
        // 108, the other 40 SMs are free for the kernels of the other streams (visual lane, weight gradients) for the whole
        // duration instead of only after the first wave.  Expected tile count (packed rows: from the caller's row estimate).
        static const bool balance = getenv("CRCT_GEMM_NO_BALANCE") == nullptr;
        const int m_eff = (a->a_rows_dev && !a->a_major && a->rows_hint > 0 && a->rows_hint < a->M) ? a->rows_hint : a->M;
        const long tiles_eff = (long)((m_eff + tile_m - 1) / tile_m) * num_n_tiles * split_k;
        if (balance && tiles_eff > sms && max_ctas_unset(a)) {
            const long waves = (tiles_eff + sms - 1) / sms;
            int g = (int)((tiles_eff + waves - 1) / waves);
            g += g / 16;                       // slack for batches with more valid rows than estimated
            if (g < grid) grid = g;
        }
    }
    if (bn == 256) return dispatch<256>(a->a_major, a->b_major, a->epilogue, tmA, tmB, p, grid, st);
    if (bn == 192) return dispatch<192>(a->a_major, a->b_major, a->epilogue, tmA, tmB, p, grid, st);
    return dispatch<128>(a->a_major, a->b_major, a->epilogue, tmA, tmB, p, grid, st);
}

// Grouped weight gradients (see gemm_wgrad_grouped_kernel): `count` problems in the wgrad form of crct_gemm_bf16
// (a_major = b_major = 1, CRCT_EPI_F32, accumulate = 1), one launch.
extern "C" CRCT_API int crct_gemm_wgrad_grouped(const crct_gemm_t* probs, int count, crct_stream_t stream) {
    if (!probs || count <= 0 || count > MAXG) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_wgrad_grouped: 1 .. %d problems", MAXG);
    const int sms = crct_num_sms();
    if (sms <= 0) return CRCT_ERR_CUDA;
    constexpr int BN = 256;
    long work = 0;
    int kb_eff[MAXG], mn[MAXG];
    for (int g = 0; g < count; ++g) {
        const crct_gemm_t* a = &probs[g];
        if (!a->A || !a->B || !a->D) CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_wgrad_grouped: null pointer in problem %d", g);
        if (!a->a_major || !a->b_major || a->epilogue != CRCT_EPI_F32 || !a->accumulate)
            CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_wgrad_grouped: problem %d is not in the weight-gradient form (a_major = b_major = 1, CRCT_EPI_F32, accumulate)", g);
        if (a->M <= 0 || a->N <= 0 || a->K <= 0 || (a->N % 8)) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_gemm_wgrad_grouped: bad shape in problem %d", g);
        if ((a->lda % 8) || (a->ldb % 8) || (a->ldd % 4) || (((uintptr_t)a->A | (uintptr_t)a->B | (uintptr_t)a->D) & 15))
            CRCT_FAIL(CRCT_ERR_ARG, "crct_gemm_wgrad_grouped: alignment of problem %d", g);
        const int k_eff = (a->a_rows_dev && a->rows_hint > 0 && a->rows_hint < a->K) ? a->rows_hint : a->K;
        kb_eff[g] = (k_eff + BLOCK_K - 1) / BLOCK_K;
        mn[g] = ((a->M + BLOCK_M - 1) / BLOCK_M) * ((a->N + BN - 1) / BN);
        work += (long)mn[g] * kb_eff[g];
    }
    // Split-K per problem: tiles of (about) equal k-depth `kb_target`.  The depth is chosen by cost over a few candidates — waves on
    // the SMs x (deepest tile + the fixed cost of a tile: fill, fp32 red.add epilogue): a text layer's four problems are
    // 216 tiles of 108 k-blocks, i.e. 1.46 waves un-split (two tile-times on 108 CTAs: ncu showed 74 % tensor-pipe when active but
    // 51 % of elapsed, 40 SMs idle); split in two they are 432 tiles = 2.92 waves of half the depth (-20 %).
    constexpr long TILE_FIXED_KB = 16;           // flat between 8 and 24 (tools/section_times.py: 2.02-2.03 ms per step; 2.25 ms before)
    auto split_for = [&](int g, long target) {
        if (probs[g].split_k > 0) return (long)probs[g].split_k;
        long sp = (kb_eff[g] + target / 2) / target;
        const long kb_total = (probs[g].K + BLOCK_K - 1) / BLOCK_K;
        if (sp < 1) sp = 1;
        if (sp > kb_total) sp = kb_total;
        return sp;
    };
    long kb_target = 0, best_cost = -1;
    for (int w = 1; w <= 6; ++w) {
        long target = (work + (long)w * sms - 1) / ((long)w * sms);
        if (target < 4) target = 4;                  // pipeline depth
        long tiles = 0, deepest = 0;
        for (int g = 0; g < count; ++g) {
            const long sp = split_for(g, target);
            tiles += (long)mn[g] * sp;
            const long depth = (kb_eff[g] + sp - 1) / sp;
            if (depth > deepest) deepest = depth;
        }
        const long waves = (tiles + sms - 1) / sms;
        const long cost = waves * (deepest + TILE_FIXED_KB);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; kb_target = target; }
    }
    GroupParams gp;
    GroupMaps tm;
    memset(&gp, 0, sizeof(gp));
    gp.count = count;
    for (int g = 0; g < count; ++g) {
        const crct_gemm_t* a = &probs[g];
        const int kb_total = (a->K + BLOCK_K - 1) / BLOCK_K;
        int split = (int)split_for(g, kb_target);
        KParams& p = gp.p[g];
        p.M = a->M; p.N = a->N; p.K = a->K;
        p.num_n_tiles = (a->N + BN - 1) / BN;
        p.split_k = split;
        p.num_tiles = mn[g] * split;
        p.kb_total = kb_total;
        p.kb_per_split = (kb_total + split - 1) / split;
        p.D = a->D; p.ldd = a->ldd; p.accumulate = 1;
        p.a_lbo = BLOCK_K * 128; p.a_sbo = 1024; p.a_kstep = UMMA_K * 128;
        p.b_lbo = BLOCK_K * 128; p.b_sbo = 1024; p.b_kstep = UMMA_K * 128;
        p.a_rows_dev = a->a_rows_dev;
        p.tile_m = BLOCK_M;
        gp.tile_start[g + 1] = gp.tile_start[g] + p.num_tiles;
        if (int rc = make_tmap(&tm.a[g], a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, BLOCK_K)) return rc;
        if (int rc = make_tmap(&tm.b[g], a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb, 64, BLOCK_K)) return rc;
    }
    for (int g = count; g < MAXG; ++g) { tm.a[g] = tm.a[0]; tm.b[g] = tm.b[0]; gp.tile_start[g + 1] = gp.tile_start[count]; }
    const int total = gp.tile_start[count];
    int grid = total < sms ? total : sms;
    if (total > sms) {
        const int waves = (total + sms - 1) / sms;
        grid = (total + waves - 1) / waves;
    }
    auto kern = gemm_wgrad_grouped_kernel<BN>;
    static bool configured = false;
    if (!configured) {
        CRCT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
        configured = true;
    }
    CRCT_CUDA(crct_launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), Cfg<BN>::SMEM_BYTES, as_stream(stream), tm, gp));
    return CRCT_OK;
}
