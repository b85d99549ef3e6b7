// Shared device/host helpers for libcrct_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/crct_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcrct_b200 targets sm_100a (B200) only"
#endif

#define CRCT_API __attribute__((visibility("default")))

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, never throws across the C ABI)
// ---------------------------------------------------------------------------------------------
void crct_set_error(const char* fmt, ...);
#define CRCT_FAIL(code, ...) do { crct_set_error(__VA_ARGS__); return (code); } while (0)
#define CRCT_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    crct_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    return CRCT_ERR_CUDA; } } while (0)
#define CRCT_LAUNCH_CHECK() CRCT_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(crct_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int crct_num_sms();
// tcgen05 / TMEM / TMA attention (attention_tc.cu): single-tile sequences (Lq, Lk <= 128); attention.cu dispatches to it
bool crct_attn_tc_eligible(int dh, int Lq, int Lk, const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int backward);
int crct_attn_fwd_tc(const crct_attn_fwd_t* a, crct_stream_t s);
int crct_attn_bwd_tc(const crct_attn_bwd_t* a, crct_stream_t s);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched through crct_launch_pdl may have its CTAs scheduled — and run
// whatever precedes pdl_wait() — while the previous kernel of the stream is still draining its last wave; pdl_wait()
// returns once that kernel has completed and its writes are visible.  The ~700 kernels of a step form long dependent
// chains, so the launch latency and the prologue (barrier init, TMEM allocation, descriptor prefetch) of every hot kernel
// otherwise sit on the critical path.  Rules: a kernel launched this way calls pdl_wait() before its first global-memory
// access; pdl_launch_dependents() at its top lets ITS successor be scheduled early (the trigger fires only when every
// CTA of the grid has started, so a successor never takes SM resources from unscheduled CTAs of its predecessor).
// Captured into CUDA graphs as programmatic dependency edges.
// MEASURED (B200, train step B=80, profiles/r01_ab_pdl_s14.txt): 17.40 / 17.22 ms with the attribute against 17.11 / 17.10 ms
// without — the early-resident successors cost more than the launch gaps they hide when three streams already keep the SMs
// busy — so the attribute is OFF unless CRCT_PDL=1; without it the two instructions are no-ops.
// ---------------------------------------------------------------------------------------------
bool crct_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t crct_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = crct_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG, identical in forward and backward and never stored.
// One 32-bit hash serves the element pair (2i, 2i+1): element idx keeps iff its 16-bit half >= threshold, with
// threshold = round(p * 65536) (0 => keep everything); kept values are scaled by 1/(1-p)
// (torch.nn.Dropout semantics; reference call sites vilbert.py:315,377,422,465,506,553,596,642,649,734,741,1045).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t crct_hash_pair(uint64_t seed, uint64_t idx) {
    const uint64_t pair = idx >> 1;
    uint32_t h = (uint32_t)pair * 0x9E3779B1u + (uint32_t)seed;
    h ^= (uint32_t)(pair >> 32) * 0x85EBCA77u + (uint32_t)(seed >> 32) * 0x27D4EB2Fu;
    h ^= h >> 15; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ bool crct_keep(uint64_t seed, uint64_t idx, uint32_t threshold) {
    const uint32_t h = crct_hash_pair(seed, idx);
    return ((idx & 1) ? (h >> 16) : (h & 0xFFFFu)) >= threshold;
}
// decisions for idx and idx+1 (one hash when idx is even)
__host__ __device__ __forceinline__ void crct_keep2(uint64_t seed, uint64_t idx, uint32_t threshold, bool& k0, bool& k1) {
    const uint32_t h = crct_hash_pair(seed, idx);
    if ((idx & 1) == 0) {
        k0 = (h & 0xFFFFu) >= threshold;
        k1 = (h >> 16) >= threshold;
    } else {
        k0 = (h >> 16) >= threshold;
        k1 = (crct_hash_pair(seed, idx + 1) & 0xFFFFu) >= threshold;
    }
}
// 32-bit fast path (element counters below 2^32): same decisions as crct_keep, with the seed mixing hoisted
struct CrctDrop32 {
    uint32_t sm, sh, thr;
};
__host__ __device__ __forceinline__ CrctDrop32 crct_drop32(uint64_t seed, uint32_t thr) {
    CrctDrop32 d;
    d.sm = (uint32_t)seed;
    d.sh = (uint32_t)(seed >> 32) * 0x27D4EB2Fu;
    d.thr = thr;
    return d;
}
__host__ __device__ __forceinline__ uint32_t crct_hash_pair32(const CrctDrop32& d, uint32_t pair) {
    uint32_t h = pair * 0x9E3779B1u + d.sm;
    h ^= d.sh;
    h ^= h >> 15; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ bool crct_keep32(const CrctDrop32& d, uint32_t idx) {
    const uint32_t h = crct_hash_pair32(d, idx >> 1);
    return ((idx & 1u) ? (h >> 16) : (h & 0xFFFFu)) >= d.thr;
}
__host__ __device__ __forceinline__ void crct_keep2_32(const CrctDrop32& d, uint32_t idx, bool& k0, bool& k1) {
    const uint32_t h = crct_hash_pair32(d, idx >> 1);
    if ((idx & 1u) == 0u) {
        k0 = (h & 0xFFFFu) >= d.thr;
        k1 = (h >> 16) >= d.thr;
    } else {
        k0 = (h >> 16) >= d.thr;
        k1 = (crct_hash_pair32(d, (idx + 1u) >> 1) & 0xFFFFu) >= d.thr;
    }
}
static inline uint32_t crct_drop_threshold(float p) {
    if (p <= 0.f) return 0u;
    double t = (double)p * 65536.0 + 0.5;
    return t >= 65535.0 ? 0xFFFFu : (uint32_t)t;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below one bf16 ulp of the result): one MUFU.RCP + one
// MUFU.EX2 instead of erff()'s ~40-instruction branchy polynomial — the GELU epilogues are issue-bound otherwise.
// e = exp(-x^2/2) is shared between erf(x/sqrt2) = 1 - poly(t) e and the Gaussian pdf of the derivative.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// hq = Phi(-|x|) = erfc(|x|/sqrt 2)/2 (A&S 7.1.26 with the 1/2 folded into the coefficients), e = exp(-x^2/2).
// Two MUFU operations per element (rcp, ex2); gelu and gelu' share them.
__device__ __forceinline__ void gelu_parts(float x, float& hq, float& e) {
    const float t = fast_rcp(fmaf(0.3275911f * 0.70710678118654752f, fabsf(x), 1.0f));
    e = fast_exp2(x * x * (-0.5f * 1.4426950408889634f));
    float poly = fmaf(0.5f * 1.061405429f, t, -0.5f * 1.453152027f);
    poly = fmaf(poly, t, 0.5f * 1.421413741f);
    poly = fmaf(poly, t, -0.5f * 0.284496736f);
    poly = fmaf(poly, t, 0.5f * 0.254829592f);
    hq = poly * t * e;
}
// x Phi(x) = max(x, 0) - |x| Phi(-|x|)                               vilbert.py:111-117 (exact erf form)
__device__ __forceinline__ float gelu_from_parts(float x, float hq) { return fmaf(-fabsf(x), hq, fmaxf(x, 0.f)); }
// Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad_from_parts(float x, float hq, float e) {
    return fmaf(x * 0.3989422804014327f, e, x >= 0.f ? 1.0f - hq : hq);
}
__device__ __forceinline__ float gelu_f(float x) {
    float hq, e;
    gelu_parts(x, hq, e);
    return gelu_from_parts(x, hq);
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    float hq, e;
    gelu_parts(x, hq, e);
    return gelu_grad_from_parts(x, hq, e);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(t);
}
// 8 consecutive bf16 <-> 8 floats through one 16-byte access
__device__ __forceinline__ void load8_bf16(const bf16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8_bf16(bf16* p, const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void load8_f32(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// dropout on 8 consecutive elements starting at element counter e0, e0 a MULTIPLE OF 8 (every caller: row * width + column
// chunk, width % 8 == 0): four pair hashes with the index/seed high words mixed once — the same decisions as crct_keep
// element by element, without its odd-index path and 64-bit arithmetic per pair.  (ncu on the RES epilogue: the hash was
// ~11 of ~20 instructions per element of an issue-bound epilogue.)
__host__ __device__ __forceinline__ void dropout8(float (&f)[8], uint64_t seed, uint64_t e0, uint32_t thr, float scale) {
    const uint64_t pair = e0 >> 1;
    const uint32_t lo = (uint32_t)pair;
    const uint32_t mix = (uint32_t)(pair >> 32) * 0x85EBCA77u + (uint32_t)(seed >> 32) * 0x27D4EB2Fu;
    const uint32_t thr_hi = thr << 16;                       // (h >> 16) >= thr  <=>  h >= thr << 16
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t h = (lo + (uint32_t)j) * 0x9E3779B1u + (uint32_t)seed;
        h ^= mix;
        h ^= h >> 15; h *= 0x85EBCA6Bu;
        h ^= h >> 13; h *= 0xC2B2AE35u;
        h ^= h >> 16;
        f[2 * j] = (h & 0xFFFFu) >= thr ? f[2 * j] * scale : 0.f;
        f[2 * j + 1] = h >= thr_hi ? f[2 * j + 1] * scale : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// PTX: mbarrier / TMA / tcgen05 (forms as in the CUTLASS sm_100 headers; validated by ptxas)
// ---------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base+i), v[j] = column (col+j)
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// ---- cta_group::2 (CTA pair) forms
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are signalled on the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// commit: arrive (once the pair's MMAs retire) on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}" :: "r"(bar), "r"(cta) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tc_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
}  // namespace ptx
#endif  // __CUDACC__
