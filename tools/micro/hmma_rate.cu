// Microbenchmark: legacy mma.sync m16n8k16 bf16 issue rate on sm_100a (how far can the warp-level attention go?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
    float c[8][4];
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    unsigned a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    for (int warps : {4, 8, 16, 32}) {
        int iters = 4096;
        k<<<148 * 2, warps * 32 / 2>>>(out, 16); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<148 * 2, warps * 32 / 2>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = 148.0 * warps * iters * 8 * 4096.0;
        printf("warps/SM=%d: %.1f TFLOP/s (%.1f FLOP/clk/SM at 1.965 GHz)\n", warps, flop / ms / 1e9, flop / ms / 1e9 * 1e12 / 148 / 1.965e9);
    }
    return 0;
}
