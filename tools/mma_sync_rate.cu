// Issue rate of the warp-level mma.sync.m16n8k8 TF32 (the form a register-resident tiny-MLP kernel can use) on sm_100a:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate tools/mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
    float c[4][4];
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0.f;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 20000;
        k<<<148, warps * 32>>>(out, 100);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double mmas = 148.0 * warps * iters * 4, flops = mmas * 16 * 8 * 8 * 2;
        printf("warps/SM %2d: %.3f ms, %.1f TFLOP/s tf32, %.2f cycles per mma per SM at 1.9 GHz\n", warps, ms, flops / ms / 1e9,
               ms * 1e-3 * 1.9e9 / (warps * iters * 4.0));
    }
    return 0;
}
