// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16, K = 16) on one SM for the shapes / operand sources the GEMM
// kernel uses.  One CTA per SM, one issuing thread, zeroed operands.  Prints cycles per MMA against the N/2 floor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/mma_rate tools/mma_rate.cu && tools/build/mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 ss N=n0 one accumulator | 1 ts N=n0 one accumulator | 2 ts alternating (n0 at col 0, n1 at col n0) | 3 ss alternating
// sw: 0 = SWIZZLE_64B tiles (rows of 64 B, K = 16 halves), 1 = SWIZZLE_128B tiles (rows of 128 B)
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int n0, int n1, int sw, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 16384);
        const uint64_t da = sw ? desc_sw128(a_s) : desc_sw64(a_s), db = sw ? desc_sw128(b_s) : desc_sw64(b_s);
        const uint32_t i0 = idesc_bf16(128, n0), i1 = idesc_bf16(128, n1 > 0 ? n1 : 16);
        const uint32_t ta = tm + 448;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint64_t ko = (uint64_t)(k * 2);
                if (mode == 0) mma_ss(tm, da + ko, db + ko, i0, 1);
                else if (mode == 1) mma_ts(tm, ta + k * 8, db + ko, i0, 1);
                else if (mode == 2) { mma_ts(tm, ta + k * 8, db + ko, i0, 1); mma_ts(tm + n0, ta + k * 8, db + ko + (uint64_t)((n0 * (sw ? 128 : 64)) >> 4), i1, 1); }
                else if (mode == 3) { mma_ss(tm, da + ko, db + ko, i0, 1); mma_ss(tm + n0, da + ko, db + ko + (uint64_t)((n0 * (sw ? 128 : 64)) >> 4), i1, 1); }
                else {
                    // modes 4 / 5: the GEMM kernel's K-block (3 passes x 2 K-steps x 2 MMAs); mode 5 commits after every K-block
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t tap = ta + (pass == 0 ? 16 : 0) + k * 8;
                        const uint64_t dbp = db + (pass == 1 ? 1216 : 0) + ko;
                        mma_ts(tm, tap, dbp, i0, 1);
                        mma_ts(tm + n0, tap, dbp + (uint64_t)((n0 * 64) >> 4), i1, 1);
                    }
                }
            }
            if (mode == 5)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

int main() {
    long long* out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    struct Case { const char* name; int mode, n0, n1, sw; };
    const Case cases[] = {
        {"ss N=256 sw128", 0, 256, 0, 1}, {"ss N=256 sw64", 0, 256, 0, 0}, {"ss N=160 sw64", 0, 160, 0, 0}, {"ss N=128 sw64", 0, 128, 0, 0},
        {"ts N=256 sw128", 1, 256, 0, 1}, {"ts N=256 sw64", 1, 256, 0, 0}, {"ts N=160 sw64", 1, 160, 0, 0}, {"ts N=128 sw64", 1, 128, 0, 0},
        {"ts N=160+144 sw64", 2, 160, 144, 0}, {"ts N=160+144 sw128", 2, 160, 144, 1}, {"ss N=160+144 sw64", 3, 160, 144, 0},
        {"ts N=256+48 sw64", 2, 256, 48, 0}, {"ts N=64 sw64", 1, 64, 0, 0},
        {"kernel K-block, no commit", 4, 160, 144, 0}, {"kernel K-block + commit", 5, 160, 144, 0},
        {"ts N=240 sw64", 1, 240, 0, 0}, {"ts N=224 sw64", 1, 224, 0, 0}, {"ts N=208 sw64", 1, 208, 0, 0}, {"ts N=192 sw64", 1, 192, 0, 0},
    };
    const int iters = 4000;
    for (const Case& c : cases) {
        for (int rep = 0; rep < 2; ++rep) rate_kernel<<<148, 128, 100 * 1024>>>(c.mode, c.n0, c.n1, c.sw, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
        long long cyc;
        cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        const int per_iter = c.mode >= 4 ? 12 : 2 * (c.mode >= 2 ? 2 : 1);
        const double per_group = (double)cyc / iters / (c.mode >= 4 ? 6 : 2);   // cycles per K = 16 step (one or two MMAs)
        const double floor_ = (c.n0 + c.n1) / 2.0;
        printf("%-22s %8.1f cycles per K=16 step (%d MMAs/iter)  floor %.0f  -> %.0f %% of the tensor-pipe rate\n", c.name, per_group,
               per_iter, floor_, 100.0 * floor_ / per_group);
    }
    return 0;
}
