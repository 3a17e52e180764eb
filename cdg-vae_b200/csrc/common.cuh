// Shared device/host helpers for libcdgvae_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "cdgvae.h"

namespace cdg {

void set_error(const char* fmt, ...);

// Experiment switches (tile variants measured slower, alternative kernels kept for A/B timing, timing probes whose results
// are INVALID) exist only in builds made with -DCDG_EXPERIMENTS; the shipped library has every one of them fixed at its
// default and does not read the environment for them.
#ifdef CDG_EXPERIMENTS
#include <stdlib.h>
static inline int exp_switch(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
#else
#define exp_switch(name, dflt) (dflt)
#endif

#define CDG_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            cdg::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,          \
                           cudaGetErrorString(_e));                                       \
            return CDG_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

extern long long g_launches;   // kernels launched by this library (host-side count)
#define CDG_CHECK_LAUNCH()                                                                \
    do {                                                                                  \
        ++cdg::g_launches;                                                                \
        CDG_CHECK_CUDA(cudaGetLastError());                                               \
    } while (0)

#define CDG_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            cdg::set_error(__VA_ARGS__);                                                  \
            return CDG_ERR_INVALID;                                                       \
        }                                                                                 \
    } while (0)

#define CDG_TRY(expr)                                                                     \
    do {                                                                                  \
        int _r = (expr);                                                                  \
        if (_r != CDG_OK) return _r;                                                      \
    } while (0)

constexpr int kNumSMs = 148;
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

enum Epi : int {
    EPI_NONE = 0,       // C = acc (or += when accumulate)
    EPI_BIAS = 1,       // C = acc + bias[n]
    EPI_BIAS_ACT = 2,   // C = act(acc + bias[n])
    EPI_MUL_DACT = 3,   // C = acc * act'(aux[m,n])   (aux holds the post-activation value)
    EPI_RECON = 4,      // pre = acc + bias[n]; xhat = tanh(pre); loss += 0.5 (xhat - x)^2; C = (xhat - x)(1 - xhat^2)/batch
};

__device__ __forceinline__ float act_fwd(float v, int act) {
    // ELU(alpha=1): modules/model.py:221 (nn.ELU);  ReLU: tabular/modules/model.py:373
    if (act == CDG_ACT_ELU) return v > 0.f ? v : expf(v) - 1.f;   // ATen's ELU computes exp(x) - 1, not expm1
    return v > 0.f ? v : 0.f;
}
// derivative expressed through the post-activation value h
__device__ __forceinline__ float act_bwd_from_out(float h, int act) {
    if (act == CDG_ACT_ELU) return h > 0.f ? 1.f : h + 1.f;
    return h > 0.f ? 1.f : 0.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024); result valid on thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        v = lane < nw ? smem[lane] : T(0);
        v = warp_sum(v);
    }
    return v;
}

// ---- optional per-category device timing (cudaEvents on the caller's stream) ----
enum ProfCat { PROF_ENC0_FWD = 0, PROF_DEC2_FWD, PROF_DEC2_DGRAD, PROF_DEC2_WGRAD, PROF_ENC0_WGRAD, PROF_GEMM_OTHER,
               PROF_LATENT, PROF_RECON, PROF_MISC, PROF_NCAT };
struct Profiler {
    bool enabled = false;
    static constexpr int kMax = 4096;
    cudaEvent_t ev[kMax];
    int cat[kMax];
    int n = 0, created = 0;
    // mark the start of a region of category c; the region ends at the next mark
    void mark(int c, cudaStream_t s) {
        if (!enabled || n >= kMax) return;
        if (n >= created) { if (cudaEventCreate(&ev[n]) != cudaSuccess) return; created = n + 1; }
        cudaEventRecord(ev[n], s);
        cat[n++] = c;
    }
};

// split-K occupancy target of the tcgen05 GEMM planner for the calling thread (gemm_tc.cu): a caller that runs several
// independent chains on streams of its own asks each small GEMM to fill only its share of the chip
void gemm_tc_set_sm_budget(int sms);

// ---- generic strided GEMM interface (implemented in gemm_simt.cu / gemm_tc.cu) ----
struct GemmDesc {
    const float* A; int64_t sa_m, sa_k;
    const float* B; int64_t sb_n, sb_k;
    float* C; int64_t ldc;
    int64_t M, N, K;
    int epi = EPI_NONE;
    int act = CDG_ACT_ELU;
    const float* bias = nullptr;     // [N]
    const float* aux = nullptr;      // [M, ld_aux]
    int64_t ld_aux = 0;
    int accumulate = 0;              // C += result (EPI_NONE only)
    // EPI_RECON (tensor-core kernel only): the reconstruction head fused into the last decoder Linear
    const float* recon_x = nullptr;  // [M, ld_x] target, same columns as C
    int64_t ld_x = 0;
    float* recon_xhat = nullptr;     // optional tanh output, laid out like C
    double* recon_acc = nullptr;     // += sum 0.5 (xhat - x)^2
    float inv_batch = 0.f;
    // Implicit-GEMM convolution (tensor-core kernel only): when conv_C > 0, A is an NHWC activation [conv_B, conv_H,
    // conv_W, conv_C] (contiguous), M = conv_B*conv_H*conv_W output pixels of a stride-1 "same" conv_k x conv_k
    // convolution, K = conv_k^2 * conv_C in (kh, kw, c) order; the halo is zero-filled by TMA.  sa_m / sa_k are ignored.
    int conv_C = 0, conv_H = 0, conv_W = 0, conv_k = 0;
    int64_t conv_B = 0;
    // Pre-split B operand (bf16x3 mode only): B(n,k) = b_hi16[n*ld_b16 + k] + b_lo16[n*ld_b16 + k] as bf16 pairs made once
    // per step by launch_split_bf16 (weights); the kernel then streams ready-made tiles and converts only A.
    const void* b_hi16 = nullptr;
    const void* b_lo16 = nullptr;
    int64_t ld_b16 = 0;
    // Same for A, usable only where the planner swaps the operands (N > 304 >= M: the first-layer weight gradient, whose
    // narrow operand is dY): A(m,k) = a_hi16[m*ld_a16 + k] + a_lo16[...].
    const void* a_hi16 = nullptr;
    const void* a_lo16 = nullptr;
    int64_t ld_a16 = 0;
    // Tensor-core kernel, no operand swap: the LAST column n = N-1 of the result goes to extra_col[m] instead of C[m, N-1].
    // A weight-gradient GEMM whose pre-split B carries a row of ones there yields the bias gradient (the column sum of dY)
    // for free, instead of a separate pass over dY.
    float* extra_col = nullptr;
    // bf16 (hi, lo) planes of the fp32 result for a following GEMM of the pre-split kernel (gemm_ps.cu): [M][ld_out16],
    // columns >= N zero except column N = 1 when out_ones (the consumer's weight planes then carry its bias in that column).
    // Kernels that can, write them from their epilogue; for the others the dispatcher adds a split pass over C (finish_planes).
    void* out_hi16 = nullptr;
    void* out_lo16 = nullptr;
    int64_t ld_out16 = 0;
    int out_ones = 0;
};
extern thread_local bool tl_planes_done;     // set by a kernel launch that wrote g.out_hi16 / out_lo16 itself
int finish_planes(const GemmDesc& g, cudaStream_t s);   // split pass over C when the routed kernel did not emit the planes

// W[rows, cols] (row stride ld) -> hi / lo bf16 copies; transpose != 0 writes them as [cols][rows] (row stride ld16 either way)
int launch_split_bf16(const float* W, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo, int64_t ld16, int transpose,
                      cudaStream_t s, int ones_row = 0);   // ones_row (transposed only): append output row `cols` = 1.0

// row-major planes [rows][ld16] with 16-byte accesses; column `cols` (if ld16 > cols) = extra[r], or 1 when `ones`, else 0
int launch_split_rows(const float* W, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo, int64_t ld16, const float* extra,
                      int ones, cudaStream_t s);

// backward of a Linear with a tiny input (in <= 4) in one pass over dY [rows, out]: dW += dY^T x, db += colsum(dY), dX = dY W
int launch_tiny_in_bwd(const float* dY, int64_t ld_dy, const float* x, int64_t ld_x, const float* W, float* dW, float* db, float* dX,
                       int64_t ld_dx, int64_t rows, int out, int in, cudaStream_t s);

int gemm_simt(const GemmDesc& g, cudaStream_t s);
int gemm_skinny(const GemmDesc& g, cudaStream_t s);   // tiny-extent shapes; CDG_ERR_UNSUPPORTED otherwise
// tcgen05 path; returns CDG_ERR_UNSUPPORTED when the shape/layout does not fit, so that the
// dispatcher can route it to the SIMT kernel.
int gemm_tc(const GemmDesc& g, int passes, void* workspace, int64_t workspace_bytes, cudaStream_t s);
// CTA-pair bf16x3 kernel for operands that BOTH arrive as bf16 (hi, lo) planes (gemm_ps.cu): needs a_hi16 / a_lo16 and
// b_hi16 / b_lo16 (no operand swap); out_hi / out_lo (optional) receive the planes of the result for the next GEMM.
int gemm_ps(const GemmDesc& g, cudaStream_t s);
// long contractions on planes, accumulated into C with fp32 atomics per <= 2048-long segment (gemm_pk.cu); mn_major: the
// planes are stored with the contraction index outermost ([K][M], [K][N]: weight gradients)
int gemm_pk(const GemmDesc& g, int mn_major, cudaStream_t s);
bool gemm_tc_can(const GemmDesc& g);   // would gemm_tc accept this contraction (without split-K)?
int gemm_dispatch(int mode, const GemmDesc& g, void* workspace, int64_t workspace_bytes, cudaStream_t s);

int launch_bias_act(float* C, int64_t ldc, int64_t M, int64_t N, const float* bias, int epi, int act,
                    const float* aux, int64_t ld_aux, cudaStream_t s, void* out_hi = nullptr, void* out_lo = nullptr,
                    int64_t ld_out16 = 0, int out_ones = 0);
int launch_colsum(const float* G, int64_t ld, int64_t M, int64_t N, float* out, cudaStream_t s);

}  // namespace cdg
