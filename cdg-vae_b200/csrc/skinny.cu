// Memory-bound kernels for the contractions with one tiny extent, where neither a tensor-core tile nor a
// 128x128 SIMT tile makes sense (all of them stream one [batch, 300] activation once):
//   smallk_elem   K <= 8   decoder first Linear (K = 1, 2: modules/model.py:245) and the dgrad of the
//                          encoder head (K = 2d = 8: model.py:224)
//   rowdot        N <= 8   encoder head forward (N = 2d) and the dgrad of the decoder first Linear (N = 1, 2)
//   skinny_wgrad  min(M,N) <= 8, contraction over the batch: wgrad of those two layers
#include "common.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace cdg {

template <bool VEC>
__global__ void __launch_bounds__(256) smallk_elem_kernel(GemmDesc g) {
    const int K = (int)g.K;
    if (VEC) {
        // 4 consecutive outputs per thread, 128-bit stores (N % 4 == 0, 16-byte aligned rows)
        const int64_t n4 = g.N / 4, total = g.M * n4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t m = i / n4, n = (i - m * n4) * 4;
            float s[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = 0; k < K; ++k) {
                const float a = g.A[m * g.sa_m + k * g.sa_k];
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = fmaf(a, __ldg(g.B + (n + e) * g.sb_n + k * g.sb_k), s[e]);
            }
            float* c = g.C + m * g.ldc + n;
            if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_ACT) {
                const float4 b = *reinterpret_cast<const float4*>(g.bias + n);
                s[0] += b.x; s[1] += b.y; s[2] += b.z; s[3] += b.w;
            }
            if (g.epi == EPI_BIAS_ACT) {
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = act_fwd(s[e], g.act);
            }
            if (g.epi == EPI_MUL_DACT) {
                const float4 h = *reinterpret_cast<const float4*>(g.aux + m * g.ld_aux + n);
                s[0] *= act_bwd_from_out(h.x, g.act); s[1] *= act_bwd_from_out(h.y, g.act);
                s[2] *= act_bwd_from_out(h.z, g.act); s[3] *= act_bwd_from_out(h.w, g.act);
            }
            if (g.accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(c);
                s[0] += o.x; s[1] += o.y; s[2] += o.z; s[3] += o.w;
            }
            *reinterpret_cast<float4*>(c) = make_float4(s[0], s[1], s[2], s[3]);
        }
        return;
    }
    const int64_t total = g.M * g.N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / g.N, n = i - m * g.N;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(g.A[m * g.sa_m + k * g.sa_k], __ldg(g.B + n * g.sb_n + k * g.sb_k), s);
        float* c = g.C + m * g.ldc + n;
        if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_ACT) s += g.bias[n];
        if (g.epi == EPI_BIAS_ACT) s = act_fwd(s, g.act);
        if (g.epi == EPI_MUL_DACT) s *= act_bwd_from_out(g.aux[m * g.ld_aux + n], g.act);
        if (g.accumulate) s += *c;
        *c = s;
    }
}

// one warp per output row; N <= 8 accumulators per lane; A rows are contiguous (sa_k == 1)
__global__ void __launch_bounds__(256) rowdot_kernel(GemmDesc g) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int N = (int)g.N;
    for (int64_t m = warp0; m < g.M; m += nwarps) {
        float acc[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[n] = 0.f;
        const float* a = g.A + m * g.sa_m;
        for (int64_t k = lane; k < g.K; k += 32) {
            const float av = a[k];
#pragma unroll
            for (int n = 0; n < 8; ++n)
                if (n < N) acc[n] = fmaf(av, __ldg(g.B + n * g.sb_n + k * g.sb_k), acc[n]);
        }
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (n < N) acc[n] = warp_sum(acc[n]);
        if (lane < N) {
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                if (n == lane) s = acc[n];
            float* c = g.C + m * g.ldc + lane;
            if (g.epi == EPI_BIAS) s += g.bias[lane];
            if (g.accumulate) s += *c;
            *c = s;
        }
    }
}

// ---- row-per-warp variants (the hot shapes: [batch, 300] activations with K or N <= 8) ----------------------------
// The element-indexed kernels above pay a 64-bit division per 4 outputs and re-read the tiny operand through the
// read-only cache for every element: ncu launch lists showed them at 1.2-1.4 TB/s.  Here the tiny operand (and the bias)
// sits in shared memory, a warp owns one batch row at a time, and every global access is a 16-byte piece of one
// contiguous 1,200-byte row.
// 4 adjacent results -> 4 bf16 of the hi plane and 4 of the lo plane (8-byte stores)
__device__ __forceinline__ void planes4(const GemmDesc& g, int64_t m, int q, const float* v) {
    uint32_t h2[2], l2[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * k] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * k + 1] - __bfloat162float(h1));
        h2[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l2[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out_hi16) + m * g.ld_out16 + 4 * q) = make_uint2(h2[0], h2[1]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out_lo16) + m * g.ld_out16 + 4 * q) = make_uint2(l2[0], l2[1]);
}

__global__ void __launch_bounds__(256) smallk_rows_kernel(GemmDesc g) {
    extern __shared__ __align__(16) float sk_sm[];        // [K][N] weights, then [N] bias
    const int K = (int)g.K, N = (int)g.N, n4 = N >> 2;
    float* ws = sk_sm;
    float* bs = sk_sm + K * N;
    const bool has_bias = g.epi == EPI_BIAS || g.epi == EPI_BIAS_ACT;
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int k = i / N, n = i - k * N;
        ws[i] = g.B[n * g.sb_n + k * g.sb_k];
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) bs[i] = has_bias ? g.bias[i] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4* ws4 = reinterpret_cast<const float4*>(ws);
    const float4* bs4 = reinterpret_cast<const float4*>(bs);
    for (int64_t m = warp0; m < g.M; m += nwarps) {
        float a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = k < K ? __ldg(g.A + m * g.sa_m + k * g.sa_k) : 0.f;
        float4* crow = reinterpret_cast<float4*>(g.C + m * g.ldc);
        const float4* hrow = g.epi == EPI_MUL_DACT ? reinterpret_cast<const float4*>(g.aux + m * g.ld_aux) : nullptr;
        for (int q = lane; q < n4; q += 32) {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < K) {
                    const float4 w = ws4[k * n4 + q];
                    s[0] = fmaf(a[k], w.x, s[0]); s[1] = fmaf(a[k], w.y, s[1]);
                    s[2] = fmaf(a[k], w.z, s[2]); s[3] = fmaf(a[k], w.w, s[3]);
                }
            }
            if (has_bias) {
                const float4 b = bs4[q];
                s[0] += b.x; s[1] += b.y; s[2] += b.z; s[3] += b.w;
            }
            if (g.epi == EPI_BIAS_ACT) {
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = act_fwd(s[e], g.act);
            }
            if (hrow) {
                const float4 h = __ldg(hrow + q);
                s[0] *= act_bwd_from_out(h.x, g.act); s[1] *= act_bwd_from_out(h.y, g.act);
                s[2] *= act_bwd_from_out(h.z, g.act); s[3] *= act_bwd_from_out(h.w, g.act);
            }
            if (g.accumulate) {
                const float4 o = crow[q];
                s[0] += o.x; s[1] += o.y; s[2] += o.z; s[3] += o.w;
            }
            crow[q] = make_float4(s[0], s[1], s[2], s[3]);
            if (g.out_hi16) planes4(g, m, q, s);
        }
        if (g.out_hi16) {                                  // padding groups: column N = 1 when asked, the rest 0
            for (int q = n4 + lane; q < (int)(g.ld_out16 >> 2); q += 32) {
                const float z[4] = {(q == n4 && g.out_ones) ? 1.f : 0.f, 0.f, 0.f, 0.f};
                planes4(g, m, q, z);
            }
        }
    }
}

// N <= 8 outputs per row, contraction over a contiguous row of K = 4 * k4 floats; B ([N][K]) in shared memory
__global__ void __launch_bounds__(256) rowdot_rows_kernel(GemmDesc g) {
    extern __shared__ __align__(16) float rd_sm[];        // [N][K]
    const int K = (int)g.K, N = (int)g.N, k4 = K >> 2;
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i - n * K;
        rd_sm[i] = g.B[n * g.sb_n + k * g.sb_k];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4* b4 = reinterpret_cast<const float4*>(rd_sm);
    for (int64_t m = warp0; m < g.M; m += nwarps) {
        float acc[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[n] = 0.f;
        const float4* arow = reinterpret_cast<const float4*>(g.A + m * g.sa_m);
        for (int q = lane; q < k4; q += 32) {
            const float4 av = __ldg(arow + q);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (n < N) {
                    const float4 w = b4[n * k4 + q];
                    acc[n] = fmaf(av.x, w.x, fmaf(av.y, w.y, fmaf(av.z, w.z, fmaf(av.w, w.w, acc[n]))));
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (n < N) acc[n] = warp_sum(acc[n]);
        if (lane < N) {
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n)
                if (n == lane) s = acc[n];
            float* c = g.C + m * g.ldc + lane;
            if (g.epi == EPI_BIAS) s += g.bias[lane];
            if (g.accumulate) s += *c;
            *c = s;
        }
    }
}

// C[m*ldc + n] += sum_k A[m + k*sa_k] * B[n + k*sb_k]; `wide_is_m` tells which side has the long extent
struct SkinnyW {
    const float* wide; int64_t wide_ld; const float* narrow; int64_t narrow_ld;
    float* C; int64_t c_sw, c_sr;          // C[w*c_sw + r*c_sr]
    int64_t W, K; int R; int k_chunk;
};
__global__ void __launch_bounds__(128) skinny_wgrad_kernel(SkinnyW a) {
    const int64_t w = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t kbeg = (int64_t)blockIdx.y * a.k_chunk, kend = min(a.K, kbeg + (int64_t)a.k_chunk);
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    if (w < a.W) {
        int64_t k = kbeg;
        for (; k + 4 <= kend; k += 4) {                 // four independent rows in flight per thread
            float wv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wv[u] = a.wide[(k + u) * a.wide_ld + w];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r < a.R) acc[r] = fmaf(wv[u], __ldg(a.narrow + (k + u) * a.narrow_ld + r), acc[r]);
        }
        for (; k < kend; ++k) {
            const float wv = a.wide[k * a.wide_ld + w];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < a.R) acc[r] = fmaf(wv, __ldg(a.narrow + k * a.narrow_ld + r), acc[r]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < a.R) atomicAdd(a.C + w * a.c_sw + r * a.c_sr, acc[r]);
    }
}

// ---- backward of a Linear with a TINY input (the decoders' first layer, in = 1 .. 4) in ONE pass over dY ---------------------
//   dW[j, i] += sum_b dY[b, j] x[b, i]      db[j] += sum_b dY[b, j]      dX[b, i] = sum_j dY[b, j] W[j, i]
// Round 1 ran three kernels here (skinny_wgrad, colsum, rowdot: 89 + 25 + 68 us per decoder at 131,072 rows), each reading
// the same [batch, 300] matrix.  A warp owns a batch row at a time: lane q holds the 16-byte pieces q, q + 32, q + 64 of the
// row, accumulates its slice of dW / db in registers over all rows it sees, and the row's dX is a warp sum.
constexpr int TIB_MAX_IN = 4, TIB_PIECES = 3;          // in <= 4, out <= 384 (3 x 32 pieces of 4 columns)
struct TinyInBwd {
    const float* dY; int64_t ld_dy;
    const float* x; int64_t ld_x;
    const float* W;                      // [out][in]
    float* dW; float* db;                // accumulated into
    float* dX; int64_t ld_dx;
    int64_t rows; int out, in;
};
template <int IN>
__global__ void __launch_bounds__(256) tiny_in_bwd_kernel(TinyInBwd a) {
    __shared__ float red[(IN + 1) * 384];
    const int lane = threadIdx.x & 31;
    const int n4 = a.out >> 2;
    float w[TIB_PIECES][4][IN], gw[TIB_PIECES][4][IN], gb[TIB_PIECES][4];
#pragma unroll
    for (int p = 0; p < TIB_PIECES; ++p)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            gb[p][e] = 0.f;
#pragma unroll
            for (int i = 0; i < IN; ++i) {
                const int j = 4 * (lane + 32 * p) + e;
                w[p][e][i] = lane + 32 * p < n4 ? a.W[(int64_t)j * IN + i] : 0.f;
                gw[p][e][i] = 0.f;
            }
        }
    for (int i = threadIdx.x; i < (IN + 1) * 384; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int R = 2;                                   // rows in flight per warp
    for (int64_t b0 = warp0; b0 < a.rows; b0 += R * nwarps) {
        float4 g[R][TIB_PIECES];
        float xv[R][IN];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t b = b0 + r * nwarps;
            const bool ok = b < a.rows;
            const float4* row = reinterpret_cast<const float4*>(a.dY + (ok ? b : 0) * a.ld_dy);
#pragma unroll
            for (int p = 0; p < TIB_PIECES; ++p)
                g[r][p] = (ok && lane + 32 * p < n4) ? __ldg(row + lane + 32 * p) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < IN; ++i) xv[r][i] = ok ? __ldg(a.x + b * a.ld_x + i) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t b = b0 + r * nwarps;
            float dx[IN];
#pragma unroll
            for (int i = 0; i < IN; ++i) dx[i] = 0.f;
#pragma unroll
            for (int p = 0; p < TIB_PIECES; ++p) {
                const float gv[4] = {g[r][p].x, g[r][p].y, g[r][p].z, g[r][p].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    gb[p][e] += gv[e];
#pragma unroll
                    for (int i = 0; i < IN; ++i) {
                        gw[p][e][i] = fmaf(gv[e], xv[r][i], gw[p][e][i]);
                        dx[i] = fmaf(gv[e], w[p][e][i], dx[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < IN; ++i) {
                const float t = warp_sum(dx[i]);
                if (lane == 0 && b < a.rows) a.dX[b * a.ld_dx + i] = t;
            }
        }
    }
    // block reduction of the register slices through shared memory, then one atomic per value and block
#pragma unroll
    for (int p = 0; p < TIB_PIECES; ++p)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * (lane + 32 * p) + e;
            if (lane + 32 * p < n4) {
                atomicAdd(&red[j], gb[p][e]);
#pragma unroll
                for (int i = 0; i < IN; ++i) atomicAdd(&red[(1 + i) * 384 + j], gw[p][e][i]);
            }
        }
    __syncthreads();
    for (int j = threadIdx.x; j < a.out; j += blockDim.x) {
        atomicAdd(a.db + j, red[j]);
        for (int i = 0; i < IN; ++i) atomicAdd(a.dW + (int64_t)j * IN + i, red[(1 + i) * 384 + j]);
    }
}
// CDG_ERR_UNSUPPORTED when the shape does not fit (the caller then takes the three separate kernels)
int launch_tiny_in_bwd(const float* dY, int64_t ld_dy, const float* x, int64_t ld_x, const float* W, float* dW, float* db, float* dX,
                       int64_t ld_dx, int64_t rows, int out, int in, cudaStream_t s) {
    if (in < 1 || in > TIB_MAX_IN || out % 4 != 0 || out > 4 * 32 * TIB_PIECES || ld_dy % 4 != 0 || (((uintptr_t)dY) & 15) != 0 || rows < 1)
        return CDG_ERR_UNSUPPORTED;
    TinyInBwd a{dY, ld_dy, x, ld_x, W, dW, db, dX, ld_dx, rows, out, in};
    const int blocks = (int)imin64((rows + 15) / 16, kNumSMs * 4);
    if (in == 1) tiny_in_bwd_kernel<1><<<blocks, 256, 0, s>>>(a);
    else if (in == 2) tiny_in_bwd_kernel<2><<<blocks, 256, 0, s>>>(a);
    else if (in == 3) tiny_in_bwd_kernel<3><<<blocks, 256, 0, s>>>(a);
    else tiny_in_bwd_kernel<4><<<blocks, 256, 0, s>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// returns CDG_ERR_UNSUPPORTED when the contraction is not one of the skinny shapes
int gemm_skinny(const GemmDesc& g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return CDG_OK;
    if (g.K <= 8) {
        const bool vec = g.N % 4 == 0 && g.ldc % 4 == 0 && (((uintptr_t)g.C) & 15) == 0 &&
                         (!(g.epi == EPI_BIAS || g.epi == EPI_BIAS_ACT) || (((uintptr_t)g.bias) & 15) == 0) &&
                         (g.epi != EPI_MUL_DACT || (g.ld_aux % 4 == 0 && (((uintptr_t)g.aux) & 15) == 0));
        static int rows_ok = -1;           // CDG_SKINNY_ROWS=0 restores the element-indexed kernels
        if (rows_ok < 0) rows_ok = exp_switch("CDG_SKINNY_ROWS", 1) != 0;
        if (rows_ok && vec && g.M >= 64 && g.K * g.N <= 8192 && g.N >= 128) {
            const size_t smem = sizeof(float) * (size_t)(g.K * g.N + g.N);
            const int blocks = (int)imin64((g.M + 7) / 8, kNumSMs * 8);
            GemmDesc gg = g;
            const bool planes = g.out_hi16 && g.out_lo16 && g.ld_out16 % 4 == 0 && g.ld_out16 >= g.N + (g.out_ones ? 1 : 0) &&
                                (((uintptr_t)g.out_hi16 | (uintptr_t)g.out_lo16) & 7) == 0;
            if (!planes) gg.out_hi16 = gg.out_lo16 = nullptr;      // (the dispatcher then adds the split pass)
            smallk_rows_kernel<<<blocks, 256, smem, s>>>(gg);
            CDG_CHECK_LAUNCH();
            if (planes) tl_planes_done = true;
            return CDG_OK;
        }
        const int64_t total = vec ? g.M * g.N / 4 : g.M * g.N;
        const int blocks = (int)imin64((total + 255) / 256, kNumSMs * 16);
        if (vec) smallk_elem_kernel<true><<<blocks, 256, 0, s>>>(g);
        else smallk_elem_kernel<false><<<blocks, 256, 0, s>>>(g);
        CDG_CHECK_LAUNCH();
        return CDG_OK;
    }
    if (g.N <= 8 && g.sa_k == 1 && (g.epi == EPI_NONE || g.epi == EPI_BIAS)) {
        static int rows_ok = -1;
        if (rows_ok < 0) rows_ok = exp_switch("CDG_SKINNY_ROWS", 1) != 0;
        if (rows_ok && g.M >= 64 && g.K % 4 == 0 && g.K >= 128 && g.N * g.K <= 8192 && g.sa_m % 4 == 0 && (((uintptr_t)g.A) & 15) == 0) {
            const int blocks = (int)imin64((g.M + 7) / 8, kNumSMs * 8);
            rowdot_rows_kernel<<<blocks, 256, sizeof(float) * (size_t)(g.N * g.K), s>>>(g);
            CDG_CHECK_LAUNCH();
            return CDG_OK;
        }
        const int blocks = (int)imin64((g.M + 7) / 8, kNumSMs * 16);
        rowdot_kernel<<<blocks, 256, 0, s>>>(g);
        CDG_CHECK_LAUNCH();
        return CDG_OK;
    }
    if ((g.M <= 8 || g.N <= 8) && g.sa_m == 1 && g.sb_n == 1 && g.epi == EPI_NONE) {
        SkinnyW a;
        const bool wide_is_m = g.M > g.N;
        a.wide = wide_is_m ? g.A : g.B; a.wide_ld = wide_is_m ? g.sa_k : g.sb_k;
        a.narrow = wide_is_m ? g.B : g.A; a.narrow_ld = wide_is_m ? g.sb_k : g.sa_k;
        a.C = g.C; a.c_sw = wide_is_m ? g.ldc : 1; a.c_sr = wide_is_m ? 1 : g.ldc;
        a.W = wide_is_m ? g.M : g.N; a.R = (int)(wide_is_m ? g.N : g.M); a.K = g.K;
        const int wb = (int)((a.W + 127) / 128);
        int ksplit = (int)imax64(1, imin64((kNumSMs * 16) / wb, (g.K + 63) / 64));
        a.k_chunk = (int)((g.K + ksplit - 1) / ksplit);
        ksplit = (int)((g.K + a.k_chunk - 1) / a.k_chunk);
        if (!g.accumulate) {
            if (g.ldc == g.N) CDG_CHECK_CUDA(cudaMemsetAsync(g.C, 0, sizeof(float) * g.M * g.N, s));
            else CDG_CHECK_CUDA(cudaMemset2DAsync(g.C, sizeof(float) * g.ldc, 0, sizeof(float) * g.N, g.M, s));
        }
        skinny_wgrad_kernel<<<dim3(wb, ksplit), 128, 0, s>>>(a);
        CDG_CHECK_LAUNCH();
        return CDG_OK;
    }
    return CDG_ERR_UNSUPPORTED;
}

}  // namespace cdg
