// FP32 CUDA-core GEMM with arbitrary operand strides.
//   C[m,n] = sum_k A[m*sa_m + k*sa_k] * B[n*sb_n + k*sb_k]
// It serves (i) every contraction whose shape cannot feed a tcgen05 tile (N = 2d = 8 encoder
// head, K in {1,2} decoder inputs: modules/model.py:224, :245), (ii) split-K reductions at the
// reference batch size, and (iii) the on-device cross-check of the tensor-core kernel.
#include "common.cuh"

#include <cuda_bf16.h>

namespace cdg {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct SimtArgs {
    GemmDesc g;
    int k_chunk;      // K range handled by one blockIdx.z
    int atomic;       // split-K: red.add into pre-zeroed / accumulated C
};

template <bool A_MCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(NT, 2) gemm_simt_kernel(SimtArgs a) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const GemmDesc& g = a.g;
    const int t = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * a.k_chunk;
    const int64_t kend = min(g.K, kbeg + (int64_t)a.k_chunk);
    const int tx = t & 15, ty = t >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- stage A tile [BM x BK] -> As[k][m] ----
        if (A_MCONTIG) {
            const int m = t & (BM - 1);
#pragma unroll
            for (int kk = t / BM; kk < BK; kk += NT / BM) {
                const int64_t gm = m0 + m, gk = k0 + kk;
                As[kk][m] = (gm < g.M && gk < kend) ? g.A[gm * g.sa_m + gk * g.sa_k] : 0.f;
            }
        } else {
            const int kk = t & (BK - 1);
#pragma unroll
            for (int m = t / BK; m < BM; m += NT / BK) {
                const int64_t gm = m0 + m, gk = k0 + kk;
                As[kk][m] = (gm < g.M && gk < kend) ? g.A[gm * g.sa_m + gk * g.sa_k] : 0.f;
            }
        }
        if (B_NCONTIG) {
            const int n = t & (BN - 1);
#pragma unroll
            for (int kk = t / BN; kk < BK; kk += NT / BN) {
                const int64_t gn = n0 + n, gk = k0 + kk;
                Bs[kk][n] = (gn < g.N && gk < kend) ? g.B[gn * g.sb_n + gk * g.sb_k] : 0.f;
            }
        } else {
            const int kk = t & (BK - 1);
#pragma unroll
            for (int n = t / BK; n < BN; n += NT / BK) {
                const int64_t gn = n0 + n, gk = k0 + kk;
                Bs[kk][n] = (gn < g.N && gk < kend) ? g.B[gn * g.sb_n + gk * g.sb_k] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* c = g.C + m * g.ldc + n;
            if (a.atomic) { atomicAdd(c, v); continue; }
            if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_ACT) v += g.bias[n];
            if (g.epi == EPI_BIAS_ACT) v = act_fwd(v, g.act);
            if (g.epi == EPI_MUL_DACT) v *= act_bwd_from_out(g.aux[m * g.ld_aux + n], g.act);
            if (g.accumulate) v += *c;
            *c = v;
        }
    }
}

__global__ void bias_act_kernel(float* C, int64_t ldc, int64_t M, int64_t N, const float* bias, int epi, int act,
                                const float* aux, int64_t ld_aux) {
    const int64_t total = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N, n = i - m * N;
        float v = C[m * ldc + n];
        if (epi == EPI_BIAS || epi == EPI_BIAS_ACT) v += bias[n];
        if (epi == EPI_BIAS_ACT) v = act_fwd(v, act);
        if (epi == EPI_MUL_DACT) v *= act_bwd_from_out(aux[m * ld_aux + n], act);
        C[m * ldc + n] = v;
    }
}
// The same pass over groups of 4 columns (16-byte accesses), optionally writing the bf16 (hi, lo) planes of the result
// [M][ld16] for a following pre-split GEMM: columns >= N are 0, column N is 1 when `ones`.
__global__ void __launch_bounds__(256) bias_act_rows_kernel(float* C, int64_t ldc, int64_t M, int N, const float* __restrict__ bias,
                                                            int epi, int act, const float* __restrict__ aux, int64_t ld_aux,
                                                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                            int64_t ld16, int ones) {
    const int n4 = N >> 2, groups = hi ? (int)(ld16 >> 2) : n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * groups; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / groups;
        const int q = (int)(i - m * groups);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (q < n4) {
            float4* cp = reinterpret_cast<float4*>(C + m * ldc) + q;
            const float4 c = *cp;
            v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w;
            if (epi == EPI_BIAS || epi == EPI_BIAS_ACT) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q);
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
            }
            if (epi == EPI_BIAS_ACT) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = act_fwd(v[e], act);
            }
            if (epi == EPI_MUL_DACT) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(aux + m * ld_aux) + q);
                v[0] *= act_bwd_from_out(h.x, act); v[1] *= act_bwd_from_out(h.y, act);
                v[2] *= act_bwd_from_out(h.z, act); v[3] *= act_bwd_from_out(h.w, act);
            }
            if (epi != EPI_NONE) *cp = make_float4(v[0], v[1], v[2], v[3]);
        } else if (q == n4 && ones) {
            v[0] = 1.f;
        }
        if (hi) {
            uint32_t h2[2], l2[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * k] - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * k + 1] - __bfloat162float(h1));
                h2[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                l2[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            *reinterpret_cast<uint2*>(hi + m * ld16 + 4 * q) = make_uint2(h2[0], h2[1]);
            *reinterpret_cast<uint2*>(lo + m * ld16 + 4 * q) = make_uint2(l2[0], l2[1]);
        }
    }
}

int launch_bias_act(float* C, int64_t ldc, int64_t M, int64_t N, const float* bias, int epi, int act,
                    const float* aux, int64_t ld_aux, cudaStream_t s, void* out_hi, void* out_lo, int64_t ld_out16, int out_ones) {
    if (M * N == 0 || (epi == EPI_NONE && !out_hi)) return CDG_OK;
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    const bool rows = N % 4 == 0 && N < (1 << 30) && ldc % 4 == 0 && al16(C) &&
                      (!(epi == EPI_BIAS || epi == EPI_BIAS_ACT) || al16(bias)) &&
                      (epi != EPI_MUL_DACT || (ld_aux % 4 == 0 && al16(aux))) &&
                      (!out_hi || (out_lo && ld_out16 % 4 == 0 && ld_out16 >= N + (out_ones ? 1 : 0) &&
                                   ((uintptr_t)out_hi & 7) == 0 && ((uintptr_t)out_lo & 7) == 0));
    if (rows) {
        const int64_t total = M * (out_hi ? ld_out16 / 4 : N / 4);
        bias_act_rows_kernel<<<(int)imin64((total + 255) / 256, kNumSMs * 16), 256, 0, s>>>(
            C, ldc, M, (int)N, bias, epi, act, aux, ld_aux, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, ld_out16, out_ones);
        CDG_CHECK_LAUNCH();
        if (out_hi) tl_planes_done = true;
        return CDG_OK;
    }
    if (epi == EPI_NONE) return CDG_OK;
    const int64_t total = M * N;
    const int blocks = (int)imin64((total + 255) / 256, kNumSMs * 8);
    bias_act_kernel<<<blocks, 256, 0, s>>>(C, ldc, M, N, bias, epi, act, aux, ld_aux);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// out[n] += sum_m G[m*ld + n]   (bias gradients: the batch sum autograd performs for nn.Linear)
// block = 32 lanes x 8 row-lanes; VEC: each lane owns 4 adjacent columns (128-bit loads), 4 rows in flight.
template <int VEC>
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ G, int64_t ld, int64_t M, int64_t N,
                                                     float* __restrict__ out, int rows_per_block) {
    __shared__ float red[8][32 * VEC + 1];
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int64_t n = ((int64_t)blockIdx.x * 32 + lane) * VEC;
    const int64_t mbeg = (int64_t)blockIdx.y * rows_per_block;
    const int64_t mend = min(M, mbeg + rows_per_block);
    float s[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) s[e] = 0.f;
    if (n < N) {
        int64_t m = mbeg + r;
        if (VEC == 4) {
            for (; m + 24 < mend; m += 32) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(G + (m + 8 * u) * ld + n);
#pragma unroll
                for (int u = 0; u < 4; ++u) { s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w; }
            }
            for (; m < mend; m += 8) {
                const float4 v = *reinterpret_cast<const float4*>(G + m * ld + n);
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            }
        } else {
            for (; m < mend; m += 8) s[0] += G[m * ld + n];
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) red[r][lane * VEC + e] = s[e];
    __syncthreads();
    if (r == 0 && n < N) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i][lane * VEC + e];
            atomicAdd(out + n + e, t);
        }
    }
}

int launch_colsum(const float* G, int64_t ld, int64_t M, int64_t N, float* out, cudaStream_t s) {
    if (M == 0 || N == 0) return CDG_OK;
    const bool vec = N % 4 == 0 && ld % 4 == 0 && (((uintptr_t)G) & 15) == 0;
    const int cols_per_block = vec ? 128 : 32;
    const int cb = (int)((N + cols_per_block - 1) / cols_per_block);
    int rsplit = (int)imin64((M + 255) / 256, imax64(1, (kNumSMs * 8) / cb));
    const int rows = (int)((M + rsplit - 1) / rsplit);
    rsplit = (int)((M + rows - 1) / rows);
    if (vec) colsum_kernel<4><<<dim3(cb, rsplit), 256, 0, s>>>(G, ld, M, N, out, rows);
    else colsum_kernel<1><<<dim3(cb, rsplit), 256, 0, s>>>(G, ld, M, N, out, rows);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

int gemm_simt(const GemmDesc& g, cudaStream_t s) {
    if (g.M == 0 || g.N == 0) return CDG_OK;
    CDG_REQUIRE(g.K >= 0 && g.M > 0 && g.N > 0, "gemm_simt: bad shape");
    const int64_t tm = (g.M + BM - 1) / BM, tn = (g.N + BN - 1) / BN;
    CDG_REQUIRE(tm <= 65535, "gemm_simt: M too large for grid.y");
    // split-K when the tile grid cannot fill the chip and K is long
    int splits = 1;
    const int64_t tiles = tm * tn;
    if (tiles < kNumSMs && g.K >= 512) {
        splits = (int)imin64((2 * kNumSMs + tiles - 1) / tiles, g.K / 128);
        splits = (int)imax64(1, imin64(splits, 256));
    }
    int k_chunk = (int)(((g.K + splits - 1) / splits + BK - 1) / BK * BK);
    if (k_chunk <= 0) k_chunk = BK;
    splits = (int)((g.K + k_chunk - 1) / k_chunk);
    if (splits < 1) splits = 1;
    SimtArgs a;
    a.g = g;
    a.k_chunk = k_chunk;
    a.atomic = splits > 1;
    if (a.atomic && !g.accumulate) {
        if (g.ldc == g.N) {
            CDG_CHECK_CUDA(cudaMemsetAsync(g.C, 0, sizeof(float) * g.M * g.N, s));
        } else {
            CDG_CHECK_CUDA(cudaMemset2DAsync(g.C, sizeof(float) * g.ldc, 0, sizeof(float) * g.N, g.M, s));
        }
    }
    dim3 grid((unsigned)tn, (unsigned)tm, (unsigned)splits);
    const bool am = (g.sa_m == 1 && g.sa_k != 1), bn = (g.sb_n == 1 && g.sb_k != 1);
    if (am && bn) gemm_simt_kernel<true, true><<<grid, NT, 0, s>>>(a);
    else if (am) gemm_simt_kernel<true, false><<<grid, NT, 0, s>>>(a);
    else if (bn) gemm_simt_kernel<false, true><<<grid, NT, 0, s>>>(a);
    else gemm_simt_kernel<false, false><<<grid, NT, 0, s>>>(a);
    CDG_CHECK_LAUNCH();
    if (a.atomic && g.epi != EPI_NONE)
        CDG_TRY(launch_bias_act(g.C, g.ldc, g.M, g.N, g.bias, g.epi, g.act, g.aux, g.ld_aux, s));
    return CDG_OK;
}

}  // namespace cdg
