// Layer kernels of the CelebA path; see conv.cuh.  All of them are memory-bound: coalesced along the channel
// dimension (NHWC), 128-bit accesses where the channel count allows, grid-stride loops sized to the 148 SMs.
#include <cuda_bf16.h>
#include "conv.cuh"

namespace cdg {

static inline int grid_for_elems(int64_t n, int per_block = 256) {
    int64_t b = (n + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > (int64_t)kNumSMs * 16) b = (int64_t)kNumSMs * 16;
    return (int)b;
}

// ---- im2col (nn.Conv2d input gathering; optional folded BatchNorm + ReLU + nearest upsampling) -------------
template <bool VEC>
__global__ void __launch_bounds__(256) im2col_kernel(Im2colArgs a) {
    const int Hin = a.Hs * a.up, Win = a.Ws * a.up;
    if (VEC) {
        const int C4 = a.C >> 2;
        const int64_t total = a.B * a.Ho * a.Wo * (int64_t)(a.k * a.k) * C4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int c4 = (int)(i % C4);
            int64_t t = i / C4;
            const int kw = (int)(t % a.k); t /= a.k;
            const int kh = (int)(t % a.k); t /= a.k;
            const int wo = (int)(t % a.Wo);
            const int64_t t2 = t / a.Wo;
            const int ho = (int)(t2 % a.Ho);
            const int64_t b = t2 / a.Ho;
            const int hi = ho * a.stride - a.pad + kh, wi = wo * a.stride - a.pad + kw;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hi >= 0 && hi < Hin && wi >= 0 && wi < Win) {
                const int hs = a.up == 2 ? hi >> 1 : hi, ws = a.up == 2 ? wi >> 1 : wi;
                v = *reinterpret_cast<const float4*>(a.src + ((b * a.Hs + hs) * a.Ws + ws) * a.ld + c4 * 4);
                if (a.scale) {
                    const float4 sc = *reinterpret_cast<const float4*>(a.scale + c4 * 4);
                    const float4 sh = *reinterpret_cast<const float4*>(a.shift + c4 * 4);
                    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                }
                if (a.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            }
            *reinterpret_cast<float4*>(a.col + t * (int64_t)a.Kp + (kh * a.k + kw) * a.C + c4 * 4) = v;
        }
        return;
    }
    const int K = a.k * a.k * a.C;
    const int64_t total = a.B * a.Ho * a.Wo * (int64_t)a.Kp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % a.Kp);
        const int64_t m = i / a.Kp;
        float v = 0.f;
        if (j < K) {
            const int c = j % a.C;
            const int kk = j / a.C;
            const int kw = kk % a.k, kh = kk / a.k;
            const int wo = (int)(m % a.Wo);
            const int64_t t2 = m / a.Wo;
            const int ho = (int)(t2 % a.Ho);
            const int64_t b = t2 / a.Ho;
            const int hi = ho * a.stride - a.pad + kh, wi = wo * a.stride - a.pad + kw;
            if (hi >= 0 && hi < Hin && wi >= 0 && wi < Win) {
                const int hs = a.up == 2 ? hi >> 1 : hi, ws = a.up == 2 ? wi >> 1 : wi;
                v = a.src[((b * a.Hs + hs) * a.Ws + ws) * a.ld + c];
                if (a.scale) v = fmaf(v, a.scale[c], a.shift[c]);
                if (a.relu) v = fmaxf(v, 0.f);
            }
        }
        a.col[i] = v;
    }
}

int launch_im2col(const Im2colArgs& a, cudaStream_t s) {
    const int K = a.k * a.k * a.C;
    const bool vec = a.C % 4 == 0 && a.ld % 4 == 0 && a.Kp == K && ((uintptr_t)a.src & 15) == 0 && ((uintptr_t)a.col & 15) == 0 &&
                     (!a.scale || ((((uintptr_t)a.scale | (uintptr_t)a.shift) & 15) == 0));
    const int64_t M = a.B * a.Ho * a.Wo;
    if (M == 0) return CDG_OK;
    if (vec) im2col_kernel<true><<<grid_for_elems(M * (K / 4)), 256, 0, s>>>(a);
    else im2col_kernel<false><<<grid_for_elems(M * a.Kp), 256, 0, s>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- weight layouts -------------------------------------------------------------------------------------
__device__ __forceinline__ void put_split(const Split16& d, int64_t row, int col, float v) {
    if (!d.hi) return;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(d.hi)[row * d.ld + col] = h;
    reinterpret_cast<__nv_bfloat16*>(d.lo)[row * d.ld + col] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__global__ void weight_prep_kernel(const float* __restrict__ w, int Co, int Ci, int k, const float* __restrict__ inv_sigma,
                                   float* __restrict__ wf, int Kpf, float* __restrict__ wd, int Kpd, Split16 f16, Split16 d16) {
    const float sc = inv_sigma ? *inv_sigma : 1.f;
    const int kk = k * k;
    const int64_t nf = (int64_t)Co * Kpf, nd = wd ? (int64_t)Ci * Kpd : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nd; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < nf) {
            const int co = (int)(i / Kpf), j = (int)(i % Kpf);
            float v = 0.f;
            if (j < kk * Ci) {
                const int ci = j % Ci, t = j / Ci;                 // t = kh * k + kw
                v = w[((int64_t)co * Ci + ci) * kk + t] * sc;
            }
            wf[i] = v;
            put_split(f16, co, j, v);
        } else {
            const int64_t r = i - nf;
            const int ci = (int)(r / Kpd), j = (int)(r % Kpd);
            float v = 0.f;
            if (j < kk * Co) {
                const int co = j % Co, t = j / Co;                 // position in the flipped kernel
                v = w[((int64_t)co * Ci + ci) * kk + (kk - 1 - t)] * sc;
            }
            wd[r] = v;
            put_split(d16, ci, j, v);
        }
    }
}
int launch_weight_prep(const float* w, int Co, int Ci, int k, const float* inv_sigma, float* wf, int Kpf, float* wd, int Kpd,
                       cudaStream_t s, Split16 wf16, Split16 wd16) {
    const int64_t n = (int64_t)Co * Kpf + (wd ? (int64_t)Ci * Kpd : 0);
    weight_prep_kernel<<<grid_for_elems(n), 256, 0, s>>>(w, Co, Ci, k, inv_sigma, wf, Kpf, wd, Kpd, wf16, wd16);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void lin0_prep_kernel(const float* __restrict__ w, const float* __restrict__ b, int C, int HW, int zd,
                                 const float* __restrict__ inv_sigma, float* __restrict__ wf, float* __restrict__ bf) {
    const float sc = inv_sigma ? *inv_sigma : 1.f;
    const int rows = C * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * (zd + 1); i += gridDim.x * blockDim.x) {
        const int r2 = i / (zd + 1), j = i % (zd + 1);             // r2 = hw * C + c  (NHWC row)
        const int c = r2 % C, hw = r2 / C;
        const int r = c * HW + hw;                                  // row of the reference's (c, h, w) view
        if (j < zd) wf[r2 * zd + j] = w[r * zd + j] * sc;
        else bf[r2] = b[r];
    }
}
int launch_lin0_prep(const float* w, const float* b, int C, int HW, int zd, const float* inv_sigma, float* wf, float* bf,
                     cudaStream_t s) {
    lin0_prep_kernel<<<grid_for_elems((int64_t)C * HW * (zd + 1)), 256, 0, s>>>(w, b, C, HW, zd, inv_sigma, wf, bf);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- spectral normalisation ------------------------------------------------------------------------------
// t = W^T u.  Wide layers (cols >= 32): block = 32 columns x 8 row lanes.  Narrow layers (the generators' first Linear:
// 8192 rows x 2..6 columns): one block, every thread strides over rows with one partial sum per column.
__global__ void __launch_bounds__(256) sn_wt_u_kernel(SnBatch sb) {
    const SnLayer L = sb.l[blockIdx.y];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (L.cols <= 8) {
        if (blockIdx.x != 0) return;
        __shared__ float red8[32];
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < L.rows; i += 256) {
            const float u = L.u[i];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < L.cols) acc[j] = fmaf(L.w[(int64_t)i * L.cols + j], u, acc[j]);
        }
        for (int j = 0; j < L.cols; ++j) {
            float v = warp_sum(acc[j]);
            __syncthreads();
            if ((tid & 31) == 0) red8[tid >> 5] = v;
            __syncthreads();
            if (tid == 0) {
                float s = 0.f;
                for (int w = 0; w < 8; ++w) s += red8[w];
                L.t[j] = s;
            }
        }
        return;
    }
    const int j = blockIdx.x * 32 + threadIdx.x;
    if (blockIdx.x * 32 >= L.cols) return;
    __shared__ float red[8][33];
    float acc = 0.f;
    if (j < L.cols)
        for (int i = threadIdx.y; i < L.rows; i += 8) acc = fmaf(L.w[(int64_t)i * L.cols + j], L.u[i], acc);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && j < L.cols) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
        L.t[j] = s;
    }
}
// v = t / max(|t|, eps);  sv = W v : one warp per row
__global__ void __launch_bounds__(256) sn_w_v_kernel(SnBatch sb) {
    const SnLayer L = sb.l[blockIdx.y];
    if (blockIdx.x * 8 >= L.rows) return;
    __shared__ float red[32];
    float n2 = 0.f;
    for (int j = threadIdx.x; j < L.cols; j += blockDim.x) n2 = fmaf(L.t[j], L.t[j], n2);
    n2 = block_sum<float>(n2, red);
    __shared__ float inv_s;
    if (threadIdx.x == 0) inv_s = 1.f / fmaxf(sqrtf(n2), 1e-12f);
    __syncthreads();
    const float inv = inv_s;
    if (blockIdx.x == 0)
        for (int j = threadIdx.x; j < L.cols; j += blockDim.x) L.v[j] = L.t[j] * inv;
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row < L.rows) {
        float acc = 0.f;
        for (int j = lane; j < L.cols; j += 32) acc = fmaf(L.w[(int64_t)row * L.cols + j], L.t[j] * inv, acc);
        acc = warp_sum(acc);
        if (lane == 0) L.sv[row] = acc;
    }
}
// u = sv / max(|sv|, eps);  sigma = u . sv
__global__ void __launch_bounds__(256) sn_finish_kernel(SnBatch sb) {
    const SnLayer L = sb.l[blockIdx.x];
    __shared__ float red[32];
    float n2 = 0.f;
    for (int i = threadIdx.x; i < L.rows; i += blockDim.x) n2 = fmaf(L.sv[i], L.sv[i], n2);
    n2 = block_sum<float>(n2, red);
    __shared__ float inv_s;
    if (threadIdx.x == 0) inv_s = 1.f / fmaxf(sqrtf(n2), 1e-12f);
    __syncthreads();
    const float inv = inv_s;
    float dot = 0.f;
    for (int i = threadIdx.x; i < L.rows; i += blockDim.x) {
        const float u = L.sv[i] * inv;
        L.u[i] = u;
        dot = fmaf(u, L.sv[i], dot);
    }
    dot = block_sum<float>(dot, red);
    if (threadIdx.x == 0) *L.inv_sigma = 1.f / dot;
}
int launch_spectral_norm(const SnBatch& b, cudaStream_t s) {
    if (b.n == 0) return CDG_OK;
    int maxc = 1, maxr = 1;
    for (int i = 0; i < b.n; ++i) { maxc = b.l[i].cols > maxc ? b.l[i].cols : maxc; maxr = b.l[i].rows > maxr ? b.l[i].rows : maxr; }
    sn_wt_u_kernel<<<dim3((maxc + 31) / 32, b.n), dim3(32, 8), 0, s>>>(b);
    CDG_CHECK_LAUNCH();
    sn_w_v_kernel<<<dim3((maxr + 7) / 8, b.n), 256, 0, s>>>(b);
    CDG_CHECK_LAUNCH();
    sn_finish_kernel<<<b.n, 256, 0, s>>>(b);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- BatchNorm (training mode) ------------------------------------------------------------------------------
// block = 32 channels x 8 row lanes; grid.y = channel slab, grid.x = row chunks
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ x, int64_t M, int C, double* __restrict__ acc) {
    const int c = blockIdx.y * 32 + threadIdx.x;
    __shared__ double r1[8][33], r2[8][33];
    double s1 = 0.0, s2 = 0.0;
    if (c < C)
        for (int64_t m = (int64_t)blockIdx.x * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.x * 8) {
            const double v = (double)x[m * C + c];
            s1 += v;
            s2 += v * v;
        }
    r1[threadIdx.y][threadIdx.x] = s1;
    r2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int r = 1; r < 8; ++r) { s1 += r1[r][threadIdx.x]; s2 += r2[r][threadIdx.x]; }
        atomicAdd(acc + c, s1);
        atomicAdd(acc + C + c, s2);
    }
}
static dim3 stats_grid(int64_t M, int C) {
    const int slabs = (C + 31) / 32;
    int64_t chunks = (M + 8 * 16 - 1) / (8 * 16);        // >= 16 rows per lane
    const int64_t cap = (int64_t)kNumSMs * 8 / slabs;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    return dim3((unsigned)chunks, (unsigned)slabs);
}
int launch_col_stats(const float* x, int64_t M, int C, double* acc, cudaStream_t s) {
    col_stats_kernel<<<stats_grid(M, C), dim3(32, 8), 0, s>>>(x, M, C, acc);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int64_t M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int n_updates,
                                   float* running_mean, float* running_var, float* scale, float* shift, float* mean_o,
                                   float* rstd_o) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = acc[c] / (double)M;
    double var = acc[C + c] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float meanf = (float)mean, varf = (float)var;
    const float rstd = 1.f / sqrtf(varf + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - meanf * sc;
    if (mean_o) mean_o[c] = meanf;
    if (rstd_o) rstd_o[c] = rstd;
    const float unbiased = (float)(var * (double)M / (double)(M > 1 ? M - 1 : 1));
    float rm = running_mean[c], rv = running_var[c];
    for (int i = 0; i < n_updates; ++i) {                  // one update per encode() call (model.py:204, :212)
        rm = momentum * meanf + (1.f - momentum) * rm;
        rv = momentum * unbiased + (1.f - momentum) * rv;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
}
int launch_bn_finalize(const double* acc, int64_t M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       int n_updates, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                       float* rstd, cudaStream_t s) {
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(acc, M, C, gamma, beta, eps, momentum, n_updates, running_mean,
                                                       running_var, scale, shift, mean, rstd);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void __launch_bounds__(256) bn_act_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, const float* __restrict__ res,
                                                     const float* __restrict__ rscale, const float* __restrict__ rshift,
                                                     int relu, float* __restrict__ y, int64_t M, int C) {
    const int C4 = C >> 2;
    const int64_t total = M * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (res) {
            float4 r = reinterpret_cast<const float4*>(res)[i];
            if (rscale) {
                const float4 rs = *reinterpret_cast<const float4*>(rscale + c), rh = *reinterpret_cast<const float4*>(rshift + c);
                r.x = fmaf(r.x, rs.x, rh.x); r.y = fmaf(r.y, rs.y, rh.y); r.z = fmaf(r.z, rs.z, rh.z); r.w = fmaf(r.w, rs.w, rh.w);
            }
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        reinterpret_cast<float4*>(y)[i] = v;
    }
}
int launch_bn_act(const float* x, const float* scale, const float* shift, const float* res, const float* rscale,
                  const float* rshift, int relu, float* y, int64_t M, int C, cudaStream_t s) {
    CDG_REQUIRE(C % 4 == 0, "bn_act: channel count %d is not a multiple of 4", C);
    bn_act_kernel<<<grid_for_elems(M * (C / 4)), 256, 0, s>>>(x, scale, shift, res, rscale, rshift, relu, y, M, C);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void __launch_bounds__(256) maxpool_bn_relu_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, float* __restrict__ y, int64_t B,
                                                              int H, int W, int C) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = B * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int64_t b = t / Ho;
        const float sc = scale[c], sh = shift[c];
        float m = -INFINITY;
        for (int kh = 0; kh < 3; ++kh) {
            const int hi = ho * 2 - 1 + kh;
            if (hi < 0 || hi >= H) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int wi = wo * 2 - 1 + kw;
                if (wi < 0 || wi >= W) continue;
                m = fmaxf(m, fmaxf(fmaf(x[((b * H + hi) * W + wi) * C + c], sc, sh), 0.f));
            }
        }
        y[i] = m;
    }
}
int launch_maxpool_bn_relu(const float* x, const float* scale, const float* shift, float* y, int64_t B, int H, int W, int C,
                           cudaStream_t s) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    maxpool_bn_relu_kernel<<<grid_for_elems(B * Ho * Wo * C), 256, 0, s>>>(x, scale, shift, y, B, H, W, C);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void avgpool_kernel(const float* __restrict__ x, int64_t B, int HW, int C, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * C; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t b = i / C;
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += x[(b * HW + p) * C + c];
        out[i] = s / (float)HW;
    }
}
int launch_avgpool(const float* x, int64_t B, int HW, int C, float* out, cudaStream_t s) {
    avgpool_kernel<<<grid_for_elems(B * C), 256, 0, s>>>(x, B, HW, C, out);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- nearest-neighbour upsampling: skip-path add and its backward ---------------------------------------------
__global__ void __launch_bounds__(256) add_up2_kernel(const float* __restrict__ y, const float* __restrict__ lo,
                                                      float* __restrict__ out, int64_t B, int H, int W, int C) {
    const int C4 = C >> 2, H2 = 2 * H, W2 = 2 * W;
    const int64_t total = B * H2 * W2 * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        int64_t t = i / C4;
        const int w = (int)(t % W2); t /= W2;
        const int h = (int)(t % H2);
        const int64_t b = t / H2;
        float4 v = reinterpret_cast<const float4*>(y)[i];
        const float4 l = reinterpret_cast<const float4*>(lo)[((b * H + (h >> 1)) * W + (w >> 1)) * C4 + c4];
        v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
        reinterpret_cast<float4*>(out)[i] = v;
    }
}
int launch_add_up2(const float* y, const float* lo, float* out, int64_t B, int H, int W, int C, cudaStream_t s) {
    CDG_REQUIRE(C % 4 == 0, "add_up2: channel count %d is not a multiple of 4", C);
    add_up2_kernel<<<grid_for_elems(B * 4 * H * W * (C / 4)), 256, 0, s>>>(y, lo, out, B, H, W, C);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void __launch_bounds__(256) downsum2_kernel(const float* __restrict__ g, float* __restrict__ out, int64_t B, int H,
                                                       int W, int C) {
    const int C4 = C >> 2, W2 = 2 * W;
    const int64_t total = B * H * W * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        int64_t t = i / C4;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int64_t b = t / H;
        const float4* p = reinterpret_cast<const float4*>(g) + ((b * 2 * H + 2 * h) * W2 + 2 * w) * C4 + c4;
        const float4 a0 = p[0], a1 = p[C4], a2 = p[(int64_t)W2 * C4], a3 = p[(int64_t)W2 * C4 + C4];
        reinterpret_cast<float4*>(out)[i] = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y),
                                                        (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
    }
}
int launch_downsum2(const float* g, float* out, int64_t B, int H, int W, int C, cudaStream_t s) {
    CDG_REQUIRE(C % 4 == 0, "downsum2: channel count %d is not a multiple of 4", C);
    downsum2_kernel<<<grid_for_elems(B * H * W * (C / 4)), 256, 0, s>>>(g, out, B, H, W, C);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- BatchNorm + ReLU backward (input gradient only: the generators are never optimised) --------
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            int64_t M, int C, double* __restrict__ acc) {
    const int c = blockIdx.y * 32 + threadIdx.x;
    __shared__ double r1[8][33], r2[8][33];
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
        for (int64_t m = (int64_t)blockIdx.x * 8 + threadIdx.y; m < M; m += (int64_t)gridDim.x * 8) {
            const float xv = x[m * C + c];
            if (fmaf(xv, sc, sh) > 0.f) {
                const float gv = g[m * C + c];
                s1 += (double)gv;
                s2 += (double)(gv * ((xv - mu) * rs));
            }
        }
    }
    r1[threadIdx.y][threadIdx.x] = s1;
    r2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int r = 1; r < 8; ++r) { s1 += r1[r][threadIdx.x]; s2 += r2[r][threadIdx.x]; }
        atomicAdd(acc + c, s1);
        atomicAdd(acc + C + c, s2);
    }
}
int launch_bn_bwd_reduce(const float* g, const float* x, const float* scale, const float* shift, const float* mean,
                         const float* rstd, int64_t M, int C, double* acc, cudaStream_t s) {
    bn_bwd_reduce_kernel<<<stats_grid(M, C), dim3(32, 8), 0, s>>>(g, x, scale, shift, mean, rstd, M, C, acc);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const double* __restrict__ acc, const float* __restrict__ add,
                                                           float* __restrict__ dx, int64_t M, int C) {
    const int64_t total = M * C;
    const float invM = 1.f / (float)M;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float sc = scale[c], xv = x[i];
        const float xh = (xv - mean[c]) * rstd[c];
        const float gm = fmaf(xv, sc, shift[c]) > 0.f ? g[i] : 0.f;
        float d = sc * (gm - (float)acc[c] * invM - xh * ((float)acc[C + c] * invM));
        if (add) d += add[i];
        dx[i] = d;
    }
}
int launch_bn_bwd_apply(const float* g, const float* x, const float* scale, const float* shift, const float* mean,
                        const float* rstd, const double* acc, const float* add, float* dx, int64_t M, int C, cudaStream_t s) {
    bn_bwd_apply_kernel<<<grid_for_elems(M * C), 256, 0, s>>>(g, x, scale, shift, mean, rstd, acc, add, dx, M, C);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

}  // namespace cdg
