// Memory-bound fused kernels of the training step: masked-sum/tanh/MSE reconstruction head
// with its backward, fused Adam over the flat arena, and the per-step log row.
#include <cuda_bf16.h>
#include "latent.cuh"
#include "elementwise.cuh"

#include <math.h>
#include <stdarg.h>

namespace cdg {

long long g_launches = 0;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// ---- reconstruction head (modules/model.py:285-287 + modules/train.py:175) ---------------
//   xhat = tanh(pre)   (pre already holds the masked sum: decoder k wrote only its live columns)
//   recon += 0.5 * (xhat - x)^2 ;  pre <- d recon / d pre = (xhat - x) (1 - xhat^2) / batch
template <int VEC>
__global__ void __launch_bounds__(256) recon_kernel(float* __restrict__ pre, const float* __restrict__ x,
                                                    float* __restrict__ xhat, int64_t total, float inv_batch,
                                                    double* acc, int write_grad) {
    __shared__ double red[32];
    double local = 0.0;
    const int64_t nvec = total / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float p[VEC], xv[VEC], xh[VEC];
        if (VEC == 4) {
            const float4 a = reinterpret_cast<const float4*>(pre)[i];
            p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w;
            if (x) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(x) + i);
                xv[0] = b.x; xv[1] = b.y; xv[2] = b.z; xv[3] = b.w;
            }
        } else {
            p[0] = pre[i];
            if (x) xv[0] = x[i];
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            xh[j] = tanhf(p[j]);
            if (x) {
                const float df = xh[j] - xv[j];
                s += 0.5f * df * df;
                p[j] = df * (1.f - xh[j] * xh[j]) * inv_batch;
            }
        }
        local += (double)s;
        if (VEC == 4) {
            if (write_grad) reinterpret_cast<float4*>(pre)[i] = make_float4(p[0], p[1], p[2], p[3]);
            if (xhat) reinterpret_cast<float4*>(xhat)[i] = make_float4(xh[0], xh[1], xh[2], xh[3]);
        } else {
            if (write_grad) pre[i] = p[0];
            if (xhat) xhat[i] = xh[0];
        }
    }
    if (acc) {
        const double s = block_sum<double>(local, red);
        if (threadIdx.x == 0) atomicAdd(acc + ACC_RECON, s);
    }
}

int launch_recon(float* pre, const float* x, float* xhat, int64_t batch, int64_t P, double* acc, int write_grad,
                 cudaStream_t s) {
    const int64_t total = batch * P;
    if (total == 0) return CDG_OK;
    const float inv_b = 1.f / (float)batch;
    const bool vec = (total % 4 == 0) && (((uintptr_t)pre | (uintptr_t)x | (uintptr_t)xhat) % 16 == 0);
    const int64_t n = vec ? total / 4 : total;
    const int blocks = (int)imin64((n + 255) / 256, kNumSMs * 16);
    if (vec) recon_kernel<4><<<blocks, 256, 0, s>>>(pre, x, xhat, total, inv_b, acc, write_grad);
    else recon_kernel<1><<<blocks, 256, 0, s>>>(pre, x, xhat, total, inv_b, acc, write_grad);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- general masks (modules/model.py:284-287 literally): pre = sum_k out_k * mask_k ; xhat = tanh(pre) ;
//      recon += 0.5 (xhat - x)^2 ; out_k <- d recon / d out_k = g * mask_k,  g = (xhat - x)(1 - xhat^2) / batch
struct MaskedArgs { float* sep[CDG_MAX_DEC]; int K; };
__global__ void __launch_bounds__(256) masked_recon_kernel(MaskedArgs m, const float* __restrict__ masks, const float* __restrict__ x,
                                                           float* __restrict__ xhat, int64_t batch, int64_t P, float inv_batch,
                                                           double* acc, int write_grad) {
    __shared__ double red[32];
    double local = 0.0;
    const int64_t total = batch * P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = i % P;
        float pre = 0.f;
        for (int k = 0; k < m.K; ++k) pre += m.sep[k][i] * masks[k * P + col];
        const float xh = tanhf(pre);
        if (xhat) xhat[i] = xh;
        if (x) {
            const float df = xh - x[i];
            local += (double)(0.5f * df * df);
            if (write_grad) {
                const float g = df * (1.f - xh * xh) * inv_batch;
                for (int k = 0; k < m.K; ++k) m.sep[k][i] = g * masks[k * P + col];
            }
        }
    }
    if (acc) {
        const double s = block_sum<double>(local, red);
        if (threadIdx.x == 0) atomicAdd(acc + ACC_RECON, s);
    }
}

int launch_masked_recon(float* const* sep, int K, const float* masks, const float* x, float* xhat, int64_t batch, int64_t P,
                        double* acc, int write_grad, cudaStream_t s) {
    MaskedArgs m;
    m.K = K;
    for (int k = 0; k < K; ++k) m.sep[k] = sep[k];
    const int64_t total = batch * P;
    if (total == 0) return CDG_OK;
    const int blocks = (int)imin64((total + 255) / 256, kNumSMs * 16);
    masked_recon_kernel<<<blocks, 256, 0, s>>>(m, masks, x, xhat, batch, P, 1.f / (float)batch, acc, write_grad);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- decoder input gather / gradient scatter for the DR variant (DR/modules/model.py:284-287) ----
// zin[b, 0:f] = z[b, off:off+f], zin[b, f] = z[b, extra]
__global__ void gather_cols_kernel(const float* __restrict__ z, int d, float* __restrict__ zin, int f, int off, int extra, int64_t B) {
    const int w = f + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * w; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / w;
        const int j = (int)(i - b * w);
        zin[i] = z[b * d + (j < f ? off + j : extra)];
    }
}
// g_z[b, off + j] += g_zin[b, j] (j < f), g_z[b, extra] += g_zin[b, f]
__global__ void scatter_add_cols_kernel(const float* __restrict__ gzin, float* __restrict__ gz, int d, int f, int off, int extra, int64_t B) {
    const int w = f + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * w; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / w;
        const int j = (int)(i - b * w);
        gz[b * d + (j < f ? off + j : extra)] += gzin[i];      // one thread per (row, column): no write conflicts
    }
}
int launch_gather_cols(const float* z, int d, float* zin, int f, int off, int extra, int64_t B, cudaStream_t s) {
    const int blocks = (int)imin64((B * (f + 1) + 255) / 256, kNumSMs * 8);
    gather_cols_kernel<<<blocks, 256, 0, s>>>(z, d, zin, f, off, extra, B);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
int launch_scatter_add_cols(const float* gzin, float* gz, int d, int f, int off, int extra, int64_t B, cudaStream_t s) {
    const int blocks = (int)imin64((B * (f + 1) + 255) / 256, kNumSMs * 8);
    scatter_add_cols_kernel<<<blocks, 256, 0, s>>>(gzin, gz, d, f, off, extra, B);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- weight pre-split for the bf16x3 GEMM: W = bf16 hi + bf16 lo, optionally transposed ------------------------------
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ W, int64_t rows, int64_t cols, int64_t ld,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld16,
                                                         int transpose, int ones_row) {
    __shared__ float tile[32][33];
    if (!transpose) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = i / cols, c = i - r * cols;
            const float v = W[r * ld + c];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[r * ld16 + c] = h;
            lo[r * ld16 + c] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
        return;
    }
    if (ones_row)       // output row `cols`: the constant 1 (hi = 1, lo = 0) over the `rows` contraction entries
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x) {
            hi[cols * ld16 + i] = __float2bfloat16_rn(1.f);
            lo[cols * ld16 + i] = __float2bfloat16_rn(0.f);
        }
    // transposed copy through a 32x32 shared-memory tile (coalesced on both sides)
    const int64_t tr = (rows + 31) / 32, tc = (cols + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int64_t t = blockIdx.x; t < tr * tc; t += gridDim.x) {
        const int64_t r0 = (t / tc) * 32, c0 = (t % tc) * 32;
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int64_t r = r0 + j, c = c0 + tx;
            tile[j][tx] = (r < rows && c < cols) ? W[r * ld + c] : 0.f;
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int64_t c = c0 + j, r = r0 + tx;               // output row = source column
            if (c < cols && r < rows) {
                const float v = tile[tx][j];
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                hi[c * ld16 + r] = h;
                lo[c * ld16 + r] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
        }
    }
}
// Row-major split with 16-byte accesses: thread = (row, group of 8 columns) -> one uint4 of hi and one of lo.  Column `cols`
// (when ld16 > cols) receives extra[r] (a Linear's bias beside its weight row) or the constant 1 (an activation row that
// multiplies that bias column): the bias then comes out of the GEMM itself.  Columns past that are zero.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ W, int64_t rows, int cols, int64_t ld,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld16,
                                                         const float* __restrict__ extra, int ones) {
    const int groups = (int)(ld16 / 8);
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * groups; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int c0 = (int)(i - r * groups) * 8;
        float v[8];
        const float* src = W + r * ld + c0;
        if (vec && c0 + 8 <= cols) {
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = c0 + k;
                v[k] = c < cols ? src[k] : (c == cols ? (extra ? extra[r] : (ones ? 1.f : 0.f)) : 0.f);
            }
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * k]), h1 = __float2bfloat16_rn(v[2 * k + 1]);
            const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * k] - __bfloat162float(h0));
            const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * k + 1] - __bfloat162float(h1));
            h[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            l[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(hi + r * ld16 + c0) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + r * ld16 + c0) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
int launch_split_rows(const float* W, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo, int64_t ld16, const float* extra,
                      int ones, cudaStream_t s) {
    if (rows == 0) return CDG_OK;
    if (ld16 % 8 != 0 || ld16 < cols || ((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) != 0 ||
        cols > (1 << 30)) {
        set_error("launch_split_rows: planes must be 16-byte aligned with a row stride that is a multiple of 8");
        return CDG_ERR_INVALID;
    }
    const int64_t work = (rows * (ld16 / 8) + 255) / 256;
    split_rows_kernel<<<(int)imin64(work, kNumSMs * 16), 256, 0, s>>>(W, rows, (int)cols, ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                    ld16, extra, ones);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

int launch_split_bf16(const float* W, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo, int64_t ld16, int transpose,
                      cudaStream_t s, int ones_row) {
    if (rows * cols == 0) return CDG_OK;
    const int64_t work = transpose ? ((rows + 31) / 32) * ((cols + 31) / 32) : (rows * cols + 255) / 256;
    split_bf16_kernel<<<(int)imin64(work, kNumSMs * 16), 256, 0, s>>>(W, rows, cols, ld, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld16,
                                                                   transpose, transpose ? ones_row : 0);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- log row (modules/train.py:198-207) --------------------------------------------------
__global__ void finalize_logs_kernel(double* acc, float* logs, int d, float recon_div, float kl_div, float align_div,
                                     float beta, float lambda_) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const float recon = (float)(acc[ACC_RECON] / (double)recon_div);
        const float kl = (float)(acc[ACC_KL] / (double)kl_div);
        const float al = (float)(acc[ACC_ALIGN] / (double)align_div);
        float loss = recon + beta * kl;
        loss += lambda_ * al;
        logs[0] = loss; logs[1] = recon; logs[2] = kl; logs[3] = al;
        for (int i = 0; i < d; ++i) logs[4 + i] = (float)(acc[ACC_VAR + i] / (double)kl_div);
        for (int i = 0; i < ACC_LEN; ++i) acc[i] = 0.0;
    }
}

int launch_finalize_logs(double* acc, float* logs, int d, float recon_div, float kl_div, float align_div, float beta,
                         float lambda_, cudaStream_t s) {
    finalize_logs_kernel<<<1, 32, 0, s>>>(acc, logs, d, recon_div, kl_div, align_div, beta, lambda_);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- Adam (torch/optim/adam.py::_single_tensor_adam, amsgrad=False) ----------------------
struct AdamK {
    int64_t off[CDG_MAX_SEG], len[CDG_MAX_SEG];
    int64_t ubeg[CDG_MAX_SEG + 1];   // first 4-float unit of segment i in the flat unit numbering (units = ceil(len / 4))
    int n_seg;
    float one_minus_b1, b2, one_minus_b2, eps, wd, gscale, step_size, bc2_sqrt;
    const int32_t* dev_step;     // graph replay: t lives on the device
    double lr, beta1, beta2;
    int64_t clamp_off, clamp_len;
    float clamp_lo, clamp_hi;
};

__global__ void bump_step_kernel(int32_t* t) { *t += 1; }

__device__ __forceinline__ void adam_one(float& pj, float gj, float& mj, float& vj, int64_t j, const AdamK& k, float step_size,
                                         float bc2_sqrt) {
    gj = gj * k.gscale;
    if (k.wd != 0.f) gj = gj + k.wd * pj;                          // coupled L2 (main_tvae.py:196-200)
    mj = mj + k.one_minus_b1 * (gj - mj);                          // exp_avg.lerp_(grad, 1 - beta1)
    vj = vj * k.b2 + k.one_minus_b2 * gj * gj;                     // mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(vj) / bc2_sqrt + k.eps;              // (sqrt(v) / sqrt(bc2)).add_(eps)
    pj = pj - step_size * (mj / denom);                            // addcdiv_(m, denom, -lr / bc1)
    if (j >= k.clamp_off && j < k.clamp_off + k.clamp_len)         // sigma.data.clamp_ (train.py:314)
        pj = fminf(fmaxf(pj, k.clamp_lo), k.clamp_hi);
}

// One flat grid over the 4-float units of ALL live segments (a block-per-segment grid sized for the longest one left most
// blocks idle and moved 4 bytes per access: 13 % of the HBM roofline in round 1): 16-byte loads / stores of p, g, m, v,
// 28 bytes of traffic per parameter.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, const __grid_constant__ AdamK k) {
    __shared__ float s_hyper[2];
    float ss = k.step_size, bs = k.bc2_sqrt;
    if (k.dev_step) {
        // bias corrections from the device-resident step count, in double like the host path
        if (threadIdx.x == 0) {
            const double t = (double)*k.dev_step;
            s_hyper[0] = (float)(k.lr / (1.0 - pow(k.beta1, t)));
            s_hyper[1] = (float)sqrt(1.0 - pow(k.beta2, t));
        }
        __syncthreads();
        ss = s_hyper[0];
        bs = s_hyper[1];
    }
    const int64_t total = k.ubeg[k.n_seg];
    int sgm = 0;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
        while (u >= k.ubeg[sgm + 1]) ++sgm;                            // u only grows: the segment index only moves forward
        const int64_t i = (u - k.ubeg[sgm]) * 4, j = k.off[sgm] + i;
        const int64_t n = k.len[sgm] - i;
        if (n >= 4 && (j & 3) == 0) {
            float4 pj = *reinterpret_cast<float4*>(p + j), mj = *reinterpret_cast<float4*>(m + j), vj = *reinterpret_cast<float4*>(v + j);
            const float4 gj = *reinterpret_cast<const float4*>(g + j);
            adam_one(pj.x, gj.x, mj.x, vj.x, j, k, ss, bs); adam_one(pj.y, gj.y, mj.y, vj.y, j + 1, k, ss, bs);
            adam_one(pj.z, gj.z, mj.z, vj.z, j + 2, k, ss, bs); adam_one(pj.w, gj.w, mj.w, vj.w, j + 3, k, ss, bs);
            *reinterpret_cast<float4*>(p + j) = pj; *reinterpret_cast<float4*>(m + j) = mj; *reinterpret_cast<float4*>(v + j) = vj;
        } else {
            for (int64_t e = 0; e < n && e < 4; ++e) {
                float pj = p[j + e], mj = m[j + e], vj = v[j + e];
                adam_one(pj, g[j + e], mj, vj, j + e, k, ss, bs);
                p[j + e] = pj; m[j + e] = mj; v[j + e] = vj;
            }
        }
    }
}

}  // namespace cdg

extern "C" int cdg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                             const cdg_adam_args* a, void* stream) {
    using namespace cdg;
    CDG_REQUIRE(params && grads && exp_avg && exp_avg_sq && a, "cdg_adam_step: null argument");
    CDG_REQUIRE(a->n_seg >= 0 && a->n_seg <= CDG_MAX_SEG, "cdg_adam_step: n_seg out of range");
    CDG_REQUIRE(a->step >= 1 || a->dev_step, "cdg_adam_step: step must be >= 1");
    if (a->n_seg == 0) return CDG_OK;
    AdamK k;
    int64_t maxlen = 0;
    k.n_seg = a->n_seg;
    k.ubeg[0] = 0;
    for (int i = 0; i < a->n_seg; ++i) {
        k.off[i] = a->seg_off[i];
        k.len[i] = a->seg_len[i];
        CDG_REQUIRE(k.off[i] >= 0 && k.len[i] >= 0, "cdg_adam_step: bad segment");
        k.ubeg[i + 1] = k.ubeg[i] + (k.len[i] + 3) / 4;
        if (k.len[i] > maxlen) maxlen = k.len[i];
    }
    const double tt = a->dev_step ? 1.0 : (double)a->step;
    const double bc1 = 1.0 - pow(a->beta1, tt);
    const double bc2 = 1.0 - pow(a->beta2, tt);
    k.dev_step = a->dev_step; k.lr = a->lr; k.beta1 = a->beta1; k.beta2 = a->beta2;
    k.one_minus_b1 = (float)(1.0 - a->beta1);
    k.b2 = (float)a->beta2;
    k.one_minus_b2 = (float)(1.0 - a->beta2);
    k.eps = (float)a->eps;
    k.wd = (float)a->weight_decay;
    k.gscale = a->grad_scale;
    k.step_size = (float)(a->lr / bc1);
    k.bc2_sqrt = (float)sqrt(bc2);
    k.clamp_off = a->clamp_off; k.clamp_len = a->clamp_len; k.clamp_lo = a->clamp_lo; k.clamp_hi = a->clamp_hi;
    if (maxlen == 0) return CDG_OK;
    if (a->dev_step) {
        bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a->dev_step);
        CDG_CHECK_LAUNCH();
    }
    const int bx = (int)imin64((k.ubeg[k.n_seg] + 255) / 256, kNumSMs * 8);
    adam_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, k);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

extern "C" const char* cdg_last_error(void) { return cdg::get_error(); }
extern "C" int cdg_version(void) { return 200; }
extern "C" int64_t cdg_abi_sizeof(int which) {
    static const int64_t sz[] = {sizeof(cdg_linear), sizeof(cdg_adam_args), sizeof(cdg_pendulum_config), sizeof(cdg_pendulum_io),
                                 sizeof(cdg_pendulum_fwd_io), sizeof(cdg_tabular_config), sizeof(cdg_tabular_io), sizeof(cdg_conv),
                                 sizeof(cdg_bnorm), sizeof(cdg_celeba_config), sizeof(cdg_celeba_io), sizeof(cdg_tvae_transform_config)};
    return which >= 0 && which < (int)(sizeof(sz) / sizeof(sz[0])) ? sz[which] : -1;
}
extern "C" long long cdg_launch_count(void) { return cdg::g_launches; }
extern "C" void cdg_launch_count_add(long long n) { cdg::g_launches += n; }
extern "C" int cdg_device_ok(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    return prop.major == 10 ? 1 : 0;
}
