// Layer kernels of the CelebA CDG-VAE path (celeba/module/model.py, celeba/module/sagan.py): convolutions as
// im2col + the tcgen05 GEMM, train-mode BatchNorm (forward and input-gradient), spectral normalisation, pooling,
// nearest-neighbour upsampling.  Activations are NHWC fp32, i.e. row-major [B*H*W, C] matrices.
#pragma once
#include "common.cuh"

namespace cdg {

static inline int round_up4(int k) { return (k + 3) & ~3; }

struct Im2colArgs {
    const float* src;      // [B, Hs, Ws, ld] (first C of ld channels are read)
    int64_t B;
    int Hs, Ws, C, ld;
    const float* scale;    // optional per-channel affine (BatchNorm folded) applied before the optional ReLU
    const float* shift;
    int relu;
    int up;                // 1, or 2 = nearest-neighbour upsampling of the (activated) source before the convolution
    int k, stride, pad;
    int Ho, Wo;
    float* col;            // [B*Ho*Wo, Kp], column order (kh, kw, c); columns >= k*k*C are zero
    int Kp;
};
int launch_im2col(const Im2colArgs& a, cudaStream_t s);

// OIHW weight (optionally times *inv_sigma) -> forward matrix wf[Co][Kpf] with K order (kh, kw, ci) and, when wd != null,
// the input-gradient matrix wd[Ci][Kpd] with K order (kh, kw, co) of the spatially flipped kernel.
// wf16 / wd16 (optional): the same matrices as bf16 (hi, lo) pairs, row stride ld (a multiple of 8), for the bf16x3 GEMM.
struct Split16 { void* hi = nullptr; void* lo = nullptr; int ld = 0; };
int launch_weight_prep(const float* w, int Co, int Ci, int k, const float* inv_sigma, float* wf, int Kpf, float* wd, int Kpd,
                       cudaStream_t s, Split16 wf16 = Split16(), Split16 wd16 = Split16());
// GenIniBlock's Linear (sagan.py:91-97): rows (c, hw) of the [C*HW, zd] weight and of the bias are re-ordered to (hw, c)
// so that the output is NHWC.
int launch_lin0_prep(const float* w, const float* b, int C, int HW, int zd, const float* inv_sigma, float* wf, float* bf,
                     cudaStream_t s);

// torch.nn.utils.spectral_norm, training mode: one power iteration per layer, in place on u and v; writes 1 / sigma.
struct SnLayer { const float* w; float* u; float* v; float* t; float* sv; float* inv_sigma; int rows, cols; };
constexpr int kSnMax = 24;
struct SnBatch { int n; SnLayer l[kSnMax]; };
int launch_spectral_norm(const SnBatch& b, cudaStream_t s);

// BatchNorm2d, training mode.  acc[2C] (zeroed by the caller) receives sum and sum of squares in double.
int launch_col_stats(const float* x, int64_t M, int C, double* acc, cudaStream_t s);
int launch_bn_finalize(const double* acc, int64_t M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       int n_updates, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                       float* rstd, cudaStream_t s);
// y = [relu](x * scale + shift + res * rscale + rshift)   (res / rscale optional; rscale null = identity)
int launch_bn_act(const float* x, const float* scale, const float* shift, const float* res, const float* rscale,
                  const float* rshift, int relu, float* y, int64_t M, int C, cudaStream_t s);
// y = maxpool3x3/s2/p1(relu(x * scale + shift))
int launch_maxpool_bn_relu(const float* x, const float* scale, const float* shift, float* y, int64_t B, int H, int W, int C,
                           cudaStream_t s);
int launch_avgpool(const float* x, int64_t B, int HW, int C, float* out, cudaStream_t s);
// out[b, h, w, c] = y[b, h, w, c] + lo[b, h/2, w/2, c]      (H, W: the low-resolution extent)
int launch_add_up2(const float* y, const float* lo, float* out, int64_t B, int H, int W, int C, cudaStream_t s);
// out[b, h, w, c] = sum of the 2x2 block of g                (backward of nearest upsampling)
int launch_downsum2(const float* g, float* out, int64_t B, int H, int W, int C, cudaStream_t s);
// backward of a = relu(bn(x)): acc[2C] += (sum gm, sum gm * xhat) with gm = g masked by the ReLU
int launch_bn_bwd_reduce(const float* g, const float* x, const float* scale, const float* shift, const float* mean,
                         const float* rstd, int64_t M, int C, double* acc, cudaStream_t s);
// dx = scale * (gm - S1/M - xhat * S2/M) (+ add)
int launch_bn_bwd_apply(const float* g, const float* x, const float* scale, const float* shift, const float* mean,
                        const float* rstd, const double* acc, const float* add, float* dx, int64_t M, int C, cudaStream_t s);

}  // namespace cdg
