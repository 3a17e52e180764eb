#pragma once
#include "common.cuh"
namespace cdg {
const char* get_error();
int launch_recon(float* pre, const float* x, float* xhat, int64_t batch, int64_t P, double* acc, int write_grad,
                 cudaStream_t s);
int launch_masked_recon(float* const* sep, int K, const float* masks, const float* x, float* xhat, int64_t batch, int64_t P,
                        double* acc, int write_grad, cudaStream_t s);
int launch_gather_cols(const float* z, int d, float* zin, int f, int off, int extra, int64_t B, cudaStream_t s);
int launch_scatter_add_cols(const float* gzin, float* gz, int d, int f, int off, int extra, int64_t B, cudaStream_t s);
int launch_finalize_logs(double* acc, float* logs, int d, float recon_div, float kl_div, float align_div, float beta,
                         float lambda_, cudaStream_t s);
}  // namespace cdg
