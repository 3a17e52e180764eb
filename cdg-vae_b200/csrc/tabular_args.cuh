#pragma once
#include "common.cuh"
namespace cdg {
constexpr int TAB_THREADS = 128;
struct TabArgs {
    cdg_tabular_config c;
    const float* params;
    float* grads;
    const float* x; const float* y; const float* noise;
    int64_t batch;
    float* xhat; float* latents;
    double* acc;
    int do_bwd, deterministic, out_total;
};
bool launch_tab_const(const TabArgs& a, cudaStream_t s);       // tabular_const.cu: loan / adult, linear SCM, canonical arena
void set_tab_const_params(int on);
bool launch_tab_fixed(const TabArgs& a, unsigned blocks, size_t smem, cudaStream_t s);
bool launch_tvae_mma(const TabArgs& a, cudaStream_t s);         // tvae_mma.cu: the same on mma.sync 3xTF32 fragments
bool launch_tvae_tile(const TabArgs& a, cudaStream_t s);        // tvae_tile.cu: warp-cooperative CDG-TVAE step
bool launch_tvae_fixed(const TabArgs& a, unsigned blocks, size_t smem, cudaStream_t s);
}  // namespace cdg
