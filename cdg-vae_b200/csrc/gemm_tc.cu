// tcgen05 / TMEM / TMA GEMM (placeholder until the kernel lands: everything is reported as
// unsupported so that the dispatcher uses the SIMT kernel).
#include "common.cuh"
namespace cdg {
int gemm_tc(const GemmDesc&, int, void*, int64_t, cudaStream_t) { return CDG_ERR_UNSUPPORTED; }
}  // namespace cdg
