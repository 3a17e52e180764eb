// Routes a contraction to the tcgen05 kernel when its shape can feed tensor-core tiles and to
// the SIMT kernel otherwise.  There is no CPU path.
#include "common.cuh"

namespace cdg {

int gemm_dispatch(int mode, const GemmDesc& g, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    if (mode == CDG_GEMM_SIMT) return gemm_simt(g, s);
    const int passes = mode == CDG_GEMM_TC1X ? 1 : (mode == CDG_GEMM_BF3X ? 2 : 3);
    int r = gemm_tc(g, passes, workspace, workspace_bytes, s);
    if (r == CDG_ERR_UNSUPPORTED) r = gemm_skinny(g, s);
    if (r == CDG_ERR_UNSUPPORTED) r = gemm_simt(g, s);
    return r;
}

}  // namespace cdg
