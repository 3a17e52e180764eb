// Routes a contraction to the tcgen05 kernel when its shape can feed tensor-core tiles and to
// the SIMT kernel otherwise.  There is no CPU path.
#include "common.cuh"

namespace cdg {

thread_local bool tl_planes_done = false;

int finish_planes(const GemmDesc& g, cudaStream_t s) {
    if (!g.out_hi16 || tl_planes_done) return CDG_OK;
    tl_planes_done = true;
    return launch_split_rows(g.C, g.M, g.N, g.ldc, g.out_hi16, g.out_lo16, g.ld_out16, nullptr, g.out_ones, s);
}

int gemm_dispatch(int mode, const GemmDesc& g, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    tl_planes_done = false;
    int r;
    if (mode == CDG_GEMM_SIMT) r = gemm_simt(g, s);
    else {
        const int passes = mode == CDG_GEMM_TC1X ? 1 : (mode == CDG_GEMM_BF3X ? 2 : 3);
        r = gemm_tc(g, passes, workspace, workspace_bytes, s);
        if (r == CDG_ERR_UNSUPPORTED) r = gemm_skinny(g, s);
        if (r == CDG_ERR_UNSUPPORTED) r = gemm_simt(g, s);
    }
    if (r != CDG_OK) return r;
    return finish_planes(g, s);
}

}  // namespace cdg
