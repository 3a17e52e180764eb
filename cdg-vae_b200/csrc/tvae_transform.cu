// CDG-TVAE data transform, apply side: mode-specific normalisation of a raw table and its inverse.
//   forward : tabular/modules/data_transformer.py:163-182 (transform), :111-129 (per-column layout),
//             tabular/modules/numerical.py:407-445 (ClusterBasedNormalizer._transform)
//   inverse : tabular/modules/data_transformer.py:184-227, :131-147; tabular/modules/numerical.py:447-457, :175-177
//   gumbel  : tabular/inference_tvae.py:232-235, :250-253
// HBM-bound byte work: rows are staged through shared memory so that both the raw table (fp64 row-major) and the transformed
// table (fp32 row-major) move as contiguous, coalesced runs (forward: one warp per 32-row strip, private staging; inverse:
// 128-row block tiles); a warp always works on 32 consecutive rows of ONE raw column at a time, so the injected random
// stream (column-major, the order in which the reference's column loop draws it) is read coalesced and the per-column
// tables are warp-uniform.
// The component index is an inverse-cdf search and must agree with NumPy's float64 result: it is decided by an fp32
// evaluation whenever that is provably enough (cbn_select_fast) and by the fp64 evaluation otherwise; values are fp64.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

namespace cdg {
namespace {

constexpr int TILE_ROWS = 128;
constexpr int TT_THREADS = 256;

struct TransformArgs {
    cdg_tvae_transform_config c;
    int32_t cont_index[CDG_MAX_TCOL];   // position of a continuous column in the injected random stream
    const double* raw; int64_t ld_raw;
    const double* rnd;                  // uniforms (forward) / standard normals (inverse), [n_cont][rows]
    const float* data; int64_t ld_data; // transformed table (inverse input)
    const float* sigmas;
    float* out; int64_t ld_out;
    double* raw_out;
    int64_t rows;
    int raw_pitch;                      // doubles per staged raw row (odd: conflict-free column walks)
    int exact_only;                     // skip the fp32 filter (tests: both routes must give identical tables)
    // tables of the fp32 filter, components PERMUTED so that the kept ones come first (in their original order): the
    // kept probabilities are then statically indexed registers instead of a dynamically indexed local array
    float log_a_f[CDG_MAX_TCOL][CDG_MAX_TCOMP], prec_f[CDG_MAX_TCOL][CDG_MAX_TCOMP];
    double mean_p[CDG_MAX_TCOL][CDG_MAX_TCOMP];
};

// Copy a [n_rows, width] tile between global (row stride ld) and shared (row stride pitch) memory, 16-byte accesses when
// the global tile is one contiguous, aligned run.
template <typename T, bool TO_SHARED>
__device__ __forceinline__ void tile_copy(T* sh, int pitch, T* gl, int64_t ld, int n_rows, int width) {
    constexpr int V = 16 / sizeof(T);
    const int64_t total = (int64_t)n_rows * width;
    if (ld == width && pitch == width && (reinterpret_cast<uintptr_t>(gl) & 15) == 0) {
        const int64_t nv = total / V;
        for (int64_t i = threadIdx.x; i < nv; i += blockDim.x) {
            if (TO_SHARED) reinterpret_cast<int4*>(sh)[i] = __ldg(reinterpret_cast<const int4*>(gl) + i);
            else reinterpret_cast<int4*>(gl)[i] = reinterpret_cast<const int4*>(sh)[i];
        }
        for (int64_t i = nv * V + threadIdx.x; i < total; i += blockDim.x) {
            if (TO_SHARED) sh[i] = gl[i]; else gl[i] = sh[i];
        }
        return;
    }
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = (int)(i / width), c = (int)(i - (int64_t)r * width);
        if (TO_SHARED) sh[r * pitch + c] = gl[r * ld + c]; else gl[r * ld + c] = sh[r * pitch + c];
    }
}

// Filtered component draw.  The draw is an interval test of u against the cdf of the kept components; evaluating the
// responsibilities in fp32 (ten MUFU-based exp instead of ten fp64 exp, the cost of the whole kernel) decides it whenever u
// is farther from every cdf boundary than a bound on the fp32 evaluation error -- otherwise the caller falls back to the
// fp64 evaluation below, so the result is ALWAYS that of the fp64 route (tests compare the two routes cell by cell).
// Error bound: components that matter have lp_k >= max - 30, so their fp32 log-densities carry at most
// eps32 * (|max| + t_max + 32) absolute error each (t = 0.5 prec d^2 of the arg-max component, d from an fp64 subtraction);
// the cdf inherits at most ~3x that; the threshold is 8x.
__device__ __forceinline__ bool cbn_select_fast(int n_all, int n_valid, const double* __restrict__ mean_p,
                                                const float* __restrict__ log_a_f, const float* __restrict__ prec_f,
                                                double x, double u, int& comp) {
    // branch-free over the CDG_MAX_TCOMP slots: unused slots carry log_a = -inf (weight 0), so no per-slot predicates
    float e[CDG_MAX_TCOMP];
    float m = -INFINITY, t_m = 0.f;
#pragma unroll
    for (int k = 0; k < CDG_MAX_TCOMP; ++k) {
        const float d = (float)(x - mean_p[k]);
        const float t = 0.5f * (d * d * prec_f[k]);
        e[k] = log_a_f[k] - t;
        t_m = e[k] > m ? t : t_m;
        m = fmaxf(m, e[k]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CDG_MAX_TCOMP; ++k) { e[k] = __expf(e[k] - m); s += e[k]; }
    const float inv_s = 1.f / s;
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < CDG_MAX_TCOMP; ++j) {
        e[j] = j < n_valid ? e[j] * inv_s + 1e-6f : 0.f;
        tot += e[j];
    }
    const float inv_tot = 1.f / tot;
    const float uf = (float)u;
    // __expf: 2 ulp + 2^-21.4 * |arg| relative; everything else a few ulp -> 6e-7 per unit of log-density magnitude
    const float tau = 8.f * 6e-7f * (fabsf(m) + t_m + 32.f);
    float run = 0.f, gap = 1.f;
    int idx = 0;
#pragma unroll
    for (int j = 0; j < CDG_MAX_TCOMP - 1; ++j) {
        run += e[j] * inv_tot;
        const bool live = j < n_valid - 1;                // the last boundary is 1: u < 1 always
        idx += (live && run <= uf) ? 1 : 0;
        gap = live ? fminf(gap, fabsf(run - uf)) : gap;
    }
    comp = idx;
    return gap > tau;                                      // NaN (every component infinitely far) -> false -> fp64 route
}

// numerical.py:407-445 for one cell.  Returns the kept-component index and the clipped normalised value.
__device__ __forceinline__ void cbn_cell(const cdg_tvae_column& col, double x, double u, int& comp, double& value) {
    // sklearn BaseMixture.predict_proba: exp(weighted_log_prob - logsumexp(weighted_log_prob)) over ALL fitted components
    double lp[CDG_MAX_TCOMP];
    double m = -INFINITY;
#pragma unroll
    for (int k = 0; k < CDG_MAX_TCOMP; ++k) {
        if (k < col.n_all) {
            const double d = x - col.mean[k];
            lp[k] = col.log_a[k] - 0.5 * (d * d * col.prec[k]);
            m = fmax(m, lp[k]);
        }
    }
    // responsibilities exp(lp - logsumexp(lp)) evaluated as exp(lp - max) / sum: one exp per component, no log
    // (differs from sklearn's form by rounding only; the component draw below is an interval test on the cdf)
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CDG_MAX_TCOMP; ++k)
        if (k < col.n_all) { lp[k] = exp(lp[k] - m); s += lp[k]; }
    const double inv_s = 1.0 / s;
    // numerical.py:424-434: keep the valid components, + 1e-6, renormalise, then np.random.choice(p=...) =
    // cdf = cumsum(p); cdf /= cdf[-1]; index = searchsorted(cdf, u, side='right')
    double p[CDG_MAX_TCOMP];
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < CDG_MAX_TCOMP; ++j) {
        if (j < col.n_valid) {
            p[j] = lp[col.valid_idx[j]] * inv_s + 1e-6;
            tot += p[j];
        }
    }
    double run = 0.0, last = 0.0;
#pragma unroll
    for (int j = 0; j < CDG_MAX_TCOMP; ++j)
        if (j < col.n_valid) { p[j] = p[j] / tot; last += p[j]; }
    int idx = 0;
#pragma unroll
    for (int j = 0; j < CDG_MAX_TCOMP; ++j) {
        if (j < col.n_valid) {
            run += p[j];
            idx += (run / last <= u) ? 1 : 0;   // side='right': count cdf entries <= u
        }
    }
    comp = idx < col.n_valid ? idx : col.n_valid - 1;
    const int k = col.valid_idx[comp];
    double v = __ddiv_rn(__dsub_rn(x, col.mean[k]), __dmul_rn(4.0, col.std[k]));   // STD_MULTIPLIER = 4 (numerical.py:365)
    value = fmin(fmax(v, -0.99), 0.99);
}

// Forward transform: ONE WARP per strip of 32 table rows, private staging in shared memory, __syncwarp only.
// (The first version used 128-row block tiles: ncu showed warps waiting at the block barriers for as long as they issued --
// 20 column tasks over 8 warps -- and 3 blocks per SM.)  A lane owns one row; the strip's columns are walked in order, so
// the mixture tables are warp-uniform and the injected uniforms (column-major) are read as one 256-byte run per column.
constexpr int WARP_ROWS = 32;
__global__ void __launch_bounds__(TT_THREADS, 3) tvae_transform_kernel(const __grid_constant__ TransformArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.c.n_col, D = a.c.out_dim;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t per_warp = (size_t)WARP_ROWS * a.raw_pitch * sizeof(double) + (((size_t)WARP_ROWS * D * sizeof(float) + 15) & ~(size_t)15);
    double* sraw = reinterpret_cast<double*>(smem_raw + warp * per_warp);
    float* sout = reinterpret_cast<float*>(sraw + (size_t)WARP_ROWS * a.raw_pitch);
    const int64_t n_strips = (a.rows + WARP_ROWS - 1) / WARP_ROWS;
    const int64_t w_stride = (int64_t)gridDim.x * (TT_THREADS / 32);
    for (int64_t t = (int64_t)blockIdx.x * (TT_THREADS / 32) + warp; t < n_strips; t += w_stride) {
        const int64_t r0 = t * WARP_ROWS;
        const int nr = (int)min((int64_t)WARP_ROWS, a.rows - r0);
        __syncwarp();                                       // previous strip's stores have read the staging buffers
        if (a.ld_raw == C) {                                // the strip is one contiguous run of nr * C doubles
            const double* g = a.raw + r0 * a.ld_raw;
            for (int i = lane; i < nr * C; i += 32) {
                const int r = i / C;
                sraw[r * a.raw_pitch + (i - r * C)] = __ldg(g + i);
            }
        } else {
            for (int i = lane; i < nr * C; i += 32) {
                const int r = i / C, c = i - r * C;
                sraw[r * a.raw_pitch + c] = __ldg(a.raw + (r0 + r) * a.ld_raw + c);
            }
        }
        {
            float4* z = reinterpret_cast<float4*>(sout);
            const int nz = (WARP_ROWS * D + 3) / 4;
            for (int i = lane; i < nz; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        if (lane < nr) {
            for (int c = 0; c < C; ++c) {
                const cdg_tvae_column& col = a.c.col[c];
                const double x = sraw[lane * a.raw_pitch + c];
                float* o = sout + lane * D + col.out_start;
                if (col.kind == CDG_TCOL_CONTINUOUS) {
                    const double u = __ldg(a.rnd + (int64_t)a.cont_index[c] * a.rows + r0 + lane);
                    int comp; double v;
                    if (!a.exact_only && cbn_select_fast(col.n_all, col.n_valid, a.mean_p[c], a.log_a_f[c], a.prec_f[c], x, u, comp)) {
                        const int k = col.valid_idx[comp];
                        v = __ddiv_rn(__dsub_rn(x, col.mean[k]), __dmul_rn(4.0, col.std[k]));
                        v = fmin(fmax(v, -0.99), 0.99);
                    } else {
                        cbn_cell(col, x, u, comp, v);
                    }
                    o[0] = (float)v;
                    o[1 + comp] = 1.f;            // data_transformer.py:121-123
                } else {
                    for (int j = 0; j < col.n_valid; ++j) o[j] = (x == col.category[j]) ? 1.f : 0.f;
                }
            }
        }
        __syncwarp();
        // the strip's nr * D floats are contiguous in the output when ld_out == D
        float* g = a.out + r0 * a.ld_out;
        if (a.ld_out == D && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            const int nv = (nr * D) >> 2;
            for (int i = lane; i < nv; i += 32) reinterpret_cast<float4*>(g)[i] = reinterpret_cast<const float4*>(sout)[i];
            for (int i = (nv << 2) + lane; i < nr * D; i += 32) g[i] = sout[i];
        } else {
            for (int i = lane; i < nr * D; i += 32) {
                const int r = i / D, c = i - r * D;
                g[(int64_t)r * a.ld_out + c] = sout[i];
            }
        }
    }
}

__global__ void __launch_bounds__(TT_THREADS) tvae_inverse_kernel(const __grid_constant__ TransformArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.c.n_col, D = a.c.out_dim;
    double* sraw = reinterpret_cast<double*>(smem_raw);
    float* sdat = reinterpret_cast<float*>(sraw + (size_t)TILE_ROWS * C);
    const int64_t n_tiles = (a.rows + TILE_ROWS - 1) / TILE_ROWS;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * TILE_ROWS;
        const int nr = (int)min((int64_t)TILE_ROWS, a.rows - r0);
        __syncthreads();
        tile_copy<float, true>(sdat, D, const_cast<float*>(a.data + r0 * a.ld_data), a.ld_data, nr, D);
        __syncthreads();
        for (int cell = threadIdx.x; cell < C * TILE_ROWS; cell += blockDim.x) {
            const int c = cell / TILE_ROWS, r = cell - c * TILE_ROWS;
            if (r >= nr) continue;
            const cdg_tvae_column& col = a.c.col[c];
            const float* d = sdat + r * D + col.out_start;
            double res;
            if (col.kind == CDG_TCOL_CONTINUOUS) {
                // data_transformer.py:134-140: component = argmax of the one-hot block (first maximum), optional draw
                int best = 0; float bv = d[1];
                for (int j = 1; j < col.n_valid; ++j) if (d[1 + j] > bv) { bv = d[1 + j]; best = j; }
                double v = (double)d[0];
                if (a.sigmas != nullptr && a.rnd != nullptr)
                    v = __dadd_rn(v, __dmul_rn((double)__ldg(a.sigmas + col.out_start),      // no FMA contraction: NumPy rounds
                                               __ldg(a.rnd + (int64_t)a.cont_index[c] * a.rows + r0 + r)));   // the product first
                v = fmin(fmax(v, -1.0), 1.0);                       // numerical.py:448
                const int k = col.valid_idx[best];
                res = __dadd_rn(__dmul_rn(__dmul_rn(v, 4.0), col.std[k]), col.mean[k]);   // numerical.py:453-455, left to right, no FMA
                if (col.round_int) res = rint(res);                 // numerical.py:175-177 (np.round: half to even)
            } else {
                int best = 0; float bv = d[0];
                for (int j = 1; j < col.n_valid; ++j) if (d[j] > bv) { bv = d[j]; best = j; }
                res = col.category[best];
            }
            sraw[r * C + c] = res;
        }
        __syncthreads();
        tile_copy<double, false>(sraw, C, a.raw_out + r0 * a.ld_raw, a.ld_raw, nr, C);
    }
}

// inference_tvae.py:232-235, :250-253 — one row per thread (n_class <= 16), fp32 like the reference's torch ops.
__global__ void gumbel_argmax_kernel(const float* __restrict__ logits, int64_t ld, int n_class, const float* __restrict__ U,
                                     int64_t rows, int64_t* __restrict__ out) {
    const float eps = 1e-20f;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const float* x = logits + r * ld;
        float m = x[0];
        for (int j = 1; j < n_class; ++j) m = fmaxf(m, x[j]);
        float s = 0.f;
        for (int j = 0; j < n_class; ++j) s += expf(x[j] - m);
        const float ls = logf(s);
        int best = 0; float bv = -INFINITY;
        for (int j = 0; j < n_class; ++j) {
            const float g = logf(-logf(U[r * n_class + j] + eps) + eps);   // the reference's sign: log(-log U)
            const float v = (x[j] - m - ls) + g;
            if (v > bv) { bv = v; best = j; }
        }
        out[r] = best;
    }
}

int check_config(const cdg_tvae_transform_config* cfg, TransformArgs& a) {
    CDG_REQUIRE(cfg != nullptr, "tvae transform: null config");
    CDG_REQUIRE(cfg->n_col >= 1 && cfg->n_col <= CDG_MAX_TCOL, "tvae transform: n_col %d out of range", cfg->n_col);
    int start = 0, nc = 0;
    for (int c = 0; c < cfg->n_col; ++c) {
        const cdg_tvae_column& col = cfg->col[c];
        CDG_REQUIRE(col.out_start == start, "tvae transform: column %d starts at %d, expected %d", c, col.out_start, start);
        if (col.kind == CDG_TCOL_CONTINUOUS) {
            CDG_REQUIRE(col.n_all >= 1 && col.n_all <= CDG_MAX_TCOMP && col.n_valid >= 1 && col.n_valid <= col.n_all,
                        "tvae transform: column %d has %d / %d components", c, col.n_valid, col.n_all);
            for (int j = 0; j < col.n_valid; ++j)
                CDG_REQUIRE(col.valid_idx[j] >= 0 && col.valid_idx[j] < col.n_all, "tvae transform: column %d bad component index", c);
            a.cont_index[c] = nc++;
            start += 1 + col.n_valid;
        } else if (col.kind == CDG_TCOL_DISCRETE) {
            CDG_REQUIRE(col.n_valid >= 1 && col.n_valid <= CDG_MAX_TCAT, "tvae transform: column %d has %d categories", c, col.n_valid);
            a.cont_index[c] = -1;
            start += col.n_valid;
        } else {
            CDG_REQUIRE(false, "tvae transform: column %d has unknown kind %d", c, col.kind);
        }
    }
    CDG_REQUIRE(cfg->out_dim == start, "tvae transform: out_dim %d, blocks sum to %d", cfg->out_dim, start);
    a.c = *cfg;
    return CDG_OK;
}

unsigned tile_grid(int64_t rows, size_t smem) {
    const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    int per_sm = (int)imin64(8, (int64_t)(200 * 1024) / (int64_t)(smem + 1024));
    if (per_sm < 1) per_sm = 1;
    return (unsigned)imax64(1, imin64(tiles, (int64_t)kNumSMs * per_sm));
}

}  // namespace
}  // namespace cdg

using namespace cdg;

extern "C" int cdg_tvae_transform(const cdg_tvae_transform_config* cfg, const double* raw, int64_t ld_raw, const double* uniforms,
                                  int64_t rows, float* out, int64_t ld_out, void* stream) {
    TransformArgs a{};
    CDG_TRY(check_config(cfg, a));
    CDG_REQUIRE(rows >= 0 && ld_raw >= cfg->n_col && ld_out >= cfg->out_dim, "tvae transform: bad extents");
    if (rows == 0) return CDG_OK;
    CDG_REQUIRE(raw && out, "tvae transform: null table pointer");
    bool any_cont = false;
    for (int c = 0; c < cfg->n_col; ++c) any_cont |= cfg->col[c].kind == CDG_TCOL_CONTINUOUS;
    CDG_REQUIRE(!any_cont || uniforms, "tvae transform: continuous columns need the injected uniforms");
    a.raw = raw; a.ld_raw = ld_raw; a.rnd = uniforms; a.out = out; a.ld_out = ld_out; a.rows = rows;
    a.raw_pitch = cfg->n_col | 1;
    for (int c = 0; c < cfg->n_col; ++c) {
        const cdg_tvae_column& col = cfg->col[c];
        int order[CDG_MAX_TCOMP], n = 0;
        bool kept[CDG_MAX_TCOMP] = {};
        if (col.kind == CDG_TCOL_CONTINUOUS) {
            for (int j = 0; j < col.n_valid; ++j) { order[n++] = col.valid_idx[j]; kept[col.valid_idx[j]] = true; }
            for (int k = 0; k < col.n_all; ++k) if (!kept[k]) order[n++] = k;
        }
        for (int k = 0; k < CDG_MAX_TCOMP; ++k) {
            const bool used = k < n;
            a.log_a_f[c][k] = used ? (float)col.log_a[order[k]] : -INFINITY;
            a.prec_f[c][k] = used ? (float)col.prec[order[k]] : 0.f;
            a.mean_p[c][k] = used ? col.mean[order[k]] : 0.0;
        }
    }
    { const char* e = getenv("CDG_TVAE_EXACT_ONLY"); a.exact_only = (e && atoi(e) != 0) ? 1 : 0; }
    const size_t per_warp = (size_t)WARP_ROWS * a.raw_pitch * sizeof(double) + (((size_t)WARP_ROWS * cfg->out_dim * sizeof(float) + 15) & ~(size_t)15);
    const size_t smem = per_warp * (TT_THREADS / 32);
    CDG_CHECK_CUDA(cudaFuncSetAttribute(tvae_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)imin64(3, (int64_t)(220 * 1024) / (int64_t)(smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int64_t strips = (rows + WARP_ROWS - 1) / WARP_ROWS;
    const unsigned blocks = (unsigned)imax64(1, imin64((strips + TT_THREADS / 32 - 1) / (TT_THREADS / 32), (int64_t)kNumSMs * per_sm));
    tvae_transform_kernel<<<blocks, TT_THREADS, smem, (cudaStream_t)stream>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

extern "C" int cdg_tvae_inverse_transform(const cdg_tvae_transform_config* cfg, const float* data, int64_t ld_data,
                                          const float* sigmas, const double* normals, int64_t rows, double* raw_out,
                                          int64_t ld_raw, void* stream) {
    TransformArgs a{};
    CDG_TRY(check_config(cfg, a));
    CDG_REQUIRE(rows >= 0 && ld_raw >= cfg->n_col && ld_data >= cfg->out_dim, "tvae inverse transform: bad extents");
    if (rows == 0) return CDG_OK;
    CDG_REQUIRE((sigmas == nullptr) == (normals == nullptr), "tvae inverse transform: sigmas and normals go together");
    CDG_REQUIRE(data && raw_out, "tvae inverse transform: null table pointer");
    a.data = data; a.ld_data = ld_data; a.sigmas = sigmas; a.rnd = normals; a.raw_out = raw_out; a.ld_raw = ld_raw; a.rows = rows;
    const size_t smem = (size_t)TILE_ROWS * cfg->n_col * sizeof(double) + (size_t)TILE_ROWS * cfg->out_dim * sizeof(float);
    CDG_CHECK_CUDA(cudaFuncSetAttribute(tvae_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tvae_inverse_kernel<<<tile_grid(rows, smem), TT_THREADS, smem, (cudaStream_t)stream>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

extern "C" int cdg_gumbel_argmax(const float* logits, int64_t ld, int32_t n_class, const float* uniforms, int64_t rows,
                                 int64_t* out_index, void* stream) {
    CDG_REQUIRE(n_class >= 1 && n_class <= 64 && ld >= n_class && rows >= 0, "gumbel argmax: bad extents");
    if (rows == 0) return CDG_OK;
    CDG_REQUIRE(logits && uniforms && out_index, "gumbel argmax: null pointer");
    const unsigned blocks = (unsigned)imax64(1, imin64((rows + 255) / 256, (int64_t)kNumSMs * 8));
    gumbel_argmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, ld, n_class, uniforms, rows, out_index);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
