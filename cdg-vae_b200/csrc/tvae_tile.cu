// CDG-TVAE step (tabular/modules/model.py:360-460, tabular/modules/train.py:245-320), warp-cooperative version.
//
// tvae_fixed_kernel gives every thread one table row: each FFMA of a Linear layer then needs its own (scalar, broadcast)
// shared-memory load of a weight, and the 2.9 k parameter-gradient products of a row are reduced over the warp with ~4
// instructions each (SASS: 4.2 k LDS + 2.7 k SHFL + 4.1 k FSEL for 4.3 k FFMA per row; 392 M rows/s = 9 % of the fp32 peak).
// Here a WARP owns a tile of 32 rows whose activations live in shared memory as [feature][32 rows], and every Linear layer
// -- forward, input gradient and weight gradient alike -- is a register-tiled product over that tile:
//   forward / input gradient: a lane computes 4 rows x 4 outputs; per contraction step one 16-byte load of 4 activations
//       and one of 4 weights feed 16 FFMA (weights: a transposed copy [in][out] for the forward pass, the parameter
//       arena's own [out][in] rows for the input gradient);
//   weight gradient: a lane owns 2 x 4 (out, in) pairs and contracts over the 32 rows with 16-byte loads along the
//       row dimension (start offset rotated per lane: conflict-free), then adds its 8 totals to the block's gradient copy
//       in shared memory -- no shuffles at all.
// The element-wise stages (reparameterisation, causal block, flows, span losses) run with lane = row on the same slabs.
// Gradient slabs overwrite the activation slabs they belong to (d loss / d pre-activation of a layer replaces the layer's
// output once its weight gradient has been taken).  Same arithmetic per element as the generic kernel; the summation
// order of a batch sum differs (parity contract 1e-4, measured ~1e-6).
#include "latent.cuh"
#include "tabular_args.cuh"

namespace cdg {

namespace {

constexpr int TT_H0 = 32, TT_H1 = 16, TT_H2 = 16, TT_D1 = 8, TT_D2 = 8, TT_D3 = 16;
constexpr int TT_MAXM = 32;           // widest decoder output handled (reference shapes: <= 13)
constexpr int TT_MAXD = 64;

__host__ __device__ constexpr int pad4(int v) { return (v + 3) & ~3; }

struct TileLayout {                    // feature-row offsets inside a warp's slab (rows of 32 floats)
    int x, h0, h1, h2, ml, lat, a1, a2, a3, xh, rows;
};
__host__ __device__ inline TileLayout tile_layout(int D, int d, int max_m) {
    TileLayout t;
    int r = 0;
    t.x = r; r += pad4(D);
    t.h0 = r; r += TT_H0;
    t.h1 = r; r += TT_H1;
    t.h2 = r; r += TT_H2;
    t.ml = r; r += pad4(2 * d);
    t.lat = r; r += pad4(5 * d) + 4;   // nz[d] u[d] gal[d] z[d] gz[d] (+4: the 1-wide decoder input is read as a 4-row group)
    t.a1 = r; r += TT_D1;
    t.a2 = r; r += TT_D2;
    t.a3 = r; r += TT_D3;
    t.xh = r; r += pad4(max_m);
    t.rows = r;
    return t;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// out[o][r] (o < OUT4, all 32 rows) = sum_{i < K} in[i][r] * W[i * ldw + o]      W: [K][ldw] with 16-byte aligned rows
//   MODE 0: + bias[o], optional ReLU                                   (forward; W = transposed copy)
//   MODE 1: times (old out[o][r] > 0)                                   (input gradient through a ReLU; W = arena rows)
//   KT > 0: the contraction length is known at compile time (fully unrolled: the loads of a whole tile are in flight at once)
template <int MODE, bool RELU, int KT = 0>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ in, int Krt, const float* __restrict__ W, int ldw,
                                          const float* __restrict__ bias, float* __restrict__ out, int OUT4) {
    const int K = KT > 0 ? KT : Krt;
    const int lane = threadIdx.x & 31, rg = lane & 7, og = lane >> 3;      // a quarter-warp: one output group, 8 row groups
    for (int o0 = og * 4; o0 < OUT4; o0 += 16) {
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[j][r] = 0.f;
        const float* ip = in + rg * 4;
        const float* wp = W + o0;
#pragma unroll 4
        for (int i = 0; i < K; ++i) {
            const float4 a = ld4(ip + i * 32);
            const float4 w = ld4(wp + i * ldw);
            acc[0][0] = fmaf(w.x, a.x, acc[0][0]); acc[0][1] = fmaf(w.x, a.y, acc[0][1]);
            acc[0][2] = fmaf(w.x, a.z, acc[0][2]); acc[0][3] = fmaf(w.x, a.w, acc[0][3]);
            acc[1][0] = fmaf(w.y, a.x, acc[1][0]); acc[1][1] = fmaf(w.y, a.y, acc[1][1]);
            acc[1][2] = fmaf(w.y, a.z, acc[1][2]); acc[1][3] = fmaf(w.y, a.w, acc[1][3]);
            acc[2][0] = fmaf(w.z, a.x, acc[2][0]); acc[2][1] = fmaf(w.z, a.y, acc[2][1]);
            acc[2][2] = fmaf(w.z, a.z, acc[2][2]); acc[2][3] = fmaf(w.z, a.w, acc[2][3]);
            acc[3][0] = fmaf(w.w, a.x, acc[3][0]); acc[3][1] = fmaf(w.w, a.y, acc[3][1]);
            acc[3][2] = fmaf(w.w, a.z, acc[3][2]); acc[3][3] = fmaf(w.w, a.w, acc[3][3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float* op = out + (o0 + j) * 32 + rg * 4;
            float4 v = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
            if (MODE == 0) {
                const float b = bias[o0 + j];
                v.x += b; v.y += b; v.z += b; v.w += b;
                if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            } else {
                const float4 h = ld4(op);
                v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f;
                v.z = h.z > 0.f ? v.z : 0.f; v.w = h.w > 0.f ? v.w : 0.f;
            }
            st4(op, v);
        }
    }
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float s) {
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); return fmaf(a.w, b.w, s);
}

// sgw[o * ldg + i] += sum_r delta[o][r] * hin[i][r]   (o < OUT, i < IN);   sgb[o] += sum_r delta[o][r]
// A lane owns TO x TI (out, in) pairs; OG lanes side by side along `out`: one pass covers OG * TO outputs x (32 / OG) * TI inputs.
// The 16-byte column a lane starts its walk over the 32 rows with is rotated by the lane index, so the eight lanes of a
// quarter-warp always hit eight different bank groups whatever rows they read.
template <int OUTT = 0, int INT_ = 0, int TO = 2, int TI = 4, int OG = 8>
__device__ __forceinline__ void warp_wgrad(const float* __restrict__ delta, int OUTrt, const float* __restrict__ hin, int INrt,
                                           float* sgw, int ldg, float* sgb) {
    const int OUT = OUTT > 0 ? OUTT : OUTrt, IN = INT_ > 0 ? INT_ : INrt;
    const int lane = threadIdx.x & 31, og = lane % OG, ig = lane / OG;
    constexpr int PO = OG * TO, PI = (32 / OG) * TI;
    for (int ob = 0; ob < OUT; ob += PO) {
        const int o0 = ob + og * TO;
        for (int ib = 0; ib < IN; ib += PI) {
            const int i0 = ib + ig * TI;
            if (o0 < OUT && i0 < IN) {
                float acc[TO][TI], bsum[TO];
#pragma unroll
                for (int a = 0; a < TO; ++a) {
                    bsum[a] = 0.f;
#pragma unroll
                    for (int j = 0; j < TI; ++j) acc[a][j] = 0.f;
                }
                const float* dp = delta + o0 * 32;
                const float* hp = hin + i0 * 32;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int q = ((it + lane) & 7) * 4;
                    float4 dv[TO];
#pragma unroll
                    for (int a = 0; a < TO; ++a) {
                        dv[a] = ld4(dp + a * 32 + q);
                        bsum[a] += (dv[a].x + dv[a].y) + (dv[a].z + dv[a].w);     // the bias gradient rides along (used by i0 == 0)
                    }
#pragma unroll
                    for (int j = 0; j < TI; ++j) {
                        const float4 h = ld4(hp + j * 32 + q);
#pragma unroll
                        for (int a = 0; a < TO; ++a) acc[a][j] = dot4(dv[a], h, acc[a][j]);
                    }
                }
#pragma unroll
                for (int a = 0; a < TO; ++a) {
#pragma unroll
                    for (int j = 0; j < TI; ++j)
                        if (o0 + a < OUT && i0 + j < IN) atomicAdd(sgw + (o0 + a) * ldg + i0 + j, acc[a][j]);
                    if (i0 == 0 && o0 + a < OUT) atomicAdd(sgb + o0 + a, bsum[a]);
                }
            }
        }
    }
}

struct TileSmem {                       // offsets (floats) into the dynamic shared memory block
    int sp, sg, wt, slab;
};

template <int DN>
__global__ void __launch_bounds__(256, 1) tvae_tile_kernel(TabArgs a, int wt_total) {
    extern __shared__ __align__(16) float smem[];
    const cdg_tabular_config& c = a.c;
    constexpr int d = DN;
    const int np = (int)c.n_params, D = c.input_dim;
    const int np4 = pad4(np);
    float* sp = smem;                                  // parameters, arena layout
    float* sg = sp + np4;                              // the block's gradient copy
    float* wt = sg + np4;                              // transposed weights [in][pad4(out)] per layer
    int max_m = 0;
    for (int k = 0; k < d; ++k) max_m = max(max_m, c.dec[k][3].out);
    const TileLayout T = tile_layout(D, d, max_m);
    float* slab = wt + wt_total + (threadIdx.x >> 5) * T.rows * 32;
    __shared__ FlowTable ft;
    __shared__ double dred[32];
    __shared__ float fred[32];
    __shared__ int wt_enc[4], wt_dec[CDG_MAX_DEC][4], span_lo[CDG_MAX_DEC + 1];

    for (int i = threadIdx.x; i < np; i += blockDim.x) { sp[i] = a.params[i]; sg[i] = 0.f; }
    {
        struct { int d, scm, flow_num; const float* params; const int64_t* flow_off; const float* A; } fa =
            {DN, c.scm, c.flow_num, a.params, c.flow_off, c.I_B_inv};
        load_flow_table(ft, fa);
    }
    if (threadIdx.x == 0) {
        int o = 0;
        for (int l = 0; l < 4; ++l) { wt_enc[l] = o; o += c.enc[l].in * pad4(c.enc[l].out); }
        for (int k = 0; k < d; ++k)
            for (int l = 0; l < 4; ++l) { wt_dec[k][l] = o; o += c.dec[k][l].in * pad4(c.dec[k][l].out); }
        // spans are listed in column order (train.py:270-285 walks them with a running offset): decoder k owns a contiguous run
        int sidx = 0, col = 0;
        for (int k = 0; k < d; ++k) {
            span_lo[k] = sidx;
            col += c.dec[k][3].out;
            while (sidx < c.n_span && c.span_start[sidx] < col) ++sidx;
        }
        span_lo[d] = sidx;
    }
    __syncthreads();
    for (int l = 0; l < 4 + 4 * d; ++l) {
        const cdg_linear& L = l < 4 ? c.enc[l] : c.dec[(l - 4) >> 2][(l - 4) & 3];
        const int off = l < 4 ? wt_enc[l] : wt_dec[(l - 4) >> 2][(l - 4) & 3];
        const int o4 = pad4(L.out);
        for (int e = threadIdx.x; e < L.in * o4; e += blockDim.x) {
            const int i = e / o4, o = e - i * o4;
            wt[off + e] = o < L.out ? sp[L.w + o * L.in + i] : 0.f;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const float invB = 1.f / (float)a.batch;
    double rec_acc = 0.0, kl_acc = 0.0, al_acc = 0.0;
    float var_acc[d];
#pragma unroll
    for (int i = 0; i < d; ++i) var_acc[i] = 0.f;
    FlowGrad fg;
    fg.clear();

    float* X = slab + T.x * 32;   float* H0 = slab + T.h0 * 32; float* H1 = slab + T.h1 * 32; float* H2 = slab + T.h2 * 32;
    float* ML = slab + T.ml * 32; float* LAT = slab + T.lat * 32;
    float* A1 = slab + T.a1 * 32; float* A2 = slab + T.a2 * 32; float* A3 = slab + T.a3 * 32; float* XH = slab + T.xh * 32;
    float* NZ = LAT; float* U = LAT + d * 32; float* GAL = LAT + 2 * d * 32; float* Z = LAT + 3 * d * 32; float* GZ = LAT + 4 * d * 32;

    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t ntiles = (a.batch + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += warps_total) {
        const int64_t b = tile * 32 + lane;
        const bool valid = b < a.batch;
        const float vm = valid ? 1.f : 0.f;
        const int64_t br = valid ? b : 0;
        const float* xrow = a.x + br * D;
        for (int i = 0; i < D; ++i) X[i * 32 + lane] = __ldg(xrow + i);
        __syncwarp();

        // ---- encoder D-32-16-16-2d (ReLU) ----
        warp_gemm<0, true>(X, D, wt + wt_enc[0], TT_H0, sp + c.enc[0].b, H0, TT_H0);
        __syncwarp();
        warp_gemm<0, true, TT_H0>(H0, TT_H0, wt + wt_enc[1], TT_H1, sp + c.enc[1].b, H1, TT_H1);
        __syncwarp();
        warp_gemm<0, true, TT_H1>(H1, TT_H1, wt + wt_enc[2], TT_H2, sp + c.enc[2].b, H2, TT_H2);
        __syncwarp();
        warp_gemm<0, false, TT_H2>(H2, TT_H2, wt + wt_enc[3], pad4(2 * d), sp + c.enc[3].b, ML, pad4(2 * d));
        __syncwarp();

        // ---- latent block, lane = row (model.py:418-437, train.py:287-303) ----
        {
            float mean[CDG_MAX_NODE], lv[CDG_MAX_NODE], nz[CDG_MAX_NODE], eps[CDG_MAX_NODE], u[CDG_MAX_NODE], z[CDG_MAX_NODE];
            float u2[CDG_MAX_NODE], z2[CDG_MAX_NODE], gal[CDG_MAX_NODE], gu2[CDG_MAX_NODE];
            float kl = 0.f, al = 0.f;
#pragma unroll
            for (int i = 0; i < CDG_MAX_NODE; ++i) {
                mean[i] = lv[i] = nz[i] = eps[i] = 0.f;
                if (i < d) {
                    mean[i] = ML[(i < d ? i : 0) * 32 + lane]; lv[i] = ML[(i < d ? d + i : 0) * 32 + lane];
                    nz[i] = a.deterministic ? 0.f : a.noise[br * d + i];
                    const float ev = expf(lv[i]);
                    eps[i] = a.deterministic ? mean[i] : mean[i] + expf(lv[i] / 2.f) * nz[i];
                    kl += mean[i] * mean[i] - lv[i] + ev;
                    var_acc[i < d ? i : 0] += vm * ev;
                }
            }
            kl_acc += (double)(vm * 0.5f * (kl - (float)d));
            matvec_A(ft, d, eps, u);
            matvec_A(ft, d, mean, u2);
            const float ascale = c.lambda_ * invB;
#pragma unroll
            for (int j = 0; j < CDG_MAX_NODE; ++j) {
                z[j] = z2[j] = gu2[j] = 0.f;
                if (j < d) {
                    z[j] = flow_fwd(ft, c.scm, c.flow_num, j, u[j]);
                    z2[j] = flow_fwd(ft, c.scm, c.flow_num, j, u2[j]);
                    if (a.y) {
                        const float yh = 1.f / (1.f + expf(-z2[j]));
                        const float yy = a.y[br * d + j];
                        al += (yy - 1.f) * fmaxf(log1pf(-yh), -100.f) - yy * fmaxf(logf(yh), -100.f);
                        const float gzz = vm * ascale * (yh - yy) / fmaxf((1.f - yh) * yh, 1e-12f) * ((1.f - yh) * yh);
                        if (a.do_bwd) gu2[j] = flow_bwd(ft, c.scm, c.flow_num, j, u2[j], gzz, fg);
                    }
                }
            }
            al_acc += (double)(vm * al);
            matvec_AT(ft, d, gu2, gal);
#pragma unroll
            for (int i = 0; i < d; ++i) {
                NZ[i * 32 + lane] = nz[i]; U[i * 32 + lane] = u[i]; GAL[i * 32 + lane] = gal[i]; Z[i * 32 + lane] = z[i];
            }
            if (a.latents && valid) {
                float* o = a.latents + b * 6 * d;
#pragma unroll
                for (int i = 0; i < d; ++i) {
                    o[i] = mean[i]; o[d + i] = lv[i]; o[2 * d + i] = eps[i]; o[3 * d + i] = u[i]; o[4 * d + i] = z[i];
                    o[5 * d + i] = z2[i];
                }
            }
        }
        __syncwarp();

        // ---- decoders 1-8-8-16-m_k, one at a time: forward, span losses, backward ----
        float rec = 0.f;
        int col = 0;
        for (int k = 0; k < d; ++k) {
            const cdg_linear& L0 = c.dec[k][0]; const cdg_linear& L1 = c.dec[k][1];
            const cdg_linear& L2 = c.dec[k][2]; const cdg_linear& L3 = c.dec[k][3];
            const int m = L3.out, m4 = pad4(m);
            warp_gemm<0, true, 1>(Z + k * 32, 1, wt + wt_dec[k][0], TT_D1, sp + L0.b, A1, TT_D1);
            __syncwarp();
            warp_gemm<0, true, TT_D1>(A1, TT_D1, wt + wt_dec[k][1], TT_D2, sp + L1.b, A2, TT_D2);
            __syncwarp();
            warp_gemm<0, true, TT_D2>(A2, TT_D2, wt + wt_dec[k][2], TT_D3, sp + L2.b, A3, TT_D3);
            __syncwarp();
            warp_gemm<0, false, TT_D3>(A3, TT_D3, wt + wt_dec[k][3], m4, sp + L3.b, XH, m4);
            __syncwarp();
            if (a.xhat && valid)
                for (int j = 0; j < m; ++j) a.xhat[b * a.out_total + col + j] = XH[j * 32 + lane];

            // span losses of this decoder's columns (tabular/modules/train.py:270-285); d loss / d xhat replaces xhat
            for (int sidx = span_lo[k]; sidx < span_lo[k + 1]; ++sidx) {
                const int st = c.span_start[sidx], dim = c.span_dim[sidx];
                // a decoder's width is a sum of whole spans (main_tvae.py:174-192); the forward-only API's placeholder span
                // (one softmax over all columns, its loss is never read) does not fit a decoder and is skipped
                if (st - col + dim > m) continue;
                float* xs = XH + (st - col) * 32 + lane;
                const float* tg = X + st * 32 + lane;
                if (c.span_kind[sidx] == CDG_SPAN_TANH) {
                    const float sd = sp[c.sigma_off + st];
                    const float th = tanhf(xs[0]);
                    const float r = tg[0] - th;
                    rec += r * r / 2.f / (sd * sd) + logf(sd);
                    xs[0] = -(r / (sd * sd)) * (1.f - th * th) * invB * vm;
                    if (a.do_bwd) {
                        const float t = warp_sum(vm * (-(r * r) / (sd * sd * sd) + 1.f / sd) * invB);
                        if (lane == 0) atomicAdd(sg + c.sigma_off + st, t);
                    }
                } else {
                    int tgt = 0;
                    float best = tg[0], mx = xs[0];
                    for (int j = 1; j < dim; ++j) {
                        const float xv = tg[j * 32];
                        if (xv > best) { best = xv; tgt = j; }
                        mx = fmaxf(mx, xs[j * 32]);
                    }
                    float se = 0.f;
                    for (int j = 0; j < dim; ++j) se += expf(xs[j * 32] - mx);
                    const float lse = mx + logf(se);
                    rec += lse - xs[tgt * 32];
                    for (int j = 0; j < dim; ++j)
                        xs[j * 32] = (expf(xs[j * 32] - lse) - (j == tgt ? 1.f : 0.f)) * invB * vm;
                }
            }
            __syncwarp();
            if (a.do_bwd) {
                warp_wgrad<0, TT_D3>(XH, m, A3, TT_D3, sg + L3.w, TT_D3, sg + L3.b);
                warp_gemm<1, false>(XH, m, sp + L3.w, TT_D3, nullptr, A3, TT_D3);
                __syncwarp();
                warp_wgrad<TT_D3, TT_D2, 2, 2, 8>(A3, TT_D3, A2, TT_D2, sg + L2.w, TT_D2, sg + L2.b);
                warp_gemm<1, false, TT_D3>(A3, TT_D3, sp + L2.w, TT_D2, nullptr, A2, TT_D2);
                __syncwarp();
                warp_wgrad<TT_D2, TT_D1, 1, 2, 8>(A2, TT_D2, A1, TT_D1, sg + L1.w, TT_D1, sg + L1.b);
                warp_gemm<1, false, TT_D2>(A2, TT_D2, sp + L1.w, TT_D1, nullptr, A1, TT_D1);
                __syncwarp();
                warp_wgrad<TT_D1, 1, 1, 1, 8>(A1, TT_D1, Z + k * 32, 1, sg + L0.w, 1, sg + L0.b);
                float gz = 0.f;
#pragma unroll
                for (int o = 0; o < TT_D1; ++o) gz = fmaf(A1[o * 32 + lane], sp[L0.w + o], gz);
                GZ[k * 32 + lane] = gz;
                __syncwarp();
            }
            col += m;
        }
        rec_acc += (double)(vm * rec);
        if (!a.do_bwd) continue;

        // ---- latent backward, lane = row: d loss / d [mean | logvar] replaces ML ----
        {
            float gu[CDG_MAX_NODE], ge[CDG_MAX_NODE];
#pragma unroll
            for (int j = 0; j < CDG_MAX_NODE; ++j)
                gu[j] = j < d ? flow_bwd(ft, c.scm, c.flow_num, j, U[(j < d ? j : 0) * 32 + lane], GZ[(j < d ? j : 0) * 32 + lane], fg) : 0.f;
            matvec_AT(ft, d, gu, ge);
            const float kscale = c.beta * invB;
#pragma unroll
            for (int i = 0; i < d; ++i) {
                const float mean = ML[i * 32 + lane], lv = ML[(d + i) * 32 + lane];
                ML[i * 32 + lane] = ge[i] + vm * kscale * mean + GAL[i * 32 + lane];
                ML[(d + i) * 32 + lane] = 0.5f * ge[i] * NZ[i * 32 + lane] * expf(lv / 2.f) + vm * 0.5f * kscale * (expf(lv) - 1.f);
            }
        }
        __syncwarp();

        // ---- encoder backward ----
        warp_wgrad<2 * d, TT_H2>(ML, 2 * d, H2, TT_H2, sg + c.enc[3].w, TT_H2, sg + c.enc[3].b);
        warp_gemm<1, false, 2 * d>(ML, 2 * d, sp + c.enc[3].w, TT_H2, nullptr, H2, TT_H2);
        __syncwarp();
        warp_wgrad<TT_H2, TT_H1>(H2, TT_H2, H1, TT_H1, sg + c.enc[2].w, TT_H1, sg + c.enc[2].b);
        warp_gemm<1, false, TT_H2>(H2, TT_H2, sp + c.enc[2].w, TT_H1, nullptr, H1, TT_H1);
        __syncwarp();
        warp_wgrad<TT_H1, TT_H0, 4, 4, 4>(H1, TT_H1, H0, TT_H0, sg + c.enc[1].w, TT_H0, sg + c.enc[1].b);
        warp_gemm<1, false, TT_H1>(H1, TT_H1, sp + c.enc[1].w, TT_H0, nullptr, H0, TT_H0);
        __syncwarp();
        warp_wgrad<TT_H0, 0, 4, 4, 8>(H0, TT_H0, X, D, sg + c.enc[0].w, D, sg + c.enc[0].b);
        __syncwarp();
    }

    // ---- block reductions ----
    if (a.acc) {
        double s = block_sum<double>(rec_acc, dred);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_RECON, s);
        s = block_sum<double>(kl_acc, dred);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_KL, s);
        s = block_sum<double>(al_acc, dred);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_ALIGN, s);
#pragma unroll
        for (int i = 0; i < d; ++i) {
            s = block_sum<double>((double)var_acc[i], dred);
            if (threadIdx.x == 0) atomicAdd(a.acc + ACC_VAR + i, s);
        }
    }
    if (a.do_bwd) {
        struct { int d, scm, flow_num; float* grads; const int64_t* flow_off; } ra = {d, c.scm, c.flow_num, a.grads, c.flow_off};
        reduce_flow_grads(fg, ft, ra, fred);
        __syncthreads();
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const float v = sg[i];
            if (v != 0.f) atomicAdd(a.grads + i, v);
        }
    }
}

bool tile_shape_ok(const cdg_tabular_config& c) {
    if (c.kind != CDG_TAB_TVAE || c.act != CDG_ACT_RELU || c.n_enc_layers != 4 || c.n_dec_layers != 4) return false;
    if (c.n_dec != c.node || (c.node != 3 && c.node != 6) || c.input_dim > TT_MAXD) return false;
    const int e[5] = {c.input_dim, TT_H0, TT_H1, TT_H2, 2 * c.node};
    for (int l = 0; l < 4; ++l)
        if (c.enc[l].in != e[l] || c.enc[l].out != e[l + 1] || (c.enc[l].w & 3) || (c.enc[l].b & 3)) return false;
    for (int k = 0; k < c.n_dec; ++k) {
        if (c.factor[k] != 1 || c.out_dim[k] > TT_MAXM) return false;
        const int dd[5] = {1, TT_D1, TT_D2, TT_D3, c.out_dim[k]};
        for (int l = 0; l < 4; ++l)
            if (c.dec[k][l].in != dd[l] || c.dec[k][l].out != dd[l + 1] || (c.dec[k][l].w & 3) || (c.dec[k][l].b & 3)) return false;
    }
    return true;
}

}  // namespace

bool launch_tvae_tile(const TabArgs& a, cudaStream_t s) {
    const cdg_tabular_config& c = a.c;
    if (!tile_shape_ok(c)) return false;
    int wt_total = 0;
    for (int l = 0; l < 4; ++l) wt_total += c.enc[l].in * pad4(c.enc[l].out);
    for (int k = 0; k < c.n_dec; ++k)
        for (int l = 0; l < 4; ++l) wt_total += c.dec[k][l].in * pad4(c.dec[k][l].out);
    int max_m = 0;
    for (int k = 0; k < c.n_dec; ++k) max_m = c.out_dim[k] > max_m ? c.out_dim[k] : max_m;
    const TileLayout T = tile_layout(c.input_dim, c.node, max_m);
    const size_t fixed = sizeof(float) * (2 * (size_t)pad4((int)c.n_params) + wt_total);
    const size_t slab = sizeof(float) * 32 * (size_t)T.rows;
    const size_t budget = 220 * 1024;
    if (fixed + slab > budget) return false;
    int warps = (int)((budget - fixed) / slab);
    if (warps > 8) warps = 8;
    const size_t smem = fixed + slab * warps;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tvae_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess ||
            cudaFuncSetAttribute(tvae_tile_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr = true;
    }
    const int64_t ntiles = (a.batch + 31) / 32;
    int64_t blocks = (ntiles + warps - 1) / warps;
    if (blocks > kNumSMs) blocks = kNumSMs;
    if (c.node == 3) tvae_tile_kernel<3><<<(unsigned)blocks, 32 * warps, smem, s>>>(a, wt_total);
    else tvae_tile_kernel<6><<<(unsigned)blocks, 32 * warps, smem, s>>>(a, wt_total);
    return true;
}

}  // namespace cdg
