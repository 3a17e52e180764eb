// The three fixed tabular CDG-VAE networks (tabular/modules/model.py:245-305) with every width known at compile
// time: loan / adult  encoder 5-4-6,  decoders 1-2-m,  m = (2,2,1) / (1,1,3)   ( 87 parameters)
//                covtype       encoder 8-4-4-4-12, decoders 1-2-2-m, m = (1,1,2,1,1,8)  (250 live parameters)
// Same algorithm as tab_step_kernel (tabular.cu), but activations live in registers and there is no loop, address
// or configuration overhead, and the parameter-gradient products of a row are reduced over the warp 32 at a
// time with a 31-shuffle reduce-scatter: ncu on the generic kernel showed 7.9 k warp instructions per 32 rows of which only
// 15 % were FP32 math (BRA 11 %, uniform-datapath integer / constant loads 30 %).  Models with other shapes keep
// using the generic kernel.
#include "latent.cuh"
#include "tabular_args.cuh"

#include <utility>

namespace cdg {

struct NetLoan {
    static constexpr int KIND = CDG_TAB_LOAN, D = 5, DN = 3, EH = 4, NE = 2, DH = 2, ND = 2, K = 3, OUT = 5;
    __host__ __device__ static constexpr int m(int k) { return k == 0 ? 2 : k == 1 ? 2 : 1; }
    __host__ __device__ static constexpr int col(int k) { return k == 0 ? 0 : k == 1 ? 2 : 4; }
};
struct NetAdult {
    static constexpr int KIND = CDG_TAB_ADULT, D = 5, DN = 3, EH = 4, NE = 2, DH = 2, ND = 2, K = 3, OUT = 5;
    __host__ __device__ static constexpr int m(int k) { return k == 2 ? 3 : 1; }
    __host__ __device__ static constexpr int col(int k) { return k; }
};
struct NetCov {
    static constexpr int KIND = CDG_TAB_COVTYPE, D = 8, DN = 6, EH = 4, NE = 4, DH = 2, ND = 3, K = 6, OUT = 14;
    __host__ __device__ static constexpr int m(int k) { return k == 2 ? 2 : k == 5 ? 8 : 1; }
    __host__ __device__ static constexpr int col(int k) { return k < 3 ? k : k == 3 ? 4 : k == 4 ? 5 : 6; }
};

// CP ("constant parameters"): the arena is copied into __constant__ memory in front of the launch and every weight is a
// compile-time offset into it (FFMA reads it as a constant-bank operand: no load instruction; the shared-memory version spends
// 280 LDS per row on weights), exp / log / reciprocal run on the SFU, the alignment BCE is evaluated in its logits form (as in
// tabular_const.cu, same saturation behaviour).  Only for the arena layout the drop-in covtype model produces (checked at
// launch); any other layout keeps the shared-memory path.
constexpr int TF_MAX_PARAMS = 2048;
__constant__ float c_tabfix[TF_MAX_PARAMS];
struct LinOff { int w, b; };
template <bool CP> __device__ __forceinline__ float par(const float* sp, int idx) {
    if constexpr (CP) return c_tabfix[idx];
    else return sp[idx];
}
template <bool CP> __device__ __forceinline__ float ex(float x) {
    if constexpr (CP) return __expf(x);
    else return expf(x);
}
// canonical covtype arena (tabular/modules/model.py:245-298 in registration order, one or two 32-float slots per tensor)
struct CovOff {
    __host__ __device__ static constexpr int enc_w(int l) { return 64 * l; }
    __host__ __device__ static constexpr int enc_b(int l) { return l == 3 ? 256 : 64 * l + 32; }
    __host__ __device__ static constexpr int flow(int j) { return 288 + 32 * j; }
    __host__ __device__ static constexpr int dec_w(int k, int l) { return 480 + 192 * k + 64 * l; }
    __host__ __device__ static constexpr int dec_b(int k, int l) { return dec_w(k, l) + 32; }
    static constexpr int NPARAMS = 1920;
};
template <bool CP> __device__ __forceinline__ LinOff enc_lin(const cdg_tabular_config& c, int l) {
    if constexpr (CP) return LinOff{CovOff::enc_w(l), CovOff::enc_b(l)};
    else return LinOff{(int)c.enc[l].w, (int)c.enc[l].b};
}
template <bool CP> __device__ __forceinline__ LinOff dec_lin(const cdg_tabular_config& c, int k, int l) {
    if constexpr (CP) return LinOff{CovOff::dec_w(k, l), CovOff::dec_b(k, l)};
    else return LinOff{(int)c.dec[k][l].w, (int)c.dec[k][l].b};
}

// flat positions of the parameter-gradient products a thread collects before the warp reduce-scatter
template <class N> struct Pos {
    __host__ __device__ static constexpr int dec_np(int k) {
        return (N::DH + N::DH) + (N::ND == 3 ? N::DH * N::DH + N::DH : 0) + (N::DH * N::m(k) + N::m(k));
    }
    __host__ __device__ static constexpr int dec_off(int k) { return k == 0 ? 0 : dec_off(k - 1) + dec_np(k - 1); }
    static constexpr int NA = dec_off(N::K);                                    // decoder parameters
    static constexpr int NA_PAD = (NA + 31) / 32 * 32;
    static constexpr int E0 = N::D * N::EH + N::EH, EM = N::EH * N::EH + N::EH, EL = N::EH * 2 * N::DN + 2 * N::DN;
    static constexpr int NB = E0 + (N::NE == 4 ? 2 * EM : 0) + EL;             // encoder parameters
    static constexpr int NB_PAD = (NB + 31) / 32 * 32;
};

// Sum v[i] over the 32 lanes for i = 0..31 with 31 shuffles (instead of 32 x 5): on return lane l holds the total
// of v[l].  Step s exchanges the half a lane does not keep with its partner lane ^ s.
__device__ __forceinline__ float reduce_scatter32(float* v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? v[i] : v[i + s];
            const float keep = up ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}
template <int NPAD>
__device__ __forceinline__ void flush_products(float (&gp)[NPAD], const int* pos2off, float* sg) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int g = 0; g < NPAD / 32; ++g) {
        const float t = reduce_scatter32(&gp[g * 32]);
        const int off = pos2off[g * 32 + lane];
        if (off >= 0) atomicAdd(sg + off, t);
    }
}

template <int IN, int OUT, bool ACT, bool CP = false>
__device__ __forceinline__ void fc(const float* sp, const LinOff L, const float (&hin)[IN], float (&hout)[OUT]) {
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        float s = par<CP>(sp, L.b + o);
#pragma unroll
        for (int i = 0; i < IN; ++i) s = fmaf(par<CP>(sp, L.w + o * IN + i), hin[i], s);
        hout[o] = ACT ? (s > 0.f ? s : ex<CP>(s) - 1.f) : s;
    }
}
// delta: dL/d(pre-activation outputs).  Accumulates dW, db; gin = dL/d(hin) times ELU'(hin) when HIN_ACT.
// ACC: the products are added to gp (a thread keeps its products over all of its rows and the warp reduction runs once).
template <int IN, int OUT, bool HIN_ACT, bool NEED_GIN, int POS, bool ACC, bool CP = false, int NPAD>
__device__ __forceinline__ void fc_bwd(const float* sp, float (&gp)[NPAD], const LinOff L, const float (&hin)[IN],
                                       const float (&delta)[OUT], float (&gin)[IN]) {
    static_assert(POS + OUT * IN + OUT <= NPAD, "gradient product array too small");
#pragma unroll
    for (int i = 0; i < IN; ++i) gin[i] = 0.f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        gp[POS + OUT * IN + o] = ACC ? gp[POS + OUT * IN + o] + delta[o] : delta[o];                      // bias
#pragma unroll
        for (int i = 0; i < IN; ++i) {
            gp[POS + o * IN + i] = ACC ? fmaf(delta[o], hin[i], gp[POS + o * IN + i]) : delta[o] * hin[i];   // weight
            if (NEED_GIN) gin[i] = fmaf(delta[o], par<CP>(sp, L.w + o * IN + i), gin[i]);
        }
    }
    if (HIN_ACT && NEED_GIN) {
#pragma unroll
        for (int i = 0; i < IN; ++i) gin[i] *= hin[i] > 0.f ? 1.f : hin[i] + 1.f;
    }
}

// decoder k (compile-time index, so that its output width m(k) and column offset col(k) are constants)
template <class N, int k, bool CP>
__device__ __forceinline__ void dec_forward(const float* sp, const cdg_tabular_config& c, const float* z,
                                            float (&a1)[N::K][N::DH], float (&a2)[N::K][N::DH], float (&xh)[N::OUT]) {
    const float zin[1] = {z[k]};
    fc<1, N::DH, true, CP>(sp, dec_lin<CP>(c, k, 0), zin, a1[k]);
    float out[N::m(k)];
    if constexpr (N::ND == 3) {
        fc<N::DH, N::DH, true, CP>(sp, dec_lin<CP>(c, k, 1), a1[k], a2[k]);
        fc<N::DH, N::m(k), false, CP>(sp, dec_lin<CP>(c, k, 2), a2[k], out);
    } else {
        fc<N::DH, N::m(k), false, CP>(sp, dec_lin<CP>(c, k, 1), a1[k], out);
    }
#pragma unroll
    for (int j = 0; j < N::m(k); ++j) xh[N::col(k) + j] = out[j];
}
template <class N, int k, bool ACC, bool CP>
__device__ __forceinline__ void dec_backward(const float* sp, float (&gp)[Pos<N>::NA_PAD], const cdg_tabular_config& c,
                                             const float* z, float (&a1)[N::K][N::DH], float (&a2)[N::K][N::DH],
                                             const float (&gx)[N::OUT], float* gz) {
    float dout[N::m(k)], g2[N::DH], g1[N::DH], gzin[1];
#pragma unroll
    for (int j = 0; j < N::m(k); ++j) dout[j] = gx[N::col(k) + j];
    const float zin[1] = {z[k]};
    // flat order inside decoder k: layer 0 (weights, bias), [layer 1], last layer
    constexpr int P0 = Pos<N>::dec_off(k), P1 = P0 + 2 * N::DH, P2 = P1 + (N::ND == 3 ? N::DH * N::DH + N::DH : 0);
    if constexpr (N::ND == 3) {
        fc_bwd<N::DH, N::m(k), true, true, P2, ACC, CP>(sp, gp, dec_lin<CP>(c, k, 2), a2[k], dout, g2);
        fc_bwd<N::DH, N::DH, true, true, P1, ACC, CP>(sp, gp, dec_lin<CP>(c, k, 1), a1[k], g2, g1);
    } else {
        fc_bwd<N::DH, N::m(k), true, true, P2, ACC, CP>(sp, gp, dec_lin<CP>(c, k, 1), a1[k], dout, g1);
    }
    fc_bwd<1, N::DH, false, true, P0, ACC, CP>(sp, gp, dec_lin<CP>(c, k, 0), zin, g1, gzin);
    gz[k] = gzin[0];
}
// arena offsets of decoder k's parameters in the same flat order
template <class N, int k>
__device__ __forceinline__ void dec_positions(const cdg_tabular_config& c, int* pos2off) {
    int p = Pos<N>::dec_off(k);
    for (int l = 0; l < N::ND; ++l) {
        const cdg_linear& L = c.dec[k][l];
        for (int i = 0; i < L.in * L.out; ++i) pos2off[p++] = (int)L.w + i;
        for (int o = 0; o < L.out; ++o) pos2off[p++] = (int)L.b + o;
    }
}
template <class N, int... Ks>
__device__ __forceinline__ void dec_positions_all(std::integer_sequence<int, Ks...>, const cdg_tabular_config& c, int* pos2off) {
    (dec_positions<N, Ks>(c, pos2off), ...);
}
template <class N, bool CP, int... Ks>
__device__ __forceinline__ void dec_forward_all(std::integer_sequence<int, Ks...>, const float* sp, const cdg_tabular_config& c,
                                                const float* z, float (&a1)[N::K][N::DH], float (&a2)[N::K][N::DH],
                                                float (&xh)[N::OUT]) {
    (dec_forward<N, Ks, CP>(sp, c, z, a1, a2, xh), ...);
}
template <class N, bool ACC, bool CP, int... Ks>
__device__ __forceinline__ void dec_backward_all(std::integer_sequence<int, Ks...>, const float* sp, float (&gp)[Pos<N>::NA_PAD],
                                                 const cdg_tabular_config& c, const float* z, float (&a1)[N::K][N::DH],
                                                 float (&a2)[N::K][N::DH], const float (&gx)[N::OUT], float* gz) {
    (dec_backward<N, Ks, ACC, CP>(sp, gp, c, z, a1, a2, gx, gz), ...);
}

template <class N, bool CP = false>
__global__ void __launch_bounds__(TAB_THREADS) tab_fixed_kernel(TabArgs a) {
    extern __shared__ float smem[];
    const cdg_tabular_config& c = a.c;
    const int np = (int)c.n_params;
    float* sp = smem;
    float* sg = smem + np;
    __shared__ FlowTable ft;
    __shared__ double dred[32];
    __shared__ float fred[32];
    using P = Pos<N>;
    __shared__ int posA[P::NA_PAD], posB[P::NB_PAD];     // flat product position -> arena offset (-1 = padding)
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sp[i] = a.params[i]; sg[i] = 0.f; }
    for (int i = threadIdx.x; i < P::NA_PAD; i += blockDim.x) posA[i] = -1;
    for (int i = threadIdx.x; i < P::NB_PAD; i += blockDim.x) posB[i] = -1;
    {
        struct { int d, scm, flow_num; const float* params; const int64_t* flow_off; const float* A; } fa =
            {N::DN, c.scm, c.flow_num, a.params, c.flow_off, c.I_B_inv};
        load_flow_table(ft, fa);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        dec_positions_all<N>(std::make_integer_sequence<int, N::K>{}, c, posA);
        int q = 0;
        for (int l = 0; l < N::NE; ++l) {
            const cdg_linear& L = c.enc[l];
            for (int i = 0; i < L.in * L.out; ++i) posB[q++] = (int)L.w + i;
            for (int o = 0; o < L.out; ++o) posB[q++] = (int)L.b + o;
        }
    }
    __syncthreads();
    constexpr int d = N::DN;
    const float invB = 1.f / (float)a.batch;
    double rec_acc = 0.0, kl_acc = 0.0, al_acc = 0.0;
    float var_acc[d];
#pragma unroll
    for (int i = 0; i < d; ++i) var_acc[i] = 0.f;
    FlowGrad fg;
    fg.clear();

    // Small networks (loan / adult: 32 + 64 product slots): a thread keeps its products in registers over ALL of its rows
    // and the warp reduce-scatter runs once per thread instead of once per row (it was ~45 % of the issued instructions).
    constexpr bool ACCUM = P::NA_PAD + P::NB_PAD <= 128;
    float gpa_acc[ACCUM ? P::NA_PAD : 1], gpb_acc[ACCUM ? P::NB_PAD : 1];
#pragma unroll
    for (int i = 0; i < (ACCUM ? P::NA_PAD : 1); ++i) gpa_acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < (ACCUM ? P::NB_PAD : 1); ++i) gpb_acc[i] = 0.f;

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nrounds = (a.batch + stride - 1) / stride;
    for (int64_t rd = 0; rd < nrounds; ++rd) {
        const int64_t b = rd * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = b < a.batch;
        const float vm = valid ? 1.f : 0.f;
        const int64_t br = valid ? b : 0;

        // ---- encoder ----
        float x[N::D], h0[N::EH], h1[N::EH], h2[N::EH], ml[2 * d];
#pragma unroll
        for (int i = 0; i < N::D; ++i) x[i] = a.x[br * N::D + i];
        fc<N::D, N::EH, true, CP>(sp, enc_lin<CP>(c, 0), x, h0);
        if constexpr (N::NE == 4) {
            fc<N::EH, N::EH, true, CP>(sp, enc_lin<CP>(c, 1), h0, h1);
            fc<N::EH, N::EH, true, CP>(sp, enc_lin<CP>(c, 2), h1, h2);
            fc<N::EH, 2 * d, false, CP>(sp, enc_lin<CP>(c, 3), h2, ml);
        } else {
            fc<N::EH, 2 * d, false, CP>(sp, enc_lin<CP>(c, 1), h0, ml);
        }

        // ---- latent block ----
        float mean[CDG_MAX_NODE], lv[CDG_MAX_NODE], nz[CDG_MAX_NODE], eps[CDG_MAX_NODE], u[CDG_MAX_NODE], z[CDG_MAX_NODE];
        float u2[CDG_MAX_NODE], z2[CDG_MAX_NODE], gal[CDG_MAX_NODE], gu2[CDG_MAX_NODE];
        float kl = 0.f, al = 0.f;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            mean[i] = lv[i] = nz[i] = eps[i] = 0.f;
            if (i < d) {
                mean[i] = ml[i < d ? i : 0]; lv[i] = ml[i < d ? d + i : 0];
                nz[i] = a.deterministic ? 0.f : a.noise[br * d + i];
                const float ev = ex<CP>(lv[i]);
                eps[i] = a.deterministic ? mean[i] : mean[i] + ex<CP>(lv[i] / 2.f) * nz[i];
                kl += mean[i] * mean[i] - lv[i] + ev;
                var_acc[i < d ? i : 0] += vm * ev;
            }
        }
        kl_acc += (double)(vm * 0.5f * (kl - (float)d));
        const float ascale = c.lambda_ * invB;
        if constexpr (CP) {
            // linear SCM, constants from the constant bank; BCE in logits form with the reference's fp32 saturation behaviour
#pragma unroll
            for (int j = 0; j < CDG_MAX_NODE; ++j) {
                u[j] = u2[j] = z[j] = z2[j] = gu2[j] = 0.f;
                if (j < d) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int i = 0; i < d; ++i) {
                        s1 = fmaf(eps[i], c.I_B_inv[i * d + (j < d ? j : 0)], s1);
                        s2 = fmaf(mean[i], c.I_B_inv[i * d + (j < d ? j : 0)], s2);
                    }
                    u[j] = s1; u2[j] = s2;
                    const float fw = c_tabfix[CovOff::flow(j < d ? j : 0)], fb = c_tabfix[CovOff::flow(j < d ? j : 0) + 1];
                    z[j] = fmaf(fw, s1, fb);
                    z2[j] = fmaf(fw, s2, fb);
                    if (a.y) {
                        const float e = __expf(-fabsf(z2[j]));
                        const float r = __fdividef(1.f, 1.f + e);
                        const float yh = z2[j] >= 0.f ? r : e * r;
                        const float om = 1.f - yh;
                        const float sp1 = __logf(1.f + e);
                        const float l_yh = yh <= 0.f ? -100.f : fmaxf(-(fmaxf(-z2[j], 0.f) + sp1), -100.f);
                        const float l_om = om <= 0.f ? -100.f : fmaxf(-(fmaxf(z2[j], 0.f) + sp1), -100.f);
                        const float yy = a.y[br * d + j];
                        al += (yy - 1.f) * l_om - yy * l_yh;
                        const float t = om * yh;
                        const float gzz = vm * ascale * (yh - yy) * (t >= 1e-12f ? 1.f : t * 1e12f);
                        if (a.do_bwd) {
                            fg.a[j][0] += gzz * s2;
                            fg.a[j][1] += gzz;
                            gu2[j] = gzz * fw;
                        }
                    }
                }
            }
        } else {
        matvec_A(ft, d, eps, u);
        matvec_A(ft, d, mean, u2);
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
        }
        al_acc += (double)(vm * al);
        if constexpr (CP) {
#pragma unroll
            for (int i = 0; i < CDG_MAX_NODE; ++i) {
                float sacc = 0.f;
#pragma unroll
                for (int j = 0; j < d; ++j) sacc = fmaf(gu2[j], c.I_B_inv[(i < d ? i : 0) * d + j], sacc);
                gal[i] = i < d ? sacc : 0.f;
            }
        } else {
            matvec_AT(ft, d, gu2, gal);
        }
        if (a.latents && valid) {
            float* o = a.latents + b * 6 * d;
#pragma unroll
            for (int i = 0; i < d; ++i) {
                o[i] = mean[i]; o[d + i] = lv[i]; o[2 * d + i] = eps[i]; o[3 * d + i] = u[i]; o[4 * d + i] = z[i];
                o[5 * d + i] = z2[i];
            }
        }

        // ---- decoders (factor 1 each): hidden activations kept for the backward ----
        float a1[N::K][N::DH], a2[N::K][N::DH], xh[N::OUT], gx[N::OUT];
        dec_forward_all<N, CP>(std::make_integer_sequence<int, N::K>{}, sp, c, z, a1, a2, xh);
        if (a.xhat && valid) {
#pragma unroll
            for (int j = 0; j < N::OUT; ++j) a.xhat[b * N::OUT + j] = xh[j];
        }

        // ---- reconstruction loss and d/d xhat (tabular/modules/train.py:199-208) ----
        float rec = 0.f;
        if constexpr (N::KIND == CDG_TAB_LOAN) {
#pragma unroll
            for (int j = 0; j < N::OUT; ++j) {
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < N::D; ++i) if (i == c.flatten_topology[j]) t = x[i];
                const float df = xh[j] - t;
                rec += 0.5f * df * df;
                gx[j] = df * invB;
            }
        } else if constexpr (N::KIND == CDG_TAB_ADULT) {
#pragma unroll
            for (int j = 0; j < N::OUT; ++j) {
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < N::D; ++i) if (i == c.flatten_topology[j]) t = x[i];
                if (j == 2) {
                    const float zz = xh[j];
                    rec += fmaxf(zz, 0.f) - zz * t + log1pf(expf(-fabsf(zz)));
                    gx[j] = (1.f / (1.f + expf(-zz)) - t) * invB;
                } else {
                    const float df = xh[j] - t;
                    rec += 0.5f * df * df;
                    gx[j] = df * invB;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float df = xh[j] - x[j];
                rec += 0.5f * df * df;
                gx[j] = df * invB;
            }
            const int cls = (int)(x[7] - 1.f);
            float mx = xh[7];
#pragma unroll
            for (int j = 1; j < 7; ++j) mx = fmaxf(mx, xh[7 + j]);
            float se = 0.f;
#pragma unroll
            for (int j = 0; j < 7; ++j) se += ex<CP>(xh[7 + j] - mx);
            const float lse = mx + (CP ? __logf(se) : logf(se));
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                if (j == cls) rec += lse - xh[7 + j];
                gx[7 + j] = (ex<CP>(xh[7 + j] - lse) - (j == cls ? 1.f : 0.f)) * invB;
            }
        }
        rec_acc += (double)(vm * rec);
        if (!a.do_bwd) continue;
#pragma unroll
        for (int j = 0; j < N::OUT; ++j) gx[j] *= vm;

        // ---- decoders backward ----
        float gz[CDG_MAX_NODE];
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) gz[i] = 0.f;
        if constexpr (ACCUM) {
            dec_backward_all<N, true, CP>(std::make_integer_sequence<int, N::K>{}, sp, gpa_acc, c, z, a1, a2, gx, gz);
        } else {
            float gpa[P::NA_PAD];
#pragma unroll
            for (int i = P::NA; i < P::NA_PAD; ++i) gpa[i] = 0.f;
            dec_backward_all<N, false, CP>(std::make_integer_sequence<int, N::K>{}, sp, gpa, c, z, a1, a2, gx, gz);
            flush_products<P::NA_PAD>(gpa, posA, sg);
        }

        // ---- latent backward ----
        float gu[CDG_MAX_NODE], ge[CDG_MAX_NODE], gml[2 * d];
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) {
            if constexpr (CP) {
                gu[j] = 0.f;
                if (j < d) {
                    fg.a[j][0] += gz[j] * u[j];
                    fg.a[j][1] += gz[j];
                    gu[j] = gz[j] * c_tabfix[CovOff::flow(j < d ? j : 0)];
                }
            } else {
                gu[j] = j < d ? flow_bwd(ft, c.scm, c.flow_num, j, u[j], gz[j], fg) : 0.f;
            }
        }
        if constexpr (CP) {
#pragma unroll
            for (int i = 0; i < CDG_MAX_NODE; ++i) {
                float sacc = 0.f;
#pragma unroll
                for (int j = 0; j < d; ++j) sacc = fmaf(gu[j], c.I_B_inv[(i < d ? i : 0) * d + j], sacc);
                ge[i] = i < d ? sacc : 0.f;
            }
        } else {
            matvec_AT(ft, d, gu, ge);
        }
        const float kscale = c.beta * invB;
#pragma unroll
        for (int i = 0; i < d; ++i) {
            gml[i] = ge[i] + vm * kscale * mean[i] + gal[i];
            gml[d + i] = 0.5f * ge[i] * nz[i] * ex<CP>(lv[i] / 2.f) + vm * 0.5f * kscale * (ex<CP>(lv[i]) - 1.f);
        }

        // ---- encoder backward (no gradient into x) ----
        float gh[N::EH], gdump[N::D];
        if constexpr (ACCUM) {
            fc_bwd<N::EH, 2 * d, true, true, P::E0, true, CP>(sp, gpb_acc, enc_lin<CP>(c, 1), h0, gml, gh);
            fc_bwd<N::D, N::EH, false, false, 0, true, CP>(sp, gpb_acc, enc_lin<CP>(c, 0), x, gh, gdump);
        } else {
            float gpb[P::NB_PAD];
#pragma unroll
            for (int i = P::NB; i < P::NB_PAD; ++i) gpb[i] = 0.f;
            if constexpr (N::NE == 4) {
                float gh2[N::EH], gh1[N::EH];
                fc_bwd<N::EH, 2 * d, true, true, P::E0 + 2 * P::EM, false, CP>(sp, gpb, enc_lin<CP>(c, 3), h2, gml, gh2);
                fc_bwd<N::EH, N::EH, true, true, P::E0 + P::EM, false, CP>(sp, gpb, enc_lin<CP>(c, 2), h1, gh2, gh1);
                fc_bwd<N::EH, N::EH, true, true, P::E0, false, CP>(sp, gpb, enc_lin<CP>(c, 1), h0, gh1, gh);
            } else {
                fc_bwd<N::EH, 2 * d, true, true, P::E0, false, CP>(sp, gpb, enc_lin<CP>(c, 1), h0, gml, gh);
            }
            fc_bwd<N::D, N::EH, false, false, 0, false, CP>(sp, gpb, enc_lin<CP>(c, 0), x, gh, gdump);
            flush_products<P::NB_PAD>(gpb, posB, sg);
        }
    }
    if constexpr (ACCUM) {
        static_assert(!ACCUM || N::NE == 2, "the accumulating path is written for the two-layer encoders");
        if (a.do_bwd) {
            flush_products<P::NA_PAD>(gpa_acc, posA, sg);
            flush_products<P::NB_PAD>(gpb_acc, posB, sg);
        }
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

template <class N>
static bool shape_is(const cdg_tabular_config& c) {
    if (c.kind != N::KIND || c.input_dim != N::D || c.node != N::DN || c.n_dec != N::K || c.act != CDG_ACT_ELU) return false;
    if (c.n_enc_layers != N::NE || c.n_dec_layers != N::ND) return false;
    for (int l = 0; l < N::NE; ++l) {
        const int in = l == 0 ? N::D : N::EH, out = l + 1 == N::NE ? 2 * N::DN : N::EH;
        if (c.enc[l].in != in || c.enc[l].out != out) return false;
    }
    for (int k = 0; k < N::K; ++k) {
        if (c.factor[k] != 1 || c.out_dim[k] != N::m(k)) return false;
        for (int l = 0; l < N::ND; ++l) {
            const int in = l == 0 ? 1 : N::DH, out = l + 1 == N::ND ? N::m(k) : N::DH;
            if (c.dec[k][l].in != in || c.dec[k][l].out != out) return false;
        }
    }
    return true;
}

static bool cov_canonical(const cdg_tabular_config& c) {
    if (c.scm != CDG_SCM_LINEAR || c.n_params != CovOff::NPARAMS || c.n_params > TF_MAX_PARAMS) return false;
    for (int l = 0; l < NetCov::NE; ++l)
        if (c.enc[l].w != CovOff::enc_w(l) || c.enc[l].b != CovOff::enc_b(l)) return false;
    for (int j = 0; j < NetCov::DN; ++j)
        if (c.flow_off[j] != CovOff::flow(j)) return false;
    for (int k = 0; k < NetCov::K; ++k)
        for (int l = 0; l < NetCov::ND; ++l)
            if (c.dec[k][l].w != CovOff::dec_w(k, l) || c.dec[k][l].b != CovOff::dec_b(k, l)) return false;
    return true;
}

static int g_fixed_const = 1;
void set_tab_fixed_const(int on) { g_fixed_const = on; }

// returns true when one of the fixed networks matched and its kernel was launched
bool launch_tab_fixed(const TabArgs& a, unsigned blocks, size_t smem, cudaStream_t s) {
    if (g_fixed_const && shape_is<NetCov>(a.c) && cov_canonical(a.c) &&
        cudaMemcpyToSymbolAsync(c_tabfix, a.params, sizeof(float) * a.c.n_params, 0, cudaMemcpyDeviceToDevice, s) == cudaSuccess) {
        tab_fixed_kernel<NetCov, true><<<blocks, TAB_THREADS, smem, s>>>(a);
        return true;
    }
    if (shape_is<NetLoan>(a.c)) tab_fixed_kernel<NetLoan><<<blocks, TAB_THREADS, smem, s>>>(a);
    else if (shape_is<NetAdult>(a.c)) tab_fixed_kernel<NetAdult><<<blocks, TAB_THREADS, smem, s>>>(a);
    else if (shape_is<NetCov>(a.c)) tab_fixed_kernel<NetCov><<<blocks, TAB_THREADS, smem, s>>>(a);
    else return false;
    return true;
}

}  // namespace cdg
