// CDG-TVAE step (tabular/modules/model.py:360-460, tabular/modules/train.py:245-320) specialised for the
// reference's fixed architecture: encoder D-32-16-16-2d (ReLU), one decoder per node 1-8-8-16-m_k (ReLU),
// sigma[D]; D, the m_k and the span table are run-time, every hidden width is compile-time.  Hidden activations
// live in registers, the 1x16 / 16x16 / 16x32 products are fully unrolled, and the parameter-gradient products of a
// row are reduced over the warp 32 at a time (31-shuffle reduce-scatter; for the first layer one group is exactly
// the 32 products of one input column).  Same arithmetic as the generic tab_step_kernel, which remains the path
// for any other shape.
#include "latent.cuh"
#include "tabular_args.cuh"

namespace cdg {

constexpr int TV_H0 = 32, TV_H1 = 16, TV_H2 = 16, TV_D1 = 8, TV_D2 = 8, TV_D3 = 16;
constexpr int TV_MAXOUT = 64;

__device__ __forceinline__ float tv_reduce_scatter32(float* v) {
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
// lane l adds the warp total of v[l] to sg[base + l * stride] (l < n)
__device__ __forceinline__ void tv_flush32(float* v, float* sg, int64_t base, int stride, int n) {
    const float t = tv_reduce_scatter32(v);
    const int lane = threadIdx.x & 31;
    if (lane < n) atomicAdd(sg + base + (int64_t)lane * stride, t);
}

template <int IN, int OUT, bool ACT>
__device__ __forceinline__ void tv_fc(const float* sp, const cdg_linear& L, const float (&hin)[IN], float (&hout)[OUT]) {
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        float s = sp[L.b + o];
#pragma unroll
        for (int i = 0; i < IN; ++i) s = fmaf(sp[L.w + o * IN + i], hin[i], s);
        hout[o] = ACT ? fmaxf(s, 0.f) : s;
    }
}
// backward of a fixed-size Linear whose input hin is a ReLU output: dW, db into sg; gin = (W^T delta) * relu'(hin)
template <int IN, int OUT, bool HIN_RELU>
__device__ __forceinline__ void tv_fc_bwd(const float* sp, float* sg, const cdg_linear& L, const float (&hin)[IN],
                                          const float (&delta)[OUT], float (&gin)[IN]) {
    static_assert((IN * OUT) % 32 == 0 || IN * OUT < 32, "weight products are flushed in groups of 32");
#pragma unroll
    for (int i = 0; i < IN; ++i) gin[i] = 0.f;
    constexpr int NW = IN * OUT;
    if constexpr (NW >= 32) {
#pragma unroll
        for (int g = 0; g < NW / 32; ++g) {
            float gp[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int p = g * 32 + e, o = p / IN, i = p % IN;
                gp[e] = delta[o] * hin[i];
            }
            tv_flush32(gp, sg, L.w + g * 32, 1, 32);
        }
    } else {
        float gp[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) gp[e] = e < NW ? delta[e / IN] * hin[e % IN] : 0.f;
        tv_flush32(gp, sg, L.w, 1, NW);
    }
    {
        float gp[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) gp[e] = e < OUT ? delta[e < OUT ? e : 0] : 0.f;
        tv_flush32(gp, sg, L.b, 1, OUT);
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o)
#pragma unroll
        for (int i = 0; i < IN; ++i) gin[i] = fmaf(delta[o], sp[L.w + o * IN + i], gin[i]);
    if (HIN_RELU) {
#pragma unroll
        for (int i = 0; i < IN; ++i) gin[i] = hin[i] > 0.f ? gin[i] : 0.f;
    }
}

template <int DN>
__global__ void __launch_bounds__(TAB_THREADS) tvae_fixed_kernel(TabArgs a) {
    extern __shared__ float smem[];
    const cdg_tabular_config& c = a.c;
    const int np = (int)c.n_params;
    float* sp = smem;
    float* sg = smem + np;
    __shared__ FlowTable ft;
    __shared__ double dred[32];
    __shared__ float fred[32];
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sp[i] = a.params[i]; sg[i] = 0.f; }
    {
        struct { int d, scm, flow_num; const float* params; const int64_t* flow_off; const float* A; } fa =
            {DN, c.scm, c.flow_num, a.params, c.flow_off, c.I_B_inv};
        load_flow_table(ft, fa);
    }
    __syncthreads();
    constexpr int d = DN;
    const int D = c.input_dim;
    const float invB = 1.f / (float)a.batch;
    double rec_acc = 0.0, kl_acc = 0.0, al_acc = 0.0;
    float var_acc[d];
#pragma unroll
    for (int i = 0; i < d; ++i) var_acc[i] = 0.f;
    FlowGrad fg;
    fg.clear();

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nrounds = (a.batch + stride - 1) / stride;
    for (int64_t rd = 0; rd < nrounds; ++rd) {
        const int64_t b = rd * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = b < a.batch;
        const float vm = valid ? 1.f : 0.f;
        const int64_t br = valid ? b : 0;
        const float* xrow = a.x + br * D;

        // ---- encoder: D-32 (run-time D), 32-16, 16-16, 16-2d ----
        float h0[TV_H0], h1[TV_H1], h2[TV_H2], ml[2 * d];
        {
            const cdg_linear& L = c.enc[0];
#pragma unroll
            for (int o = 0; o < TV_H0; ++o) h0[o] = sp[L.b + o];
            for (int i = 0; i < D; ++i) {
                const float xi = __ldg(xrow + i);
#pragma unroll
                for (int o = 0; o < TV_H0; ++o) h0[o] = fmaf(sp[L.w + o * D + i], xi, h0[o]);
            }
#pragma unroll
            for (int o = 0; o < TV_H0; ++o) h0[o] = fmaxf(h0[o], 0.f);
        }
        tv_fc<TV_H0, TV_H1, true>(sp, c.enc[1], h0, h1);
        tv_fc<TV_H1, TV_H2, true>(sp, c.enc[2], h1, h2);
        tv_fc<TV_H2, 2 * d, false>(sp, c.enc[3], h2, ml);

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
        if (a.latents && valid) {
            float* o = a.latents + b * 6 * d;
#pragma unroll
            for (int i = 0; i < d; ++i) {
                o[i] = mean[i]; o[d + i] = lv[i]; o[2 * d + i] = eps[i]; o[3 * d + i] = u[i]; o[4 * d + i] = z[i];
                o[5 * d + i] = z2[i];
            }
        }

        // ---- decoders forward: 1-8-8-16-m_k; xhat indexed with run-time columns (local memory, <= 64 words) ----
        float xh[TV_MAXOUT], gx[TV_MAXOUT];
        {
            int col = 0;
#pragma unroll
            for (int k = 0; k < d; ++k) {
                float a1[TV_D1], a2[TV_D2], a3[TV_D3];
                const float zin[1] = {z[k]};
                tv_fc<1, TV_D1, true>(sp, c.dec[k][0], zin, a1);
                tv_fc<TV_D1, TV_D2, true>(sp, c.dec[k][1], a1, a2);
                tv_fc<TV_D2, TV_D3, true>(sp, c.dec[k][2], a2, a3);
                const cdg_linear& L = c.dec[k][3];
                for (int j = 0; j < L.out; ++j) {
                    float s = sp[L.b + j];
#pragma unroll
                    for (int i = 0; i < TV_D3; ++i) s = fmaf(sp[L.w + j * TV_D3 + i], a3[i], s);
                    xh[col + j] = s;
                }
                col += L.out;
            }
        }
        if (a.xhat && valid)
            for (int j = 0; j < a.out_total; ++j) a.xhat[b * a.out_total + j] = xh[j];

        // ---- span losses (tabular/modules/train.py:270-285) ----
        float rec = 0.f;
        for (int sidx = 0; sidx < c.n_span; ++sidx) {
            const int st = c.span_start[sidx], dim = c.span_dim[sidx];
            if (c.span_kind[sidx] == CDG_SPAN_TANH) {
                const float sd = sp[c.sigma_off + st];
                const float th = tanhf(xh[st]);
                const float r = __ldg(xrow + st) - th;
                rec += r * r / 2.f / (sd * sd) + logf(sd);
                gx[st] = -(r / (sd * sd)) * (1.f - th * th) * invB * vm;
                if (a.do_bwd) {
                    float t = warp_sum(vm * (-(r * r) / (sd * sd * sd) + 1.f / sd) * invB);
                    if ((threadIdx.x & 31) == 0) atomicAdd(sg + c.sigma_off + st, t);
                }
            } else {
                int tgt = 0;
                float best = __ldg(xrow + st), mx = xh[st];
                for (int j = 1; j < dim; ++j) {
                    const float xv = __ldg(xrow + st + j);
                    if (xv > best) { best = xv; tgt = j; }
                    mx = fmaxf(mx, xh[st + j]);
                }
                float se = 0.f;
                for (int j = 0; j < dim; ++j) se += expf(xh[st + j] - mx);
                const float lse = mx + logf(se);
                rec += lse - xh[st + tgt];
                for (int j = 0; j < dim; ++j) gx[st + j] = (expf(xh[st + j] - lse) - (j == tgt ? 1.f : 0.f)) * invB * vm;
            }
        }
        rec_acc += (double)(vm * rec);
        if (!a.do_bwd) continue;

        // ---- decoders backward (hidden activations recomputed) ----
        float gz[CDG_MAX_NODE];
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) gz[i] = 0.f;
        {
            int col = 0;
#pragma unroll
            for (int k = 0; k < d; ++k) {
                float a1[TV_D1], a2[TV_D2], a3[TV_D3], g3[TV_D3], g2[TV_D2], g1[TV_D1], gzin[1];
                const float zin[1] = {z[k]};
                tv_fc<1, TV_D1, true>(sp, c.dec[k][0], zin, a1);
                tv_fc<TV_D1, TV_D2, true>(sp, c.dec[k][1], a1, a2);
                tv_fc<TV_D2, TV_D3, true>(sp, c.dec[k][2], a2, a3);
                const cdg_linear& L = c.dec[k][3];
#pragma unroll
                for (int i = 0; i < TV_D3; ++i) g3[i] = 0.f;
                // last layer, two output rows per 32-wide product group
                for (int j = 0; j < L.out; j += 2) {
                    const float d0 = gx[col + j];
                    const float d1 = j + 1 < L.out ? gx[col + j + 1] : 0.f;
                    float gp[32];
#pragma unroll
                    for (int i = 0; i < TV_D3; ++i) {
                        gp[i] = d0 * a3[i];
                        gp[TV_D3 + i] = d1 * a3[i];
                        g3[i] = fmaf(d0, sp[L.w + j * TV_D3 + i], g3[i]);
                        if (j + 1 < L.out) g3[i] = fmaf(d1, sp[L.w + (j + 1) * TV_D3 + i], g3[i]);
                    }
                    tv_flush32(gp, sg, L.w + j * TV_D3, 1, j + 1 < L.out ? 32 : TV_D3);
                }
                for (int j = 0; j < L.out; j += 32) {
                    float gp[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) gp[e] = j + e < L.out ? gx[col + j + e] : 0.f;
                    tv_flush32(gp, sg, L.b + j, 1, min(32, L.out - j));
                }
#pragma unroll
                for (int i = 0; i < TV_D3; ++i) g3[i] = a3[i] > 0.f ? g3[i] : 0.f;
                tv_fc_bwd<TV_D2, TV_D3, true>(sp, sg, c.dec[k][2], a2, g3, g2);
                tv_fc_bwd<TV_D1, TV_D2, true>(sp, sg, c.dec[k][1], a1, g2, g1);
                tv_fc_bwd<1, TV_D1, false>(sp, sg, c.dec[k][0], zin, g1, gzin);
                gz[k] = gzin[0];
                col += L.out;
            }
        }

        // ---- latent backward ----
        float gu[CDG_MAX_NODE], ge[CDG_MAX_NODE], gml[2 * d];
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) gu[j] = j < d ? flow_bwd(ft, c.scm, c.flow_num, j, u[j], gz[j], fg) : 0.f;
        matvec_AT(ft, d, gu, ge);
        const float kscale = c.beta * invB;
#pragma unroll
        for (int i = 0; i < d; ++i) {
            gml[i] = ge[i] + vm * kscale * mean[i] + gal[i];
            gml[d + i] = 0.5f * ge[i] * nz[i] * expf(lv[i] / 2.f) + vm * 0.5f * kscale * (expf(lv[i]) - 1.f);
        }

        // ---- encoder backward ----
        float gh2[TV_H2], gh1[TV_H1], gh0[TV_H0];
        {
            // 16 x 2d head: 2d rows of 16 products, two rows per group
            const cdg_linear& L = c.enc[3];
#pragma unroll
            for (int i = 0; i < TV_H2; ++i) gh2[i] = 0.f;
#pragma unroll
            for (int j = 0; j < 2 * d; j += 2) {
                float gp[32];
#pragma unroll
                for (int i = 0; i < TV_H2; ++i) {
                    gp[i] = gml[j] * h2[i];
                    gp[TV_H2 + i] = gml[j + 1] * h2[i];
                    gh2[i] = fmaf(gml[j], sp[L.w + j * TV_H2 + i], gh2[i]);
                    gh2[i] = fmaf(gml[j + 1], sp[L.w + (j + 1) * TV_H2 + i], gh2[i]);
                }
                tv_flush32(gp, sg, L.w + j * TV_H2, 1, 32);
            }
            float gp[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) gp[e] = e < 2 * d ? gml[e < 2 * d ? e : 0] : 0.f;
            tv_flush32(gp, sg, L.b, 1, 2 * d);
#pragma unroll
            for (int i = 0; i < TV_H2; ++i) gh2[i] = h2[i] > 0.f ? gh2[i] : 0.f;
        }
        tv_fc_bwd<TV_H1, TV_H2, true>(sp, sg, c.enc[2], h1, gh2, gh1);
        tv_fc_bwd<TV_H0, TV_H1, true>(sp, sg, c.enc[1], h0, gh1, gh0);
        {
            // first layer: for input column i the 32 products gh0[o] * x_i are exactly one group (lane o -> W0[o][i])
            const cdg_linear& L = c.enc[0];
            for (int i = 0; i < D; ++i) {
                const float xi = __ldg(xrow + i);
                float gp[32];
#pragma unroll
                for (int o = 0; o < TV_H0; ++o) gp[o] = gh0[o] * xi;
                tv_flush32(gp, sg, L.w + i, D, 32);
            }
            float gp[32];
#pragma unroll
            for (int o = 0; o < TV_H0; ++o) gp[o] = gh0[o];
            tv_flush32(gp, sg, L.b, 1, 32);
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

static bool tvae_shape_ok(const cdg_tabular_config& c) {
    if (c.kind != CDG_TAB_TVAE || c.act != CDG_ACT_RELU || c.n_enc_layers != 4 || c.n_dec_layers != 4) return false;
    if (c.n_dec != c.node || (c.node != 3 && c.node != 6)) return false;
    const int e[5] = {c.input_dim, TV_H0, TV_H1, TV_H2, 2 * c.node};
    for (int l = 0; l < 4; ++l)
        if (c.enc[l].in != e[l] || c.enc[l].out != e[l + 1]) return false;
    int total = 0;
    for (int k = 0; k < c.n_dec; ++k) {
        if (c.factor[k] != 1) return false;
        const int dd[5] = {1, TV_D1, TV_D2, TV_D3, c.out_dim[k]};
        for (int l = 0; l < 4; ++l)
            if (c.dec[k][l].in != dd[l] || c.dec[k][l].out != dd[l + 1]) return false;
        total += c.out_dim[k];
    }
    return total <= TV_MAXOUT && c.input_dim <= TV_MAXOUT;
}

bool launch_tvae_fixed(const TabArgs& a, unsigned blocks, size_t smem, cudaStream_t s) {
    if (!tvae_shape_ok(a.c)) return false;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(tvae_fixed_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 * 4);
        cudaFuncSetAttribute(tvae_fixed_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 * 4);
        attr = true;
    }
    if (a.c.node == 3) tvae_fixed_kernel<3><<<blocks, TAB_THREADS, smem, s>>>(a);
    else tvae_fixed_kernel<6><<<blocks, TAB_THREADS, smem, s>>>(a);
    return true;
}

}  // namespace cdg
