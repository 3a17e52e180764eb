// Tabular CDG-VAE and CDG-TVAE training step as ONE kernel: per row, the whole forward, the
// dataset-specific reconstruction loss, KL, alignment and the whole backward run in a thread's
// registers / local memory; weights sit in shared memory; parameter gradients are reduced with
// warp shuffles into a shared-memory gradient image and flushed once per block.
//
//   models : tabular/modules/model.py:234-358 (CDGVAE), :360-460 (TVAE)
//   losses : tabular/modules/train.py:199-210 (loan / adult / covtype), :270-285 (TVAE spans),
//            :213-225 (KL, alignment on all of y)
//
// The MLPs are 2-32 wide (87 .. ~3k parameters): they are per-row arithmetic, not GEMMs, so no
// tensor-core tiles are used here.
#include "latent.cuh"
#include "elementwise.cuh"
#include "tabular_args.cuh"

#include <new>
#include <stdlib.h>

struct cdg_tabular_plan {
    cdg_tabular_config c;
    int out_total;
};

namespace cdg {

constexpr int TAB_MAX_IN = 64;     // widest input (TVAE D)
constexpr int TAB_MAX_H = 32;      // widest hidden layer
constexpr int TAB_MAX_OUT = 64;    // xhat width

// sum over the warp, then one shared-memory atomic
__device__ __forceinline__ void wacc(float* sg, int64_t idx, float v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) atomicAdd(sg + idx, v);
}

// h_out = act(W h_in + b) for one Linear (+ activation unless last)
__device__ __forceinline__ void lin_fwd(const float* sp, const cdg_linear& L, const float* hin, float* hout, bool act, int kind) {
    for (int o = 0; o < L.out; ++o) {
        float s = sp[L.b + o];
        const float* w = sp + L.w + (int64_t)o * L.in;
        for (int i = 0; i < L.in; ++i) s = fmaf(w[i], hin[i], s);
        hout[o] = act ? act_fwd(s, kind) : s;
    }
}

// Given delta = dL/d(pre-activation output) of Linear L: accumulate dW, db; return dL/d(hin) in gin (already
// multiplied by act'(hin) when hin is itself a post-activation).
__device__ __forceinline__ void lin_bwd(const float* sp, float* sg, const cdg_linear& L, const float* hin, const float* delta,
                                        float* gin, bool hin_is_act, int kind) {
    for (int i = 0; i < L.in; ++i) gin[i] = 0.f;
    for (int o = 0; o < L.out; ++o) {
        const float dl = delta[o];
        wacc(sg, L.b + o, dl);
        const float* w = sp + L.w + (int64_t)o * L.in;
        for (int i = 0; i < L.in; ++i) {
            wacc(sg, L.w + (int64_t)o * L.in + i, dl * hin[i]);
            gin[i] = fmaf(dl, w[i], gin[i]);
        }
    }
    if (hin_is_act)
        for (int i = 0; i < L.in; ++i) gin[i] *= act_bwd_from_out(hin[i], kind);
}

__global__ void __launch_bounds__(TAB_THREADS) tab_step_kernel(TabArgs a) {
    extern __shared__ float smem[];
    const cdg_tabular_config& c = a.c;
    const int np = (int)c.n_params;
    float* sp = smem;                 // parameters
    float* sg = smem + np;            // gradient image of this block
    __shared__ FlowTable ft;
    __shared__ double dred[32];
    __shared__ float fred[32];
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sp[i] = a.params[i]; sg[i] = 0.f; }
    {
        struct { int d, scm, flow_num; const float* params; const int64_t* flow_off; const float* A; } fa =
            {c.node, c.scm, c.flow_num, a.params, c.flow_off, c.I_B_inv};
        load_flow_table(ft, fa);
    }
    __syncthreads();

    const int d = c.node, D = c.input_dim, act = c.act;
    const int LE = c.n_enc_layers, LD = c.n_dec_layers;
    const float invB = 1.f / (float)a.batch;
    double rec_acc = 0.0, kl_acc = 0.0, al_acc = 0.0;
    float var_acc[CDG_MAX_NODE];
#pragma unroll
    for (int i = 0; i < CDG_MAX_NODE; ++i) var_acc[i] = 0.f;
    FlowGrad fg;
    fg.clear();

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nrounds = (a.batch + stride - 1) / stride;
    for (int64_t rd = 0; rd < nrounds; ++rd) {
        const int64_t b = rd * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = b < a.batch;
        const float vm = valid ? 1.f : 0.f;
        const int64_t br = valid ? b : 0;

        // ---------------- encoder forward (activations kept for the backward) ----------------
        float xin[TAB_MAX_IN];
        float eh[CDG_MAX_LAYERS][TAB_MAX_H];
        for (int i = 0; i < D; ++i) xin[i] = a.x[br * D + i];
        for (int l = 0; l < LE; ++l)
            lin_fwd(sp, c.enc[l], l == 0 ? xin : eh[l - 1], eh[l], l + 1 < LE, act);
        const float* ml = eh[LE - 1];

        // ---------------- latent block ----------------
        float mean[CDG_MAX_NODE], lv[CDG_MAX_NODE], nz[CDG_MAX_NODE], eps[CDG_MAX_NODE], u[CDG_MAX_NODE], z[CDG_MAX_NODE];
        float u2[CDG_MAX_NODE], z2[CDG_MAX_NODE], gal[CDG_MAX_NODE], gu2[CDG_MAX_NODE];
        float kl = 0.f, al = 0.f;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            mean[i] = lv[i] = nz[i] = eps[i] = 0.f;
            if (i < d) {
                mean[i] = ml[i]; lv[i] = ml[d + i];
                nz[i] = a.deterministic ? 0.f : a.noise[br * d + i];
                const float ev = expf(lv[i]);
                eps[i] = a.deterministic ? mean[i] : mean[i] + expf(lv[i] / 2.f) * nz[i];
                kl += mean[i] * mean[i] - lv[i] + ev;
                var_acc[i] += vm * ev;
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
                    const float gz = vm * ascale * (yh - yy) / fmaxf((1.f - yh) * yh, 1e-12f) * ((1.f - yh) * yh);
                    if (a.do_bwd) gu2[j] = flow_bwd(ft, c.scm, c.flow_num, j, u2[j], gz, fg);
                }
            }
        }
        al_acc += (double)(vm * al);
        matvec_AT(ft, d, gu2, gal);
        if (a.latents && valid) {
            float* o = a.latents + b * 6 * d;
            for (int i = 0; i < d; ++i) {
                o[i] = mean[i]; o[d + i] = lv[i]; o[2 * d + i] = eps[i]; o[3 * d + i] = u[i]; o[4 * d + i] = z[i];
                o[5 * d + i] = z2[i];
            }
        }

        // ---------------- decoders forward: xhat = cat_k D_k(z_k) ----------------
        float xh[TAB_MAX_OUT], gx[TAB_MAX_OUT];
        {
            int col = 0, zoff = 0;
            for (int k = 0; k < c.n_dec; ++k) {
                float dh[2][TAB_MAX_H];
                const float* hin = z + zoff;
                for (int l = 0; l < LD; ++l) {
                    float* hout = (l + 1 < LD) ? dh[l & 1] : xh + col;
                    lin_fwd(sp, c.dec[k][l], hin, hout, l + 1 < LD, act);
                    hin = hout;
                }
                col += c.out_dim[k];
                zoff += c.factor[k];
            }
        }
        if (a.xhat && valid)
            for (int j = 0; j < a.out_total; ++j) a.xhat[b * a.out_total + j] = xh[j];

        // ---------------- reconstruction loss and d/d xhat ----------------
        float rec = 0.f;
        for (int j = 0; j < a.out_total; ++j) gx[j] = 0.f;
        if (c.kind == CDG_TAB_LOAN) {                       // train.py:199-200
            for (int j = 0; j < a.out_total; ++j) {
                const float df = xh[j] - xin[c.flatten_topology[j]];
                rec += 0.5f * df * df;
                gx[j] = df * invB;
            }
        } else if (c.kind == CDG_TAB_ADULT) {               // train.py:201-205
            for (int j = 0; j < a.out_total; ++j) {
                const float t = xin[c.flatten_topology[j]];
                if (j == 2) {                               // income: BCE-with-logits, mean over the batch
                    const float zz = xh[j];
                    rec += fmaxf(zz, 0.f) - zz * t + log1pf(expf(-fabsf(zz)));
                    gx[j] = (1.f / (1.f + expf(-zz)) - t) * invB;
                } else {
                    const float df = xh[j] - t;
                    rec += 0.5f * df * df;
                    gx[j] = df * invB;
                }
            }
        } else if (c.kind == CDG_TAB_COVTYPE) {             // train.py:206-208
            for (int j = 0; j < 7; ++j) {
                const float df = xh[j] - xin[j];
                rec += 0.5f * df * df;
                gx[j] = df * invB;
            }
            const int ncls = a.out_total - 7;
            const int cls = (int)(xin[7] - 1.f);
            float mx = -INFINITY;
            for (int j = 0; j < ncls; ++j) mx = fmaxf(mx, xh[7 + j]);
            float se = 0.f;
            for (int j = 0; j < ncls; ++j) se += expf(xh[7 + j] - mx);
            const float lse = mx + logf(se);
            if (cls >= 0 && cls < ncls) rec += lse - xh[7 + cls];
            for (int j = 0; j < ncls; ++j) gx[7 + j] = (expf(xh[7 + j] - lse) - (j == cls ? 1.f : 0.f)) * invB;
        } else {                                            // CDG-TVAE spans, train.py:270-285
            for (int sidx = 0; sidx < c.n_span; ++sidx) {
                const int st = c.span_start[sidx], dim = c.span_dim[sidx];
                if (c.span_kind[sidx] == CDG_SPAN_TANH) {
                    const float sd = sp[c.sigma_off + st];
                    const float th = tanhf(xh[st]);
                    const float r = xin[st] - th;
                    rec += r * r / 2.f / (sd * sd) + logf(sd);
                    gx[st] = -(r / (sd * sd)) * (1.f - th * th) * invB;
                    if (a.do_bwd) wacc(sg, c.sigma_off + st, vm * (-(r * r) / (sd * sd * sd) + 1.f / sd) * invB);
                } else {
                    int tgt = 0;
                    float best = xin[st], mx = xh[st];
                    for (int j = 1; j < dim; ++j) {
                        if (xin[st + j] > best) { best = xin[st + j]; tgt = j; }     // torch.argmax: first maximum
                        mx = fmaxf(mx, xh[st + j]);
                    }
                    float se = 0.f;
                    for (int j = 0; j < dim; ++j) se += expf(xh[st + j] - mx);
                    const float lse = mx + logf(se);
                    rec += lse - xh[st + tgt];
                    for (int j = 0; j < dim; ++j) gx[st + j] = (expf(xh[st + j] - lse) - (j == tgt ? 1.f : 0.f)) * invB;
                }
            }
        }
        rec_acc += (double)(vm * rec);
        if (!a.do_bwd) continue;
        for (int j = 0; j < a.out_total; ++j) gx[j] *= vm;

        // ---------------- decoders backward (hidden activations recomputed) ----------------
        float gz[CDG_MAX_NODE];
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) gz[i] = 0.f;
        {
            int col = 0, zoff = 0;
            for (int k = 0; k < c.n_dec; ++k) {
                float dh[CDG_MAX_LAYERS][TAB_MAX_H];
                const float* hin = z + zoff;
                for (int l = 0; l + 1 < LD; ++l) {
                    lin_fwd(sp, c.dec[k][l], hin, dh[l], true, act);
                    hin = dh[l];
                }
                float delta[TAB_MAX_H], gin[TAB_MAX_H];
                for (int j = 0; j < c.out_dim[k]; ++j) delta[j] = gx[col + j];
                for (int l = LD - 1; l >= 0; --l) {
                    const float* in_l = l == 0 ? z + zoff : dh[l - 1];
                    lin_bwd(sp, sg, c.dec[k][l], in_l, delta, gin, l > 0, act);
                    for (int i = 0; i < c.dec[k][l].in; ++i) delta[i] = gin[i];
                }
                for (int i = 0; i < c.factor[k]; ++i) gz[zoff + i] = delta[i];
                col += c.out_dim[k];
                zoff += c.factor[k];
            }
        }

        // ---------------- latent backward ----------------
        float gu[CDG_MAX_NODE], ge[CDG_MAX_NODE], gml[2 * CDG_MAX_NODE];
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) gu[j] = j < d ? flow_bwd(ft, c.scm, c.flow_num, j, u[j], gz[j], fg) : 0.f;
        matvec_AT(ft, d, gu, ge);
        const float kscale = c.beta * invB;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            if (i < d) {
                gml[i] = ge[i] + vm * kscale * mean[i] + gal[i];
                gml[d + i] = 0.5f * ge[i] * nz[i] * expf(lv[i] / 2.f) + vm * 0.5f * kscale * (expf(lv[i]) - 1.f);
            }
        }

        // ---------------- encoder backward ----------------
        {
            float delta[TAB_MAX_H], gin[TAB_MAX_IN];
            for (int j = 0; j < 2 * d; ++j) delta[j] = gml[j];
            for (int l = LE - 1; l >= 0; --l) {
                const float* in_l = l == 0 ? xin : eh[l - 1];
                if (l == 0) {
                    // no gradient flows into x: only dW, db
                    const cdg_linear& L = c.enc[0];
                    for (int o = 0; o < L.out; ++o) {
                        wacc(sg, L.b + o, delta[o]);
                        for (int i = 0; i < L.in; ++i) wacc(sg, L.w + (int64_t)o * L.in + i, delta[o] * in_l[i]);
                    }
                } else {
                    lin_bwd(sp, sg, c.enc[l], in_l, delta, gin, true, act);
                    for (int i = 0; i < c.enc[l].in; ++i) delta[i] = gin[i];
                }
            }
        }
    }

    // ---------------- block reductions ----------------
    if (a.acc) {
        double s = block_sum<double>(rec_acc, dred);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_RECON, s);
        s = block_sum<double>(kl_acc, dred);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_KL, s);
        s = block_sum<double>(al_acc, dred);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_ALIGN, s);
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            if (i < d) {
                s = block_sum<double>((double)var_acc[i], dred);
                if (threadIdx.x == 0) atomicAdd(a.acc + ACC_VAR + i, s);
            }
        }
    }
    if (a.do_bwd) {
        struct { int d, scm, flow_num; float* grads; const int64_t* flow_off; } ra = {c.node, c.scm, c.flow_num, a.grads, c.flow_off};
        reduce_flow_grads(fg, ft, ra, fred);
        __syncthreads();
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const float v = sg[i];
            if (v != 0.f) atomicAdd(a.grads + i, v);
        }
    }
}

}  // namespace cdg

using namespace cdg;

extern "C" int cdg_tabular_create(const cdg_tabular_config* cfg, cdg_tabular_plan** out) {
    CDG_REQUIRE(cfg && out, "cdg_tabular_create: null argument");
    const cdg_tabular_config& c = *cfg;
    CDG_REQUIRE(c.kind >= CDG_TAB_LOAN && c.kind <= CDG_TAB_TVAE, "Not supported dataset!");     // train.py:210
    CDG_REQUIRE(c.node >= 1 && c.node <= CDG_MAX_NODE, "node out of range");
    CDG_REQUIRE(c.n_dec >= 1 && c.n_dec <= CDG_MAX_DEC, "n_dec out of range");
    CDG_REQUIRE(c.scm == CDG_SCM_LINEAR || c.scm == CDG_SCM_PLANAR, "Not supported SCM!");
    CDG_REQUIRE(c.flow_num >= 1 && c.flow_num <= CDG_MAX_FLOW, "flow_num out of range");
    CDG_REQUIRE(c.input_dim >= 1 && c.input_dim <= TAB_MAX_IN, "input_dim %d exceeds %d", c.input_dim, TAB_MAX_IN);
    CDG_REQUIRE(c.n_enc_layers >= 1 && c.n_enc_layers <= CDG_MAX_LAYERS && c.n_dec_layers >= 1 && c.n_dec_layers <= CDG_MAX_LAYERS,
                "layer count out of range");
    CDG_REQUIRE(c.n_params > 0 && c.n_params <= 16384, "n_params out of range");
    int s = 0, total = 0;
    for (int k = 0; k < c.n_dec; ++k) {
        s += c.factor[k];
        total += c.out_dim[k];
        CDG_REQUIRE(c.out_dim[k] >= 1 && c.out_dim[k] <= TAB_MAX_H, "decoder %d output width out of range", k);
        for (int l = 0; l < c.n_dec_layers; ++l)
            CDG_REQUIRE(c.dec[k][l].in <= TAB_MAX_H && c.dec[k][l].out <= TAB_MAX_H, "decoder layer too wide");
        CDG_REQUIRE(c.dec[k][0].in == c.factor[k] && c.dec[k][c.n_dec_layers - 1].out == c.out_dim[k], "decoder %d shape mismatch", k);
    }
    CDG_REQUIRE(s == c.node, "sum(factor) != node");
    CDG_REQUIRE(total <= TAB_MAX_OUT, "xhat too wide");
    for (int l = 0; l < c.n_enc_layers; ++l)
        CDG_REQUIRE(c.enc[l].out <= TAB_MAX_H && (l == 0 || c.enc[l].in <= TAB_MAX_H), "encoder layer too wide");
    CDG_REQUIRE(c.enc[0].in == c.input_dim && c.enc[c.n_enc_layers - 1].out == 2 * c.node, "encoder shape mismatch");
    if (c.kind == CDG_TAB_TVAE) {
        CDG_REQUIRE(c.sigma_off >= 0 && c.n_span >= 1 && c.n_span <= CDG_MAX_SPANS && total == c.input_dim, "TVAE span table invalid");
        for (int i = 0; i < c.n_span; ++i)
            CDG_REQUIRE(c.span_start[i] >= 0 && c.span_start[i] + c.span_dim[i] <= total, "span %d out of range", i);
    } else if (c.kind == CDG_TAB_COVTYPE) {
        CDG_REQUIRE(total > 7 && c.input_dim >= 8, "covtype shape mismatch");
    } else {
        CDG_REQUIRE(total <= 16, "flatten_topology too long");
        for (int j = 0; j < total; ++j)
            CDG_REQUIRE(c.flatten_topology[j] >= 0 && c.flatten_topology[j] < c.input_dim, "flatten_topology out of range");
    }
    cdg_tabular_plan* p = new (std::nothrow) cdg_tabular_plan;
    CDG_REQUIRE(p, "out of host memory");
    p->c = c;
    p->out_total = total;
    *out = p;
    return CDG_OK;
}

extern "C" void cdg_tabular_destroy(cdg_tabular_plan* p) { delete p; }

extern "C" int64_t cdg_tabular_workspace_bytes(const cdg_tabular_plan* p, int64_t batch) {
    if (!p || batch < 0) return -1;
    return 256;     // the double-precision loss accumulators
}

static int g_tvae_tile = 1;        // 2 = mma.sync fragments (same speed on loan-shaped, slower on covtype-shaped tables: DESIGN 4.8)

static int tab_run(cdg_tabular_plan* p, const cdg_tabular_io* io, int do_bwd, int deterministic, void* stream) {
    CDG_REQUIRE(p && io, "null argument");
    CDG_REQUIRE(io->params && io->x && io->workspace, "null buffer");
    CDG_REQUIRE(io->workspace_bytes >= 256, "workspace too small");
    CDG_REQUIRE(io->batch > 0, "empty batch");
    CDG_REQUIRE(deterministic || io->noise, "noise missing");
    if (do_bwd) CDG_REQUIRE(io->grads && io->y && io->logs, "grads / y / logs missing");
    cudaStream_t s = (cudaStream_t)stream;
    TabArgs a;
    a.c = p->c;
    a.params = io->params; a.grads = io->grads; a.x = io->x; a.y = io->y; a.noise = io->noise; a.batch = io->batch;
    a.xhat = io->xhat; a.latents = io->latents; a.acc = io->logs ? (double*)io->workspace : nullptr;
    a.do_bwd = do_bwd; a.deterministic = deterministic; a.out_total = p->out_total;
    if (do_bwd) CDG_CHECK_CUDA(cudaMemsetAsync(io->grads, 0, sizeof(float) * p->c.n_params, s));
    if (a.acc) CDG_CHECK_CUDA(cudaMemsetAsync(a.acc, 0, sizeof(double) * ACC_LEN, s));
    const size_t smem = sizeof(float) * 2 * (size_t)p->c.n_params;
    static bool attr = false;
    if (!attr) {
        CDG_CHECK_CUDA(cudaFuncSetAttribute(tab_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 * 4));
        attr = true;
    }
    int64_t blocks = (io->batch + TAB_THREADS - 1) / TAB_THREADS;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    // the three fixed CDG-VAE tables have a compile-time specialised kernel (tabular_fixed.cu)
    static const bool generic_only = exp_switch("CDG_TAB_GENERIC", 0) != 0;
    if (generic_only ||
        !(launch_tab_const(a, s) || launch_tab_fixed(a, (unsigned)blocks, smem, s) ||
          (g_tvae_tile == 2 && launch_tvae_mma(a, s)) || (g_tvae_tile >= 1 && launch_tvae_tile(a, s)) ||
          launch_tvae_fixed(a, (unsigned)blocks, smem, s)))
        tab_step_kernel<<<(unsigned)blocks, TAB_THREADS, smem, s>>>(a);
    CDG_CHECK_LAUNCH();
    if (io->logs)
        CDG_TRY(launch_finalize_logs(a.acc, io->logs, p->c.node, (float)io->batch, (float)io->batch, (float)io->batch,
                                     p->c.beta, p->c.lambda_, s));
    return CDG_OK;
}

extern "C" void cdg_tabular_const_params(int32_t on) { set_tab_const_params(on); }
extern "C" void cdg_tabular_tvae_tile(int32_t on) { g_tvae_tile = on; }

extern "C" int cdg_tabular_forward_backward(cdg_tabular_plan* p, const cdg_tabular_io* io, void* stream) {
    return tab_run(p, io, 1, 0, stream);
}
extern "C" int cdg_tabular_forward(cdg_tabular_plan* p, const cdg_tabular_io* io, int32_t deterministic, void* stream) {
    return tab_run(p, io, 0, deterministic, stream);
}
