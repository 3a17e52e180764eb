// The causal latent block shared by every CDG-VAE family, fused per row:
//   reparameterisation (modules/model.py:276-277), z = eps @ (I-B)^-1 (model.py:262),
//   per-node flows (InvertiblePriorLinear model.py:20-25 | PlanarFlows model.py:87-100),
//   Gaussian KL (modules/train.py:180-185), label alignment BCE (train.py:189-190),
//   posterior-variance logging (train.py:194-196) and the backward of all of it.
// Memory-bound, one thread per row, warp-shuffle + block reductions for the loss scalars and
// the (<= 48) flow-parameter gradients.
#include "latent.cuh"

namespace cdg {

__global__ void __launch_bounds__(256) latent_fwd_kernel(LatentArgs a) {
    __shared__ FlowTable ft;
    __shared__ double red[32];
    load_flow_table(ft, a);
    __syncthreads();
    const int d = a.d;
    double kl_acc = 0.0;
    float var_acc[CDG_MAX_NODE];
#pragma unroll
    for (int i = 0; i < CDG_MAX_NODE; ++i) var_acc[i] = 0.f;

    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.batch; b += (int64_t)gridDim.x * blockDim.x) {
        float mean[CDG_MAX_NODE], eps[CDG_MAX_NODE], u[CDG_MAX_NODE];
        float kl = 0.f;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            if (i < d) {
                mean[i] = a.ml[b * 2 * d + i];
                const float lv = a.ml[b * 2 * d + d + i];
                const float ev = expf(lv);
                eps[i] = a.deterministic ? mean[i] : mean[i] + expf(lv / 2.f) * a.noise[b * d + i];
                kl += mean[i] * mean[i] - lv + ev;
                var_acc[i] += ev;
            }
        }
        kl_acc += 0.5 * (double)(kl - (float)d);
        matvec_A(ft, d, eps, u);
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) {
            if (j < d) {
                const float z = flow_fwd(ft, a.scm, a.flow_num, j, u[j]);
                if (a.eps_out) a.eps_out[b * d + j] = eps[j];
                if (a.u_out) a.u_out[b * d + j] = u[j];
                if (a.z_out) a.z_out[b * d + j] = z;
            }
        }
    }
    if (a.acc) {
        double s = block_sum<double>(kl_acc, red);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_KL, s);
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            if (i < d) {
                double v = block_sum<double>((double)var_acc[i], red);
                if (threadIdx.x == 0) atomicAdd(a.acc + ACC_VAR + i, v);
            }
        }
    }
}

// Deterministic path (model.py:300: eps = mean) + alignment loss + its whole backward:
// writes g_align[b, 0:d] = d(lambda*align)/d(mean) (and zeros for the logvar half), and
// accumulates the flow-parameter gradients of this path.
__global__ void __launch_bounds__(256) align_kernel(LatentArgs a) {
    __shared__ FlowTable ft;
    __shared__ double red[32];
    __shared__ float fred[32];
    load_flow_table(ft, a);
    __syncthreads();
    const int d = a.d;
    double al_acc = 0.0;
    FlowGrad fg;
    fg.clear();
    const float gscale = a.lambda_ / (float)a.batch;

    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.batch; b += (int64_t)gridDim.x * blockDim.x) {
        float mean[CDG_MAX_NODE], u[CDG_MAX_NODE], gu[CDG_MAX_NODE], gm[CDG_MAX_NODE];
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) mean[i] = i < d ? a.ml[b * 2 * d + i] : 0.f;
        matvec_A(ft, d, mean, u);
        float al = 0.f;
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) {
            gu[j] = 0.f;
            if (j < d) {
                const float z = flow_fwd(ft, a.scm, a.flow_num, j, u[j]);
                if (a.z_out) a.z_out[b * d + j] = z;
                if (a.y) {
                    // F.binary_cross_entropy on probabilities, log clamped at -100 (train.py:189-190)
                    const float yh = 1.f / (1.f + expf(-z));
                    const float y = a.y[b * a.ld_y + j];
                    al += (y - 1.f) * fmaxf(log1pf(-yh), -100.f) - y * fmaxf(logf(yh), -100.f);
                    // autograd: BCE' = (yh - y) / max(yh (1-yh), 1e-12); sigmoid' = yh (1-yh)
                    const float gz = gscale * (yh - y) / fmaxf((1.f - yh) * yh, 1e-12f) * ((1.f - yh) * yh);
                    gu[j] = flow_bwd(ft, a.scm, a.flow_num, j, u[j], gz, fg);
                }
            }
        }
        al_acc += (double)al;
        if (a.g_out) {
            matvec_AT(ft, d, gu, gm);
#pragma unroll
            for (int i = 0; i < CDG_MAX_NODE; ++i) {
                if (i < d) {
                    a.g_out[b * 2 * d + i] = gm[i];
                    a.g_out[b * 2 * d + d + i] = 0.f;
                }
            }
        }
    }
    if (a.acc) {
        double s = block_sum<double>(al_acc, red);
        if (threadIdx.x == 0) atomicAdd(a.acc + ACC_ALIGN, s);
    }
    if (a.grads && a.y) reduce_flow_grads(fg, ft, a, fred);
}

// Backward of the stochastic path given g_z = d loss / d latent (from the decoders).
__global__ void __launch_bounds__(256) latent_bwd_kernel(LatentArgs a) {
    __shared__ FlowTable ft;
    __shared__ float fred[32];
    load_flow_table(ft, a);
    __syncthreads();
    const int d = a.d;
    FlowGrad fg;
    fg.clear();
    const float kscale = a.beta / (float)a.batch;

    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.batch; b += (int64_t)gridDim.x * blockDim.x) {
        float gu[CDG_MAX_NODE], ge[CDG_MAX_NODE];
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j)
            gu[j] = j < d ? flow_bwd(ft, a.scm, a.flow_num, j, a.u_in[b * d + j], a.g_z[b * d + j], fg) : 0.f;
        matvec_AT(ft, d, gu, ge);
        if (a.g_eps) {
#pragma unroll
            for (int i = 0; i < CDG_MAX_NODE; ++i)
                if (i < d) ge[i] += a.g_eps[b * d + i];
        }
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            if (i < d) {
                const float mean = a.ml[b * 2 * d + i];
                const float lv = a.ml[b * 2 * d + d + i];
                float gm = ge[i] + kscale * mean;
                if (a.g_align) gm += a.g_align[b * 2 * d + i];
                const float glv = 0.5f * ge[i] * a.noise[b * d + i] * expf(lv / 2.f) + 0.5f * kscale * (expf(lv) - 1.f);
                a.g_out[b * 2 * d + i] = gm;
                a.g_out[b * 2 * d + d + i] = glv;
            }
        }
    }
    if (a.grads) reduce_flow_grads(fg, ft, a, fred);
}

static int grid_for(int64_t batch) {
    int64_t blocks = (batch + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    return (int)blocks;
}

int launch_latent_fwd(const LatentArgs& a, cudaStream_t s) {
    if (a.batch == 0) return CDG_OK;
    latent_fwd_kernel<<<grid_for(a.batch), 256, 0, s>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
int launch_align(const LatentArgs& a, cudaStream_t s) {
    if (a.batch == 0) return CDG_OK;
    align_kernel<<<grid_for(a.batch), 256, 0, s>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
int launch_latent_bwd(const LatentArgs& a, cudaStream_t s) {
    if (a.batch == 0) return CDG_OK;
    latent_bwd_kernel<<<grid_for(a.batch), 256, 0, s>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// ---- stand-alone flow evaluation (evaluation scripts: inference.py:302-317, metric.py:226-255) -----------------------
// One thread per (row, node).  direction 0: out = flow_j(in) [+ log|det|], direction 1: out = flow_j^-1(in).
//   InvertiblePriorLinear (modules/model.py:20-29): o = p0 e + p1, log|p0|;  inverse (o - p1) / p0.
//   PlanarFlows (modules/model.py:77-100), input_dim = 1: forward h <- h + u_hat ELU(h w + b) with
//   logdet += log|1 + ELU'(h w + b) w u_hat| (evaluated before the update, :91-98); inverse = flows in reverse order,
//   each `inverse_loop` fixed-point iterations z <- h - u_hat ELU(z w + b) started at z = h (:80-84).
struct FlowApplyArgs {
    int d, scm, flow_num, loops, direction;
    int64_t batch, ld_in, ld_out, ld_logdet;
    const float* params;
    int64_t flow_off[CDG_MAX_NODE];
    const float* in;
    float* out;
    float* logdet;
};

__global__ void __launch_bounds__(256) flow_apply_kernel(FlowApplyArgs a) {
    __shared__ FlowTable ft;
    load_flow_table<false>(ft, a);
    __syncthreads();
    const int d = a.d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.batch * d; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int j = (int)(i - r * d);
        float h = a.in[r * a.ld_in + j];
        float ld = 0.f;
        if (a.scm == CDG_SCM_LINEAR) {
            if (a.direction == 0) { h = ft.w[j][0] * h + ft.b[j][0]; ld = logf(fabsf(ft.w[j][0])); }
            else h = (h - ft.b[j][0]) / ft.w[j][0];
        } else if (a.direction == 0) {
            for (int f = 0; f < a.flow_num; ++f) {
                const float w = ft.w[j][f], uh = ft.uhat[j][f];
                const float x = h * w + ft.b[j][f];
                const float grad = x > 0.f ? 1.f : expf(x);
                ld += logf(fabsf(1.f + (grad * w) * uh));
                h = h + uh * elu_ref(x);
            }
        } else {
            for (int f = a.flow_num - 1; f >= 0; --f) {
                const float w = ft.w[j][f], uh = ft.uhat[j][f], b = ft.b[j][f];
                float z = h;
                for (int it = 0; it < a.loops; ++it) z = h - uh * elu_ref(z * w + b);
                h = z;
            }
        }
        a.out[r * a.ld_out + j] = h;
        if (a.logdet) a.logdet[r * a.ld_logdet + j] = ld;
    }
}

}  // namespace cdg

extern "C" int cdg_flow_apply(int scm, int flow_num, int inverse_loop, int node, const float* params, const int64_t* flow_off,
                              const float* in, int64_t ld_in, float* out, int64_t ld_out, float* logdet, int64_t ld_logdet,
                              int64_t batch, int direction, void* stream) {
    using namespace cdg;
    CDG_REQUIRE(params && flow_off && in && out, "cdg_flow_apply: null argument");
    CDG_REQUIRE(scm == CDG_SCM_LINEAR || scm == CDG_SCM_PLANAR, "Not supported SCM!");
    CDG_REQUIRE(node >= 1 && node <= CDG_MAX_NODE, "cdg_flow_apply: node=%d out of range", node);
    CDG_REQUIRE(scm == CDG_SCM_LINEAR || (flow_num >= 1 && flow_num <= CDG_MAX_FLOW), "cdg_flow_apply: flow_num=%d out of range", flow_num);
    CDG_REQUIRE(direction == 0 || direction == 1, "cdg_flow_apply: direction must be 0 (forward) or 1 (inverse)");
    CDG_REQUIRE(inverse_loop >= 0 && ld_in >= node && ld_out >= node && (!logdet || ld_logdet >= node), "cdg_flow_apply: bad extents");
    if (batch <= 0) return CDG_OK;
    FlowApplyArgs a;
    memset(&a, 0, sizeof(a));
    a.d = node; a.scm = scm; a.flow_num = scm == CDG_SCM_LINEAR ? 1 : flow_num; a.loops = inverse_loop; a.direction = direction;
    a.batch = batch; a.ld_in = ld_in; a.ld_out = ld_out; a.ld_logdet = ld_logdet;
    a.params = params; a.in = in; a.out = out; a.logdet = logdet;
    for (int i = 0; i < node; ++i) a.flow_off[i] = flow_off[i];
    const int64_t n = batch * node;
    flow_apply_kernel<<<(int)imin64((n + 255) / 256, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(a);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
