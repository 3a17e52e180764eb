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

}  // namespace cdg
