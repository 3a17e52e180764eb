// Device pieces of the causal latent block, shared by latent.cu (pendulum) and tabular.cu.
#pragma once
#include "common.cuh"

namespace cdg {

// slots of the double-precision loss accumulator
enum { ACC_RECON = 0, ACC_KL = 1, ACC_ALIGN = 2, ACC_VAR = 3, ACC_LEN = 3 + CDG_MAX_NODE };

struct LatentArgs {
    int d, scm, flow_num, deterministic;
    int64_t batch;
    const float* params;
    float* grads;                 // flow-parameter gradients are accumulated here (may be null)
    int64_t flow_off[CDG_MAX_NODE];
    float A[CDG_MAX_NODE * CDG_MAX_NODE];
    float beta, lambda_;
    const float* ml;              // [batch, 2d] = [mean | logvar]
    const float* noise;           // [batch, d]
    const float* y; int ld_y;     // [batch, ld_y]
    const float* u_in;            // [batch, d] saved orig_latent (bwd)
    const float* g_z;             // [batch, d]
    const float* g_align;         // [batch, 2d] or null
    const float* g_eps;           // [batch, d] or null: extra d loss / d epsilon (InfoMax discriminator path)
    float* eps_out; float* u_out; float* z_out;
    float* g_out;                 // [batch, 2d]
    double* acc;                  // loss accumulators (may be null)
};

struct FlowTable {
    float A[CDG_MAX_NODE * CDG_MAX_NODE];
    float w[CDG_MAX_NODE][CDG_MAX_FLOW], b[CDG_MAX_NODE][CDG_MAX_FLOW], u[CDG_MAX_NODE][CDG_MAX_FLOW];
    float uhat[CDG_MAX_NODE][CDG_MAX_FLOW], sig[CDG_MAX_NODE][CDG_MAX_FLOW], duw[CDG_MAX_NODE][CDG_MAX_FLOW];
};

struct FlowGrad {
    float a[CDG_MAX_NODE][3 * CDG_MAX_FLOW];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i)
#pragma unroll
            for (int s = 0; s < 3 * CDG_MAX_FLOW; ++s) a[i][s] = 0.f;
    }
};

__device__ __forceinline__ float elu_ref(float t) { return t > 0.f ? t : expf(t) - 1.f; }

// WITH_A = false: the caller has no causal matrix (stand-alone flow evaluation); a.A is then never read.  (A.A may be an
// array member or a pointer: no sizeof-based bound here -- one once truncated the tabular kernels' I_B_inv to 2 floats.)
template <bool WITH_A = true, typename Args>
__device__ __forceinline__ void load_flow_table(FlowTable& ft, const Args& a) {
    const int d = a.d;
    if constexpr (WITH_A)
        for (int i = threadIdx.x; i < d * d; i += blockDim.x) ft.A[i] = a.A[i];
    for (int t = threadIdx.x; t < d * CDG_MAX_FLOW; t += blockDim.x) {
        const int i = t / CDG_MAX_FLOW, f = t % CDG_MAX_FLOW;
        const float* p = a.params + a.flow_off[i];
        if (a.scm == CDG_SCM_LINEAR) {
            if (f == 0) { ft.w[i][0] = p[0]; ft.b[i][0] = p[1]; }       // p[0] * eps + p[1]
        } else if (f < a.flow_num) {
            const float w = p[f], b = p[a.flow_num + f], u = p[2 * a.flow_num + f];
            // PlanarFlows.build_u (modules/model.py:70-75) with input_dim = 1:
            //   u_hat = u + (log(1 + exp(w u)) - 1 - w u) * w / |w|^2
            const float wu = w * u;
            const float r = (-1.f + logf(1.f + expf(wu))) - wu;
            const float nrm = fabsf(w);
            ft.w[i][f] = w; ft.b[i][f] = b; ft.u[i][f] = u;
            ft.uhat[i][f] = u + r * (w / (nrm * nrm));
            const float sg = 1.f / (1.f + expf(-wu));
            ft.sig[i][f] = sg;                                          // d u_hat / d u
            ft.duw[i][f] = ((sg - 1.f) * u * w - r) / (w * w);          // d u_hat / d w
        }
    }
}

__device__ __forceinline__ void matvec_A(const FlowTable& ft, int d, const float* e, float* u) {
#pragma unroll
    for (int j = 0; j < CDG_MAX_NODE; ++j) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i)
            if (i < d && j < d) s = fmaf(e[i], ft.A[i * d + j], s);
        u[j] = s;
    }
}
__device__ __forceinline__ void matvec_AT(const FlowTable& ft, int d, const float* gu, float* ge) {
#pragma unroll
    for (int i = 0; i < CDG_MAX_NODE; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j)
            if (i < d && j < d) s = fmaf(gu[j], ft.A[i * d + j], s);
        ge[i] = s;
    }
}

__device__ __forceinline__ float flow_fwd(const FlowTable& ft, int scm, int F, int j, float h) {
    if (scm == CDG_SCM_LINEAR) return ft.w[j][0] * h + ft.b[j][0];
#pragma unroll
    for (int f = 0; f < CDG_MAX_FLOW; ++f)
        if (f < F) h = h + ft.uhat[j][f] * elu_ref(h * ft.w[j][f] + ft.b[j][f]);   // model.py:99
    return h;
}

// Returns d loss / d (flow input); accumulates this row's flow-parameter gradients.
__device__ __forceinline__ float flow_bwd(const FlowTable& ft, int scm, int F, int j, float h0, float g, FlowGrad& fg) {
    if (scm == CDG_SCM_LINEAR) {
        fg.a[j][0] += g * h0;
        fg.a[j][1] += g;
        return g * ft.w[j][0];
    }
    float hs[CDG_MAX_FLOW];
    float h = h0;
#pragma unroll
    for (int f = 0; f < CDG_MAX_FLOW; ++f) {
        hs[f] = h;
        if (f < F) h = h + ft.uhat[j][f] * elu_ref(h * ft.w[j][f] + ft.b[j][f]);
    }
#pragma unroll
    for (int f = CDG_MAX_FLOW - 1; f >= 0; --f) {
        if (f < F) {
            const float w = ft.w[j][f], uh = ft.uhat[j][f];
            const float t = hs[f] * w + ft.b[j][f];
            const float e = elu_ref(t);
            const float de = t > 0.f ? 1.f : expf(t);
            const float g_uhat = g * e;
            const float g_t = g * uh * de;
            fg.a[j][f * 3 + 0] += g_t * hs[f] + g_uhat * ft.duw[j][f];   // d/dw
            fg.a[j][f * 3 + 1] += g_t;                                    // d/db
            fg.a[j][f * 3 + 2] += g_uhat * ft.sig[j][f];                  // d/du
            g = g + g_t * w;
        }
    }
    return g;
}

template <typename Args>
__device__ __forceinline__ void reduce_flow_grads(const FlowGrad& fg, const FlowTable&, const Args& a, float* fred) {
    const int d = a.d, F = a.flow_num;
#pragma unroll
    for (int i = 0; i < CDG_MAX_NODE; ++i) {
        if (i >= d) continue;
        if (a.scm == CDG_SCM_LINEAR) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const float v = block_sum<float>(fg.a[i][s], fred);
                if (threadIdx.x == 0) atomicAdd(a.grads + a.flow_off[i] + s, v);
            }
        } else {
#pragma unroll
            for (int f = 0; f < CDG_MAX_FLOW; ++f) {
                if (f >= F) continue;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const float v = block_sum<float>(fg.a[i][f * 3 + s], fred);
                    if (threadIdx.x == 0) atomicAdd(a.grads + a.flow_off[i] + s * F + f, v);
                }
            }
        }
    }
}

int launch_latent_fwd(const LatentArgs& a, cudaStream_t s);
int launch_align(const LatentArgs& a, cudaStream_t s);
int launch_latent_bwd(const LatentArgs& a, cudaStream_t s);

}  // namespace cdg
