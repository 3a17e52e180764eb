/*
 * cdgvae.h — C ABI of libcdgvae_sm100.so, the B200 (sm_100a) implementation of the
 * CDG-VAE / CDG-TVAE training step.
 *
 * The reference (an-seunghwan/CDG-VAE) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md §8b): the boundary it offers is the Python API of modules/model.py,
 * modules/train.py, tabular/modules/model.py and tabular/modules/train.py.  The host-side
 * mirror of that API lives in cdg-vae_b200/ and binds the entry points below with ctypes
 * (INTEGRATION.md shows the stub).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain C: pointers, sizes, POD structs; no torch / C++ types cross the boundary;
 *  - every tensor is fp32, row-major, resident in device memory owned by the CALLER; the
 *    library allocates nothing on the device;
 *  - all work is enqueued on the caller's stream (cudaStream_t passed as void*); no call
 *    synchronises; every call is CUDA-graph capturable;
 *  - return value 0 = OK, otherwise an error code; cdg_last_error() gives the message.
 */
#ifndef CDGVAE_H
#define CDGVAE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDG_MAX_NODE 8      /* latent nodes d (reference uses 3, 4, 5, 6)            */
#define CDG_MAX_DEC 8       /* decoders / causal factors K                            */
#define CDG_MAX_FLOW 2      /* planar flows per node (reference default flow_num = 1) */
#define CDG_MAX_LAYERS 4    /* Linear layers per MLP in the tabular families          */
#define CDG_MAX_SPANS 64    /* CDG-TVAE output spans                                  */
#define CDG_MAX_SEG 32      /* Adam segments                                          */

enum { CDG_OK = 0, CDG_ERR_INVALID = 1, CDG_ERR_CUDA = 2, CDG_ERR_WORKSPACE = 3, CDG_ERR_UNSUPPORTED = 4 };
enum { CDG_SCM_LINEAR = 0, CDG_SCM_PLANAR = 1 };
enum { CDG_GEMM_AUTO = 0, CDG_GEMM_SIMT = 1, CDG_GEMM_TC3X = 2, CDG_GEMM_TC1X = 3 };

const char* cdg_last_error(void);
int cdg_version(void);
/* 1 when the current device is sm_100 (the only target this library is built for). */
int cdg_device_ok(void);
/* Number of kernels this library has launched so far in this process (host-side counter). */
long long cdg_launch_count(void);

/* A Linear layer's weight [out,in] and bias [out] as float offsets into the parameter arena
 * (the gradient / exp_avg / exp_avg_sq arenas share the layout). */
typedef struct { int64_t w, b; int32_t in, out; } cdg_linear;

/* ------------------------------------------------------------------------------------------
 * Adam over a flat arena.  Replaces torch.optim.Adam.step() as constructed at main.py:189-192,
 * tabular/main.py:205-208 and tabular/main_tvae.py:196-200 (coupled weight decay), plus the
 * sigma clamp of tabular/modules/train.py:314.  Segments whose gradient the reference leaves
 * None (covtype decoder.6) are simply not listed.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_seg;
    int64_t seg_off[CDG_MAX_SEG], seg_len[CDG_MAX_SEG];
    double lr, beta1, beta2, eps, weight_decay; /* Python floats of the optimizer's param_group */
    float grad_scale;            /* multiplies every gradient first (1/world_size under DP) */
    int32_t step;                /* 1-based step count t of this update                      */
    int64_t clamp_off, clamp_len; /* optional clamp of params[clamp_off : +len] after update  */
    float clamp_lo, clamp_hi;
    int32_t* dev_step;           /* optional device counter: when non-NULL the call first increments it and uses
                                    its value as t (`step` is ignored), so a captured CUDA graph can be replayed */
} cdg_adam_args;

int cdg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  const cdg_adam_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pendulum CDG-VAE (modules/model.py:208-304 CDGVAE; modules/train.py:150-282).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t node;                         /* d = config["node"]                                 */
    int32_t n_dec;                        /* K = len(config["factor"])                          */
    int32_t factor[CDG_MAX_DEC];          /* latents per decoder (model.py:283)                 */
    int32_t dec_extra[CDG_MAX_DEC];       /* latent column appended to decoder k's input, or -1: the DR variant feeds
                                             every decoder the last ("spurious") latent too (DR/modules/model.py:284-287) */
    int32_t col_lo[CDG_MAX_DEC];          /* live flat output columns of decoder k: the support */
    int32_t col_hi[CDG_MAX_DEC];          /*   [lo,hi) of its {0,1} mask (main.py:167-179)      */
    int32_t scm;                          /* CDG_SCM_LINEAR | CDG_SCM_PLANAR (model.py:233-240) */
    int32_t flow_num;
    int32_t input_dim;                    /* P = 3 * image_size^2                               */
    int32_t hidden;                       /* H = 300                                            */
    int32_t gemm_mode;                    /* CDG_GEMM_*                                         */
    int32_t general_mask;                 /* 1: masks are arbitrary [K,P] weights (io->masks), every decoder computes all
                                             P columns and xhat = tanh(sum_k out_k * mask_k) literally (model.py:284-287);
                                             0: band masks given as col_lo/col_hi (the fast path)                        */
    int64_t n_params;                     /* arena length in floats                             */
    cdg_linear enc[3];                    /* encoder.{0,2,4}          (model.py:219-225)        */
    cdg_linear dec[CDG_MAX_DEC][3];       /* decoder.k.{0,2,4}        (model.py:243-250)        */
    int64_t flow_off[CDG_MAX_NODE];       /* linear: p[2]; planar: w[F], b[F], u[F] contiguous  */
    float I_B_inv[CDG_MAX_NODE * CDG_MAX_NODE]; /* row-major d x d (model.py:228-230)           */
    float beta, lambda_;                  /* loss = recon + beta*KL + lambda*align (train.py:198)*/
} cdg_pendulum_config;

typedef struct cdg_pendulum_plan cdg_pendulum_plan;

int cdg_pendulum_create(const cdg_pendulum_config* cfg, cdg_pendulum_plan** out);
void cdg_pendulum_destroy(cdg_pendulum_plan* p);
/* Workspace the caller must provide for a step on `batch` (+ `batch_l` labeled) samples. */
int64_t cdg_pendulum_workspace_bytes(const cdg_pendulum_plan* p, int64_t batch, int64_t batch_l);

typedef struct {
    const float* params;     /* parameter arena                                               */
    float* grads;            /* gradient arena (written: d loss / d params)                   */
    void* workspace;
    int64_t workspace_bytes;
    const float* x;          /* [batch, P]   (images flattened HWC, model.py:257)             */
    const float* y;          /* [batch, ld_y] labels in [0,1]; first d columns used (train.py:190) */
    int32_t ld_y;
    const float* noise;      /* [batch, d]   injected N(0,1) draw (model.py:276)              */
    int64_t batch;
    const float* x_l;        /* semi-supervised labeled batch (train.py:261) or NULL          */
    const float* y_l;
    int32_t ld_y_l;
    int64_t batch_l;
    float* logs;             /* [4 + d]: loss, recon, KL, alignment, posterior_variance1..d   */
    float* xhat;             /* optional [batch, P] reconstruction output (train.py:209), or NULL */
    const float* masks;      /* [K, P] decoder masks on the device; required when general_mask = 1 */
} cdg_pendulum_io;

/* zero_grad + forward + losses + backward of one batch: train.py:168-202 (:235-278 when x_l != NULL).
 * Gradients land in io->grads; no optimizer update. */
int cdg_pendulum_forward_backward(cdg_pendulum_plan* p, const cdg_pendulum_io* io, void* stream);

/* Forward only (model.py:290-304), for model.forward()/encode()/decode() outside the train loop.
 * Any output pointer may be NULL.  deterministic != 0 uses eps = mean (model.py:273-274). */
typedef struct {
    const float* params;
    void* workspace;
    int64_t workspace_bytes;
    const float* x;          /* [batch, P] or NULL when `latent_in` is given                   */
    const float* noise;      /* [batch, d] (ignored when deterministic)                        */
    const float* latent_in;  /* decode-only entry: [batch, d] concatenated latents (model.py:281) */
    int64_t batch;
    int32_t deterministic;
    float* mean; float* logvar; float* epsilon; float* orig_latent; float* latent; float* align_latent;
    float* xhat_separated;   /* [K, batch, P] unmasked per-decoder outputs (model.py:284)       */
    float* xhat;             /* [batch, P]                                                      */
    const float* masks;      /* [K, P] (general_mask = 1)                                       */
} cdg_pendulum_fwd_io;

int cdg_pendulum_forward(cdg_pendulum_plan* p, const cdg_pendulum_fwd_io* io, void* stream);

/* Optional device-side timing of the step by kernel category (cudaEvents recorded on the caller's stream
 * between the launches of cdg_pendulum_forward_backward).  cdg_pendulum_profile_read synchronises on the
 * recorded events, ADDS the elapsed milliseconds per category to out_ms[CDG_PROF_NCAT] and clears the record. */
enum { CDG_PROF_ENC0_FWD = 0, CDG_PROF_DEC2_FWD, CDG_PROF_DEC2_DGRAD, CDG_PROF_DEC2_WGRAD, CDG_PROF_ENC0_WGRAD,
       CDG_PROF_GEMM_OTHER, CDG_PROF_LATENT, CDG_PROF_RECON, CDG_PROF_MISC, CDG_PROF_NCAT };
int cdg_pendulum_profile_enable(cdg_pendulum_plan* p, int enable);
int cdg_pendulum_profile_read(cdg_pendulum_plan* p, double* out_ms);

/* ------------------------------------------------------------------------------------------
 * Tabular CDG-VAE and CDG-TVAE (tabular/modules/model.py:234-460; tabular/modules/train.py:173-320):
 * whole forward + losses + backward per row in registers, one launch.
 * ---------------------------------------------------------------------------------------- */
enum { CDG_TAB_LOAN = 0, CDG_TAB_ADULT = 1, CDG_TAB_COVTYPE = 2, CDG_TAB_TVAE = 3 };
enum { CDG_ACT_ELU = 0, CDG_ACT_RELU = 1 };
enum { CDG_SPAN_TANH = 0, CDG_SPAN_SOFTMAX = 1 };

typedef struct {
    int32_t kind;                          /* CDG_TAB_*                                        */
    int32_t node, n_dec;
    int32_t factor[CDG_MAX_DEC];
    int32_t out_dim[CDG_MAX_DEC];          /* `mask` of tabular/main.py:189-198                */
    int32_t scm, flow_num;
    int32_t input_dim;                     /* D                                                */
    int32_t act;                           /* CDG_ACT_ELU (CDGVAE) | CDG_ACT_RELU (TVAE)        */
    int32_t n_enc_layers, n_dec_layers;
    cdg_linear enc[CDG_MAX_LAYERS];
    cdg_linear dec[CDG_MAX_DEC][CDG_MAX_LAYERS];
    int64_t flow_off[CDG_MAX_NODE];
    int64_t sigma_off;                     /* TVAE sigma[D] (model.py:407), else -1            */
    int32_t flatten_topology[16];          /* loan/adult column permutation (train.py:199-205) */
    int32_t n_span;                        /* TVAE spans (train.py:270-285)                    */
    int32_t span_start[CDG_MAX_SPANS], span_dim[CDG_MAX_SPANS], span_kind[CDG_MAX_SPANS];
    int64_t n_params;                      /* arena length in floats                           */
    float I_B_inv[CDG_MAX_NODE * CDG_MAX_NODE];
    float beta, lambda_;
} cdg_tabular_config;

typedef struct cdg_tabular_plan cdg_tabular_plan;

int cdg_tabular_create(const cdg_tabular_config* cfg, cdg_tabular_plan** out);
void cdg_tabular_destroy(cdg_tabular_plan* p);
int64_t cdg_tabular_workspace_bytes(const cdg_tabular_plan* p, int64_t batch);

typedef struct {
    const float* params;
    float* grads;
    void* workspace;
    int64_t workspace_bytes;
    const float* x;          /* [batch, D]   */
    const float* y;          /* [batch, d]   */
    const float* noise;      /* [batch, d]   */
    int64_t batch;
    float* logs;             /* [4 + d]      */
    float* xhat;             /* optional [batch, sum(out_dim)] */
    float* latents;          /* optional [batch, 6*d]: mean, logvar, epsilon, orig_latent, latent, align_latent */
} cdg_tabular_io;

int cdg_tabular_forward_backward(cdg_tabular_plan* p, const cdg_tabular_io* io, void* stream);
int cdg_tabular_forward(cdg_tabular_plan* p, const cdg_tabular_io* io, int32_t deterministic, void* stream);

/* ------------------------------------------------------------------------------------------
 * The dense contraction used by every pendulum Linear layer, exposed for tests and profiling:
 *   C[M,N] (ldc) = sum_k A(m,k) * B(n,k),  A(m,k) = A[m*sa_m + k*sa_k],  B(n,k) = B[n*sb_n + k*sb_k]
 * mode selects the SIMT fp32 kernel or the tcgen05 3xTF32 (fp32-faithful) / 1xTF32 kernels.
 * ---------------------------------------------------------------------------------------- */
int cdg_gemm(int mode, const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_n, int64_t sb_k,
             float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate,
             void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDGVAE_H */
