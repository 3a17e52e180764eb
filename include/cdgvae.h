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
enum { CDG_GEMM_AUTO = 0, CDG_GEMM_SIMT = 1, CDG_GEMM_TC3X = 2, CDG_GEMM_TC1X = 3, CDG_GEMM_BF3X = 4 };

const char* cdg_last_error(void);
int cdg_version(void);
/* 1 when the current device is sm_100 (the only target this library is built for). */
int cdg_device_ok(void);
/* Number of kernels this library has launched so far in this process (host-side counter). */
long long cdg_launch_count(void);
/* A caller that replays a captured CUDA graph of this library's launches adds the graph's kernel count here (launches made
 * by a replay never pass through the library's host code). */
void cdg_launch_count_add(long long n);
/* sizeof() of the i-th struct of this header as the library was compiled (order: cdg_linear, cdg_adam_args,
 * cdg_pendulum_config, cdg_pendulum_io, cdg_pendulum_fwd_io, cdg_tabular_config, cdg_tabular_io, cdg_conv, cdg_bnorm,
 * cdg_celeba_config, cdg_celeba_io, cdg_tvae_transform_config); -1 past the end.  A binding compares these with its own
 * struct declarations at load time, so a stale library is refused instead of being fed mismatched layouts. */
int64_t cdg_abi_sizeof(int which);

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
 * Data parallel (no reference counterpart: the reference is single-device).  One-shot all-reduce(sum) of a SMALL gradient
 * arena (tabular 2.4 KB, CDG-TVAE 12 KB, CelebA's trainable part 49 KB) over NVLink peer memory, one kernel per rank:
 * every rank owns a symmetric buffer of 2 * n_max floats + 64 uint32 flags (zeroed once, mapped into every peer process);
 * peer_ptrs[p] (HOST array, `world` entries) is rank p's buffer as mapped into THIS process.  `step_counter` is one device
 * uint32 (zeroed once) the kernel advances itself, so the launch can be captured and replayed.  Every rank must enqueue the
 * same sequence of calls.  Sums are taken in rank order: all ranks end with bit-identical gradients.
 * ---------------------------------------------------------------------------------------- */
int cdg_allreduce_oneshot(float* grads, int64_t n, int64_t n_max, const uint64_t* peer_ptrs, int32_t world, int32_t rank,
                          uint32_t* step_counter, void* stream);

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
/* Test / diagnostic hook: offset (in floats) of a named region of the step's workspace for the given batch sizes:
 * 0 = pre (after the step: d loss / d pre), 1 + k = a1 of decoder k, 5 + k = a2 of decoder k, 9 = h1, 10 = h2, 11 = ga2,
 * 12 = bf16 planes of a2 of decoder 0; 100 = (not an offset) 1 when the last cdg_pendulum_forward_backward left d loss / d pre
 * as bf16 planes [batch][P] (hi plane, then lo plane, in the bytes of `pre`) instead of fp32; -1 for an unknown id. */
int64_t cdg_pendulum_workspace_offset(const cdg_pendulum_plan* plan, int64_t batch, int64_t batch_l, int which);
/* ... for a step that also runs the InfoMax discriminator (io->d_params != NULL). */
int64_t cdg_pendulum_workspace_bytes_infomax(const cdg_pendulum_plan* p, int64_t batch);

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
    /* InfoMax baseline (modules/model.py:191-206 Discriminator; modules/train.py:71-148 train_InfoMax), when d_params != NULL:
     * a discriminator Linear(P+d,300)-ELU-Linear(300,300)-ELU-Linear(300,1) on (x, epsilon) and on (x, epsilon[perm]);
     * MI = -(mean D_joint - mean exp(D_marginal - 1)); loss += gamma * MI; gradients of loss + MI (train.py:139-140) land
     * in `grads` (model) and `d_grads` (discriminator); logs become [loss, recon, KL, alignment, MutualInfo, variances]. */
    const float* d_params;   /* discriminator arena                                            */
    float* d_grads;
    cdg_linear d_net[3];     /* net.{0,2,4} offsets in that arena                              */
    int64_t d_n_params;
    const int64_t* perm;     /* [batch] row permutation (torch.randperm, train.py:75)          */
    float gamma;
} cdg_pendulum_io;

/* zero_grad + forward + losses + backward of one batch: train.py:168-202 (:235-278 when x_l != NULL).
 * Gradients land in io->grads; no optimizer update. */
int cdg_pendulum_forward_backward(cdg_pendulum_plan* p, const cdg_pendulum_io* io, void* stream);
/* Data parallel: when enabled, cdg_pendulum_forward_backward records an event on its stream as soon as the gradients of a
 * bucket are final -- bucket k < n_dec: decoder k (the backward pass finishes the decoders in index order), bucket n_dec:
 * everything (encoder and flows come last).  The caller makes its communication stream wait for the event
 * (cdg_stream_wait_event) and starts that bucket's all-reduce while the rest of the backward pass is still running.
 * The events belong to the plan (created on first enable, destroyed with it). */
int cdg_pendulum_ready_events_enable(cdg_pendulum_plan* p, int enable);
int cdg_pendulum_ready_event(cdg_pendulum_plan* p, int bucket, void** event_out);
int cdg_stream_wait_event(void* stream, void* event);

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

/* Stand-alone evaluation of the per-node flows on [batch, node] matrices (one launch for all nodes): the forward map with
 * optional log|det| (InvertiblePriorLinear.forward modules/model.py:20-25, PlanarFlows.forward :87-100) or the inverse
 * (InvertiblePriorLinear.inverse :27-29; PlanarFlows.inverse :77-85: `inverse_loop` fixed-point iterations per flow).
 * `params` + flow_off[j] points at node j's parameters ({p0, p1} | {w[F], b[F], u[F]}); direction 0 = forward, 1 = inverse.
 * Replaces the evaluation scripts' `model.inverse(...)` / `layer(x)` calls (inference.py:302-317, metric.py:226-255). */
int cdg_flow_apply(int scm, int flow_num, int inverse_loop, int node, const float* params, const int64_t* flow_off,
                   const float* in, int64_t ld_in, float* out, int64_t ld_out, float* logdet, int64_t ld_logdet,
                   int64_t batch, int direction, void* stream);

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
/* The loan / adult step keeps its parameters in __constant__ memory (copied in front of every launch, stream-ordered).
 * A process that steps two tabular models concurrently on DIFFERENT streams turns this off (0): shared-memory kernels. */
void cdg_tabular_const_params(int32_t on);
/* CDG-TVAE step: 1 (default) = a warp owns 32 rows, every Linear layer a register-tiled fp32 product over shared-memory
 * slabs (tvae_tile.cu); 2 = the same organisation with mma.sync 3xTF32 fragments (tvae_mma.cu: measured no faster);
 * 0 = one row per thread (tvae_fixed.cu).  Kept switchable for the route-equivalence tests. */
void cdg_tabular_tvae_tile(int32_t on);
int cdg_tabular_forward(cdg_tabular_plan* p, const cdg_tabular_io* io, int32_t deterministic, void* stream);

/* ------------------------------------------------------------------------------------------
 * CelebA CDG-VAE (celeba/module/model.py:106-218 CDGVAE; celeba/module/sagan.py:137-210 Generator;
 * celeba/module/train.py:10-76 train_CDGVAE): frozen train-mode ResNet-18 encoder with a trainable
 * fc, two Gaussian posteriors, 5 SAGAN generators (never optimised: only their input gradient is
 * computed), L1 reconstruction.  Two arenas: `params` (trainable: encoder.fc, flows) and `frozen`
 * (everything else, including the state the reference mutates on every forward: BatchNorm running
 * statistics and the spectral-norm u / v vectors).  Activations are NHWC.
 * ---------------------------------------------------------------------------------------- */
#define CDG_CELEBA_GEN 5        /* decoders (model.py:146-151)          */
#define CDG_CELEBA_BLOCKS 5     /* GenBlocks per generator at 128 px    */
#define CDG_CELEBA_RES 8        /* BasicBlocks of ResNet-18             */

/* nn.Conv2d / nn.Linear weight in torch layout (OIHW) at float offset `w` of the frozen arena; bias / spectral-norm
 * u / v offsets are -1 when the layer has none. */
typedef struct { int64_t w, b, u, v; int32_t cin, cout, k, stride, pad; } cdg_conv;
/* nn.BatchNorm2d: weight, bias, running_mean, running_var offsets in the frozen arena */
typedef struct { int64_t weight, bias, running_mean, running_var; int32_t c; float momentum, eps; } cdg_bnorm;
typedef struct { cdg_conv conv1, conv2, conv0; cdg_bnorm bn1, bn2; } cdg_gen_block;          /* sagan.py:104-140 */
typedef struct {
    int32_t z_dim;
    int32_t z_src[CDG_MAX_NODE];           /* latent column i >= 0 that input j of this decoder reads (model.py:190-194);
                                              -(i+1) = column i of epsilon2 (model.py:195)                                 */
    cdg_conv lin0;                         /* GenIniBlock.snlinear0: cin = z_dim, cout = 512*16 (sagan.py:91)            */
    cdg_gen_block blk[CDG_CELEBA_BLOCKS];
    cdg_conv attn[4];                      /* Self_Attn convolutions: sigma == 0, only their power iteration runs        */
    cdg_bnorm bn;                          /* sagan.py:184 (momentum 1e-4) */
    cdg_conv to_rgb;
} cdg_generator;
typedef struct { cdg_conv conv1, conv2, down; cdg_bnorm bn1, bn2, bn_down; int32_t has_down; } cdg_res_block;

typedef struct {
    int32_t node, latent_dim, scm, flow_num;
    int32_t image_size;                    /* 128 */
    int32_t gemm_mode;
    int64_t n_params, n_frozen;
    cdg_linear fc;                         /* encoder.fc in the TRAINABLE arena: 512 -> 2*node + 2*latent_dim (model.py:118) */
    int64_t flow_off[CDG_MAX_NODE];        /* trainable arena */
    float I_B_inv[CDG_MAX_NODE * CDG_MAX_NODE];
    float beta, lambda_;
    cdg_conv rn_conv1; cdg_bnorm rn_bn1;   /* torchvision resnet18 stem */
    cdg_res_block rn_blk[CDG_CELEBA_RES];
    cdg_generator gen[CDG_CELEBA_GEN];
} cdg_celeba_config;

typedef struct cdg_celeba_plan cdg_celeba_plan;
int cdg_celeba_create(const cdg_celeba_config* cfg, cdg_celeba_plan** out);
void cdg_celeba_destroy(cdg_celeba_plan* p);
int64_t cdg_celeba_workspace_bytes(cdg_celeba_plan* p, int64_t batch);

typedef struct {
    const float* params;     /* trainable arena */
    float* grads;            /* its gradient arena (written when backward != 0) */
    float* frozen;           /* frozen arena; running statistics and u / v are updated in place */
    void* workspace;
    int64_t workspace_bytes;
    const float* x;          /* [batch, S, S, ld_x]: channels 0-2 = image in [0,1] (train.py:31) */
    int32_t ld_x;
    const float* y;          /* [batch, ld_y]; first `node` columns used (train.py:55) */
    int32_t ld_y;
    const float* masks;      /* [5][batch, S, S] decoder masks (celeba/main.py:111; model.py:198) */
    const float* noise1;     /* [batch, node]  (model.py:182) */
    const float* noise2;     /* [batch, node]  (model.py:184) */
    int64_t batch;
    int32_t backward;        /* 1: losses + gradients (train.py:27-69); 0: forward only (model.py:202-218) */
    int32_t deterministic;   /* eps = mean (model.py:178-180) */
    int32_t encoder_passes;  /* encode() calls this evaluation stands for: each one updates the ResNet BatchNorm
                                running statistics (2 for forward(), model.py:204 + :212) */
    float* logs;             /* [5]: loss, recon, KL, alignment, active */
    float* xhat;             /* optional [batch, S, S, 3] */
    float* xhat_separated;   /* optional [5][batch, S, S, 3] generator outputs, NHWC */
    float* latents;          /* optional [9][batch, node]: mean1, logvar1, epsilon1, orig_latent, latent, align_latent,
                                mean2, logvar2, epsilon2 */
    int32_t encode_only;     /* 1: stop after the latent block (model.encode / get_posterior): the generators and their
                                BatchNorm / spectral-norm state are not touched */
    const float* latent_in;  /* decode-only entry (model.py:188): [batch, node] latents ... */
    const float* epsilon2_in;/* ... and [batch, latent_dim] epsilon2; when both are given the encoder is skipped */
} cdg_celeba_io;

int cdg_celeba_step(cdg_celeba_plan* p, const cdg_celeba_io* io, void* stream);
/* The five generators of a step run on five internal streams forked from / joined to `stream` with events (default 1);
 * 0 enqueues everything on `stream` alone.  Bits 8 and up, when non-zero, override the number of SMs a small GEMM of a generator
 * chain is split to fill while the chains run side by side (default 148; lower is faster and less accurate).  The workspace layout does not depend on this switch. */
void cdg_celeba_generator_streams(int32_t on);

/* Single layers of that path, exposed for unit tests against torch (tests/test_celeba_gpu.py):
 * NHWC convolution  out[B,Ho,Wo,Co] = conv(x[B,H,W,Ci], w[Co,Ci,k,k]) + bias, and its input gradient. */
int64_t cdg_conv2d_workspace_bytes(int64_t batch, int32_t h, int32_t w, const cdg_conv* cv, int32_t up);
int cdg_conv2d_forward(int mode, const float* x, int64_t batch, int32_t h, int32_t w, const float* weight, const float* bias,
                       const cdg_conv* cv, int32_t up, float* out, void* workspace, int64_t workspace_bytes, void* stream);
int cdg_conv2d_dgrad(int mode, const float* gout, int64_t batch, int32_t h, int32_t w, const float* weight, const cdg_conv* cv,
                     float* gin, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * The dense contraction used by every pendulum Linear layer, exposed for tests and profiling:
 *   C[M,N] (ldc) = sum_k A(m,k) * B(n,k),  A(m,k) = A[m*sa_m + k*sa_k],  B(n,k) = B[n*sb_n + k*sb_k]
 * mode selects the SIMT fp32 kernel or the tcgen05 3xTF32 (fp32-faithful) / 1xTF32 kernels.
 * ---------------------------------------------------------------------------------------- */
int cdg_gemm(int mode, const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_n, int64_t sb_k,
             float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int accumulate,
             void* workspace, int64_t workspace_bytes, void* stream);

/* bf16x3 fast path for weight operands: W = bf16 hi + bf16 lo made once per step (optionally transposed, row stride ld16
 * elements, a multiple of 8), then C[M,N] = sum_k A(m,k) * (b_hi[n*ld16+k] + b_lo[n*ld16+k]) on the tcgen05 kernel. */
int cdg_split_bf16(const float* W, int64_t rows, int64_t cols, int64_t ld, void* hi, void* lo, int64_t ld16, int transpose,
                   void* stream);
int cdg_gemm_bsplit(const float* A, int64_t sa_m, int64_t sa_k, const void* b_hi, const void* b_lo, int64_t ld16, float* C,
                    int64_t ldc, int64_t M, int64_t N, int64_t K, void* stream);
/* The same product with BOTH operands as bf16 (hi, lo) planes (A: [M][ld_a16], B: [N][ld_b16], made by cdg_split_bf16 or by
 * a previous call's out_hi / out_lo): the CTA-pair kernel of csrc/gemm_ps.cu, TMA-fed with no in-kernel conversion.
 * epi: 0 = none, 1 = + bias[n], 2 = ELU(+ bias[n]), 3 = * ELU'(aux[m,n]) (aux = post-activation values, row stride ld_aux).
 * C (row stride ldc, 32-byte aligned rows) and / or out_hi / out_lo ([M][ld_out16] bf16 planes of the fp32 result) are written.
 * Returns CDG_ERR_UNSUPPORTED for shapes the kernel does not take (M < 128, N % 4 != 0, K > 2048, misaligned operands). */
int cdg_gemm_planes(const void* a_hi, const void* a_lo, int64_t ld_a16, const void* b_hi, const void* b_lo, int64_t ld_b16,
                    float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int epi, const float* bias, const float* aux,
                    int64_t ld_aux, void* out_hi, void* out_lo, int64_t ld_out16, void* stream);

/* Long contractions on planes (csrc/gemm_pk.cu): C[M,N] += sum_k A(m,k) B(n,k), N <= 304, K cut into segments of <= 2048
 * whose partial tiles are added to C with fp32 atomics (C must hold the value to accumulate into, e.g. zeros).
 * mn_major = 0: planes are [M][ld_a16] / [N][ld_b16] (contraction index contiguous: input gradients);
 * mn_major = 1: planes are [K][ld_a16] / [K][ld_b16] (contraction index outermost: weight gradients, contraction over the
 * batch, both operands read as they lie in memory).  extra_col (optional, [M]): column N-1 of the product is added there
 * instead of to C (a ones column in B then yields the bias gradient); C then has N-1 columns. */
int cdg_gemm_planes_acc(const void* a_hi, const void* a_lo, int64_t ld_a16, const void* b_hi, const void* b_lo, int64_t ld_b16,
                        float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int mn_major, float* extra_col, void* stream);

/* ------------------------------------------------------------------------------------------
 * CDG-TVAE data transform, APPLY side: the step on either side of train_TVAE.
 * Fitting (BayesianGaussianMixture, category discovery) stays on the host; its result arrives here as tables.
 *
 * cdg_tvae_transform replaces DataTransformer.transform (tabular/modules/data_transformer.py:163-182) =
 *   per continuous column ClusterBasedNormalizer._transform (tabular/modules/numerical.py:407-445: GMM
 *   responsibilities, +1e-6, renormalise, component drawn by inverse-cdf from ONE uniform per cell, value
 *   (x - mean_c) / (4 std_c) clipped to +-0.99) laid out by _transform_continuous (data_transformer.py:111-125:
 *   scalar, then one-hot of the component), per discrete column the one-hot of _transform_discrete (:127-129).
 * cdg_tvae_inverse_transform replaces DataTransformer.inverse_transform (data_transformer.py:184-227) =
 *   argmax of the component block, optional N(value, sigmas[start]) draw (:138-140), clip to +-1 and
 *   value * 4 std_c + mean_c (numerical.py:447-457), rounding of integer-typed columns (numerical.py:175-177);
 *   discrete columns: argmax -> category value.
 * cdg_gumbel_argmax replaces the Cover_Type draw of tabular/inference_tvae.py:232-235, :250-253:
 *   argmax_j( log_softmax(logits)_j + log(-log(U_j + 1e-20) + 1e-20) ), fp32, first maximum wins.
 *
 * All randomness is INJECTED (the reference draws it from NumPy's / torch's global generators on the host):
 *   uniforms[cc * rows + r] / normals[cc * rows + r] for the cc-th CONTINUOUS column and row r, i.e. the order in
 *   which the reference's column-by-column loop consumes its stream.  raw tables are float64 row-major (NumPy's
 *   dtype in the reference), transformed tables fp32 row-major (what the training step reads).
 * Per-cell arithmetic is fp64 so that component indices and values agree with the reference bit for bit.
 * ---------------------------------------------------------------------------------------- */
#define CDG_MAX_TCOL 16     /* raw columns of a table                         */
#define CDG_MAX_TCOMP 10    /* GMM components (DataTransformer max_clusters)  */
#define CDG_MAX_TCAT 16     /* categories of a discrete column                */
enum { CDG_TCOL_CONTINUOUS = 0, CDG_TCOL_DISCRETE = 1 };
typedef struct {
    int32_t kind;                       /* CDG_TCOL_*                                                             */
    int32_t out_start;                  /* first column of this raw column's block in the transformed table       */
    int32_t n_all;                      /* continuous: components the mixture was fitted with (<= CDG_MAX_TCOMP)  */
    int32_t n_valid;                    /* continuous: components kept (weight > threshold); discrete: categories */
    int32_t valid_idx[CDG_MAX_TCOMP];   /* continuous: mixture index of the j-th kept component                   */
    int32_t round_int;                  /* inverse: raw dtype is integer -> round half to even                    */
    /* continuous: log( weight_k N(x; .) ) = log_a[k] - 0.5 * prec[k] * (x - mean[k])^2 for ALL n_all components
     * (for sklearn's BayesianGaussianMixture: precisions_cholesky_^2 * degrees_of_freedom_ and the digamma terms of
     * _estimate_log_weights / _estimate_log_prob folded into log_a by the host), std[k] = sqrt(covariances_[k]). */
    double mean[CDG_MAX_TCOMP], std[CDG_MAX_TCOMP], prec[CDG_MAX_TCOMP], log_a[CDG_MAX_TCOMP];
    double category[CDG_MAX_TCAT];      /* discrete: value of the j-th one-hot position                           */
} cdg_tvae_column;
typedef struct {
    int32_t n_col;                      /* raw columns                       */
    int32_t out_dim;                    /* transformed columns (sum of blocks) */
    cdg_tvae_column col[CDG_MAX_TCOL];
} cdg_tvae_transform_config;

int cdg_tvae_transform(const cdg_tvae_transform_config* cfg, const double* raw, int64_t ld_raw, const double* uniforms,
                       int64_t rows, float* out, int64_t ld_out, void* stream);
/* sigmas (fp32 [out_dim]) and normals may both be NULL: no draw, as inverse_transform(data) without sigmas. */
int cdg_tvae_inverse_transform(const cdg_tvae_transform_config* cfg, const float* data, int64_t ld_data, const float* sigmas,
                               const double* normals, int64_t rows, double* raw_out, int64_t ld_raw, void* stream);
int cdg_gumbel_argmax(const float* logits, int64_t ld, int32_t n_class, const float* uniforms, int64_t rows,
                      int64_t* out_index, void* stream);

/* ------------------------------------------------------------------------------------------
 * Image bytes -> training-step input.  Replaces the host-side conversion of modules/datasets.py:28, :57
 * (`(np.array(train_x).astype(float) - 127.5) / 127.5`, float64) + `torch.FloatTensor(...)` (:42, :64): out[i] =
 * (float)(((double)pixels[i] - 127.5) / 127.5), bit-identical, so that the dataset's native bytes (1 per pixel instead
 * of 4) are what crosses PCIe every step (data.py::DevicePrefetcher(pixels=True)).
 * ---------------------------------------------------------------------------------------- */
int cdg_pixels_to_float(const uint8_t* pixels, int64_t n, float* out, void* stream);
/* Batch assembly from a device-resident uint8 dataset [n_images][row_bytes]: out[r] = the same conversion of image idx[r]
 * (what DataLoader(shuffle=True) + collate + modules/datasets.py:28 do on the host), one pass, coalesced on both sides.
 * idx: int64 on the device, values in [0, n_images) (not checked); row_bytes a multiple of 16. */
int cdg_pixels_gather_to_float(const uint8_t* images, int64_t n_images, int64_t row_bytes, const int64_t* idx, int64_t rows,
                               float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDGVAE_H */
