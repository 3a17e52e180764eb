// CelebA CDG-VAE step (celeba/module/model.py:106-218, celeba/module/sagan.py:74-210, celeba/module/train.py:10-76)
// behind cdg_celeba_step(): frozen train-mode ResNet-18 -> fc -> two posteriors -> causal latent block -> 5 SAGAN
// generators -> masked sum, tanh, L1 reconstruction -> input-gradient-only backward through the generators -> fc / flow
// gradients.  Convolutions are im2col + the tcgen05 3xTF32 GEMM (gemm_tc.cu); everything else is the memory-bound
// kernels of conv.cu and the per-row latent kernels below.  One host pass enqueues the whole step on the caller's
// stream; the same pass run "dry" sizes the workspace.
#include <map>
#include <new>

#include "conv.cuh"
#include "latent.cuh"

namespace cdg {

namespace {

constexpr int kAccDoubles = 1 << 17;      // BatchNorm statistic accumulators of one step (zeroed once)
constexpr int kNGen = CDG_CELEBA_GEN, kNBlk = CDG_CELEBA_BLOCKS;

struct Bump {
    char* base = nullptr;
    int64_t off = 0, peak = 0;
    template <typename T>
    T* take(int64_t n) {
        off = (off + 255) & ~int64_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * (int64_t)sizeof(T);
        if (off > peak) peak = off;
        return p;
    }
};

struct Cx {
    cudaStream_t s = nullptr;
    bool dry = true;
    int mode = CDG_GEMM_AUTO;
    Bump ws;
    float* col = nullptr;
    int64_t col_cap = 0, col_need = 0;
    double* acc = nullptr;
    int64_t acc_used = 0;
    float* frozen = nullptr;
    double* take_acc(int n) {
        double* p = acc ? acc + acc_used : nullptr;
        acc_used += n;
        return p;
    }
};

#define RUN(expr)                           \
    do {                                    \
        if (!cx.dry) CDG_TRY(expr);         \
    } while (0)

struct BnState { float *scale, *shift, *mean, *rstd; };

// ---- convolution = im2col + GEMM ----------------------------------------------------------------------------
struct Geom { int Ho, Wo, K, Kp; int64_t M; };
static Geom conv_geom(int64_t B, int Hs, int Ws, int up, int C, int k, int stride, int pad) {
    Geom g;
    const int Hin = Hs * up, Win = Ws * up;
    g.Ho = (Hin + 2 * pad - k) / stride + 1;
    g.Wo = (Win + 2 * pad - k) / stride + 1;
    g.K = k * k * C;
    g.Kp = round_up4(g.K);
    g.M = B * g.Ho * g.Wo;
    return g;
}

// a prepared weight matrix [N][ld] fp32 and, optionally, its bf16 (hi, lo) split for the bf16x3 kernel
struct WMat {
    const float* w = nullptr;
    const void* hi = nullptr; const void* lo = nullptr; int ld16 = 0;
    WMat(const float* p = nullptr) : w(p) {}
};
// weights as the B operand.  gemm_mode "bf3x": the bf16x3 kernel with ready-made weight tiles (18.0 vs 19.2 ms per step at
// batch 16).  It is NOT the default here: this step amplifies GEMM rounding by 1e3 .. 1e4 (tools/celeba_grad_noise.py), and
// bf16x3's 4-8e-6 per GEMM lands the fc gradient at 1.8e-2 of an fp64 evaluation where 3xTF32 (and the reference's own fp32)
// stay within 1e-2 -- so "auto" keeps the 3xTF32 arithmetic for the CelebA convolutions.
static int gemm_w(Cx& cx, GemmDesc& g, const WMat& wm) {
    if (wm.hi && cx.mode == CDG_GEMM_BF3X) {
        g.b_hi16 = wm.hi; g.b_lo16 = wm.lo; g.ld_b16 = wm.ld16;
        const int r = gemm_tc(g, 2, nullptr, 0, cx.s);
        if (r != CDG_ERR_UNSUPPORTED) return r;
        g.b_hi16 = g.b_lo16 = nullptr;
    }
    if (g.conv_C > 0) return gemm_tc(g, cx.mode == CDG_GEMM_TC1X ? 1 : (cx.mode == CDG_GEMM_BF3X ? 2 : 3), nullptr, 0, cx.s);
    return gemm_dispatch(cx.mode, g, nullptr, 0, cx.s);
}
static int gemm_nt(Cx& cx, const float* A, int64_t lda, const WMat& Bm, int64_t ldb, float* C, int64_t ldc, int64_t M, int64_t N,
                   int64_t K, const float* bias) {
    GemmDesc g{};
    g.A = A; g.sa_m = lda; g.sa_k = 1;
    g.B = Bm.w; g.sb_n = ldb; g.sb_k = 1;
    g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
    g.epi = bias ? EPI_BIAS : EPI_NONE;
    g.bias = bias;
    return gemm_w(cx, g, Bm);
}

// out[M, Co] = conv(act(src)) with the prepared forward matrix wf[Co][Kp]
static int conv_fwd(Cx& cx, const float* src, int64_t B, int Hs, int Ws, int C, int ld, const BnState* bn, int relu, int up,
                    int k, int stride, int pad, const WMat& wf, int Co, const float* bias, float* out) {
    const Geom g = conv_geom(B, Hs, Ws, up, C, k, stride, pad);
    if (k == 1 && stride == 1 && up == 1 && !bn && !relu && ld == C && C % 4 == 0) {
        RUN(gemm_nt(cx, src, C, wf, g.Kp, out, Co, g.M, Co, g.Kp, bias));
        return CDG_OK;
    }
    // implicit GEMM (gemm_tc.cu conv mode): stride-1 "same" convolutions over >= 32 channels read the activation itself
    // through 4-D TMA boxes with a zero-filled halo; only the activated / upsampled input is materialised (1x, not k*k x)
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    const int Hin = Hs * up, Win = Ws * up;
    if (cx.mode != CDG_GEMM_SIMT && k > 1 && (k & 1) && stride == 1 && pad == (k - 1) / 2 && C % 32 == 0 && ld == C &&
        pow2(Hin) && pow2(Win) && (Win < 128 || Win % 128 == 0) && g.M * Co >= 4096) {
        const bool materialise = bn || relu || up != 1;
        if (materialise && g.M * C > cx.col_need) cx.col_need = g.M * C;
        if (cx.dry) return CDG_OK;
        const float* act = src;
        if (materialise) {
            CDG_REQUIRE(g.M * C <= cx.col_cap, "activation scratch too small");
            Im2colArgs a{};
            a.src = src; a.B = B; a.Hs = Hs; a.Ws = Ws; a.C = C; a.ld = ld;
            a.scale = bn ? bn->scale : nullptr; a.shift = bn ? bn->shift : nullptr;
            a.relu = relu; a.up = up; a.k = 1; a.stride = 1; a.pad = 0; a.Ho = Hin; a.Wo = Win;
            a.col = cx.col; a.Kp = C;
            CDG_TRY(launch_im2col(a, cx.s));
            act = cx.col;
        }
        GemmDesc d{};
        d.A = act; d.sa_m = C; d.sa_k = 1;
        d.B = wf.w; d.sb_n = g.Kp; d.sb_k = 1;
        d.C = out; d.ldc = Co; d.M = g.M; d.N = Co; d.K = g.K;
        d.epi = bias ? EPI_BIAS : EPI_NONE; d.bias = bias;
        d.conv_C = C; d.conv_H = Hin; d.conv_W = Win; d.conv_k = k; d.conv_B = B;
        const int r = gemm_w(cx, d, wf);
        if (r != CDG_ERR_UNSUPPORTED) return r;
        CDG_REQUIRE(false, "implicit-GEMM convolution rejected a shape the planner accepted (C=%d H=%d W=%d Co=%d)", C, Hin, Win, Co);
    }
    if (g.M * g.Kp > cx.col_need) cx.col_need = g.M * g.Kp;
    if (cx.dry) return CDG_OK;
    CDG_REQUIRE(g.M * g.Kp <= cx.col_cap, "im2col scratch too small");
    Im2colArgs a{};
    a.src = src; a.B = B; a.Hs = Hs; a.Ws = Ws; a.C = C; a.ld = ld;
    a.scale = bn ? bn->scale : nullptr; a.shift = bn ? bn->shift : nullptr;
    a.relu = relu; a.up = up; a.k = k; a.stride = stride; a.pad = pad; a.Ho = g.Ho; a.Wo = g.Wo;
    a.col = cx.col; a.Kp = g.Kp;
    CDG_TRY(launch_im2col(a, cx.s));
    CDG_TRY(gemm_nt(cx, cx.col, g.Kp, wf, g.Kp, out, Co, g.M, Co, g.Kp, bias));
    return CDG_OK;
}

// gin[M, Ci] = input gradient of a stride-1 "same" convolution, with wd[Ci][Kpd] (flipped kernel, K order (kh,kw,co))
static int conv_dgrad(Cx& cx, const float* gout, int64_t B, int H, int W, int Co, int k, const WMat& wd, int Ci, float* gin) {
    return conv_fwd(cx, gout, B, H, W, Co, Co, nullptr, 0, 1, k, 1, (k - 1) / 2, wd, Ci, nullptr, gin);
}

static int bn_forward(Cx& cx, const float* x, int64_t M, const cdg_bnorm& bn, int n_updates, BnState* st, bool keep_stats) {
    st->scale = cx.ws.take<float>(bn.c);
    st->shift = cx.ws.take<float>(bn.c);
    st->mean = keep_stats ? cx.ws.take<float>(bn.c) : nullptr;
    st->rstd = keep_stats ? cx.ws.take<float>(bn.c) : nullptr;
    double* acc = cx.take_acc(2 * bn.c);
    if (cx.dry) return CDG_OK;
    float* f = cx.frozen;
    CDG_TRY(launch_col_stats(x, M, bn.c, acc, cx.s));
    CDG_TRY(launch_bn_finalize(acc, M, bn.c, f + bn.weight, f + bn.bias, bn.eps, bn.momentum, n_updates, f + bn.running_mean,
                               f + bn.running_var, st->scale, st->shift, st->mean, st->rstd, cx.s));
    return CDG_OK;
}

// dx = backward of relu(bn(x)) given g = d/d(relu output); optional `add`; in place over g allowed
static int bn_backward(Cx& cx, const float* g, const float* x, int64_t M, int C, const BnState& st, const float* add, float* dx) {
    double* acc = cx.take_acc(2 * C);
    if (cx.dry) return CDG_OK;
    CDG_TRY(launch_bn_bwd_reduce(g, x, st.scale, st.shift, st.mean, st.rstd, M, C, acc, cx.s));
    CDG_TRY(launch_bn_bwd_apply(g, x, st.scale, st.shift, st.mean, st.rstd, acc, add, dx, M, C, cx.s));
    return CDG_OK;
}

// ---- prepared weights -------------------------------------------------------------------------------------------
struct PConv {
    float* wf = nullptr; float* wd = nullptr; int Kpf = 0, Kpd = 0; float* inv_sigma = nullptr;
    Split16 f16, d16;
    WMat fw() const { WMat m(wf); m.hi = f16.hi; m.lo = f16.lo; m.ld16 = f16.ld; return m; }
    WMat dw() const { WMat m(wd); m.hi = d16.hi; m.lo = d16.lo; m.ld16 = d16.ld; return m; }
};

static int prep_conv(Cx& cx, const cdg_conv& cv, bool need_wd, PConv* p) {
    p->Kpf = round_up4(cv.k * cv.k * cv.cin);
    p->Kpd = round_up4(cv.k * cv.k * cv.cout);
    p->wf = cx.ws.take<float>((int64_t)cv.cout * p->Kpf);
    p->wd = need_wd ? cx.ws.take<float>((int64_t)cv.cin * p->Kpd) : nullptr;
    p->f16.ld = (p->Kpf + 7) & ~7; p->d16.ld = (p->Kpd + 7) & ~7;
    p->f16.hi = cx.ws.take<uint16_t>((int64_t)cv.cout * p->f16.ld);
    p->f16.lo = cx.ws.take<uint16_t>((int64_t)cv.cout * p->f16.ld);
    if (need_wd) {
        p->d16.hi = cx.ws.take<uint16_t>((int64_t)cv.cin * p->d16.ld);
        p->d16.lo = cx.ws.take<uint16_t>((int64_t)cv.cin * p->d16.ld);
    }
    if (cx.dry) return CDG_OK;
    const bool want16 = cx.mode == CDG_GEMM_BF3X;            // the bf16 (hi, lo) copies only feed the opt-in bf16x3 path
    if (!want16) { p->f16 = Split16(); p->d16 = Split16(); }
    CDG_TRY(launch_weight_prep(cx.frozen + cv.w, cv.cout, cv.cin, cv.k, p->inv_sigma, p->wf, p->Kpf, p->wd, p->Kpd, cx.s, p->f16,
                               need_wd ? p->d16 : Split16()));
    return CDG_OK;
}

// ---- ResNet-18 (torchvision) forward, train-mode BatchNorm, no gradient (model.py:117-125: frozen) ---------------
static int resnet_forward(Cx& cx, const cdg_celeba_config& c, const float* x, int ld_x, int64_t B, int n_upd, float** feat_out) {
    const int S = c.image_size;
    PConv p0;
    CDG_TRY(prep_conv(cx, c.rn_conv1, false, &p0));
    const Geom g0 = conv_geom(B, S, S, 1, 3, 7, 2, 3);
    float* y0 = cx.ws.take<float>(g0.M * 64);
    CDG_TRY(conv_fwd(cx, x, B, S, S, 3, ld_x, nullptr, 0, 1, 7, 2, 3, p0.fw(), 64, nullptr, y0));
    BnState b0;
    CDG_TRY(bn_forward(cx, y0, g0.M, c.rn_bn1, n_upd, &b0, false));
    int H = (g0.Ho - 1) / 2 + 1;
    float* h = cx.ws.take<float>(B * H * H * 64);
    RUN(launch_maxpool_bn_relu(y0, b0.scale, b0.shift, h, B, g0.Ho, g0.Wo, 64, cx.s));
    int C = 64;
    for (int i = 0; i < CDG_CELEBA_RES; ++i) {
        const cdg_res_block& rb = c.rn_blk[i];
        const int Co = rb.conv1.cout, st = rb.conv1.stride;
        PConv p1, p2, pd;
        CDG_TRY(prep_conv(cx, rb.conv1, false, &p1));
        CDG_TRY(prep_conv(cx, rb.conv2, false, &p2));
        const Geom g1 = conv_geom(B, H, H, 1, C, 3, st, 1);
        float* o1 = cx.ws.take<float>(g1.M * Co);
        CDG_TRY(conv_fwd(cx, h, B, H, H, C, C, nullptr, 0, 1, 3, st, 1, p1.fw(), Co, nullptr, o1));
        BnState s1, s2, sd;
        CDG_TRY(bn_forward(cx, o1, g1.M, rb.bn1, n_upd, &s1, false));
        float* o2 = cx.ws.take<float>(g1.M * Co);
        CDG_TRY(conv_fwd(cx, o1, B, g1.Ho, g1.Wo, Co, Co, &s1, 1, 1, 3, 1, 1, p2.fw(), Co, nullptr, o2));
        CDG_TRY(bn_forward(cx, o2, g1.M, rb.bn2, n_upd, &s2, false));
        float* out = cx.ws.take<float>(g1.M * Co);
        if (rb.has_down) {
            CDG_TRY(prep_conv(cx, rb.down, false, &pd));
            float* idn = cx.ws.take<float>(g1.M * Co);
            CDG_TRY(conv_fwd(cx, h, B, H, H, C, C, nullptr, 0, 1, 1, st, 0, pd.fw(), Co, nullptr, idn));
            CDG_TRY(bn_forward(cx, idn, g1.M, rb.bn_down, n_upd, &sd, false));
            RUN(launch_bn_act(o2, s2.scale, s2.shift, idn, sd.scale, sd.shift, 1, out, g1.M, Co, cx.s));
        } else {
            RUN(launch_bn_act(o2, s2.scale, s2.shift, h, nullptr, nullptr, 1, out, g1.M, Co, cx.s));
        }
        h = out; H = g1.Ho; C = Co;
    }
    float* feat = cx.ws.take<float>(B * C);
    RUN(launch_avgpool(h, B, H * H, C, feat, cx.s));
    *feat_out = feat;
    return CDG_OK;
}

// ---- SAGAN generator (sagan.py:137-210, image_size 128) ---------------------------------------------------------
struct GenRun {
    PConv lin0, c1[kNBlk], c2[kNBlk], c0[kNBlk], rgb;
    float* lin_wf = nullptr; float* lin_bf = nullptr;
    float* h_in[kNBlk]; float* y1[kNBlk];
    BnState bn1[kNBlk], bn2[kNBlk], bn_out;
    float* h_last = nullptr;        // [B*S*S, 32]
    float* rgb_pre = nullptr;       // [B*S*S, 3]
};

static int generator_prepare(Cx& cx, const cdg_generator& G, GenRun* r) {
    // one power iteration for every spectral-norm layer (21 per generator), then the scaled weight layouts
    SnBatch sb{};
    const cdg_conv* all[kSnMax];
    int n = 0;
    all[n++] = &G.lin0;
    for (int b = 0; b < kNBlk; ++b) { all[n++] = &G.blk[b].conv1; all[n++] = &G.blk[b].conv2; all[n++] = &G.blk[b].conv0; }
    for (int i = 0; i < 4; ++i) all[n++] = &G.attn[i];
    all[n++] = &G.to_rgb;
    float* inv = cx.ws.take<float>(n);
    sb.n = n;
    for (int i = 0; i < n; ++i) {
        const cdg_conv& cv = *all[i];
        SnLayer& L = sb.l[i];
        L.rows = cv.cout; L.cols = cv.cin * cv.k * cv.k;
        L.t = cx.ws.take<float>(L.cols);
        L.sv = cx.ws.take<float>(L.rows);
        L.inv_sigma = inv ? inv + i : nullptr;
        if (!cx.dry) { L.w = cx.frozen + cv.w; L.u = cx.frozen + cv.u; L.v = cx.frozen + cv.v; }
    }
    RUN(launch_spectral_norm(sb, cx.s));
    auto sig = [&](int i) { return inv ? inv + i : nullptr; };
    const int C0 = G.blk[0].conv1.cin, HW = 16;
    r->lin_wf = cx.ws.take<float>((int64_t)C0 * HW * G.z_dim);
    r->lin_bf = cx.ws.take<float>((int64_t)C0 * HW);
    RUN(launch_lin0_prep(cx.frozen + G.lin0.w, cx.frozen + G.lin0.b, C0, HW, G.z_dim, sig(0), r->lin_wf, r->lin_bf, cx.s));
    for (int b = 0; b < kNBlk; ++b) {
        r->c1[b].inv_sigma = sig(1 + 3 * b); r->c2[b].inv_sigma = sig(2 + 3 * b); r->c0[b].inv_sigma = sig(3 + 3 * b);
        CDG_TRY(prep_conv(cx, G.blk[b].conv1, true, &r->c1[b]));
        CDG_TRY(prep_conv(cx, G.blk[b].conv2, true, &r->c2[b]));
        CDG_TRY(prep_conv(cx, G.blk[b].conv0, true, &r->c0[b]));
    }
    r->rgb.inv_sigma = sig(n - 1);
    CDG_TRY(prep_conv(cx, G.to_rgb, true, &r->rgb));
    return CDG_OK;
}

static int generator_forward(Cx& cx, const cdg_generator& G, GenRun* r, const float* zin, int64_t B) {
    float* f = cx.frozen;
    const int C0 = G.blk[0].conv1.cin;
    float* h = cx.ws.take<float>(B * 16 * C0);
    RUN(gemm_nt(cx, zin, G.z_dim, r->lin_wf, G.z_dim, h, (int64_t)16 * C0, B, (int64_t)16 * C0, G.z_dim, r->lin_bf));
    int H = 4;
    for (int b = 0; b < kNBlk; ++b) {
        const cdg_gen_block& gb = G.blk[b];
        const int Ci = gb.conv1.cin, Co = gb.conv1.cout;
        const int64_t Mlo = B * H * H, Mhi = Mlo * 4;
        r->h_in[b] = h;
        CDG_TRY(bn_forward(cx, h, Mlo, gb.bn1, 1, &r->bn1[b], true));
        float* y1 = r->y1[b] = cx.ws.take<float>(Mhi * Co);
        CDG_TRY(conv_fwd(cx, h, B, H, H, Ci, Ci, &r->bn1[b], 1, 2, 3, 1, 1, r->c1[b].fw(), Co, f + gb.conv1.b, y1));
        CDG_TRY(bn_forward(cx, y1, Mhi, gb.bn2, 1, &r->bn2[b], true));
        float* y2 = cx.ws.take<float>(Mhi * Co);
        CDG_TRY(conv_fwd(cx, y1, B, 2 * H, 2 * H, Co, Co, &r->bn2[b], 1, 1, 3, 1, 1, r->c2[b].fw(), Co, f + gb.conv2.b, y2));
        // skip path: a 1x1 convolution commutes with nearest upsampling, so conv_0 runs at the low resolution
        // (not handed back to the allocator: the next taker would be another generator's chain on another stream)
        float* sk = cx.ws.take<float>(Mlo * Co);
        CDG_TRY(conv_fwd(cx, h, B, H, H, Ci, Ci, nullptr, 0, 1, 1, 1, 0, r->c0[b].fw(), Co, f + gb.conv0.b, sk));
        RUN(launch_add_up2(y2, sk, y2, B, H, H, Co, cx.s));
        h = y2; H *= 2;
    }
    r->h_last = h;
    const int Cl = G.to_rgb.cin;
    CDG_TRY(bn_forward(cx, h, B * H * H, G.bn, 1, &r->bn_out, true));
    r->rgb_pre = cx.ws.take<float>(B * H * H * 3);
    CDG_TRY(conv_fwd(cx, h, B, H, H, Cl, Cl, &r->bn_out, 1, 1, 3, 1, 1, r->rgb.fw(), 3, f + G.to_rgb.b, r->rgb_pre));
    return CDG_OK;
}

// g_pre: [B*S*S, 3] gradient w.r.t. the toRGB pre-activation;  g_zin: [B, z_dim]
static int generator_backward(Cx& cx, const cdg_generator& G, const GenRun& r, const float* g_pre, int64_t B, int S, float* g_zin) {
    const size_t mark = cx.ws.off;
    const int Cl = G.to_rgb.cin;
    int H = S;
    float* g = cx.ws.take<float>(B * H * H * Cl);
    CDG_TRY(conv_dgrad(cx, g_pre, B, H, H, 3, 3, r.rgb.dw(), Cl, g));
    CDG_TRY(bn_backward(cx, g, r.h_last, B * H * H, Cl, r.bn_out, nullptr, g));
    for (int b = kNBlk - 1; b >= 0; --b) {
        const cdg_gen_block& gb = G.blk[b];
        const int Ci = gb.conv1.cin, Co = gb.conv1.cout;
        const int Hl = H / 2;
        const int64_t Mhi = B * H * H, Mlo = B * Hl * Hl;
        float* ga2 = cx.ws.take<float>(Mhi * Co);
        CDG_TRY(conv_dgrad(cx, g, B, H, H, Co, 3, r.c2[b].dw(), Co, ga2));
        CDG_TRY(bn_backward(cx, ga2, r.y1[b], Mhi, Co, r.bn2[b], nullptr, ga2));
        float* gup = cx.ws.take<float>(Mhi * Ci);
        CDG_TRY(conv_dgrad(cx, ga2, B, H, H, Co, 3, r.c1[b].dw(), Ci, gup));
        float* ga1 = cx.ws.take<float>(Mlo * Ci);
        RUN(launch_downsum2(gup, ga1, B, Hl, Hl, Ci, cx.s));
        float* gs = cx.ws.take<float>(Mlo * Co);
        RUN(launch_downsum2(g, gs, B, Hl, Hl, Co, cx.s));
        float* gskip = cx.ws.take<float>(Mlo * Ci);
        CDG_TRY(conv_dgrad(cx, gs, B, Hl, Hl, Co, 1, r.c0[b].dw(), Ci, gskip));
        CDG_TRY(bn_backward(cx, ga1, r.h_in[b], Mlo, Ci, r.bn1[b], gskip, ga1));
        g = ga1; H = Hl;
    }
    // GenIniBlock: g [B*16, C0] is [B, 8192] in the prepared (hw, c) order; g_zin = g @ lin_wf
    const int64_t N0 = (int64_t)16 * G.blk[0].conv1.cin;
    if (!cx.dry) {
        GemmDesc d{};
        d.A = g; d.sa_m = N0; d.sa_k = 1;
        d.B = r.lin_wf; d.sb_n = 1; d.sb_k = G.z_dim;
        d.C = g_zin; d.ldc = G.z_dim; d.M = B; d.N = G.z_dim; d.K = N0;
        CDG_TRY(gemm_dispatch(cx.mode, d, nullptr, 0, cx.s));
    }
    (void)mark;          // the temporaries stay allocated: the next generator's chain runs concurrently on another stream
    return CDG_OK;
}

// ---- latent block with two posteriors (model.py:157-186, train.py:36-63) -----------------------------------------
enum { CACC_RECON = 0, CACC_KL1, CACC_KL2, CACC_ALIGN, CACC_VAR1, CACC_VAR2 = CACC_VAR1 + CDG_MAX_NODE, CACC_LEN = CACC_VAR2 + CDG_MAX_NODE };

struct CLatArgs {
    int d, d2, scm, flow_num, deterministic;
    int64_t batch;
    const float* params;
    float* grads;
    int64_t flow_off[CDG_MAX_NODE];
    float A[CDG_MAX_NODE * CDG_MAX_NODE];
    float beta, lambda_;
    const float* h; int ldh;            // fc output [B, 2d + 2d2]
    const float *n1, *n2;
    const float* y; int ld_y;
    float* lat;                         // optional [9][B, d]
    float* zin[kNGen]; const float* gzin[kNGen];
    int zdim[kNGen]; int zsrc[kNGen][CDG_MAX_NODE];
    float* g_h;                         // [B, ldh]
    double* acc;
};

__global__ void __launch_bounds__(128) celeba_latent_fwd_kernel(CLatArgs a) {
    __shared__ FlowTable ft;
    __shared__ double red[32];
    load_flow_table(ft, a);
    __syncthreads();
    const int d = a.d, d2 = a.d2;
    double kl1 = 0.0, kl2 = 0.0, al = 0.0;
    float v1[CDG_MAX_NODE], v2[CDG_MAX_NODE];
#pragma unroll
    for (int i = 0; i < CDG_MAX_NODE; ++i) v1[i] = v2[i] = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.batch; b += (int64_t)gridDim.x * blockDim.x) {
        const float* h = a.h + b * a.ldh;
        float m1[CDG_MAX_NODE], e1[CDG_MAX_NODE], e2[CDG_MAX_NODE], u[CDG_MAX_NODE], ud[CDG_MAX_NODE], z[CDG_MAX_NODE];
        float k1 = 0.f, k2 = 0.f;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            m1[i] = e1[i] = e2[i] = 0.f;
            if (i < d) {
                m1[i] = h[i];
                const float lv = h[d + i], ev = expf(lv);
                e1[i] = a.deterministic ? m1[i] : m1[i] + expf(lv / 2.f) * a.n1[b * d + i];
                k1 += m1[i] * m1[i] - lv + ev;
                v1[i] += ev;
                if (a.lat) { a.lat[(0 * a.batch + b) * d + i] = m1[i]; a.lat[(1 * a.batch + b) * d + i] = lv; a.lat[(2 * a.batch + b) * d + i] = e1[i]; }
            }
            if (i < d2) {
                const float m2 = h[2 * d + i], lv = h[2 * d + d2 + i], ev = expf(lv);
                e2[i] = a.deterministic ? m2 : m2 + expf(lv / 2.f) * a.n2[b * d + i];      // noise2 is [B, node] (model.py:184)
                k2 += m2 * m2 - lv + ev;
                v2[i] += ev;
                if (a.lat) { a.lat[(6 * a.batch + b) * d + i] = m2; a.lat[(7 * a.batch + b) * d + i] = lv; a.lat[(8 * a.batch + b) * d + i] = e2[i]; }
            }
        }
        kl1 += 0.5 * (double)(k1 - (float)d);
        kl2 += 0.5 * (double)(k2 - (float)d);                          // train.py:48 subtracts node here too
        matvec_A(ft, d, e1, u);
        matvec_A(ft, d, m1, ud);
        float als = 0.f;
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) {
            z[j] = 0.f;
            if (j < d) {
                z[j] = flow_fwd(ft, a.scm, a.flow_num, j, u[j]);
                const float zd = flow_fwd(ft, a.scm, a.flow_num, j, ud[j]);
                if (a.lat) { a.lat[(3 * a.batch + b) * d + j] = u[j]; a.lat[(4 * a.batch + b) * d + j] = z[j]; a.lat[(5 * a.batch + b) * d + j] = zd; }
                if (a.y) {
                    const float yh = 1.f / (1.f + expf(-zd)), y = a.y[b * a.ld_y + j];
                    als += (y - 1.f) * fmaxf(log1pf(-yh), -100.f) - y * fmaxf(logf(yh), -100.f);
                }
            }
        }
        al += (double)als;
#pragma unroll
        for (int k = 0; k < kNGen; ++k) {
            if (!a.zin[k]) continue;
            for (int j = 0; j < a.zdim[k]; ++j) {
                const int src = a.zsrc[k][j];
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < CDG_MAX_NODE; ++i) {
                    if (src == i) v = z[i];
                    if (src == -1 - i) v = e2[i];
                }
                a.zin[k][b * a.zdim[k] + j] = v;
            }
        }
    }
    if (!a.acc) return;
    double s = block_sum<double>(kl1, red);
    if (threadIdx.x == 0) atomicAdd(a.acc + CACC_KL1, s);
    s = block_sum<double>(kl2, red);
    if (threadIdx.x == 0) atomicAdd(a.acc + CACC_KL2, s);
    s = block_sum<double>(al, red);
    if (threadIdx.x == 0) atomicAdd(a.acc + CACC_ALIGN, s);
#pragma unroll
    for (int i = 0; i < CDG_MAX_NODE; ++i) {
        if (i < d) { s = block_sum<double>((double)v1[i], red); if (threadIdx.x == 0) atomicAdd(a.acc + CACC_VAR1 + i, s); }
        if (i < d2) { s = block_sum<double>((double)v2[i], red); if (threadIdx.x == 0) atomicAdd(a.acc + CACC_VAR2 + i, s); }
    }
}

__global__ void __launch_bounds__(128) celeba_latent_bwd_kernel(CLatArgs a) {
    __shared__ FlowTable ft;
    __shared__ float fred[32];
    load_flow_table(ft, a);
    __syncthreads();
    const int d = a.d, d2 = a.d2;
    FlowGrad fg;
    fg.clear();
    const float kscale = a.beta / (float)a.batch, ascale = a.lambda_ / (float)a.batch;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.batch; b += (int64_t)gridDim.x * blockDim.x) {
        const float* h = a.h + b * a.ldh;
        float m1[CDG_MAX_NODE], e1[CDG_MAX_NODE], u[CDG_MAX_NODE], ud[CDG_MAX_NODE], gz[CDG_MAX_NODE], ge2[CDG_MAX_NODE];
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            m1[i] = e1[i] = gz[i] = ge2[i] = 0.f;
            if (i < d) {
                m1[i] = h[i];
                e1[i] = a.deterministic ? m1[i] : m1[i] + expf(h[d + i] / 2.f) * a.n1[b * d + i];
            }
        }
#pragma unroll
        for (int k = 0; k < kNGen; ++k) {
            if (!a.gzin[k]) continue;
            for (int j = 0; j < a.zdim[k]; ++j) {
                const int src = a.zsrc[k][j];
                const float v = a.gzin[k][b * a.zdim[k] + j];
#pragma unroll
                for (int i = 0; i < CDG_MAX_NODE; ++i) {
                    if (src == i) gz[i] += v;
                    if (src == -1 - i) ge2[i] += v;
                }
            }
        }
        matvec_A(ft, d, e1, u);
        matvec_A(ft, d, m1, ud);
        float gu[CDG_MAX_NODE], gud[CDG_MAX_NODE], ge[CDG_MAX_NODE], gmd[CDG_MAX_NODE];
#pragma unroll
        for (int j = 0; j < CDG_MAX_NODE; ++j) {
            gu[j] = gud[j] = 0.f;
            if (j < d) {
                gu[j] = flow_bwd(ft, a.scm, a.flow_num, j, u[j], gz[j], fg);
                const float zd = flow_fwd(ft, a.scm, a.flow_num, j, ud[j]);
                const float yh = 1.f / (1.f + expf(-zd)), y = a.y[b * a.ld_y + j];
                const float gzd = ascale * (yh - y) / fmaxf((1.f - yh) * yh, 1e-12f) * ((1.f - yh) * yh);
                gud[j] = flow_bwd(ft, a.scm, a.flow_num, j, ud[j], gzd, fg);
            }
        }
        matvec_AT(ft, d, gu, ge);
        matvec_AT(ft, d, gud, gmd);
        float* gh = a.g_h + b * a.ldh;
#pragma unroll
        for (int i = 0; i < CDG_MAX_NODE; ++i) {
            if (i < d) {
                const float lv = h[d + i];
                const float gstoch = a.deterministic ? 0.f : 0.5f * ge[i] * a.n1[b * d + i] * expf(lv / 2.f);
                gh[i] = ge[i] + kscale * m1[i] + gmd[i];
                gh[d + i] = gstoch + 0.5f * kscale * (expf(lv) - 1.f);
            }
            if (i < d2) {
                const float m2 = h[2 * d + i], lv = h[2 * d + d2 + i];
                const float gstoch = a.deterministic ? 0.f : 0.5f * ge2[i] * a.n2[b * d + i] * expf(lv / 2.f);
                gh[2 * d + i] = ge2[i] + kscale * m2;
                gh[2 * d + d2 + i] = gstoch + 0.5f * kscale * (expf(lv) - 1.f);
            }
        }
    }
    reduce_flow_grads(fg, ft, a, fred);
}

// decode-only entry (model.py:188-195): decoder inputs gathered from caller-supplied latents / epsilon2
__global__ void celeba_gather_kernel(CLatArgs a, const float* __restrict__ z, const float* __restrict__ e2) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.batch; b += (int64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < kNGen; ++k)
            for (int j = 0; j < a.zdim[k]; ++j) {
                const int src = a.zsrc[k][j];
                a.zin[k][b * a.zdim[k] + j] = src >= 0 ? z[b * a.d + src] : e2[b * a.d2 + (-1 - src)];
            }
}

// ---- masked sum, tanh, L1 reconstruction and its gradient (model.py:197-199, train.py:31-33) -----------------------
struct ReconArgs {
    const float* pre[kNGen];   // [M, 3] toRGB pre-activations
    float* gpre[kNGen];        // [M, 3] d loss / d pre (null: forward only)
    const float* masks;        // [5][M]
    const float* x; int ld_x;  // [M, ld_x], channels 0..2 in [0,1]
    float* xhat;               // optional [M, 3]
    float* sep;                // optional [5][M, 3]
    int64_t M;
    float inv_batch;
    double* acc;
};
__global__ void __launch_bounds__(256) celeba_recon_kernel(ReconArgs a) {
    __shared__ double red[32];
    double loss = 0.0;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += (int64_t)gridDim.x * blockDim.x) {
        float t[kNGen][3], mk[kNGen], s[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kNGen; ++k) {
            mk[k] = a.masks[k * a.M + m];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                t[k][c] = tanhf(a.pre[k][m * 3 + c]);                    // Generator's own tanh (sagan.py:209)
                s[c] += t[k][c] * mk[k];
                if (a.sep) a.sep[(k * a.M + m) * 3 + c] = t[k][c];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float xh = tanhf(s[c]);
            const float diff = a.x ? xh - (a.x[m * a.ld_x + c] * 2.f - 1.f) : 0.f;   // train.py:31-32
            loss += (double)fabsf(diff);
            if (a.xhat) a.xhat[m * 3 + c] = xh;
            if (a.gpre[0]) {
                const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                const float gs = sg * (1.f - xh * xh) * a.inv_batch;
#pragma unroll
                for (int k = 0; k < kNGen; ++k) a.gpre[k][m * 3 + c] = gs * mk[k] * (1.f - t[k][c] * t[k][c]);
            }
        }
    }
    const double s = block_sum<double>(loss, red);
    if (threadIdx.x == 0 && a.acc) atomicAdd(a.acc + CACC_RECON, s);
}

__global__ void celeba_logs_kernel(const double* acc, float* logs, int d, int d2, int64_t batch, float beta, float lambda_) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double B = (double)batch;
    const float recon = (float)(acc[CACC_RECON] / B);
    const float kl = (float)(acc[CACC_KL1] / B) + (float)(acc[CACC_KL2] / B);
    const float align = (float)(acc[CACC_ALIGN] / B);
    float active = 0.f;                                              // train.py:60-63
    for (int i = 0; i < d; ++i) active += (float)(acc[CACC_VAR1 + i] / B) < 0.1f ? 1.f : 0.f;
    for (int i = 0; i < d2; ++i) active += (float)(acc[CACC_VAR2 + i] / B) < 0.1f ? 1.f : 0.f;
    active /= (float)(d + d2);
    logs[0] = recon + beta * kl + lambda_ * align;                   // train.py:65-66
    logs[1] = recon; logs[2] = kl; logs[3] = align; logs[4] = active;
}

}  // namespace

}  // namespace cdg

using namespace cdg;

struct cdg_celeba_plan {
    cdg_celeba_config c;
    std::map<int64_t, std::pair<int64_t, int64_t>> layout;     // batch -> (workspace bytes, im2col floats)
    // The five generators are independent between the latent block and the reconstruction head (and again through their
    // input-gradient chains), and at batch 16 most of their ~700 launches are far too small to fill 148 SMs: each generator
    // is enqueued on a stream of its own, forked from / joined to the caller's stream with events (legal under capture).
    cudaStream_t gs[kNGen] = {};
    cudaEvent_t ev_fork[3] = {}, ev_join[2][kNGen] = {};
    bool streams_ready = false;
    bool ensure_streams() {
        if (streams_ready) return true;
        for (int k = 0; k < kNGen; ++k) {
            if (cudaStreamCreateWithFlags(&gs[k], cudaStreamNonBlocking) != cudaSuccess) return false;
            for (int j = 0; j < 2; ++j)
                if (cudaEventCreateWithFlags(&ev_join[j][k], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        for (int j = 0; j < 3; ++j)
            if (cudaEventCreateWithFlags(&ev_fork[j], cudaEventDisableTiming) != cudaSuccess) return false;
        streams_ready = true;
        return true;
    }
    ~cdg_celeba_plan() {
        for (int k = 0; k < kNGen; ++k) {
            if (gs[k]) cudaStreamDestroy(gs[k]);
            for (int j = 0; j < 2; ++j) if (ev_join[j][k]) cudaEventDestroy(ev_join[j][k]);
        }
        for (int j = 0; j < 3; ++j) if (ev_fork[j]) cudaEventDestroy(ev_fork[j]);
    }
};

// Split-K occupancy target per GEMM while the five generator chains run side by side.  A lower target is faster (148: 12.5 ms,
// 74: 11.7, 24: 11.4 at batch 16) but means longer TMEM accumulations, whose truncating fp32 adds this step amplifies: the fc
// gradient's error against an fp64 evaluation grows 0.0018 -> 0.0026 -> 0.0042 (batch 2, step 1) and the golden test's second step
// crossed its 1e-2 bound in 3 runs of 6 at 24.  Parity first: the default keeps the single-chain split (148).
constexpr int kGenSmBudget = kNumSMs;
static int g_generator_streams = 7;        // bit 0: forward chains, bit 1: input-gradient chains, bit 2: weight preparation beside the encoder
extern "C" void cdg_celeba_generator_streams(int32_t on) { g_generator_streams = on; }

static int celeba_pass(cdg_celeba_plan* p, const cdg_celeba_io* io, Cx& cx) {
    const cdg_celeba_config& c = p->c;
    const int64_t B = io->batch;
    const int S = c.image_size, d = c.node, d2 = c.latent_dim, ldh = 2 * d + 2 * d2;
    const int64_t M = B * S * S;
    cx.col = cx.ws.take<float>(cx.col_cap);
    cx.acc = cx.ws.take<double>(kAccDoubles);
    double* lacc = cx.ws.take<double>(CACC_LEN);
    // one im2col / activation scratch per generator (they run side by side); the layout is the same with streams off
    float* gen_col[kNGen];
    for (int k = 0; k < kNGen; ++k) gen_col[k] = cx.ws.take<float>(cx.col_cap);
    float* const main_col = cx.col;
    const cudaStream_t main_s = cx.s;
    const bool conc = !cx.dry && g_generator_streams && !io->encode_only && p->ensure_streams();
    if (!cx.dry) {
        CDG_CHECK_CUDA(cudaMemsetAsync(cx.acc, 0, sizeof(double) * kAccDoubles, cx.s));
        CDG_CHECK_CUDA(cudaMemsetAsync(lacc, 0, sizeof(double) * CACC_LEN, cx.s));
        if (io->backward) CDG_CHECK_CUDA(cudaMemsetAsync(io->grads, 0, sizeof(float) * c.n_params, cx.s));
    }
    const int budget = (g_generator_streams >> 8) > 0 ? (g_generator_streams >> 8) : kGenSmBudget;
    struct BudgetGuard { ~BudgetGuard() { gemm_tc_set_sm_budget(0); } } budget_guard;
    const bool conc_prep = conc && (g_generator_streams & 4), conc_fwd = conc && (g_generator_streams & 1);
    const bool conc_bwd = conc && (g_generator_streams & 2);
    if (conc_prep) {
        CDG_CHECK_CUDA(cudaEventRecord(p->ev_fork[0], main_s));
        for (int k = 0; k < kNGen; ++k) CDG_CHECK_CUDA(cudaStreamWaitEvent(p->gs[k], p->ev_fork[0], 0));
    }
    // encoder (model.py:157-158): one evaluation; the deterministic pass sees the same input (model.py:212)
    const bool decode_only = io->latent_in != nullptr;
    float* feat = nullptr;
    float* h = nullptr;
    if (!decode_only) {
        CDG_TRY(resnet_forward(cx, c, io->x, io->ld_x, B, io->encoder_passes, &feat));
        h = cx.ws.take<float>(B * ldh);
        RUN(gemm_nt(cx, feat, 512, io->params + c.fc.w, 512, h, ldh, B, ldh, 512, io->params + c.fc.b));
    }
    CLatArgs la{};
    la.d = d; la.d2 = d2; la.scm = c.scm; la.flow_num = c.flow_num; la.deterministic = io->deterministic;
    la.batch = B; la.params = io->params; la.grads = io->grads;
    memcpy(la.flow_off, c.flow_off, sizeof(la.flow_off));
    memcpy(la.A, c.I_B_inv, sizeof(la.A));
    la.beta = c.beta; la.lambda_ = c.lambda_;
    la.h = h; la.ldh = ldh; la.n1 = io->noise1; la.n2 = io->noise2; la.y = io->y; la.ld_y = io->ld_y;
    la.lat = io->latents; la.acc = lacc;
    float* gzin[kNGen];
    for (int k = 0; k < kNGen; ++k) {
        la.zdim[k] = c.gen[k].z_dim;
        for (int j = 0; j < CDG_MAX_NODE; ++j) la.zsrc[k][j] = c.gen[k].z_src[j];
        la.zin[k] = cx.ws.take<float>(B * la.zdim[k]);
        gzin[k] = cx.ws.take<float>(B * la.zdim[k]);
    }
    const int lgrid = (int)imin64((B + 127) / 128, kNumSMs * 4);
    if (!cx.dry) {
        if (decode_only) celeba_gather_kernel<<<lgrid, 128, 0, cx.s>>>(la, io->latent_in, io->epsilon2_in);
        else celeba_latent_fwd_kernel<<<lgrid, 128, 0, cx.s>>>(la);
        CDG_CHECK_LAUNCH();
    }
    if (io->encode_only) return CDG_OK;
    // decoders (model.py:188-200); the weight preparation (spectral norm, layouts) of a generator does not depend on the
    // latent block, so on its own stream it overlaps the encoder
    static thread_local GenRun runs[kNGen];
    ReconArgs ra{};
    if (conc_fwd) CDG_CHECK_CUDA(cudaEventRecord(p->ev_fork[1], main_s));
    if (conc_fwd) gemm_tc_set_sm_budget(budget);
    for (int k = 0; k < kNGen; ++k) {
        runs[k] = GenRun();
        if (conc_fwd) cx.s = p->gs[k];
        cx.col = gen_col[k];
        if (conc_fwd && !conc_prep) CDG_CHECK_CUDA(cudaStreamWaitEvent(p->gs[k], p->ev_fork[1], 0));
        CDG_TRY(generator_prepare(cx, c.gen[k], &runs[k]));
        if (conc_fwd) CDG_CHECK_CUDA(cudaStreamWaitEvent(p->gs[k], p->ev_fork[1], 0));
        CDG_TRY(generator_forward(cx, c.gen[k], &runs[k], la.zin[k], B));
        if (conc_fwd) CDG_CHECK_CUDA(cudaEventRecord(p->ev_join[0][k], p->gs[k]));
        ra.pre[k] = runs[k].rgb_pre;
        ra.gpre[k] = io->backward ? cx.ws.take<float>(M * 3) : nullptr;
    }
    cx.s = main_s; cx.col = main_col;
    gemm_tc_set_sm_budget(0);
    if (conc_fwd)
        for (int k = 0; k < kNGen; ++k) CDG_CHECK_CUDA(cudaStreamWaitEvent(main_s, p->ev_join[0][k], 0));
    ra.masks = io->masks; ra.x = io->x; ra.ld_x = io->ld_x; ra.xhat = io->xhat; ra.sep = io->xhat_separated;
    ra.M = M; ra.inv_batch = 1.f / (float)B; ra.acc = lacc;
    if (!cx.dry) {
        celeba_recon_kernel<<<(int)imin64((M + 255) / 256, kNumSMs * 8), 256, 0, cx.s>>>(ra);
        CDG_CHECK_LAUNCH();
        if (io->logs) {
            celeba_logs_kernel<<<1, 32, 0, cx.s>>>(lacc, io->logs, d, d2, B, c.beta, c.lambda_);
            CDG_CHECK_LAUNCH();
        }
    }
    if (!io->backward) return CDG_OK;
    if (conc_bwd) CDG_CHECK_CUDA(cudaEventRecord(p->ev_fork[2], main_s));
    if (conc_bwd) gemm_tc_set_sm_budget(budget);
    for (int k = 0; k < kNGen; ++k) {
        if (conc_bwd) {
            cx.s = p->gs[k];
            CDG_CHECK_CUDA(cudaStreamWaitEvent(p->gs[k], p->ev_fork[2], 0));
        }
        cx.col = gen_col[k];
        CDG_TRY(generator_backward(cx, c.gen[k], runs[k], ra.gpre[k], B, S, gzin[k]));
        if (conc_bwd) CDG_CHECK_CUDA(cudaEventRecord(p->ev_join[1][k], p->gs[k]));
        la.gzin[k] = gzin[k];
    }
    cx.s = main_s; cx.col = main_col;
    gemm_tc_set_sm_budget(0);
    if (conc_bwd)
        for (int k = 0; k < kNGen; ++k) CDG_CHECK_CUDA(cudaStreamWaitEvent(main_s, p->ev_join[1][k], 0));
    la.g_h = cx.ws.take<float>(B * ldh);
    if (!cx.dry) {
        celeba_latent_bwd_kernel<<<lgrid, 128, 0, cx.s>>>(la);
        CDG_CHECK_LAUNCH();
        // encoder.fc gradients (the only trainable part of the encoder, model.py:121-125)
        GemmDesc g{};
        g.A = la.g_h; g.sa_m = 1; g.sa_k = ldh;
        g.B = feat; g.sb_n = 1; g.sb_k = 512;
        g.C = io->grads + c.fc.w; g.ldc = 512; g.M = ldh; g.N = 512; g.K = B;
        CDG_TRY(gemm_simt(g, cx.s));
        CDG_TRY(launch_colsum(la.g_h, ldh, B, ldh, io->grads + c.fc.b, cx.s));
    }
    return CDG_OK;
}

static int celeba_layout(cdg_celeba_plan* p, int64_t batch, int64_t* bytes, int64_t* col) {
    auto it = p->layout.find(batch);
    if (it == p->layout.end()) {
        cdg_celeba_io io{};
        io.batch = batch; io.backward = 1; io.encoder_passes = 2; io.ld_x = 8;
        Cx a;                                    // first dry pass: im2col scratch requirement
        a.mode = p->c.gemm_mode;
        CDG_TRY(celeba_pass(p, &io, a));
        Cx b;
        b.mode = p->c.gemm_mode;
        b.col_cap = a.col_need;                  // second: total with that scratch in place
        CDG_TRY(celeba_pass(p, &io, b));
        CDG_REQUIRE(b.acc_used <= kAccDoubles, "BatchNorm accumulator region too small");
        it = p->layout.emplace(batch, std::make_pair(b.ws.peak + 256, a.col_need)).first;
    }
    *bytes = it->second.first;
    *col = it->second.second;
    return CDG_OK;
}

extern "C" int cdg_celeba_create(const cdg_celeba_config* cfg, cdg_celeba_plan** out) {
    CDG_REQUIRE(cfg && out, "null argument");
    const cdg_celeba_config& c = *cfg;
    CDG_REQUIRE(c.node >= 1 && c.node <= CDG_MAX_NODE && c.latent_dim >= 1 && c.latent_dim <= CDG_MAX_NODE, "node / latent_dim out of range");
    CDG_REQUIRE(c.latent_dim == c.node, "both noise draws are shaped [B, node] (model.py:182-185): latent_dim must equal node");
    CDG_REQUIRE(c.scm == CDG_SCM_LINEAR || c.scm == CDG_SCM_PLANAR, "Not supported SCM!");
    CDG_REQUIRE(c.flow_num >= 1 && c.flow_num <= CDG_MAX_FLOW, "flow_num out of range");
    CDG_REQUIRE(c.image_size == 128, "the generators are the 128-pixel SAGAN variant (sagan.py:173-178)");
    CDG_REQUIRE(c.fc.in == 512 && c.fc.out == 2 * c.node + 2 * c.latent_dim, "encoder.fc shape mismatch");
    for (int k = 0; k < kNGen; ++k) {
        const cdg_generator& G = c.gen[k];
        CDG_REQUIRE(G.z_dim >= 1 && G.z_dim <= CDG_MAX_NODE && G.lin0.cin == G.z_dim, "generator %d: z_dim mismatch", k);
        CDG_REQUIRE(G.lin0.cout == 16 * G.blk[0].conv1.cin, "generator %d: first block shape mismatch", k);
        for (int j = 0; j < G.z_dim; ++j)
            CDG_REQUIRE(G.z_src[j] < c.node && G.z_src[j] >= -c.latent_dim, "generator %d: latent source out of range", k);
        for (int b = 0; b + 1 < kNBlk; ++b) CDG_REQUIRE(G.blk[b].conv1.cout == G.blk[b + 1].conv1.cin, "generator %d: channel chain", k);
        CDG_REQUIRE(G.to_rgb.cin == G.blk[kNBlk - 1].conv1.cout && G.to_rgb.cout == 3, "generator %d: toRGB shape", k);
    }
    cdg_celeba_plan* p = new (std::nothrow) cdg_celeba_plan;
    CDG_REQUIRE(p, "out of host memory");
    p->c = c;
    *out = p;
    return CDG_OK;
}

extern "C" void cdg_celeba_destroy(cdg_celeba_plan* p) { delete p; }

extern "C" int64_t cdg_celeba_workspace_bytes(cdg_celeba_plan* p, int64_t batch) {
    int64_t bytes = 0, col = 0;
    if (!p || batch < 1 || celeba_layout(p, batch, &bytes, &col) != CDG_OK) return -1;
    return bytes;
}

extern "C" int cdg_celeba_step(cdg_celeba_plan* p, const cdg_celeba_io* io, void* stream) {
    CDG_REQUIRE(p && io, "null argument");
    CDG_REQUIRE(io->batch >= 1 && io->params && io->frozen && io->workspace, "missing input");
    if (io->latent_in) {
        CDG_REQUIRE(io->epsilon2_in && !io->backward && !io->encode_only, "decode-only entry: latent_in and epsilon2_in, forward only");
    } else {
        CDG_REQUIRE(io->x && io->ld_x >= 3, "x with at least 3 channels required");
        CDG_REQUIRE(io->deterministic || (io->noise1 && io->noise2), "noise required");
    }
    CDG_REQUIRE(io->encode_only || io->masks, "masks required");
    CDG_REQUIRE(!io->backward || (io->grads && io->y), "backward needs grads and y");
    int64_t bytes = 0, col = 0;
    CDG_TRY(celeba_layout(p, io->batch, &bytes, &col));
    if (io->workspace_bytes < bytes) {
        set_error("workspace too small: %lld < %lld", (long long)io->workspace_bytes, (long long)bytes);
        return CDG_ERR_WORKSPACE;
    }
    Cx cx;
    cx.dry = false;
    cx.s = (cudaStream_t)stream;
    cx.mode = p->c.gemm_mode;
    cx.ws.base = (char*)io->workspace;
    cx.col_cap = col;
    cx.frozen = io->frozen;
    return celeba_pass(p, io, cx);
}

// ---- single-layer entry points for unit tests --------------------------------------------------------------------
extern "C" int64_t cdg_conv2d_workspace_bytes(int64_t batch, int32_t h, int32_t w, const cdg_conv* cv, int32_t up) {
    const Geom g = conv_geom(batch, h, w, up, cv->cin, cv->k, cv->stride, cv->pad);
    const Geom gd = conv_geom(batch, h, w, 1, cv->cout, cv->k, 1, (cv->k - 1) / 2);
    const int64_t col = imax64(g.M * g.Kp, gd.M * gd.Kp);
    return 4 * (col + (int64_t)cv->cout * g.Kp + (int64_t)cv->cin * gd.Kp) + 4096;
}

extern "C" int cdg_conv2d_forward(int mode, const float* x, int64_t batch, int32_t h, int32_t w, const float* weight,
                                  const float* bias, const cdg_conv* cv, int32_t up, float* out, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
    CDG_REQUIRE(workspace_bytes >= cdg_conv2d_workspace_bytes(batch, h, w, cv, up), "workspace too small");
    Cx cx;
    cx.dry = false; cx.s = (cudaStream_t)stream; cx.mode = mode; cx.ws.base = (char*)workspace;
    const Geom g = conv_geom(batch, h, w, up, cv->cin, cv->k, cv->stride, cv->pad);
    float* wf = cx.ws.take<float>((int64_t)cv->cout * g.Kp);
    cx.col_cap = g.M * g.Kp;
    cx.col = cx.ws.take<float>(cx.col_cap);
    CDG_TRY(launch_weight_prep(weight, cv->cout, cv->cin, cv->k, nullptr, wf, g.Kp, nullptr, 0, cx.s));
    return conv_fwd(cx, x, batch, h, w, cv->cin, cv->cin, nullptr, 0, up, cv->k, cv->stride, cv->pad, wf, cv->cout, bias, out);
}

extern "C" int cdg_conv2d_dgrad(int mode, const float* gout, int64_t batch, int32_t h, int32_t w, const float* weight,
                                const cdg_conv* cv, float* gin, void* workspace, int64_t workspace_bytes, void* stream) {
    CDG_REQUIRE(cv->stride == 1 && cv->pad == (cv->k - 1) / 2, "dgrad is provided for stride-1 'same' convolutions");
    CDG_REQUIRE(workspace_bytes >= cdg_conv2d_workspace_bytes(batch, h, w, cv, 1), "workspace too small");
    Cx cx;
    cx.dry = false; cx.s = (cudaStream_t)stream; cx.mode = mode; cx.ws.base = (char*)workspace;
    const int Kpf = round_up4(cv->k * cv->k * cv->cin), Kpd = round_up4(cv->k * cv->k * cv->cout);
    float* wf = cx.ws.take<float>((int64_t)cv->cout * Kpf);
    float* wd = cx.ws.take<float>((int64_t)cv->cin * Kpd);
    cx.col_cap = batch * h * w * (int64_t)Kpd;
    cx.col = cx.ws.take<float>(cx.col_cap);
    CDG_TRY(launch_weight_prep(weight, cv->cout, cv->cin, cv->k, nullptr, wf, Kpf, wd, Kpd, cx.s));
    return conv_dgrad(cx, gout, batch, h, w, cv->cout, cv->k, wd, cv->cin, gin);
}
