// CDG-TVAE step (tabular/modules/model.py:360-460, tabular/modules/train.py:245-320) with every Linear layer on the tensor pipe.
//
// Same organisation as tvae_tile.cu -- a warp owns a tile of 32 table rows whose activations live in shared memory as
// [feature][32 rows], gradient slabs overwrite the activation slabs they belong to, the element-wise stages run with lane =
// row -- but the three products of a layer (forward, input gradient, weight gradient) are issued as warp-level
// mma.sync.m16n8k8 TF32 instructions with 3xTF32 operands (x = hi + lo, hi = tf32(x); lo.hi + hi.lo + hi.hi, fp32 accumulators:
// ~1e-6 per product, the same recipe as the pendulum step's 3xTF32 mode).  ncu on the SIMT tile kernel showed why: a 16-byte
// shared-memory load costs four wavefronts however much of it is a broadcast, so its 4 x 4 register tiles moved 8 wavefronts per
// 16 FFMA and the kernel sat on the shared-memory pipe (55 % busy, issue slots 38 %).  An MMA fragment is fetched with 4-byte
// loads that are conflict-free by construction: feature row k of a slab is stored with its 32 row slots XOR-ed by 8 (k mod 4),
// so the 8 x 4 (row, feature) pattern of an A fragment touches 32 different banks; the B fragments of the forward and input-
// gradient products come from fragment-ordered copies of the weights built once per block (one 8-byte load per lane and MMA
// triple).  tcgen05 is not the tool here: its smallest M is 64 rows per instruction with operands staged through shared-memory
// descriptors, for layers that are 8 to 32 wide; tools/mma_sync_rate.cu measures the warp-level form at 2.07 cycles per
// m16n8k8 and SM on B200 (278 TFLOP/s TF32), 1.3 x the FFMA rate after the 3 x and without the per-FFMA operand loads.
#include "latent.cuh"
#include "tabular_args.cuh"

namespace cdg {

namespace {

constexpr int TM_H0 = 32, TM_H1 = 16, TM_H2 = 16, TM_D1 = 8, TM_D2 = 8, TM_D3 = 16;
constexpr int TM_MAXM = 32;           // widest decoder output handled (reference shapes: <= 13)
constexpr int TM_MAXD = 64;

__host__ __device__ constexpr int pad4(int v) { return (v + 3) & ~3; }
__host__ __device__ constexpr int pad8(int v) { return (v + 7) & ~7; }
__host__ __device__ constexpr int cdiv8(int v) { return (v + 7) >> 3; }

// element (feature k, row r) of a slab
__device__ __forceinline__ int SW(int k, int r) { return k * 32 + (r ^ ((k & 3) << 3)); }

struct MmaLayout {                     // feature-row offsets inside a warp's slab (rows of 32 floats); every region starts at a
    int x, h0, h1, h2, ml, lat, a1, a2, a3, xh, rows;      // multiple of 4 rows (the swizzle uses the region-local feature index)
};
__host__ __device__ inline MmaLayout mma_layout(int D, int d, int max_m) {
    MmaLayout t;
    int r = 0;
    t.x = r; r += pad8(D);
    t.h0 = r; r += TM_H0;
    t.h1 = r; r += TM_H1;
    t.h2 = r; r += TM_H2;
    t.ml = r; r += pad8(2 * d);
    t.lat = r; r += pad8(5 * d);       // nz[d] u[d] gal[d] z[d] gz[d]
    t.a1 = r; r += TM_D1;
    t.a2 = r; r += TM_D2;
    t.a3 = r; r += TM_D3;
    t.xh = r; r += pad8(max_m);
    t.rows = r;
    return t;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    lo = __float_as_uint(x - __uint_as_float(hi));       // the MMA reads the top 19 bits of it: 2^-22 of x is dropped
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (ahi + alo) (bhi + blo) without the lo.lo term, small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
    mma_tf32(c, alo, bh0, bh1);
    mma_tf32(c, ahi, bl0, bl1);
    mma_tf32(c, ahi, bh0, bh1);
}

// Fragment layouts of mma.m16n8k8 (g = lane / 4, t = lane % 4):
//   A (16 x 8):  a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)
//   B ( 8 x 8):  b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   C (16 x 8):  c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)

// out[o][r] = sum_{k < K} in[koff + k][r] * B[k][o]   for the 32 rows of the tile and o < OUT,
//   B given as the fragment-ordered copy `frag` (cdiv8(K) x cdiv8(OUT) blocks of 64 floats: lane l holds b0, b1 at 2 l);
//   MODE 0: + bias[o], optional ReLU (forward);   MODE 1: times (old out[o][r] > 0) (input gradient through a ReLU)
template <int MAXNT, int MODE, bool RELU>
__device__ __forceinline__ void mma_rows(const float* __restrict__ in, int koff, int K, const float* __restrict__ frag, int OUT,
                                         const float* __restrict__ bias, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NT = cdiv8(OUT), KS = cdiv8(K);
    float acc[MAXNT][2][4];
#pragma unroll
    for (int nt = 0; nt < MAXNT; ++nt) {
        float b0 = 0.f, b1 = 0.f;
        if (MODE == 0 && nt < NT) {
            const int o = nt * 8 + 2 * t;
            b0 = o < OUT ? bias[o] : 0.f;
            b1 = o + 1 < OUT ? bias[o + 1] : 0.f;
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { acc[nt][mt][0] = b0; acc[nt][mt][1] = b1; acc[nt][mt][2] = b0; acc[nt][mt][3] = b1; }
    }
    for (int ks = 0; ks < KS; ++ks) {
        const int k0 = ks * 8 + t, k1 = k0 + 4;
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int r = mt * 16 + g;
            const float a0 = k0 < K ? in[SW(koff + k0, r)] : 0.f, a1 = k0 < K ? in[SW(koff + k0, r + 8)] : 0.f;
            const float a2 = k1 < K ? in[SW(koff + k1, r)] : 0.f, a3 = k1 < K ? in[SW(koff + k1, r + 8)] : 0.f;
            split_tf32(a0, ahi[mt][0], alo[mt][0]); split_tf32(a1, ahi[mt][1], alo[mt][1]);
            split_tf32(a2, ahi[mt][2], alo[mt][2]); split_tf32(a3, ahi[mt][3], alo[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt) {
            if (nt < NT) {
                const float2 b = *reinterpret_cast<const float2*>(frag + ((ks * NT + nt) * 32 + lane) * 2);
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(b.x, bh0, bl0); split_tf32(b.y, bh1, bl1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) mma3(acc[nt][mt], ahi[mt], alo[mt], bh0, bh1, bl0, bl1);
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < MAXNT; ++nt) {
        if (nt < NT) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int o = nt * 8 + 2 * t + (e & 1), r = mt * 16 + g + ((e >> 1) << 3);
                    if (o < OUT) {
                        float v = acc[nt][mt][e];
                        float* op = out + SW(o, r);
                        if (MODE == 0) { if (RELU) v = fmaxf(v, 0.f); }
                        else v = *op > 0.f ? v : 0.f;
                        *op = v;
                    }
                }
            }
        }
    }
}

// sgw[o * ldg + i] += sum_r delta[o][r] * hin[ioff + i][r]   (o < OUT <= 32, i < IN <= 8 MAXNT);   sgb[o] += sum_r delta[o][r]
// (the bias gradient is one more column of the same product: B = a column of ones, exact in TF32)
template <int MAXNT>
__device__ __forceinline__ void mma_wgrad(const float* __restrict__ delta, int OUT, const float* __restrict__ hin, int ioff, int IN,
                                          float* sgw, int ldg, float* sgb) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int MT = (OUT + 15) >> 4, NT = cdiv8(IN);
    float acc[2][MAXNT][4], accb[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) accb[mt][e] = 0.f;
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    }
    const uint32_t one = g == 0 ? __float_as_uint(1.f) : 0u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int r0 = ks * 8 + t, r1 = r0 + 4;
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (mt < MT) {
                const int o = mt * 16 + g;
                const float a0 = o < OUT ? delta[SW(o, r0)] : 0.f, a1 = o + 8 < OUT ? delta[SW(o + 8, r0)] : 0.f;
                const float a2 = o < OUT ? delta[SW(o, r1)] : 0.f, a3 = o + 8 < OUT ? delta[SW(o + 8, r1)] : 0.f;
                split_tf32(a0, ahi[mt][0], alo[mt][0]); split_tf32(a1, ahi[mt][1], alo[mt][1]);
                split_tf32(a2, ahi[mt][2], alo[mt][2]); split_tf32(a3, ahi[mt][3], alo[mt][3]);
                mma_tf32(accb[mt], alo[mt], one, one);
                mma_tf32(accb[mt], ahi[mt], one, one);
            }
        }
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt) {
            if (nt < NT) {
                const int i = nt * 8 + g;
                const float b0 = i < IN ? hin[SW(ioff + i, r0)] : 0.f, b1 = i < IN ? hin[SW(ioff + i, r1)] : 0.f;
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(b0, bh0, bl0); split_tf32(b1, bh1, bl1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    if (mt < MT) mma3(acc[mt][nt], ahi[mt], alo[mt], bh0, bh1, bl0, bl1);
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        if (mt < MT) {
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt) {
                if (nt < NT) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int o = mt * 16 + g + ((e >> 1) << 3), i = nt * 8 + 2 * t + (e & 1);
                        if (o < OUT && i < IN) atomicAdd(sgw + o * ldg + i, acc[mt][nt][e]);
                    }
                }
            }
            if (t == 0) {
                const int o = mt * 16 + g;
                if (o < OUT) atomicAdd(sgb + o, accb[mt][0]);
                if (o + 8 < OUT) atomicAdd(sgb + o + 8, accb[mt][2]);
            }
        }
    }
}


template <int DN>
__global__ void __launch_bounds__(256, 1) tvae_mma_kernel(TabArgs a, int frag_total) {
    extern __shared__ __align__(16) float smem[];
    const cdg_tabular_config& c = a.c;
    constexpr int d = DN;
    const int np = (int)c.n_params, D = c.input_dim;
    const int np4 = pad4(np);
    float* sp = smem;                                  // parameters, arena layout
    float* sg = sp + np4;                              // the block's gradient copy
    float* fr = sg + np4;                              // fragment-ordered weights: forward copies, then input-gradient copies
    int max_m = 0;
    for (int k = 0; k < d; ++k) max_m = max(max_m, c.dec[k][3].out);
    const MmaLayout T = mma_layout(D, d, max_m);
    float* slab = fr + frag_total + (threadIdx.x >> 5) * T.rows * 32;
    __shared__ FlowTable ft;
    __shared__ double dred[32];
    __shared__ float fred[32];
    __shared__ int f_off[4 + 4 * CDG_MAX_DEC], g_off[4 + 4 * CDG_MAX_DEC], span_lo[CDG_MAX_DEC + 1];

    for (int i = threadIdx.x; i < np; i += blockDim.x) { sp[i] = a.params[i]; sg[i] = 0.f; }
    {
        struct { int d, scm, flow_num; const float* params; const int64_t* flow_off; const float* A; } fa =
            {DN, c.scm, c.flow_num, a.params, c.flow_off, c.I_B_inv};
        load_flow_table(ft, fa);
    }
    if (threadIdx.x == 0) {
        int o = 0;
        for (int l = 0; l < 4 + 4 * d; ++l) {
            const cdg_linear& L = l < 4 ? c.enc[l] : c.dec[(l - 4) >> 2][(l - 4) & 3];
            f_off[l] = o; o += cdiv8(L.in) * cdiv8(L.out) * 64;
        }
        for (int l = 0; l < 4 + 4 * d; ++l) {
            const cdg_linear& L = l < 4 ? c.enc[l] : c.dec[(l - 4) >> 2][(l - 4) & 3];
            g_off[l] = o; o += cdiv8(L.out) * cdiv8(L.in) * 64;
        }
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
        {   // forward: B[k = in][n = out] = W[out][in]
            const int NT = cdiv8(L.out), n = cdiv8(L.in) * NT * 64;
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                const int blk = e >> 6, ln = (e >> 1) & 31, hi = e & 1, ks = blk / NT, nt = blk - ks * NT;
                const int k = ks * 8 + (ln & 3) + 4 * hi, o = nt * 8 + (ln >> 2);
                fr[f_off[l] + e] = (k < L.in && o < L.out) ? sp[L.w + o * L.in + k] : 0.f;
            }
        }
        {   // input gradient: B[k = out][n = in] = W[out][in]
            const int NT = cdiv8(L.in), n = cdiv8(L.out) * NT * 64;
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                const int blk = e >> 6, ln = (e >> 1) & 31, hi = e & 1, ks = blk / NT, nt = blk - ks * NT;
                const int k = ks * 8 + (ln & 3) + 4 * hi, i = nt * 8 + (ln >> 2);
                fr[g_off[l] + e] = (k < L.out && i < L.in) ? sp[L.w + k * L.in + i] : 0.f;
            }
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
    constexpr int NZ = 0, U = d, GAL = 2 * d, Z = 3 * d, GZ = 4 * d;       // feature rows of LAT

    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t ntiles = (a.batch + 31) / 32;
    for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += warps_total) {
        const int64_t b = tile * 32 + lane;
        const bool valid = b < a.batch;
        const float vm = valid ? 1.f : 0.f;
        const int64_t br = valid ? b : 0;
        const float* xrow = a.x + br * D;
        for (int i = 0; i < D; ++i) X[SW(i, lane)] = __ldg(xrow + i);
        __syncwarp();

        // ---- encoder D-32-16-16-2d (ReLU) ----
        mma_rows<4, 0, true>(X, 0, D, fr + f_off[0], TM_H0, sp + c.enc[0].b, H0);
        __syncwarp();
        mma_rows<2, 0, true>(H0, 0, TM_H0, fr + f_off[1], TM_H1, sp + c.enc[1].b, H1);
        __syncwarp();
        mma_rows<2, 0, true>(H1, 0, TM_H1, fr + f_off[2], TM_H2, sp + c.enc[2].b, H2);
        __syncwarp();
        mma_rows<2, 0, false>(H2, 0, TM_H2, fr + f_off[3], 2 * d, sp + c.enc[3].b, ML);
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
                    mean[i] = ML[SW(i < d ? i : 0, lane)]; lv[i] = ML[SW(i < d ? d + i : 0, lane)];
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
                LAT[SW(NZ + i, lane)] = nz[i]; LAT[SW(U + i, lane)] = u[i]; LAT[SW(GAL + i, lane)] = gal[i];
                LAT[SW(Z + i, lane)] = z[i];
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
            const int m = L3.out, lb = 4 + 4 * k;
            mma_rows<1, 0, true>(LAT, Z + k, 1, fr + f_off[lb], TM_D1, sp + L0.b, A1);
            __syncwarp();
            mma_rows<1, 0, true>(A1, 0, TM_D1, fr + f_off[lb + 1], TM_D2, sp + L1.b, A2);
            __syncwarp();
            mma_rows<2, 0, true>(A2, 0, TM_D2, fr + f_off[lb + 2], TM_D3, sp + L2.b, A3);
            __syncwarp();
            mma_rows<4, 0, false>(A3, 0, TM_D3, fr + f_off[lb + 3], m, sp + L3.b, XH);
            __syncwarp();
            if (a.xhat && valid)
                for (int j = 0; j < m; ++j) a.xhat[b * a.out_total + col + j] = XH[SW(j, lane)];

            // span losses of this decoder's columns (tabular/modules/train.py:270-285); d loss / d xhat replaces xhat
            for (int sidx = span_lo[k]; sidx < span_lo[k + 1]; ++sidx) {
                const int st = c.span_start[sidx], dim = c.span_dim[sidx];
                // a decoder's width is a sum of whole spans (main_tvae.py:174-192); the forward-only API's placeholder span
                // (one softmax over all columns, its loss is never read) does not fit a decoder and is skipped
                if (st - col + dim > m) continue;
                const int j0 = st - col;
                if (c.span_kind[sidx] == CDG_SPAN_TANH) {
                    const float sd = sp[c.sigma_off + st];
                    const float th = tanhf(XH[SW(j0, lane)]);
                    const float r = X[SW(st, lane)] - th;
                    rec += r * r / 2.f / (sd * sd) + logf(sd);
                    XH[SW(j0, lane)] = -(r / (sd * sd)) * (1.f - th * th) * invB * vm;
                    if (a.do_bwd) {
                        const float t = warp_sum(vm * (-(r * r) / (sd * sd * sd) + 1.f / sd) * invB);
                        if (lane == 0) atomicAdd(sg + c.sigma_off + st, t);
                    }
                } else {
                    int tgt = 0;
                    float best = X[SW(st, lane)], mx = XH[SW(j0, lane)];
                    for (int j = 1; j < dim; ++j) {
                        const float xv = X[SW(st + j, lane)];
                        if (xv > best) { best = xv; tgt = j; }
                        mx = fmaxf(mx, XH[SW(j0 + j, lane)]);
                    }
                    float se = 0.f;
                    for (int j = 0; j < dim; ++j) se += expf(XH[SW(j0 + j, lane)] - mx);
                    const float lse = mx + logf(se);
                    rec += lse - XH[SW(j0 + tgt, lane)];
                    for (int j = 0; j < dim; ++j)
                        XH[SW(j0 + j, lane)] = (expf(XH[SW(j0 + j, lane)] - lse) - (j == tgt ? 1.f : 0.f)) * invB * vm;
                }
            }
            __syncwarp();
            if (a.do_bwd) {
                mma_wgrad<2>(XH, m, A3, 0, TM_D3, sg + L3.w, TM_D3, sg + L3.b);
                mma_rows<2, 1, false>(XH, 0, m, fr + g_off[lb + 3], TM_D3, nullptr, A3);
                __syncwarp();
                mma_wgrad<1>(A3, TM_D3, A2, 0, TM_D2, sg + L2.w, TM_D2, sg + L2.b);
                mma_rows<1, 1, false>(A3, 0, TM_D3, fr + g_off[lb + 2], TM_D2, nullptr, A2);
                __syncwarp();
                mma_wgrad<1>(A2, TM_D2, A1, 0, TM_D1, sg + L1.w, TM_D1, sg + L1.b);
                mma_rows<1, 1, false>(A2, 0, TM_D2, fr + g_off[lb + 1], TM_D1, nullptr, A1);
                __syncwarp();
                mma_wgrad<1>(A1, TM_D1, LAT, Z + k, 1, sg + L0.w, 1, sg + L0.b);
                float gz = 0.f;
#pragma unroll
                for (int o = 0; o < TM_D1; ++o) gz = fmaf(A1[SW(o, lane)], sp[L0.w + o], gz);
                LAT[SW(GZ + k, lane)] = gz;
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
                gu[j] = j < d ? flow_bwd(ft, c.scm, c.flow_num, j, LAT[SW(U + (j < d ? j : 0), lane)], LAT[SW(GZ + (j < d ? j : 0), lane)], fg) : 0.f;
            matvec_AT(ft, d, gu, ge);
            const float kscale = c.beta * invB;
#pragma unroll
            for (int i = 0; i < d; ++i) {
                const float mean = ML[SW(i, lane)], lv = ML[SW(d + i, lane)];
                ML[SW(i, lane)] = ge[i] + vm * kscale * mean + LAT[SW(GAL + i, lane)];
                ML[SW(d + i, lane)] = 0.5f * ge[i] * LAT[SW(NZ + i, lane)] * expf(lv / 2.f) + vm * 0.5f * kscale * (expf(lv) - 1.f);
            }
        }
        __syncwarp();

        // ---- encoder backward ----
        mma_wgrad<2>(ML, 2 * d, H2, 0, TM_H2, sg + c.enc[3].w, TM_H2, sg + c.enc[3].b);
        mma_rows<2, 1, false>(ML, 0, 2 * d, fr + g_off[3], TM_H2, nullptr, H2);
        __syncwarp();
        mma_wgrad<2>(H2, TM_H2, H1, 0, TM_H1, sg + c.enc[2].w, TM_H1, sg + c.enc[2].b);
        mma_rows<2, 1, false>(H2, 0, TM_H2, fr + g_off[2], TM_H1, nullptr, H1);
        __syncwarp();
        mma_wgrad<4>(H1, TM_H1, H0, 0, TM_H0, sg + c.enc[1].w, TM_H0, sg + c.enc[1].b);
        mma_rows<4, 1, false>(H1, 0, TM_H1, fr + g_off[1], TM_H0, nullptr, H0);
        __syncwarp();
        mma_wgrad<8>(H0, TM_H0, X, 0, D, sg + c.enc[0].w, D, sg + c.enc[0].b);
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

bool mma_shape_ok(const cdg_tabular_config& c) {
    if (c.kind != CDG_TAB_TVAE || c.act != CDG_ACT_RELU || c.n_enc_layers != 4 || c.n_dec_layers != 4) return false;
    if (c.n_dec != c.node || (c.node != 3 && c.node != 6) || c.input_dim > TM_MAXD) return false;
    const int e[5] = {c.input_dim, TM_H0, TM_H1, TM_H2, 2 * c.node};
    for (int l = 0; l < 4; ++l)
        if (c.enc[l].in != e[l] || c.enc[l].out != e[l + 1]) return false;
    for (int k = 0; k < c.n_dec; ++k) {
        if (c.factor[k] != 1 || c.out_dim[k] > TM_MAXM) return false;
        const int dd[5] = {1, TM_D1, TM_D2, TM_D3, c.out_dim[k]};
        for (int l = 0; l < 4; ++l)
            if (c.dec[k][l].in != dd[l] || c.dec[k][l].out != dd[l + 1]) return false;
    }
    return true;
}

}  // namespace

bool launch_tvae_mma(const TabArgs& a, cudaStream_t s) {
    const cdg_tabular_config& c = a.c;
    if (!mma_shape_ok(c)) return false;
    int frag_total = 0, max_m = 0;
    for (int l = 0; l < 4; ++l) frag_total += 2 * cdiv8(c.enc[l].in) * cdiv8(c.enc[l].out) * 64;
    for (int k = 0; k < c.n_dec; ++k) {
        for (int l = 0; l < 4; ++l) frag_total += 2 * cdiv8(c.dec[k][l].in) * cdiv8(c.dec[k][l].out) * 64;
        max_m = c.out_dim[k] > max_m ? c.out_dim[k] : max_m;
    }
    const MmaLayout T = mma_layout(c.input_dim, c.node, max_m);
    const size_t fixed = sizeof(float) * (2 * (size_t)pad4((int)c.n_params) + frag_total);
    const size_t slab = sizeof(float) * 32 * (size_t)T.rows;
    const size_t budget = 220 * 1024;
    if (fixed + slab > budget) return false;
    int warps = (int)((budget - fixed) / slab);
    if (warps > 8) warps = 8;
    const size_t smem = fixed + slab * warps;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tvae_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess ||
            cudaFuncSetAttribute(tvae_mma_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        attr = true;
    }
    const int64_t ntiles = (a.batch + 31) / 32;
    int64_t blocks = (ntiles + warps - 1) / warps;
    if (blocks > kNumSMs) blocks = kNumSMs;
    if (c.node == 3) tvae_mma_kernel<3><<<(unsigned)blocks, 32 * warps, smem, s>>>(a, frag_total);
    else tvae_mma_kernel<6><<<(unsigned)blocks, 32 * warps, smem, s>>>(a, frag_total);
    return true;
}

}  // namespace cdg
