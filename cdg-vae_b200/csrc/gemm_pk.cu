// Long-contraction companion of gemm_ps.cu: C[M, N <= 304] += sum_k A(m,k) B(n,k) with both operands as bf16 (hi, lo)
// planes, the contraction cut into segments of <= 2048 (one TMEM accumulation each: the tensor core accumulates in fp32
// with truncation, ~3.5e-9 * K relative) whose partial tiles are added to C with red.global.add.f32 (round to nearest).
//
// Two operand layouts:
//   K-major   A(m,k) = a[m * lda + k], B(n,k) = b[n * ldb + k]          input gradients: dX = dY W with dY's planes
//                                                                        ([batch][n_out]) and the transposed weight planes
//   MN-major  A(m,k) = a[k * lda + m], B(n,k) = b[k * ldb + n]          weight gradients: dW = dY^T X, contraction over the
//                                                                        batch, BOTH operands read as they lie in memory
// The MN-major form is what removes the per-call "split + transpose" passes of gemm_tc.cu's weight-gradient route: the UMMA
// shared-memory descriptor takes tiles whose contiguous direction is M / N (64 elements = one 128-byte swizzle row per
// K index, 8 K-rows per 1024-byte atom; leading-dimension offset = distance between 64-wide blocks), which is exactly the
// image a TMA box {64 mn, BK k} of a row-major [K][mn] plane leaves in shared memory.
//
// One 256 x 304 tile per cluster and work item (CTA pair, cta_group::2: each CTA 128 rows of A, half of the B columns of
// each of the two MMAs N = 256 and N = 48), a single 304-column accumulator (a work item's main loop is >= 32 K-blocks long,
// so the un-overlapped atomic epilogue costs a few percent), eight epilogue warps.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace cdg {
namespace pk {

using namespace tc;

constexpr int BM = 128, NE = 8, THREADS = 32 * (2 + NE);
constexpr int BN = 304, N0 = 256, N1 = 48;          // two MMAs per K step
constexpr int H0 = N0 / 2, H1 = N1 / 2;             // B columns (rows of the K-major tile) per CTA and MMA

template <bool MN>
struct Cfg {
    static constexpr int BK = MN ? 32 : 64;                                      // K per stage
    static constexpr int BLK = BK * 128;                                         // MN-major: one 64-wide block, BK K-rows
    static constexpr int A_BYTES = MN ? 2 * BLK : BM * 128;                      // one plane of this CTA's 128 rows
    static constexpr int B0_BYTES = MN ? 2 * BLK : H0 * 128;                     // MMA 0: 128 columns of B
    static constexpr int B1_BYTES = MN ? BLK : 32 * 128;                         // MMA 1: 24 columns (a 32-row / 1-block slot)
    static constexpr int B1_TX = MN ? BLK : H1 * 128;                            // bytes TMA actually delivers for it
    static constexpr int STAGE_BYTES = 2 * (A_BYTES + B0_BYTES + B1_BYTES);
    static constexpr int STAGE_TX = 2 * (A_BYTES + B0_BYTES + B1_TX);
    static constexpr int STAGES = MN ? 5 : 3;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 256;
    static_assert(SMEM <= 232448, "tile does not fit shared memory");
};

struct Params {
    float* C; int64_t ldc;
    int64_t M; int N;                    // N <= 304 result columns; column n >= nc goes to extra[m] (n == nc) or nowhere
    int nc;                              // columns of C (N or N - 1)
    float* extra;                        // bias gradient: += column nc of the product (weight gradients with a ones column)
    int kb_total, kb_per_split, splits;
    int64_t work_total;
};

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
// SWIZZLE_128B descriptors.  K-major: rows of 128 B, 8-row groups 1024 B apart (stride offset).  MN-major: per K index one
// 128-byte row of 64 M/N elements, 8 K-rows per 1024-byte atom (stride offset), 64-wide blocks `lbo` bytes apart.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool mn_major) {
    // c_format F32 @4, a_format / b_format BF16 @7 / @10, a_major @15, b_major @16 (1 = M/N contiguous), N>>3 @17, M>>4 @24
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (1u << 15) | (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool MN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_pk_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const __grid_constant__ CUtensorMap tmBh1, const __grid_constant__ CUtensorMap tmBl1, const Params p) {
    using C_ = Cfg<MN>;
    constexpr int STAGES = C_::STAGES, BK = C_::BK;
    const uint32_t rank = cluster_ctarank();
    const int64_t w_first = (int64_t)(blockIdx.x >> 1), w_step = (int64_t)(gridDim.x >> 1);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C_::STAGE_BYTES);
    uint64_t* full = bars;                       // [STAGES] (leader's copy is used)
    uint64_t* empty = bars + STAGES;             // [STAGES] multicast commit
    uint64_t* acc_full = empty + STAGES;         // [1]
    uint64_t* acc_empty = acc_full + 1;          // [1] (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // stage: [A hi | A lo | B0 hi | B0 lo | B1 hi | B1 lo]
    auto a_pl = [&](int s, int pl) { return smem + (size_t)s * C_::STAGE_BYTES + pl * C_::A_BYTES; };
    auto b0_pl = [&](int s, int pl) { return smem + (size_t)s * C_::STAGE_BYTES + 2 * C_::A_BYTES + pl * C_::B0_BYTES; };
    auto b1_pl = [&](int s, int pl) {
        return smem + (size_t)s * C_::STAGE_BYTES + 2 * C_::A_BYTES + 2 * C_::B0_BYTES + pl * C_::B1_BYTES;
    };
    auto decode = [&](int64_t w, int& m_pair, int& kb_beg, int& nkb) {
        const int sp = (int)(w % p.splits);
        m_pair = (int)(w / p.splits);
        kb_beg = sp * p.kb_per_split;
        nkb = min(p.kb_total, kb_beg + p.kb_per_split) - kb_beg;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        mbar_init(smem_u32(acc_full), 1);
        mbar_init(smem_u32(acc_empty), 2 * NE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAh)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAl)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBh)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBl)) : "memory");
        }
        uint32_t it = 0;
        for (int64_t w = w_first; w < p.work_total; w += w_step) {
            int m_pair, kb_beg, nkb;
            decode(w, m_pair, kb_beg, nkb);
            const int arow = (2 * m_pair + (int)rank) * BM;
            for (int i = 0; i < nkb; ++i, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
                if (!elect_one()) continue;
                const uint32_t fb = smem_u32(&full[s]);
                if (rank == 0) mbar_expect_tx(fb, 2 * C_::STAGE_TX);
                const uint32_t leader = fb & 0xFEFFFFFFu;
                const int k0 = (kb_beg + i) * BK;
                const CUtensorMap* mA[2] = {&tmAh, &tmAl};
                const CUtensorMap* mB[2] = {&tmBh, &tmBl};
                const CUtensorMap* mB1[2] = {&tmBh1, &tmBl1};
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    if (MN) {
                        // boxes {64 m/n, BK k}: one 64-wide block each
                        tma_load_2d_pair(smem_u32(a_pl(s, pl)), mA[pl], leader, arow, k0);
                        tma_load_2d_pair(smem_u32(a_pl(s, pl) + C_::BLK), mA[pl], leader, arow + 64, k0);
                        tma_load_2d_pair(smem_u32(b0_pl(s, pl)), mB[pl], leader, (int)rank * H0, k0);
                        tma_load_2d_pair(smem_u32(b0_pl(s, pl) + C_::BLK), mB[pl], leader, (int)rank * H0 + 64, k0);
                        tma_load_2d_pair(smem_u32(b1_pl(s, pl)), mB[pl], leader, N0 + (int)rank * H1, k0);
                    } else {
                        tma_load_2d_pair(smem_u32(a_pl(s, pl)), mA[pl], leader, k0, arow);
                        tma_load_2d_pair(smem_u32(b0_pl(s, pl)), mB[pl], leader, k0, (int)rank * H0);
                        tma_load_2d_pair(smem_u32(b1_pl(s, pl)), mB1[pl], leader, k0, N0 + (int)rank * H1);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA) =================
        if (rank == 0) {
            constexpr uint32_t id0 = idesc_bf16(2 * BM, N0, MN), id1 = idesc_bf16(2 * BM, N1, MN);
            constexpr uint32_t LBO = MN ? (uint32_t)C_::BLK : 16u;
            constexpr uint32_t KSTEP = MN ? (2048u >> 4) : (32u >> 4);     // descriptor step per K = 16: two 8-row atoms / 32 bytes
            uint32_t it = 0, j = 0;
            for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
                int m_pair, kb_beg, nkb;
                decode(w, m_pair, kb_beg, nkb);
                mbar_wait(smem_u32(acc_empty), (j & 1u) ^ 1u);
                tc_fence_after();
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(smem_u32(&full[s]), ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t dah = desc_sw128(smem_u32(a_pl(s, 0)), LBO), dal = desc_sw128(smem_u32(a_pl(s, 1)), LBO);
                        const uint64_t db0h = desc_sw128(smem_u32(b0_pl(s, 0)), LBO), db0l = desc_sw128(smem_u32(b0_pl(s, 1)), LBO);
                        const uint64_t db1h = desc_sw128(smem_u32(b1_pl(s, 0)), LBO), db1l = desc_sw128(smem_u32(b1_pl(s, 1)), LBO);
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {               // small terms first: Al*Bh, Ah*Bl, Ah*Bh
                            const uint64_t da = pass == 0 ? dal : dah;
                            const uint64_t db0 = pass == 1 ? db0l : db0h;
                            const uint64_t db1 = pass == 1 ? db1l : db1h;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint32_t acc = (i > 0 || pass > 0 || k > 0) ? 1u : 0u;
                                const uint64_t ko = (uint64_t)(k * KSTEP);
                                umma_bf16_ss2(tmem_base, da + ko, db0 + ko, id0, acc);
                                umma_bf16_ss2(tmem_base + N0, da + ko, db1 + ko, id1, acc);
                            }
                        }
                        umma_commit2(smem_u32(&empty[s]));
                        if (i == nkb - 1) umma_commit2(smem_u32(acc_full));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================= epilogue: partial tile -> C with red.global.add =================
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int c_beg = half == 0 ? 0 : 160, c_end = half == 0 ? 160 : BN;
        uint32_t j = 0;
        for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
            int m_pair, kb_beg, nkb;
            decode(w, m_pair, kb_beg, nkb);
            const int64_t m = (int64_t)(2 * m_pair + (int)rank) * BM + q * 32 + lane;
            const bool row_ok = m < p.M;
            mbar_wait(smem_u32(acc_full), j & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            uint32_t vraw[2][16];
            tmem_ld16_issue(trow + (uint32_t)c_beg, vraw[0]);
#pragma unroll 1
            for (int c0 = c_beg; c0 < c_end; c0 += 32) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c = c0 + 16 * u;
                    if (c >= c_end) break;
                    tmem_ld16_wait(vraw[u]);
                    if (c + 16 < c_end) tmem_ld16_issue(trow + (uint32_t)(c + 16), vraw[u ^ 1]);
                    else {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank != 0) mbar_arrive_rank0(smem_u32(acc_empty));
                            else mbar_arrive(smem_u32(acc_empty));
                        }
                    }
                    if (!row_ok) continue;
                    float* crow = p.C + m * p.ldc + c;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = c + 4 * g;
                        const float v0 = __uint_as_float(vraw[u][4 * g]), v1 = __uint_as_float(vraw[u][4 * g + 1]);
                        const float v2 = __uint_as_float(vraw[u][4 * g + 2]), v3 = __uint_as_float(vraw[u][4 * g + 3]);
                        if (n + 4 <= p.nc) red_add_v4(crow + 4 * g, v0, v1, v2, v3);
                        else if (n < p.N) {
                            const float v[4] = {v0, v1, v2, v3};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (n + e < p.nc) atomicAdd(crow + 4 * g + e, v[e]);
                                else if (n + e == p.nc && p.extra && n + e < p.N) atomicAdd(p.extra + m, v[e]);
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &f, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}
struct Key {
    const void* ptr; int64_t d0, d1, ld; int b0, b1;
    bool operator<(const Key& o) const { return std::tie(ptr, d0, d1, ld, b0, b1) < std::tie(o.ptr, o.d0, o.d1, o.ld, o.b0, o.b1); }
};
static std::map<Key, CUtensorMap> g_maps;
static std::mutex g_mu;
// bf16 plane, inner (contiguous) extent d0, outer extent d1 (row stride ld elements): box {b0, b1}, SWIZZLE_128B, zero fill
static int plane_map(const void* X, int64_t d0, int64_t d1, int64_t ld, int b0, int b1, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    Key key{X, d0, d1, ld, b0, b1};
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1}, strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1}, estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (bf16 plane, pk) failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}

template <bool MN>
static int launch(const CUtensorMap* t, const Params& p, cudaStream_t s) {
    auto kern = gemm_pk_kernel<MN>;
    static bool attr_done = false;
    if (!attr_done) {
        CDG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MN>::SMEM));
        attr_done = true;
    }
    const unsigned clusters = (unsigned)imin64(p.work_total, kNumSMs / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = Cfg<MN>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CDG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, t[0], t[1], t[2], t[3], t[4], t[5], p));
    ++g_launches;
    return CDG_OK;
}

}  // namespace pk

// C[M, N] += A B^T (+ column N-1 of the product into extra_col when set), planes of both operands; mn_major: the planes are
// stored [K][M] / [K][N] (contraction index outermost).  C is accumulated into (the caller zeroes it when it wants a plain
// result).  CDG_ERR_UNSUPPORTED when the shape does not fit.
int gemm_pk(const GemmDesc& g, int mn_major, cudaStream_t s) {
    using namespace pk;
    if (!g.a_hi16 || !g.a_lo16 || !g.b_hi16 || !g.b_lo16 || !g.C) return CDG_ERR_UNSUPPORTED;
    if (g.epi != EPI_NONE || g.conv_C > 0 || g.out_hi16) return CDG_ERR_UNSUPPORTED;
    const int nc = (int)(g.extra_col ? g.N - 1 : g.N);
    if (g.N < 16 || g.N > BN || g.M < 128 || g.K < 64 || nc % 4 != 0 || g.ldc % 4 != 0 || ((uintptr_t)g.C & 15) != 0)
        return CDG_ERR_UNSUPPORTED;
    if (g.ld_a16 % 8 != 0 || g.ld_b16 % 8 != 0 || (((uintptr_t)g.a_hi16 | (uintptr_t)g.a_lo16 | (uintptr_t)g.b_hi16 | (uintptr_t)g.b_lo16) & 15))
        return CDG_ERR_UNSUPPORTED;
    if (g.M >= (1ll << 31) || g.K >= (1ll << 31)) return CDG_ERR_UNSUPPORTED;
    const int BK = mn_major ? Cfg<true>::BK : Cfg<false>::BK;
    const int kb_total = (int)((g.K + BK - 1) / BK);
    const int KB_CAP = 2048 / BK;
    const int64_t tm = (g.M + 2 * BM - 1) / (2 * BM);
    int splits = (kb_total + KB_CAP - 1) / KB_CAP;
    if (tm * splits < kNumSMs / 2 && kb_total >= 8) {                 // fill the chip when the tile grid is small
        const int want = (int)imin64((kNumSMs / 2 + tm - 1) / tm, kb_total / 4);
        splits = (int)imax64(splits, imin64(want, 256));
    }
    int kb_per = (kb_total + splits - 1) / splits;
    splits = (kb_total + kb_per - 1) / kb_per;
    Params p;
    memset(&p, 0, sizeof(p));
    p.C = g.C; p.ldc = g.ldc; p.M = g.M; p.N = (int)g.N; p.nc = nc; p.extra = g.extra_col;
    p.kb_total = kb_total; p.kb_per_split = kb_per; p.splits = splits; p.work_total = tm * splits;
    CUtensorMap t[6];
    if (mn_major) {
        // planes [K][M] and [K][N]: inner extent = M / N, boxes {64, BK}
        CDG_TRY(plane_map(g.a_hi16, g.M, g.K, g.ld_a16, 64, BK, &t[0]));
        CDG_TRY(plane_map(g.a_lo16, g.M, g.K, g.ld_a16, 64, BK, &t[1]));
        CDG_TRY(plane_map(g.b_hi16, g.N, g.K, g.ld_b16, 64, BK, &t[2]));
        CDG_TRY(plane_map(g.b_lo16, g.N, g.K, g.ld_b16, 64, BK, &t[3]));
        t[4] = t[2]; t[5] = t[3];
        return launch<true>(t, p, s);
    }
    CDG_TRY(plane_map(g.a_hi16, g.K, g.M, g.ld_a16, BK, BM, &t[0]));
    CDG_TRY(plane_map(g.a_lo16, g.K, g.M, g.ld_a16, BK, BM, &t[1]));
    CDG_TRY(plane_map(g.b_hi16, g.K, g.N, g.ld_b16, BK, H0, &t[2]));
    CDG_TRY(plane_map(g.b_lo16, g.K, g.N, g.ld_b16, BK, H0, &t[3]));
    CDG_TRY(plane_map(g.b_hi16, g.K, g.N, g.ld_b16, BK, H1, &t[4]));
    CDG_TRY(plane_map(g.b_lo16, g.K, g.N, g.ld_b16, BK, H1, &t[5]));
    return launch<false>(t, p, s);
}

}  // namespace cdg
