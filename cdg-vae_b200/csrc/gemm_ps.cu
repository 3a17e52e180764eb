// CTA-pair tcgen05 GEMM whose operands BOTH arrive pre-split into bf16 (hi, lo) planes  ("ps" = pre-split).
//
//   C[m,n] = sum_k (Ah + Al)(m,k) * (Bh + Bl)(n,k)  ~  Al*Bh + Ah*Bl + Ah*Bh        (bf16x3, fp32 accumulate in TMEM)
//
// gemm_tc.cu converts its streamed fp32 operand inside the main loop (8 converter warps between TMA and MMA).  ncu of the
// fused reconstruction head (decoder output Linear, K = 300; profiles/r02_ncu_dec2_fwd_before.txt) showed what that costs
// when K is short: the MMA warp waits 70 % of its time for the converters, the converters 55 % of theirs for TMA, and the
// four epilogue warps (one per scheduler, never waiting) set the pace at ~37 k cycles per 128 x 256 tile against 7.7 k of
// MMA work.  Here the producer of an activation writes its bf16 planes once (or a split pass does), so that
//   * TMA feeds the MMA directly (no converter warps, one hand-off less per K-block; the peer CTA's loads complete on the
//     leader's mbarrier, the `.cta_group::2` form);
//   * the freed warps are epilogue warps: eight of them, two per TMEM lane quarter, each on half of the tile's columns,
//     so two epilogue warps share every scheduler and cover each other's latencies;
//   * K-blocks are 64 bf16 = one 128-byte swizzle row (conflict-free UMMA operand reads), three 64 KB stages.
// One 256 x BN tile per cluster and work item (M = 256 spans the pair, each CTA holds 128 rows of A and BN / 2 rows of B);
// accumulators double-buffered in TMEM (2 x BN columns), so a tile's epilogue overlaps the next tile's main loop.
//
// Epilogues: the reconstruction head of the pendulum step (modules/model.py:287 tanh, modules/train.py:175 and its
// gradient), bias (+ ELU), and ELU' scaling for input gradients; each can also emit the bf16 planes of its fp32 result for
// the next GEMM.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace cdg {
namespace ps {

using namespace tc;

constexpr int BM = 128;                 // rows of A per CTA (the pair's MMA has M = 256)
constexpr int BK = 64;                  // bf16 per K-block: one 128-byte swizzle row
constexpr int NE = 8;                   // epilogue warps: two per TMEM lane quarter
constexpr int THREADS = 32 * (2 + NE);
// Staged epilogue (reconstruction head): every epilogue warp moves its 32 rows x 32 columns chunks through shared memory
// with TMA -- target chunks in (EPI_NIN deep, running ahead across tiles), gradient chunks out (bulk stores) -- so the
// threads touch only shared and tensor memory.  Row-per-thread global accesses (32 bytes of 32 different rows per
// instruction) cost the L1 data pipe one cycle per 32-byte sector: ncu of the direct version showed that pipe 58 % busy and
// every epilogue warp waiting on it (72 % long-scoreboard stalls) at 33 k cycles per tile.  Paid for with one ring stage.
constexpr int EPI_NIN = 2, EPI_NOUT = 1, EPI_CHUNK_BYTES = 32 * 128;

// Staged OUTPUT (the other epilogues, STG with EPI != EPI_RECON): the fp32 result chunk (4 KB) and the two bf16 plane chunks
// (2 KB each) of a warp leave through shared memory as bulk stores; the side operand of EPI_MUL_DACT is still read per row.
constexpr int SO_BYTES = EPI_CHUNK_BYTES + EPI_CHUNK_BYTES;

template <int BN, bool STG = false, bool RECON = true>
struct Cfg {
    static constexpr bool SO = STG && !RECON;
    static constexpr int STAGES = (STG && RECON) ? 2 : 3;
    static constexpr int EPI_BYTES = SO ? NE * SO_BYTES : STG ? NE * (EPI_NIN + EPI_NOUT) * EPI_CHUNK_BYTES : 0;
    static constexpr int A_BYTES = BM * BK * 2;               // one plane of this CTA's A rows
    static constexpr int B_ROWS = BN / 2;                     // this CTA's half of the tile's B rows
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int NACC = 2 * BN <= 512 ? 2 : 1;
    static constexpr int TMEM_COLS = NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;
    static constexpr int BAR_BYTES = 512;
    static constexpr int BIAS_BYTES = 2 * BN * 4;             // per-tile bias values, two tiles in flight
    static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + BIAS_BYTES;   // (1024-aligned dynamic base)
    static_assert(BN % 32 == 0 && BN <= 256, "one MMA per tile: N <= 256, whole 16-column chunks per epilogue half");
    static_assert(!(STG && RECON) || BN % 64 == 0, "staged reconstruction head: whole 32-column chunks per epilogue half");
    static_assert(SMEM <= 232448, "tile does not fit shared memory");
};

struct Params {
    float* C; int64_t ldc;                       // fp32 result (may be null when only the planes are wanted)
    int64_t M, N;
    int act;
    const float* bias;
    const float* aux; int64_t ld_aux;            // EPI_MUL_DACT: post-activation values
    const float* rx; int64_t rx_ld; float* rxhat; double* racc; float inv_batch;   // EPI_RECON
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int64_t ld16;                    // optional bf16 planes of the result
    int ones_col;                                // planes: column N = 1 (the consumer's bias column), later columns 0
    int c8, side8;                               // C / xhat rows (side operand rows) are 32-byte aligned: 256-bit accesses
    int kb_total, tiles_n;
    int last_ksteps;                             // K = 16 steps of the last K-block that hold real columns
    int64_t work_total;
};

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t leader_bar, int c0, int c1) {
    // the load lands in THIS CTA's shared memory and completes on the LEADER CTA's mbarrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
// K-major bf16 tile, rows of 128 B, SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void st_v8_u32(void* ptr, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ld_nc_v8f(const float* ptr, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(ptr));
}
__device__ __forceinline__ void st_v8f(float* ptr, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// 8 fp32 values -> 8 bf16 hi and 8 bf16 lo (a = hi + lo), each one 16-byte store
__device__ __forceinline__ void store_planes8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

constexpr float K2LOG2E = 2.885390081777927f;    // 2 * log2(e): tanh(x) = 1 - 2 / (1 + 2^(K2LOG2E x))

// PL (staged reconstruction head only): the gradient leaves as bf16 (hi, lo) planes (tmC / tmC2) instead of fp32 -- the same
// bytes, in the form the decoder's input- and weight-gradient GEMMs (gemm_pk.cu) consume without converting anything.
template <int BN, int EPI, bool STG, bool PL = false>
__global__ void __launch_bounds__(THREADS, 1)
gemm_ps_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmC2, const Params p) {
    static_assert(!PL || (STG && EPI == EPI_RECON), "planes of the gradient: staged reconstruction head only");
    using C_ = Cfg<BN, STG, EPI == EPI_RECON>;
    constexpr int NACC = C_::NACC;
    constexpr int STAGES = C_::STAGES;
    const uint32_t rank = cluster_ctarank();
    const int64_t w_first = (int64_t)(blockIdx.x >> 1), w_step = (int64_t)(gridDim.x >> 1);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* epi_base = smem + STAGES * C_::STAGE_BYTES;             // staged-epilogue buffers (1024-byte aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + C_::EPI_BYTES);
    uint64_t* full = bars;                       // [STAGES]  both CTAs' loads of a stage landed (used in the leader)
    uint64_t* empty = bars + STAGES;             // [STAGES]  the MMAs reading a stage retired (multicast to both CTAs)
    uint64_t* acc_full = empty + STAGES;         // [NACC]    accumulator complete (multicast)
    uint64_t* acc_empty = acc_full + NACC;       // [NACC]    drained by both CTAs' epilogue warps (leader's copy is used)
    uint64_t* in_full = acc_empty + NACC;        // [NE][EPI_NIN] staged epilogue: a warp's target chunk landed
    uint64_t* in_empty = in_full + (STG ? NE * EPI_NIN : 0);   // [NE][EPI_NIN] ... and has been read by all 32 lanes
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_empty + (STG ? NE * EPI_NIN : 0));
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C_::BAR_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto a_hi = [&](int s) { return smem + (size_t)s * C_::STAGE_BYTES; };
    auto a_lo = [&](int s) { return smem + (size_t)s * C_::STAGE_BYTES + C_::A_BYTES; };
    auto b_hi = [&](int s) { return smem + (size_t)s * C_::STAGE_BYTES + 2 * C_::A_BYTES; };
    auto b_lo = [&](int s) { return smem + (size_t)s * C_::STAGE_BYTES + 2 * C_::A_BYTES + C_::B_BYTES; };
    auto decode = [&](int64_t w, int& m_pair, int& n_blk) {
        n_blk = (int)(w % p.tiles_n);
        m_pair = (int)(w / p.tiles_n);
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), 1);
        }
        for (int b = 0; b < NACC; ++b) {
            mbar_init(smem_u32(&acc_full[b]), 1);
            mbar_init(smem_u32(&acc_empty[b]), 2 * NE);
        }
        if (STG)
            for (int b = 0; b < NE * EPI_NIN; ++b) {
                mbar_init(smem_u32(&in_full[b]), 1);
                mbar_init(smem_u32(&in_empty[b]), 32);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C_::TMEM_COLS));
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
            int m_pair, n_blk;
            decode(w, m_pair, n_blk);
            const int arow = (2 * m_pair + (int)rank) * BM;
            const int brow = n_blk * BN + (int)rank * C_::B_ROWS;
            for (int kb = 0; kb < p.kb_total; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
                if (!elect_one()) continue;
                const uint32_t fb = smem_u32(&full[s]);
                if (rank == 0) mbar_expect_tx(fb, 2 * C_::STAGE_BYTES);       // both CTAs' bytes complete on this barrier
                const uint32_t leader = fb & 0xFEFFFFFFu;                    // same offset in CTA rank 0 of the pair
                const int k0 = kb * BK;
                tma_load_2d_pair(smem_u32(a_hi(s)), &tmAh, leader, k0, arow);
                tma_load_2d_pair(smem_u32(a_lo(s)), &tmAl, leader, k0, arow);
                tma_load_2d_pair(smem_u32(b_hi(s)), &tmBh, leader, k0, brow);
                tma_load_2d_pair(smem_u32(b_lo(s)), &tmBl, leader, k0, brow);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA) =================
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
            uint32_t it = 0, j = 0;
            for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
                const uint32_t buf = j % NACC;
                mbar_wait(smem_u32(&acc_empty[buf]), ((j / NACC) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * BN;
                for (int kb = 0; kb < p.kb_total; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(smem_u32(&full[s]), ph);
                    tc_fence_after();
                    const int ksteps = kb == p.kb_total - 1 ? p.last_ksteps : BK / 16;
                    if (elect_one()) {
                        const uint64_t dah = desc_sw128(smem_u32(a_hi(s))), dal = desc_sw128(smem_u32(a_lo(s)));
                        const uint64_t dbh = desc_sw128(smem_u32(b_hi(s))), dbl = desc_sw128(smem_u32(b_lo(s)));
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {               // small terms first
                            const uint64_t da = pass == 0 ? dal : dah;
                            const uint64_t db = pass == 1 ? dbl : dbh;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)                 // UMMA K = 16 bf16 = 32 bytes along the row
                                if (k < ksteps)                               // (the last K-block may be short: K = 301 -> 3 steps)
                                    umma_bf16_ss2(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                                  (kb > 0 || pass > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit2(smem_u32(&empty[s]));
                        if (kb == p.kb_total - 1) umma_commit2(smem_u32(&acc_full[buf]));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================= epilogue (both CTAs): warp -> (TMEM lane quarter, column half) =================
        const int ew = warp - 2, q = warp & 3, half = ew >> 2;
        constexpr int HC = BN / 2;                     // columns per warp
        if constexpr (STG && EPI != EPI_RECON) {
            // ---- staged output: 32-column chunks through shared memory, bulk stores (tmX: fp32 result, tmC / tmC2: planes) ----
            constexpr int NCH32 = (BN + 31) / 32;                    // chunks per tile; the first half of them -> warp half 0
            constexpr int C0 = (NCH32 + 1) / 2;
            const int c_lo = half == 0 ? 0 : C0, c_hi = half == 0 ? C0 : NCH32;
            uint8_t* obuf = epi_base + (size_t)ew * SO_BYTES;        // [fp32 chunk 4 KB | hi 2 KB | lo 2 KB]
            const int et = threadIdx.x - 64;
            const bool has_bias = p.bias != nullptr;
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC2)) : "memory");
            }
            uint32_t j = 0;
            for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
                int m_pair, n_blk;
                decode(w, m_pair, n_blk);
                const uint32_t buf = j % NACC;
                const int row0 = (2 * m_pair + (int)rank) * BM + q * 32;
                const int64_t m = (int64_t)row0 + lane;
                const bool row_ok = m < p.M;
                float* bs = bias_s + (j & 1u) * BN;
                if ((EPI == EPI_BIAS || EPI == EPI_BIAS_ACT) && has_bias) {
                    for (int c = et; c < BN; c += 32 * NE) {
                        const int64_t n = (int64_t)n_blk * BN + c;
                        bs[c] = n < p.N ? p.bias[n] : 0.f;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
                }
                const uint32_t bsa = smem_u32(bs);
                const float* side = EPI == EPI_MUL_DACT ? p.aux + m * p.ld_aux + (int64_t)n_blk * BN : nullptr;
                float sd[32];
                auto fetch = [&](int c) {                            // side operand of chunk c: 16-byte pieces, groups past N skipped
                    if (EPI != EPI_MUL_DACT || c >= c_hi || !row_ok) return;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if ((int64_t)n_blk * BN + 32 * c + 4 * g + 4 <= p.N) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(side + 32 * c + 4 * g));
                            sd[4 * g] = t.x; sd[4 * g + 1] = t.y; sd[4 * g + 2] = t.z; sd[4 * g + 3] = t.w;
                        }
                };
                fetch(c_lo);
                mbar_wait(smem_u32(&acc_full[buf]), (j / NACC) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t vraw[32];
                    tmem_ld16_issue(trow + (uint32_t)(c * 32), vraw);
                    if (c * 32 + 16 < BN) tmem_ld16_issue(trow + (uint32_t)(c * 32 + 16), vraw + 16);
                    tmem_ld16_wait(vraw);
                    tmem_ld16_wait(vraw + 16);
                    if (c == c_hi - 1) {                             // this warp's part of the accumulator is in registers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank != 0) mbar_arrive_rank0(smem_u32(&acc_empty[buf]));
                            else mbar_arrive(smem_u32(&acc_empty[buf]));
                        }
                    }
                    const int64_t n0 = (int64_t)n_blk * BN + 32 * c;
                    if (n0 >= p.N + 4 && !(p.out_hi && n0 < p.ld16)) { fetch(c + 1); continue; }   // nothing of this chunk exists
                    float o[32];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                        if ((EPI == EPI_BIAS || EPI == EPI_BIAS_ACT) && has_bias)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                                         : "r"(bsa + (uint32_t)(32 * c + 4 * g) * 4u));
                        const float b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = 4 * g + e;
                            float x = __uint_as_float(vraw[i]);
                            if (EPI == EPI_BIAS || EPI == EPI_BIAS_ACT) x += b4[e];
                            if (EPI == EPI_BIAS_ACT) x = act_fwd(x, p.act);
                            if (EPI == EPI_MUL_DACT) x *= act_bwd_from_out(sd[i], p.act);
                            // columns past N: 0, except column N itself = 1 when the consumer's bias column needs it
                            if (n0 + i >= p.N) x = (p.ones_col && n0 + i == p.N) ? 1.f : 0.f;
                            o[i] = x;
                        }
                    }
                    fetch(c + 1);                                    // (sd[] has been consumed)
                    // the bulk stores that used these buffers have read them
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                    if (p.C) {
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            *reinterpret_cast<float4*>(obuf + sw_chunk<32>((uint32_t)lane, (uint32_t)g)) =
                                make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                    }
                    if (p.out_hi) {
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            uint32_t hh[4], ll[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float v0 = o[8 * gg + 2 * k], v1 = o[8 * gg + 2 * k + 1];
                                const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                                const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                                const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                                hh[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                                ll[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                            }
                            const uint32_t off = sw_chunk<16>((uint32_t)lane, (uint32_t)gg);
                            *reinterpret_cast<uint4*>(obuf + EPI_CHUNK_BYTES + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            *reinterpret_cast<uint4*>(obuf + EPI_CHUNK_BYTES + EPI_CHUNK_BYTES / 2 + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int col0 = (int)n0;
                        if (p.C)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(obuf)), "r"(col0), "r"(row0) : "memory");
                        if (p.out_hi) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(obuf + EPI_CHUNK_BYTES)), "r"(col0), "r"(row0) : "memory");
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&tmC2)), "r"(smem_u32(obuf + EPI_CHUNK_BYTES + EPI_CHUNK_BYTES / 2)),
                                         "r"(col0), "r"(row0) : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else if constexpr (STG) {
            static_assert(EPI == EPI_RECON, "the staged target / gradient epilogue is the reconstruction head's");
            constexpr int CPT = HC / 32;               // 32-column chunks per warp and tile
            uint8_t* ebuf = epi_base + (size_t)ew * (EPI_NIN + EPI_NOUT) * EPI_CHUNK_BYTES;
            uint64_t* infull = in_full + ew * EPI_NIN;
            uint64_t* inempty = in_empty + ew * EPI_NIN;
            const int et = threadIdx.x - 64;
            auto nvalid = [&](int n_blk) {             // chunks of this warp's column half that hold real columns
                const int64_t nv = (p.N - ((int64_t)n_blk * BN + half * HC) + 31) / 32;
                return nv < 0 ? 0 : nv < CPT ? (int)nv : CPT;
            };
            // producer side (lane 0): this warp's next target chunk, in its own (tile, chunk) order
            int64_t pw = w_first;
            int pc = 0;
            uint32_t pcount = 0;
            auto issue_next = [&]() {
                while (pw < p.work_total) {
                    int m_pair, n_blk;
                    decode(pw, m_pair, n_blk);
                    if (pc >= nvalid(n_blk)) { pc = 0; pw += w_step; continue; }
                    const uint32_t b = pcount % EPI_NIN;
                    const uint32_t bar = smem_u32(&infull[b]);
                    mbar_expect_tx(bar, EPI_CHUNK_BYTES);
                    tma_load_2d(smem_u32(ebuf + b * EPI_CHUNK_BYTES), &tmX, bar, n_blk * BN + half * HC + pc * 32,
                                (2 * m_pair + (int)rank) * BM + q * 32);
                    ++pcount;
                    ++pc;
                    return;
                }
            };
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
#pragma unroll
                for (int i = 0; i < EPI_NIN; ++i) issue_next();
            }
            uint32_t j = 0, ccount = 0;
            float rloss = 0.f;
            double dloss = 0.0;
            const bool has_bias = p.bias != nullptr;
            uint8_t* xo = ebuf + (size_t)EPI_NIN * EPI_CHUNK_BYTES;
            for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
                int m_pair, n_blk;
                decode(w, m_pair, n_blk);
                const uint32_t buf = j % NACC;
                const int row0 = (2 * m_pair + (int)rank) * BM + q * 32;
                const bool row_ok = (int64_t)row0 + lane < p.M;
                float* bs = bias_s + (j & 1u) * BN;
                if (has_bias) {
                    for (int c = et; c < BN; c += 32 * NE) {
                        const int64_t n = (int64_t)n_blk * BN + c;
                        bs[c] = (n < p.N ? p.bias[n] : 0.f) * K2LOG2E;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
                }
                const uint32_t bsh = smem_u32(bs + half * HC);
                float* xh_row = (p.rxhat && row_ok) ? p.rxhat + ((int64_t)row0 + lane) * p.ldc : nullptr;
                mbar_wait(smem_u32(&acc_full[buf]), (j / NACC) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + buf * BN + (uint32_t)(half * HC) + ((uint32_t)(q * 32) << 16);
                const int nv = nvalid(n_blk);
#pragma unroll 1
                for (int c = 0; c < CPT; ++c) {
                    uint32_t vraw[32];
                    tmem_ld16_issue(trow + (uint32_t)(c * 32), vraw);
                    tmem_ld16_issue(trow + (uint32_t)(c * 32 + 16), vraw + 16);
                    tmem_ld16_wait(vraw);
                    tmem_ld16_wait(vraw + 16);
                    if (c == CPT - 1) {                              // this warp's part of the accumulator is in registers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank != 0) mbar_arrive_rank0(smem_u32(&acc_empty[buf]));
                            else mbar_arrive(smem_u32(&acc_empty[buf]));
                        }
                    }
                    if (c >= nv) continue;                           // columns past N: nothing was loaded, nothing to store
                    const uint32_t b = ccount % EPI_NIN;
                    mbar_wait(smem_u32(&infull[b]), (ccount / EPI_NIN) & 1u);
                    ++ccount;
                    const uint32_t xin = smem_u32(ebuf + b * EPI_CHUNK_BYTES);
                    float xs[32];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        float4 t;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                     : "r"(xin + sw_chunk<32>((uint32_t)lane, (uint32_t)ch)));
                        xs[4 * ch] = t.x; xs[4 * ch + 1] = t.y; xs[4 * ch + 2] = t.z; xs[4 * ch + 3] = t.w;
                    }
                    // Refill the buffer just read -- but only once every lane's loads have been PERFORMED: a shared-memory load
                    // that has merely been issued may still be queued in the pipe when an early refill lands (seen on B200 as
                    // a few stale 32-byte row tails per million elements; __syncwarp orders execution, not completion).  The
                    // consumer-release idiom: every lane arrives (release) on the buffer's "empty" barrier behind its loads,
                    // the issuing lane waits for that phase (acquire) before the TMA write.
                    mbar_arrive(smem_u32(&inempty[b]));
                    if (lane == 0) {
                        mbar_wait(smem_u32(&inempty[b]), ((ccount - 1) / EPI_NIN) & 1u);
                        issue_next();
                    }
                    const int col0 = n_blk * BN + half * HC + c * 32;
                    // the bulk store that used the output buffer has read it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {                 // 8 columns at a time
                        float o[8];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int g = 2 * gg + h;
                            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (has_bias)
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                                             : "r"(bsh + (uint32_t)(c * 32 + 4 * g) * 4u));
                            const float b4[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int i = 4 * g + e;
                                const float ex = ex2_approx(fmaf(__uint_as_float(vraw[i]), K2LOG2E, b4[e]));
                                const float r = rcp_approx(1.f + ex);
                                const float t = fmaf(-2.f, r, 1.f);
                                const float df = t - xs[i];
                                rloss = fmaf(df, df, rloss);
                                o[4 * h + e] = (df * p.inv_batch) * fmaf(-t, t, 1.f);
                            }
                        }
                        if (PL) {
                            // rows of 64 bytes (32 bf16), SWIZZLE_64B: hi tile, then lo tile
                            uint32_t hh[4], ll[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat16 h0 = __float2bfloat16_rn(o[2 * k]), h1 = __float2bfloat16_rn(o[2 * k + 1]);
                                const __nv_bfloat16 l0 = __float2bfloat16_rn(o[2 * k] - __bfloat162float(h0));
                                const __nv_bfloat16 l1 = __float2bfloat16_rn(o[2 * k + 1] - __bfloat162float(h1));
                                hh[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                                ll[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                            }
                            const uint32_t off = sw_chunk<16>((uint32_t)lane, (uint32_t)gg);
                            *reinterpret_cast<uint4*>(xo + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                            *reinterpret_cast<uint4*>(xo + EPI_CHUNK_BYTES / 2 + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                        } else {
                            *reinterpret_cast<float4*>(xo + sw_chunk<32>((uint32_t)lane, (uint32_t)(2 * gg))) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4*>(xo + sw_chunk<32>((uint32_t)lane, (uint32_t)(2 * gg + 1))) = make_float4(o[4], o[5], o[6], o[7]);
                        }
                    }
                    if (xh_row) {
                        // only the step whose xhat is returned (the last of a train_* call): tanh recomputed from the accumulator
                        // registers, row-per-thread stores -- kept out of the loop above so that it costs the other steps nothing
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float t4[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float b = has_bias ? bs[half * HC + c * 32 + 4 * g + e] : 0.f;
                                const float ex = ex2_approx(fmaf(__uint_as_float(vraw[4 * g + e]), K2LOG2E, b));
                                t4[e] = fmaf(-2.f, rcp_approx(1.f + ex), 1.f);
                            }
                            *reinterpret_cast<float4*>(xh_row + col0 + 4 * g) = make_float4(t4[0], t4[1], t4[2], t4[3]);
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&tmC)), "r"(smem_u32(xo)), "r"(col0), "r"(row0) : "memory");
                        if (PL)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&tmC2)), "r"(smem_u32(xo + EPI_CHUNK_BYTES / 2)), "r"(col0),
                                         "r"(row0) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (row_ok) dloss += (double)rloss;                  // (rows past M: zero-filled operands, not part of the loss)
                rloss = 0.f;
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            if (p.racc) {
                const double s = warp_sum(dloss);
                if (lane == 0) atomicAdd(p.racc, 0.5 * s);
            }
        } else {
        constexpr int NCH = HC / 16;                   // 16-column chunks per warp
        const int et = threadIdx.x - 64;               // 0 .. 32 NE - 1
        uint32_t j = 0;
        float rloss = 0.f;
        double dloss = 0.0;
        for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
            int m_pair, n_blk;
            decode(w, m_pair, n_blk);
            const uint32_t buf = j % NACC;
            const int64_t m = (int64_t)(2 * m_pair + (int)rank) * BM + q * 32 + lane;
            const bool row_ok = m < p.M;
            const int64_t n_base = (int64_t)n_blk * BN + half * HC;
            float* bs = bias_s + (j & 1u) * BN;
            const bool has_bias = p.bias != nullptr;
            if ((EPI == EPI_RECON || EPI == EPI_BIAS || EPI == EPI_BIAS_ACT) && has_bias) {
                for (int c = et; c < BN; c += 32 * NE) {
                    const int64_t n = (int64_t)n_blk * BN + c;
                    const float b = n < p.N ? p.bias[n] : 0.f;
                    bs[c] = EPI == EPI_RECON ? b * K2LOG2E : b;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * NE) : "memory");
            }
            const uint32_t bsh = smem_u32(bs + half * HC);           // (explicit ld.shared below: a generic pointer costs LD.E)
            // side operand (target / post-activation), two chunks ahead of its use
            const float* side = nullptr;
            if (EPI == EPI_RECON) side = p.rx + m * p.rx_ld + n_base;
            if (EPI == EPI_MUL_DACT) side = p.aux + m * p.ld_aux + n_base;
            constexpr bool SIDE = EPI == EPI_RECON || EPI == EPI_MUL_DACT;
            constexpr int U = NCH % 4 == 0 ? 4 : NCH;  // chunks per unrolled group: sd[] / vraw[] indices are static inside it
            static_assert(NCH % U == 0 && (U >= 3 || U == NCH), "whole groups of chunks, side buffers reused two chunks later");
            float sd[U][16];
            // side operand of chunk c (two chunks ahead of its use): 32-byte pieces when the rows allow it, else 16-byte ones;
            // groups of 4 columns past N are not touched (N is a multiple of 4)
            auto fetch = [&](int c, float* dst) {
                if (c >= NCH || !row_ok) return;
                const int64_t n0 = n_base + 16 * c;
                if (p.side8 && n0 + 16 <= p.N) {
                    ld_nc_v8f(side + 16 * c, dst);
                    ld_nc_v8f(side + 16 * c + 8, dst + 8);
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (n0 + 4 * g + 4 <= p.N) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(side + 16 * c + 4 * g));
                            dst[4 * g] = t.x; dst[4 * g + 1] = t.y; dst[4 * g + 2] = t.z; dst[4 * g + 3] = t.w;
                        }
                }
            };
            if (SIDE) { fetch(0, sd[0]); if (U > 1) fetch(1, sd[1 % U]); }
            mbar_wait(smem_u32(&acc_full[buf]), (j / NACC) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + buf * BN + (uint32_t)(half * HC) + ((uint32_t)(q * 32) << 16);
            uint32_t vraw[2][16];
            tmem_ld16_issue(trow, vraw[0]);
#pragma unroll 1
            for (int c0 = 0; c0 < NCH; c0 += U) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u;
                    float v[16];
                    tmem_ld16_wait(vraw[u & 1]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vraw[u & 1][i]);
                    if (c + 1 < NCH) tmem_ld16_issue(trow + (uint32_t)(16 * (c + 1)), vraw[(u + 1) & 1]);
                    else {
                        // this warp's part of the accumulator is in registers: hand the buffer back as early as possible
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank != 0) mbar_arrive_rank0(smem_u32(&acc_empty[buf]));
                            else mbar_arrive(smem_u32(&acc_empty[buf]));
                        }
                    }
                    const int64_t n0 = n_base + 16 * c;
                    // results leave in halves of 8 columns (one 256-bit store each): keeps the live temporaries small
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int64_t n8 = n0 + 8 * h;                    // first of these 8 columns
                        const bool v0 = row_ok && n8 + 4 <= p.N, v1 = row_ok && n8 + 8 <= p.N;   // its two groups of 4
                        float o[8], bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if ((EPI == EPI_RECON || EPI == EPI_BIAS || EPI == EPI_BIAS_ACT) && has_bias) {
#pragma unroll
                            for (int g = 0; g < 2; ++g)
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb[4 * g]), "=f"(bb[4 * g + 1]),
                                             "=f"(bb[4 * g + 2]), "=f"(bb[4 * g + 3]) : "r"(bsh + (uint32_t)(16 * c + 8 * h + 4 * g) * 4u));
                        }
                        if (EPI == EPI_RECON) {
                            float t[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float e = ex2_approx(fmaf(v[8 * h + i], K2LOG2E, bb[i]));
                                const float r = rcp_approx(1.f + e);
                                t[i] = fmaf(-2.f, r, 1.f);
                                const float df = t[i] - sd[u][8 * h + i];
                                if (i < 4 ? v0 : v1) rloss = fmaf(df, df, rloss);
                                o[i] = (df * p.inv_batch) * fmaf(-t[i], t[i], 1.f);
                            }
                            if (p.rxhat) {
                                float* xh = p.rxhat + m * p.ldc + n8;
                                if (v1 && p.c8) st_v8f(xh, t);
                                else {
                                    if (v0) *reinterpret_cast<float4*>(xh) = make_float4(t[0], t[1], t[2], t[3]);
                                    if (v1) *reinterpret_cast<float4*>(xh + 4) = make_float4(t[4], t[5], t[6], t[7]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float x = v[8 * h + i];
                                if (EPI == EPI_BIAS || EPI == EPI_BIAS_ACT) x += bb[i];
                                if (EPI == EPI_BIAS_ACT) x = act_fwd(x, p.act);
                                if (EPI == EPI_MUL_DACT) x *= act_bwd_from_out(sd[u][8 * h + i], p.act);
                                o[i] = x;
                            }
                        }
                        if (p.C) {
                            float* crow = p.C + m * p.ldc + n8;
                            if (v1 && p.c8) st_v8f(crow, o);
                            else {
                                if (v0) *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
                                if (v1) *reinterpret_cast<float4*>(crow + 4) = make_float4(o[4], o[5], o[6], o[7]);
                            }
                        }
                        if (p.out_hi && row_ok && n8 < p.ld16) {
                            // planes of the result for the next GEMM: columns past N are 0, except column N itself when the
                            // consumer wants its bias column multiplied by 1 there
                            if (!v1) {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (n8 + i >= p.N) o[i] = (p.ones_col && n8 + i == p.N) ? 1.f : 0.f;
                            }
                            store_planes8(p.out_hi + m * p.ld16 + n8, p.out_lo + m * p.ld16 + n8, o);
                        }
                    }
                    if (SIDE) fetch(c + 2, sd[(u + 2) % U]);     // its buffer was consumed two chunks ago (this one when U = 2)
                }
            }
            if (EPI == EPI_RECON) { dloss += (double)rloss; rloss = 0.f; }
        }
        if (EPI == EPI_RECON && p.racc) {
            const double s = warp_sum(dloss);
            if (lane == 0) atomicAdd(p.racc, 0.5 * s);
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_::TMEM_COLS));
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
    const void* ptr; int64_t rows, k, ld; int box_rows;
    bool operator<(const Key& o) const { return std::tie(ptr, rows, k, ld, box_rows) < std::tie(o.ptr, o.rows, o.k, o.ld, o.box_rows); }
};
static std::map<Key, CUtensorMap> g_maps;
static std::mutex g_mu;
// bf16 plane [rows][K] (row stride ld elements): box {64 k, box_rows}, SWIZZLE_128B, out-of-range elements read as zero
static int plane_map(const void* X, int64_t rows, int64_t K, int64_t ld, int box_rows, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    Key key{X, rows, K, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (bf16 plane) failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}

// fp32 matrix [rows][cols] (row stride ld floats) as 32 x 32 boxes, SWIZZLE_128B: the staged epilogue's target / gradient chunks
static int chunk_map_f32(const float* X, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    Key key{X, rows, cols, ld, -32};
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32u, 32u}, estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32 chunks) failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}

// bf16 plane [rows][cols] (row stride ld elements) as 32 x 32 boxes of 64-byte rows, SWIZZLE_64B: the gradient planes the
// staged reconstruction head stores
static int chunk_map_bf16(const void* X, int64_t rows, int64_t cols, int64_t ld, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    Key key{X, rows, cols, ld, -16};
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {32u, 32u}, estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (bf16 chunks) failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}

template <int BN, int EPI, bool STG = false, bool PL = false>
static int launch(const CUtensorMap& tah, const CUtensorMap& tal, const CUtensorMap& tbh, const CUtensorMap& tbl, const Params& p,
                  cudaStream_t s, const CUtensorMap* tx = nullptr, const CUtensorMap* tc_ = nullptr, const CUtensorMap* tc2 = nullptr) {
    auto kern = gemm_ps_kernel<BN, EPI, STG, PL>;
    static bool attr_done = false;
    if (!attr_done) {
        CDG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, STG, EPI == EPI_RECON>::SMEM));
        attr_done = true;
    }
    const unsigned clusters = (unsigned)imin64(p.work_total, kNumSMs / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = Cfg<BN, STG, EPI == EPI_RECON>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CDG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tah, tal, tbh, tbl, tx ? *tx : tah, tc_ ? *tc_ : tah, tc2 ? *tc2 : tah, p));
    ++g_launches;
    return CDG_OK;
}

template <int BN>
static int launch_epi(int epi, const CUtensorMap& tah, const CUtensorMap& tal, const CUtensorMap& tbh, const CUtensorMap& tbl,
                      const Params& p, cudaStream_t s) {
    switch (epi) {
        case EPI_RECON: return launch<BN, EPI_RECON>(tah, tal, tbh, tbl, p, s);
        case EPI_BIAS: return launch<BN, EPI_BIAS>(tah, tal, tbh, tbl, p, s);
        case EPI_BIAS_ACT: return launch<BN, EPI_BIAS_ACT>(tah, tal, tbh, tbl, p, s);
        case EPI_MUL_DACT: return launch<BN, EPI_MUL_DACT>(tah, tal, tbh, tbl, p, s);
        case EPI_NONE: return launch<BN, EPI_NONE>(tah, tal, tbh, tbl, p, s);
    }
    return CDG_ERR_UNSUPPORTED;
}

template <int BN>
static int launch_epi_so(int epi, const CUtensorMap& tah, const CUtensorMap& tal, const CUtensorMap& tbh, const CUtensorMap& tbl,
                         const Params& p, cudaStream_t s, const CUtensorMap* tc_, const CUtensorMap* th, const CUtensorMap* tl) {
    switch (epi) {
        case EPI_BIAS: return launch<BN, EPI_BIAS, true>(tah, tal, tbh, tbl, p, s, tc_, th, tl);
        case EPI_BIAS_ACT: return launch<BN, EPI_BIAS_ACT, true>(tah, tal, tbh, tbl, p, s, tc_, th, tl);
        case EPI_MUL_DACT: return launch<BN, EPI_MUL_DACT, true>(tah, tal, tbh, tbl, p, s, tc_, th, tl);
        case EPI_NONE: return launch<BN, EPI_NONE, true>(tah, tal, tbh, tbl, p, s, tc_, th, tl);
    }
    return CDG_ERR_UNSUPPORTED;
}

}  // namespace ps

static bool al32p(const void* q) { return ((uintptr_t)q & 31) == 0; }

// C = A B^T with A(m,k) = a_hi16 + a_lo16 ([M][ld_a16] bf16 planes) and B(n,k) = b_hi16 + b_lo16 ([N][ld_b16]); out_hi / out_lo
// (optional, [M][ld_out16]) receive the bf16 planes of the fp32 result.  CDG_ERR_UNSUPPORTED when the shape / alignment does
// not fit (the caller then takes the converting kernel).
int gemm_ps(const GemmDesc& g, cudaStream_t s) {
    using namespace ps;
    void* out_hi = g.out_hi16; void* out_lo = g.out_lo16;
    const int64_t ld_out16 = g.ld_out16;
    const int ones_col = g.out_ones;
    if (!g.a_hi16 || !g.a_lo16 || !g.b_hi16 || !g.b_lo16) return CDG_ERR_UNSUPPORTED;
    static const int64_t min_m = exp_switch("CDG_PS_MIN_M", 128);
    if (g.M < min_m || g.N < 16 || g.N % 4 != 0 || g.K < 16 || g.K > 2048 || g.accumulate || g.extra_col || g.conv_C > 0)
        return CDG_ERR_UNSUPPORTED;
    auto al16p = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (g.ld_a16 % 8 != 0 || g.ld_b16 % 8 != 0 || (((uintptr_t)g.a_hi16 | (uintptr_t)g.a_lo16 | (uintptr_t)g.b_hi16 | (uintptr_t)g.b_lo16) & 15))
        return CDG_ERR_UNSUPPORTED;
    if (g.C && !(g.ldc % 4 == 0 && al16p(g.C))) return CDG_ERR_UNSUPPORTED;
    if (!g.C && !out_hi) return CDG_ERR_UNSUPPORTED;
    if (out_hi && !(out_lo && ld_out16 % 8 == 0 && ld_out16 >= g.N + (ones_col ? 1 : 0) && al16p(out_hi) && al16p(out_lo)))
        return CDG_ERR_UNSUPPORTED;
    if (g.epi == EPI_RECON && g.recon_xhat && !g.C && g.ldc % 4 != 0) return CDG_ERR_UNSUPPORTED;   // xhat rows use ldc
    // (a null bias with a bias epilogue = the bias is folded into the contraction: a ones column in A, the bias column in B)
    if (g.epi == EPI_MUL_DACT && !(g.aux && g.ld_aux % 4 == 0 && al16p(g.aux))) return CDG_ERR_UNSUPPORTED;
    if (g.epi == EPI_RECON) {
        if (!g.recon_x || (!g.C && !out_hi)) { set_error("EPI_RECON without target / gradient buffer"); return CDG_ERR_INVALID; }
        if (!(g.ld_x % 4 == 0 && al16p(g.recon_x) && (!g.recon_xhat || al16p(g.recon_xhat)))) return CDG_ERR_UNSUPPORTED;
    }
    // tile width: 256 for wide outputs; the hidden layers (N = 300) take two 160-wide tiles
    const int BN = g.N > 320 ? 256 : g.N > 256 ? 160 : g.N > 160 ? 256 : g.N > 128 ? 160 : g.N > 64 ? 128 : 64;
    const int64_t tm = (g.M + 2 * BM - 1) / (2 * BM), tn = (g.N + BN - 1) / BN;
    if (tn > (1 << 30) || tm > (1 << 30) || g.M >= (1ll << 31) || g.N >= (1ll << 31)) return CDG_ERR_UNSUPPORTED;
    Params p;
    memset(&p, 0, sizeof(p));
    p.C = g.C; p.ldc = g.ldc; p.M = g.M; p.N = g.N; p.act = g.act; p.bias = g.bias; p.aux = g.aux; p.ld_aux = g.ld_aux;
    p.rx = g.recon_x; p.rx_ld = g.ld_x; p.rxhat = g.recon_xhat; p.racc = g.recon_acc; p.inv_batch = g.inv_batch;
    p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo; p.ld16 = ld_out16; p.ones_col = ones_col;
    p.c8 = (!g.C || (g.ldc % 8 == 0 && al32p(g.C))) && (!g.recon_xhat || al32p(g.recon_xhat)) ? 1 : 0;
    p.side8 = g.epi == EPI_RECON ? (g.ld_x % 8 == 0 && al32p(g.recon_x)) : g.epi == EPI_MUL_DACT ? (g.ld_aux % 8 == 0 && al32p(g.aux)) : 0;
    p.kb_total = (int)((g.K + BK - 1) / BK); p.tiles_n = (int)tn; p.work_total = tm * tn;
    p.last_ksteps = (int)((g.K - (int64_t)(p.kb_total - 1) * BK + 15) / 16);
    CUtensorMap tah, tal, tbh, tbl;
    CDG_TRY(plane_map(g.a_hi16, g.M, g.K, g.ld_a16, BM, &tah));
    CDG_TRY(plane_map(g.a_lo16, g.M, g.K, g.ld_a16, BM, &tal));
    CDG_TRY(plane_map(g.b_hi16, g.N, g.K, g.ld_b16, BN / 2, &tbh));
    CDG_TRY(plane_map(g.b_lo16, g.N, g.K, g.ld_b16, BN / 2, &tbl));
    static const int staged = exp_switch("CDG_PS_STG", 1);
    if (out_hi && g.epi == EPI_RECON) {
        // reconstruction head with the gradient as bf16 planes (no fp32 copy): staged epilogue only
        if (!(BN == 256 && g.N % 32 == 0 && p.side8 && ld_out16 % 8 == 0)) return CDG_ERR_UNSUPPORTED;
        CUtensorMap tx, th, tl;
        CDG_TRY(chunk_map_f32(g.recon_x, g.M, g.N, g.ld_x, &tx));
        CDG_TRY(chunk_map_bf16(out_hi, g.M, g.N, ld_out16, &th));
        CDG_TRY(chunk_map_bf16(out_lo, g.M, g.N, ld_out16, &tl));
        tl_planes_done = true;
        return launch<256, EPI_RECON, true, true>(tah, tal, tbh, tbl, p, s, &tx, &th, &tl);
    }
    if (out_hi) tl_planes_done = true;
    if (g.epi == EPI_RECON && BN == 256 && g.N % 32 == 0 && staged && p.c8 && p.side8) {
        // reconstruction head: target in / gradient out through shared memory by TMA
        CUtensorMap tx, tcm;
        CDG_TRY(chunk_map_f32(g.recon_x, g.M, g.N, g.ld_x, &tx));
        CDG_TRY(chunk_map_f32(g.C, g.M, g.N, g.ldc, &tcm));
        return launch<256, EPI_RECON, true>(tah, tal, tbh, tbl, p, s, &tx, &tcm);
    }
    static const int staged_out = exp_switch("CDG_PS_SO", 1);
    if (staged_out && BN <= 160 && g.epi != EPI_RECON && g.N % 4 == 0 && (!g.C || g.ldc % 4 == 0)) {
        // hidden-layer GEMMs: result and planes leave as bulk stores of 32-column chunks
        CUtensorMap tcm = tah, th = tah, tl = tah;
        if (g.C) CDG_TRY(chunk_map_f32(g.C, g.M, g.N, g.ldc, &tcm));
        if (out_hi) {
            CDG_TRY(chunk_map_bf16(out_hi, g.M, ld_out16, ld_out16, &th));
            CDG_TRY(chunk_map_bf16(out_lo, g.M, ld_out16, ld_out16, &tl));
        }
        if (BN == 160) return launch_epi_so<160>(g.epi, tah, tal, tbh, tbl, p, s, &tcm, &th, &tl);
        if (BN == 128) return launch_epi_so<128>(g.epi, tah, tal, tbh, tbl, p, s, &tcm, &th, &tl);
        return launch_epi_so<64>(g.epi, tah, tal, tbh, tbl, p, s, &tcm, &th, &tl);
    }
    if (BN == 256) return launch_epi<256>(g.epi, tah, tal, tbh, tbl, p, s);
    if (BN == 160) return launch_epi<160>(g.epi, tah, tal, tbh, tbl, p, s);
    if (BN == 128) return launch_epi<128>(g.epi, tah, tal, tbh, tbl, p, s);
    return launch_epi<64>(g.epi, tah, tal, tbh, tbl, p, s);
}

}  // namespace cdg
