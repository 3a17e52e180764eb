// tcgen05 / TMEM / TMA GEMM for sm_100a with FP32-faithful arithmetic (3xTF32).
//
//   C[m,n] = sum_k A(m,k) * B(n,k)        A(m,k) = A[m*sa_m + k*sa_k],  B(n,k) = B[n*sb_n + k*sb_k]
//
// The reference's Linear layers are true-FP32 cuBLAS/MKL GEMMs (allow_tf32 = False) and parity is
// judged at 1e-4 on gradients, so single-pass TF32 (10-bit mantissa) is not enough.  Each fp32
// operand is split on the fly into hi = tf32(a) and lo = a - hi, and every K-step issues
//   D += A_lo*B_hi;  D += A_hi*B_lo;  D += A_hi*B_hi        (tcgen05.mma.kind::tf32, fp32 accumulate in TMEM)
// which recovers ~21 mantissa bits per product.
//
// Persistent kernel, one CTA per SM (shared memory bound), looping over (tile, K-split) work items of
// 128 x BN outputs; 14 warps:
//   warp 0       TMA producer: cp.async.bulk.tensor of the raw fp32 tiles (2-stage ring, mbarrier tx-count),
//                running ahead across work items
//   warp 1       allocates TMEM, issues tcgen05.mma (one lane), commits to the ring's "empty" barriers and
//                to the accumulator-full barrier; accumulators are double-buffered in TMEM when 2*BN <= 512
//   warps 2-9    converter: split raw tiles into hi/lo K-major SWIZZLE_128B tiles in shared memory (operands
//                whose contraction index is not the contiguous one -- dgrad weights, both wgrad operands --
//                arrive as [k][mn] boxes and are transposed here)
//   warps 10-13  epilogue: tcgen05.ld the accumulator (one TMEM lane = one output row per thread), bias /
//                ELU / ELU' / accumulate / fused tanh-MSE reconstruction head, 128-bit stores; overlaps the
//                next work item's main loop
// Partial tiles rely on TMA zero fill.  K-splits add into C with red.global.add.f32.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

namespace cdg {

#ifdef CDG_EXPERIMENTS
#define CDG_PROBE(p) ((p).probe)
#define CDG_SKIP_FENCE_OFF(p) ((p).probe & 4)       // experiments: CDG_TC_PROBE=4 keeps the proxy fence (A/B timing)
#else
#define CDG_SKIP_FENCE_OFF(p) 0
#define CDG_PROBE(p) 0          // the timing probes (skip conversion / skip TMA: results invalid) are not in the shipped kernels
#endif

namespace tc {

constexpr int BM = 128;
// K-block = one swizzle row: 32 floats (SWIZZLE_128B, 2-stage ring) or 16 floats (SWIZZLE_64B, 4-stage ring).
// The ring holds the same bytes either way; the finer blocks keep more TMA loads in flight per byte.
#ifndef CDG_TC_BF3X_A_TMEM
#define CDG_TC_BF3X_A_TMEM 1
#endif
#define CDG_TC_A_TMEM_BF16_OK(bf3x) (!(bf3x) || CDG_TC_BF3X_A_TMEM)
constexpr int NUM_CONV_WARPS = 8;
#ifndef CDG_TC_CONV_GROUPS
#define CDG_TC_CONV_GROUPS 2
#endif
constexpr int CONV_GROUPS = CDG_TC_CONV_GROUPS;   // 1: all eight converter warps share every stage; 2: two groups alternate stages
#ifndef CDG_TC_EPI_WARPS
#define CDG_TC_EPI_WARPS 4
#endif
// 4 epilogue warps (one per TMEM lane quarter).  8 (two per quarter, each taking half of the tile's columns) were measured
// on the short-K fused-head GEMM (dec2 forward, K = 300): 7.4 vs 6.8 ms per step -- the epilogue is not what paces it
constexpr int NUM_EPI_WARPS = CDG_TC_EPI_WARPS;
constexpr int THREADS = 32 * (2 + NUM_CONV_WARPS + NUM_EPI_WARPS);

struct Params {
    float* C; int64_t sc_m, sc_n;
    int64_t M, N;
    int epi, act, accumulate, atomic, vec;      // vec: row-major output, 16-byte aligned rows -> float4 path
    int vec8;                                   // rows 32-byte aligned as well -> 256-bit LDG/STG (full sectors)
    const float* bias; int bias_on_m;
    const float* aux; int64_t aux_sm, aux_sn;
    const float* rx; int64_t rx_ld; float* rxhat; double* racc; float inv_batch;   // EPI_RECON
    int kb_total, kb_per_split;                 // K-blocks (of BK) in total / per split
    int tiles_n, splits;
    int probe;                                  // timing experiments only (results invalid): 1 = no conversion, 2 = no TMA
    int rawhi;                                  // K-major operands: leave the raw tile as `hi` (tensor core truncates)
    int64_t work_total;                         // tiles_m * tiles_n * splits
    float* extra;                               // column N-1 of the result goes to extra[m] (bias gradient via a ones row)
    int a16swap;                                // experiment switch: order of the two bf16 in a packed TMEM column
    int b_pre;                                  // bf16x3: B arrives pre-split (hi / lo bf16 tiles by TMA, no conversion)
    int sw32;                                   // bf16x3: bf16 tiles as two K = 16 sub-tiles of 32-byte rows (SWIZZLE_32B)
    int tr;                                     // fused reconstruction head: transposing (row-coalesced) epilogue
    int conv_cb, conv_W, conv_H, conv_k;        // implicit-GEMM convolution: K-blocks per tap (0 = plain GEMM), extent, kernel
};


// Staged reconstruction epilogue (STG): every epilogue warp owns EPI_NIN target buffers and EPI_NOUT output buffers of
// 32 rows x 32 floats (SWIZZLE_128B) moved by TMA, paid for with one ring stage.
constexpr int EPI_NOUT = 2, EPI_CHUNK_BYTES = 32 * 128;

template <int BN, int BK, int PASSES = 3, bool STG = false, bool CTA2 = false, bool TR = false>
struct Cfg {
    // PASSES == 2 is the bf16x3 arithmetic (see the BF3X notes at the converter): the raw fp32 tile is converted IN PLACE
    // into a hi and a lo bf16 tile (a 128-byte fp32 row becomes two 64-byte bf16 rows), so a stage holds one copy of
    // each operand and the ring is 4 deep at every tile width.
    static constexpr bool BF3X = PASSES == 2;
    // (a 6-stage ring was tried for BN = 160 and changed nothing: the ring depth is not what limits the main loop)
    // narrow tiles (BN <= 64: the convolutions with 32 / 64 output channels) leave room for a 4-deep ring of 32-float blocks
    static constexpr int STAGES = CTA2 ? (STG ? 4 : 6) : BF3X ? (STG ? 3 : 4) : (BK == 32 ? (BN <= 64 ? 4 : 2) : 4);
    static constexpr int EPI_NIN = CTA2 ? 4 : 3;
    static constexpr int EPI_BYTES = STG ? NUM_EPI_WARPS * (EPI_NIN + EPI_NOUT) * EPI_CHUNK_BYTES : 0;
    static constexpr int ROW_BYTES = BK * 4;
    static constexpr int A_BYTES = BM * BK * 4;
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int B_CTA_BYTES = CTA2 ? B_BYTES / 2 : B_BYTES;   // CTA pairs: each CTA stages half of the B rows
    static constexpr int STAGE_BYTES = BF3X ? (A_BYTES + B_CTA_BYTES) : 2 * (A_BYTES + B_BYTES);  // hi + lo for both operands
    static constexpr int BIAS_LD = 320;                               // epilogue's bias tile (two tiles in flight)
    static constexpr int BIAS_BYTES = STG ? 0 : 2 * BIAS_LD * 4;
    // transposing epilogue of the fused reconstruction head (256-wide tiles): one 32 x 33 float scratch per epilogue warp
    static constexpr int TR_LD = 33;
    static constexpr int TR_BYTES = TR ? NUM_EPI_WARPS * 32 * TR_LD * 4 : 0;
    static_assert(!TR || (!STG && BN % 32 == 0 && NUM_EPI_WARPS == 4), "transposing epilogue: whole 32-column chunks, one warp per lane quarter");
    static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 384 /*barriers*/ + BIAS_BYTES + TR_BYTES;
    static_assert(!STG || (BF3X && BN % 32 == 0 && NUM_EPI_WARPS == 4), "staged epilogue: bf16x3 tiles, one warp per lane quarter");
    static constexpr int B_ROWS_PER_BOX = BN <= 256 ? BN : BN / 2;  // K-major TMA box rows (<= 256)
    static constexpr int B_CW = BN <= 128 ? BN : (BN % 128 == 0 ? 128 : (BN == 160 ? 80 : BN / 4));   // MN-major chunk width (<= 128)
    static constexpr int N0 = BN <= 256 ? BN : 160;                 // first MMA's N
    static constexpr int N1 = BN - N0;                              // second MMA's N (0 = none)
    static constexpr int NACC = (2 * BN <= 512) ? 2 : 1;            // TMEM accumulator buffers
    // With a single accumulator (BN = 304) the remaining TMEM columns hold the A operand ring (hi and lo,
    // BK columns each per stage): the MMA then reads A from tensor memory, which removes A from the shared
    // memory data path -- the path ncu shows saturated (LSU + tensor wavefronts at 92 % of peak).
    static constexpr int A_COL0 = (NACC * BN + 31) / 32 * 32;
    // bf16x3: hi and lo of the A tile as packed bf16 pairs, BK/2 columns each per stage (the converter then writes no shared
    // memory at all and the UMMA reads only B from it: ncu showed the L1/shared pipe at 68-75 % with A in shared memory)
    static constexpr int A_STAGE_COLS = BF3X ? BK : 2 * BK;
    static constexpr bool A_TMEM = (A_COL0 + STAGES * A_STAGE_COLS <= 512) && CDG_TC_A_TMEM_BF16_OK(BF3X);
    static constexpr int TMEM_USED = (NACC * BN + 31) / 32 * 32 + (A_TMEM ? STAGES * A_STAGE_COLS : 0);
    static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static_assert(BN % 16 == 0 && N0 % 16 == 0 && N1 % 16 == 0, "UMMA N must be a multiple of 16 at M = 128");
    static_assert(SMEM <= 232448, "tile does not fit shared memory");
    static_assert(!BF3X || BK == 32, "bf16x3 tiles are built from 32-float K-blocks");
    static_assert(!CTA2 || BF3X, "CTA pairs: bf16x3 tiles");
    static_assert(!CTA2 || (N0 % 16 == 0 && N1 % 16 == 0), "UMMA N is a multiple of 16 at M = 256; half-B boxes are whole 8-row groups");
};

// ---- converter --------------------------------------------------------------------------------
// K-major source: the raw tile already sits in the `hi` buffer in its final (swizzled) place.  Any
// element-wise rewrite keeps the layout, so the tile is processed as a flat float4 array.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int PASSES, int NT>
__device__ __forceinline__ void convert_kmajor(uint8_t* hi, uint8_t* lo, int bytes, int ct, int rawhi) {
    if (PASSES == 1) return;
    float4* h = reinterpret_cast<float4*>(hi);
    float4* l = reinterpret_cast<float4*>(lo);
    const int n = bytes / 16;
    if (rawhi) {
        // the tensor core ignores the 13 low mantissa bits of a tf32 operand: hi = trunc(raw) is implicit
#pragma unroll 4
        for (int i = ct; i < n; i += NT) {
            const float4 r = h[i];
            l[i] = make_float4(r.x - tf32_trunc(r.x), r.y - tf32_trunc(r.y), r.z - tf32_trunc(r.z), r.w - tf32_trunc(r.w));
        }
        return;
    }
#pragma unroll 4
    for (int i = ct; i < n; i += NT) {
        const float4 r = h[i];
        float4 a, b;
        a.x = tf32_rna(r.x); a.y = tf32_rna(r.y); a.z = tf32_rna(r.z); a.w = tf32_rna(r.w);
        b.x = r.x - a.x; b.y = r.y - a.y; b.z = r.z - a.z; b.w = r.w - a.w;
        h[i] = a;
        l[i] = b;
    }
}

// A operand -> tensor memory.  Converter warp w owns TMEM lane quarter w % 4 (rows 32q..32q+31, one per lane); the
// two warps of a quarter split the K-block.  Source: K-major raw tile (swizzled, in `hi`) or MN-major raw box
// ([k][128], in `lo`).
template <int PASSES, int BK, bool MN>
__device__ __forceinline__ void convert_a_tmem(const uint8_t* hi, const uint8_t* lo, uint32_t tmem_hi, int warp, int lane, int khalf) {
    constexpr int NV = BK / 2;                       // values per thread: 16 (BK = 32) or 8 (BK = 16)
    const int q = warp & 3;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const int kbeg = khalf * NV;
    float v[NV];
    if (!MN) {
#pragma unroll
        for (int c = 0; c < NV / 4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(hi + sw_chunk<BK>(row, (uint32_t)(kbeg / 4 + c)));
            v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
        }
    } else {
        const float* raw = reinterpret_cast<const float*>(lo);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = raw[(kbeg + j) * BM + row];
    }
    uint32_t h[NV], l[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float a = PASSES == 1 ? v[j] : tf32_rna(v[j]);
        h[j] = __float_as_uint(a);
        l[j] = __float_as_uint(v[j] - a);
    }
    const uint32_t base = tmem_hi + ((uint32_t)(q * 32) << 16) + (uint32_t)kbeg;
#pragma unroll
    for (int j = 0; j < NV; j += 8) {
        tmem_st8(base + j, h + j);
        if (PASSES == 3) tmem_st8(base + BK + j, l + j);
    }
    tmem_st_wait();
}

// MN-major source: raw boxes [BK k-rows][CW mn] landed in the `lo` buffer (chunk c at byte c*CW*128).
// Each thread owns one mn column of a chunk: read its BK values (conflict-free), wait until every
// converter thread has read the chunk (its bytes are about to be overwritten), then write the K-major
// rows of hi and lo.
template <int PASSES, int ROWS, int CW, int BK, int NT>
__device__ __forceinline__ void convert_mnmajor(uint8_t* hi, uint8_t* lo, int ct, int bar_id) {
    constexpr int IT = (ROWS + NT - 1) / NT;
    float v[IT][BK];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int j = ct + it * NT;                       // mn index inside the tile
        if (j < ROWS) {
            const int c = j / CW, col = j - c * CW;
            const float* raw = reinterpret_cast<const float*>(lo + (size_t)c * CW * (BK * 4));
#pragma unroll
            for (int k = 0; k < BK; ++k) v[it][k] = raw[k * CW + col];
        }
    }
    // every raw byte has been read before any of them is overwritten by the lo rows
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int j = ct + it * NT;
        if (j < ROWS) {
#pragma unroll
            for (int ch = 0; ch < BK / 4; ++ch) {
                const float x0 = v[it][4 * ch], x1 = v[it][4 * ch + 1], x2 = v[it][4 * ch + 2], x3 = v[it][4 * ch + 3];
                const uint32_t off = sw_chunk<BK>((uint32_t)j, ch);
                if (PASSES == 1) {
                    *reinterpret_cast<float4*>(hi + off) = make_float4(x0, x1, x2, x3);
                } else {
                    float4 a;
                    a.x = tf32_rna(x0); a.y = tf32_rna(x1); a.z = tf32_rna(x2); a.w = tf32_rna(x3);
                    *reinterpret_cast<float4*>(hi + off) = a;
                    *reinterpret_cast<float4*>(lo + off) = make_float4(x0 - a.x, x1 - a.y, x2 - a.z, x3 - a.w);
                }
            }
        }
    }
}

// ---- bf16x3 (PASSES == 2) -------------------------------------------------------------------------
// a = hi + lo with hi = bf16(a), lo = bf16(a - hi): 16 mantissa bits per operand.  Every K-block issues
//   D += A_lo*B_hi;  D += A_hi*B_lo;  D += A_hi*B_hi        (tcgen05.mma.kind::f16 on bf16, fp32 accumulate in TMEM)
// i.e. 3 bf16 MMAs (= 1.5 TF32 MMAs) per product instead of 3 TF32 MMAs, and half the operand bytes through the
// shared-memory pipe.  Dropped term A_lo*B_lo ~ 2^-18; representation error 2^-17 per operand: ~1e-5 relative per GEMM.
// Tiles: raw fp32 rows of 128 B (SWIZZLE_128B) become, in place, a hi tile (rows of 64 B, SWIZZLE_64B) in the first half
// of the buffer and a lo tile in the second half.  All reads of a tile complete (named barrier) before any write.
__device__ __forceinline__ void bf16_split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
        h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);   // lower k in the low half
        l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// A tile -> tensor memory as packed bf16 pairs (hi: columns [0, 16), lo: [16, 32) of the stage's slot).  Converter warp w owns
// TMEM lane quarter w % 4 (one row per lane); the two warps of a quarter take 16 of the 32 K entries each.  Source: K-major
// raw tile (SWIZZLE_128B) or MN-major raw box ([32 k][128 m]).  No shared-memory writes, no barrier.
template <bool MN>
__device__ __forceinline__ void bf3x_a_tmem(const uint8_t* raw, uint32_t tmem_slot, int warp, int lane, int swap_halves) {
    const int q = warp & 3, khalf = (warp - 2) >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    float v[16];
    if (!MN) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(raw + sw_chunk<32>(row, (uint32_t)(khalf * 4 + c)));
            v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
        }
    } else {
        const float* r = reinterpret_cast<const float*>(raw);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = r[(khalf * 16 + j) * BM + row];
    }
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
        const uint32_t a0 = __bfloat16_as_ushort(h0), a1 = __bfloat16_as_ushort(h1);
        const uint32_t b0 = __bfloat16_as_ushort(l0), b1 = __bfloat16_as_ushort(l1);
        h[i] = swap_halves ? (a1 | (a0 << 16)) : (a0 | (a1 << 16));
        l[i] = swap_halves ? (b1 | (b0 << 16)) : (b0 | (b1 << 16));
    }
    const uint32_t base = tmem_slot + ((uint32_t)(q * 32) << 16) + (uint32_t)(khalf * 8);
    tmem_st8(base, h);
    tmem_st8(base + 16, l);
    tmem_st_wait();
}

// Byte offset of the 16-byte chunk g (8 bf16: k = 8g .. 8g+7) of row r inside a bf16 tile of ROWS rows x 32 k.
//   sw32 == 0: rows of 64 B, SWIZZLE_64B.  An MMA (K = 16) then reads half of every row: the 32-byte pieces of rows r and
//              r + 2 fall on the same banks (the swizzle only permutes chunks inside the half), so every operand read costs
//              two shared-memory wavefronts per useful one.
//   sw32 != 0: two K = 16 sub-tiles of 32-byte rows, SWIZZLE_32B (address bit 4 ^= bit 7): 8 rows = 256 contiguous bytes.
template <int ROWS>
__device__ __forceinline__ uint32_t bf16_tile_off(uint32_t r, uint32_t g, int sw32) {
    if (sw32) return (g >> 1) * (uint32_t)(ROWS * 32) + r * 32u + (((g & 1u) ^ ((r >> 2) & 1u)) << 4);
    return sw_chunk<16>(r, g);
}

// K-major raw tile [ROWS][32 floats]: unit u = (row, 8-float group g) -> one 16-byte bf16 chunk of hi and of lo
template <int ROWS, int NT>
__device__ __forceinline__ void bf3x_kmajor(uint8_t* buf, int ct, int bar_id, int sw32) {
    constexpr int UNITS = ROWS * 4;
    constexpr int IT = (UNITS + NT - 1) / NT;
    float v[IT][8];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int u = ct + it * NT;
        if (u < UNITS) {
            const uint32_t r = (uint32_t)(u >> 2), g = (uint32_t)(u & 3);
            const float4 a = *reinterpret_cast<const float4*>(buf + sw_chunk<32>(r, 2 * g));
            const float4 b = *reinterpret_cast<const float4*>(buf + sw_chunk<32>(r, 2 * g + 1));
            v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
            v[it][4] = b.x; v[it][5] = b.y; v[it][6] = b.z; v[it][7] = b.w;
        }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
    uint8_t* lo_tile = buf + ROWS * 64;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int u = ct + it * NT;
        if (u < UNITS) {
            const uint32_t r = (uint32_t)(u >> 2), g = (uint32_t)(u & 3);
            uint4 hi, lo;
            bf16_split8(v[it], hi, lo);
            const uint32_t off = bf16_tile_off<ROWS>(r, g, sw32);
            *reinterpret_cast<uint4*>(buf + off) = hi;
            *reinterpret_cast<uint4*>(lo_tile + off) = lo;
        }
    }
}

// MN-major raw boxes [32 k-rows][CW mn] (chunk c at byte c*CW*128): thread j owns mn column j, i.e. K-major row j
template <int ROWS, int CW, int NT>
__device__ __forceinline__ void bf3x_mnmajor(uint8_t* buf, int ct, int bar_id, int sw32) {
    constexpr int IT = (ROWS + NT - 1) / NT;
    float v[IT][32];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int j = ct + it * NT;
        if (j < ROWS) {
            const int c = j / CW, col = j - c * CW;
            const float* raw = reinterpret_cast<const float*>(buf + (size_t)c * CW * 128);
#pragma unroll
            for (int k = 0; k < 32; ++k) v[it][k] = raw[k * CW + col];
        }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
    uint8_t* lo_tile = buf + ROWS * 64;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
        const int j = ct + it * NT;
        if (j < ROWS) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint4 hi, lo;
                bf16_split8(&v[it][8 * g], hi, lo);
                const uint32_t off = bf16_tile_off<ROWS>((uint32_t)j, (uint32_t)g, sw32);
                *reinterpret_cast<uint4*>(buf + off) = hi;
                *reinterpret_cast<uint4*>(lo_tile + off) = lo;
            }
        }
    }
}

// Branch-free tanh for the fused reconstruction head: tanh(x) = 1 - 2 / (1 + e^{2x}).  Absolute error <= ~3e-7
// (ex2.approx + rcp.approx), i.e. well inside the 1e-4 parity budget of xhat in [-1, 1]; tanhf()'s branchy
// polynomial/exp split serialises the single epilogue warp per scheduler (measured: 66k cycles per 128x256 tile).
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, 1.f + e);
}

// 256-bit global accesses (sm_100): one instruction moves a full 32-byte sector per thread.  With 128-bit stores
// ncu showed 2x DRAM write amplification on the row-per-thread epilogue (two partial-sector writes per sector).
__device__ __forceinline__ void ld_nc_v8(const float* ptr, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(ptr));
}
__device__ __forceinline__ void st_v8(float* ptr, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w),
                 "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// one output element group of the epilogue -------------------------------------------------------
__device__ __forceinline__ float epi_scalar(const Params& p, float val, int64_t m, int64_t n, float* cptr) {
    if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_ACT) val += p.bias[p.bias_on_m ? m : n];
    if (p.epi == EPI_BIAS_ACT) val = act_fwd(val, p.act);
    if (p.epi == EPI_MUL_DACT) val *= act_bwd_from_out(p.aux[m * p.aux_sm + n * p.aux_sn], p.act);
    if (p.accumulate) val += *cptr;
    return val;
}

template <int BN, int BK, int PASSES, bool A_MN, bool B_MN, bool STG = false, bool CTA2 = false, bool TR = false>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmX,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmXh, const Params p) {
    using C_ = Cfg<BN, BK, PASSES, STG, CTA2, TR>;
    constexpr int EPI_NIN = C_::EPI_NIN;
    // CTA pairs: cluster c = blockIdx.x / 2 owns work items c, c + #clusters, ...; a work item covers TWO 128-row blocks
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
    const int64_t w_first = CTA2 ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
    const int64_t w_step = CTA2 ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
    constexpr bool BF3X = C_::BF3X;
    constexpr int NACC = C_::NACC;
    constexpr int STAGES = C_::STAGES;
    constexpr int A_BYTES = C_::A_BYTES;
    constexpr int ROWB = C_::ROW_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi_base = smem + STAGES * C_::STAGE_BYTES;             // staged-epilogue buffers (1024-byte aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + C_::EPI_BYTES);
    // bars[0..S) full (TMA landed), [S..2S) converted, [2S..3S) empty, then NACC acc-full, NACC acc-empty,
    // then (staged epilogue) EPI_NIN "target chunk landed" barriers per epilogue warp
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 384);
    float* tr_s = bias_s + 2 * C_::BIAS_LD;                          // transposing-epilogue scratch (TR_BYTES)
    uint64_t* acc_full = bars + 3 * STAGES;
    uint64_t* acc_empty = acc_full + NACC;
    uint64_t* in_full = acc_empty + NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_full + (STG ? NUM_EPI_WARPS * EPI_NIN : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto stage_ptr = [&](int s) { return smem + (size_t)s * C_::STAGE_BYTES; };
    // tf32 modes: [A hi | A lo | B hi | B lo], each a full fp32 tile.  bf16x3: [A | B], each raw fp32 tile converted in
    // place into [hi bf16 | lo bf16] halves.
    auto a_hi = [&](int s) { return stage_ptr(s); };
    auto a_lo = [&](int s) { return stage_ptr(s) + (BF3X ? A_BYTES / 2 : A_BYTES); };
    auto b_hi = [&](int s) { return stage_ptr(s) + (BF3X ? A_BYTES : 2 * A_BYTES); };
    auto b_lo = [&](int s) { return stage_ptr(s) + (BF3X ? A_BYTES + C_::B_CTA_BYTES / 2 : 2 * A_BYTES + C_::B_BYTES); };
    // work item -> (m block, n block, first K-block, number of K-blocks)
    auto decode = [&](int64_t w, int& m_blk, int& n_blk, int& kb_beg, int& nkb) {
        const int sp = (int)(w % p.splits);
        const int64_t t = w / p.splits;
        n_blk = (int)(t % p.tiles_n);
        m_blk = (int)(t / p.tiles_n);
        if (CTA2) m_blk = 2 * m_blk + (int)rank;
        kb_beg = sp * p.kb_per_split;
        nkb = min(p.kb_total, kb_beg + p.kb_per_split) - kb_beg;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[STAGES + s]), (CTA2 ? 2 : 1) * NUM_CONV_WARPS / (BF3X ? 1 : CONV_GROUPS));
            mbar_init(smem_u32(&bars[2 * STAGES + s]), 1);
        }
        for (int b = 0; b < NACC; ++b) {
            mbar_init(smem_u32(&acc_full[b]), 1);
            mbar_init(smem_u32(&acc_empty[b]), (CTA2 ? 2 : 1) * 32 * NUM_EPI_WARPS);
        }
        if (STG)
            for (int b = 0; b < NUM_EPI_WARPS * EPI_NIN; ++b) mbar_init(smem_u32(&in_full[b]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C_::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C_::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        // (whole warp in the loops, one elected lane issues: see the MMA issuer)
        {
            if (elect_one()) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
            }
            uint32_t it = 0;
            for (int64_t w = w_first; w < p.work_total; w += w_step) {
                int m_blk, n_blk, kb_beg, nkb;
                decode(w, m_blk, n_blk, kb_beg, nkb);
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(smem_u32(&bars[2 * STAGES + s]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bars[s]);
                    if (!elect_one()) continue;
                    if (CDG_PROBE(p) & 2) { mbar_arrive(full); continue; }
                    mbar_expect_tx(full, A_BYTES + C_::B_CTA_BYTES);
                    const int k0 = (kb_beg + i) * BK;
                    // the streamed operand comes from HBM: pull the tile needed PF K-blocks from now into L2
                    constexpr int PF = 2 * STAGES;
                    if (i + PF < nkb && p.conv_cb == 0) {
                        if (!A_MN) tma_prefetch_2d(&tmA, k0 + PF * BK, m_blk * BM);
                        else tma_prefetch_2d(&tmA, m_blk * BM, k0 + PF * BK);
                    }
                    if (!A_MN && p.conv_cb > 0) {
                        // implicit GEMM: K-block -> (tap, channel block); the tile's 128 pixels are a {w, h, b} box
                        const int kb = kb_beg + i, tap = kb / p.conv_cb, c0 = (kb - tap * p.conv_cb) * BK;
                        const int kh = tap / p.conv_k, kw = tap - kh * p.conv_k, pad = (p.conv_k - 1) >> 1;
                        const int64_t m0 = (int64_t)m_blk * BM;
                        const int w0 = (int)(m0 % p.conv_W);
                        const int64_t r = m0 / p.conv_W;
                        tma_load_4d(smem_u32(a_hi(s)), &tmA, full, c0, w0 + kw - pad, (int)(r % p.conv_H) + kh - pad, (int)(r / p.conv_H));
                    } else if (!A_MN) tma_load_2d(smem_u32(a_hi(s)), &tmA, full, k0, m_blk * BM);
                    else tma_load_2d(smem_u32(BF3X ? a_hi(s) : a_lo(s)), &tmA, full, m_blk * BM, k0);
                    if (CTA2) {
                        // this CTA's half of each MMA's B rows (pre-split bf16, rows of 64 B): [N0/2 rows | N1/2 rows]
                        constexpr int H0 = C_::N0 / 2, H1 = C_::N1 / 2;
                        tma_load_2d(smem_u32(b_hi(s)), &tmB, full, k0, n_blk * BN + (int)rank * H0);
                        tma_load_2d(smem_u32(b_lo(s)), &tmB2, full, k0, n_blk * BN + (int)rank * H0);
                        if (H1 > 0) {
                            tma_load_2d(smem_u32(b_hi(s) + H0 * 64), &tmX, full, k0, n_blk * BN + C_::N0 + (int)rank * H1);
                            tma_load_2d(smem_u32(b_lo(s) + H0 * 64), &tmC, full, k0, n_blk * BN + C_::N0 + (int)rank * H1);
                        }
                    } else if (BF3X && p.b_pre) {
                        // ready-made bf16 tiles: rows of 64 B straight into the hi / lo halves
#pragma unroll
                        for (int r = 0; r < BN; r += C_::B_ROWS_PER_BOX) {
                            if (p.sw32) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    tma_load_2d(smem_u32(b_hi(s) + h * BN * 32 + r * 32), &tmB, full, k0 + 16 * h, n_blk * BN + r);
                                    tma_load_2d(smem_u32(b_lo(s) + h * BN * 32 + r * 32), &tmB2, full, k0 + 16 * h, n_blk * BN + r);
                                }
                            } else {
                                tma_load_2d(smem_u32(b_hi(s) + r * 64), &tmB, full, k0, n_blk * BN + r);
                                tma_load_2d(smem_u32(b_lo(s) + r * 64), &tmB2, full, k0, n_blk * BN + r);
                            }
                        }
                    } else if (!B_MN) {
#pragma unroll
                        for (int r = 0; r < BN; r += C_::B_ROWS_PER_BOX)
                            tma_load_2d(smem_u32(b_hi(s) + r * ROWB), &tmB, full, k0, n_blk * BN + r);
                    } else {
#pragma unroll
                        for (int r = 0; r < BN; r += C_::B_CW)
                            tma_load_2d(smem_u32((BF3X ? b_hi(s) : b_lo(s)) + r * ROWB), &tmB, full, n_blk * BN + r, k0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the loops (warp-uniform control flow and operands); one elected lane issues the MMAs.
        // With `if (lane == 0)` around everything the compiler kept descriptors in per-thread registers and wrapped EVERY
        // tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop: ~150 issue cycles per MMA, i.e. the tensor pipe idled
        // more than half of the time (pure-MMA probe: 45 % of its rate; tools/mma_rate.cu).
        if (rank == 0) {
            const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
            constexpr uint32_t idesc0 = make_idesc(BM, C_::N0);
            constexpr uint32_t idesc1 = make_idesc(BM, C_::N1 > 0 ? C_::N1 : 16);
            uint32_t it = 0, j = 0;
            for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
                int m_blk, n_blk, kb_beg, nkb;
                decode(w, m_blk, n_blk, kb_beg, nkb);
                const uint32_t buf = j % NACC;
                mbar_wait(smem_u32(&acc_empty[buf]), ((j / NACC) & 1u) ^ 1u);     // epilogue drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * BN;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(smem_u32(&bars[STAGES + s]), ph);
                    tc_fence_after();
                    if (elect_one()) {
                    if (BF3X) {
                        // bf16 tiles: rows of 64 B (SWIZZLE_64B descriptors), UMMA K = 16 elements = 32 B
                        constexpr uint32_t ib0 = make_idesc_bf16(CTA2 ? 2 * BM : BM, C_::N0);
                        constexpr uint32_t ib1 = make_idesc_bf16(CTA2 ? 2 * BM : BM, C_::N1 > 0 ? C_::N1 : 16);
                        const bool s32 = p.sw32 != 0;
                        const uint64_t dah = s32 ? make_kmajor_desc_sw32(smem_u32(a_hi(s))) : make_kmajor_desc<16>(smem_u32(a_hi(s)));
                        const uint64_t dal = s32 ? make_kmajor_desc_sw32(smem_u32(a_lo(s))) : make_kmajor_desc<16>(smem_u32(a_lo(s)));
                        const uint64_t dbh = s32 ? make_kmajor_desc_sw32(smem_u32(b_hi(s))) : make_kmajor_desc<16>(smem_u32(b_hi(s)));
                        const uint64_t dbl = s32 ? make_kmajor_desc_sw32(smem_u32(b_lo(s))) : make_kmajor_desc<16>(smem_u32(b_lo(s)));
                        // descriptor steps (16-byte units): second K = 16 half of a tile, second MMA's rows (N0 ..)
                        const uint64_t ak1 = (uint64_t)((s32 ? BM * 32 : 32) >> 4), bk1 = (uint64_t)((s32 ? BN * 32 : 32) >> 4);
                        const uint64_t bn1 = (uint64_t)(((CTA2 ? C_::N0 / 2 : C_::N0) * (s32 ? 32 : 64)) >> 4);
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint64_t da = pass == 0 ? dal : dah;
                            const uint64_t db = pass == 1 ? dbl : dbh;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint32_t acc = (i > 0 || pass > 0 || k > 0) ? 1u : 0u;
                                const uint64_t dbk = db + k * bk1;
                                if (CTA2 && C_::A_TMEM) {
                                    const uint32_t ta = tmem_base + C_::A_COL0 + s * C_::A_STAGE_COLS + (pass == 0 ? 16 : 0) + k * 8;
                                    umma_bf16_ts2(tacc, ta, dbk, ib0, acc);
                                    if (C_::N1 > 0) umma_bf16_ts2(tacc + C_::N0, ta, dbk + bn1, ib1, acc);
                                } else if (CTA2) {
                                    umma_bf16_ss2(tacc, da + k * ak1, dbk, ib0, acc);
                                    if (C_::N1 > 0) umma_bf16_ss2(tacc + C_::N0, da + k * ak1, dbk + bn1, ib1, acc);
                                } else if (C_::A_TMEM) {
                                    const uint32_t ta = tmem_base + C_::A_COL0 + s * C_::A_STAGE_COLS + (pass == 0 ? 16 : 0) + k * 8;
                                    umma_bf16_ts(tacc, ta, dbk, ib0, acc);
                                    if (C_::N1 > 0) umma_bf16_ts(tacc + C_::N0, ta, dbk + bn1, ib1, acc);
                                } else {
                                    umma_bf16(tacc, da + k * ak1, dbk, ib0, acc);
                                    if (C_::N1 > 0) umma_bf16(tacc + C_::N0, da + k * ak1, dbk + bn1, ib1, acc);
                                }
                            }
                        }
                    } else {
                    const uint64_t dah = make_kmajor_desc<BK>(smem_u32(a_hi(s))), dal = make_kmajor_desc<BK>(smem_u32(a_lo(s)));
                    const uint64_t dbh = make_kmajor_desc<BK>(smem_u32(b_hi(s))), dbl = make_kmajor_desc<BK>(smem_u32(b_lo(s)));
#pragma unroll
                    for (int pass = 0; pass < PASSES; ++pass) {
                        // small terms first: A_lo*B_hi, A_hi*B_lo, then A_hi*B_hi
                        const uint64_t da = (PASSES == 3 && pass == 0) ? dal : dah;
                        const uint64_t db = (PASSES == 3 && pass == 1) ? dbl : dbh;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint32_t acc = (i > 0 || pass > 0 || k > 0) ? 1u : 0u;
                            const uint64_t koff = (uint64_t)((k * 32) >> 4);
                            if (C_::A_TMEM) {
                                const uint32_t ta = tmem_base + C_::A_COL0 + s * 2 * BK + ((PASSES == 3 && pass == 0) ? BK : 0) + k * 8;
                                umma_tf32_ts(tacc, ta, db + koff, idesc0, acc);
                                if (C_::N1 > 0)
                                    umma_tf32_ts(tacc + C_::N0, ta, db + koff + (uint64_t)((C_::N0 * ROWB) >> 4), idesc1, acc);
                            } else {
                                umma_tf32(tacc, da + koff, db + koff, idesc0, acc);
                                if (C_::N1 > 0)
                                    umma_tf32(tacc + C_::N0, da + koff, db + koff + (uint64_t)((C_::N0 * ROWB) >> 4), idesc1, acc);
                            }
                        }
                    }
                    }
                    if (CTA2) umma_commit2(smem_u32(&bars[2 * STAGES + s]));
                    else umma_commit(smem_u32(&bars[2 * STAGES + s]));  // frees the stage when these MMAs retire
                    if (i == nkb - 1) {
                        if (CTA2) umma_commit2(smem_u32(&acc_full[buf]));
                        else umma_commit(smem_u32(&acc_full[buf]));     // accumulator of this work item complete
                    }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 2 + NUM_CONV_WARPS) {
        // ================= converter =================
        // Two groups of four warps alternate ring stages: a stage's conversion is a latency chain (shared-memory read,
        // tcgen05.st + wait, proxy fence, arrive), so two stages in flight hide most of it.  Each group covers all four
        // TMEM lane quarters (warps 2-5 and 6-9: warp % 4 = 2,3,0,1).
        constexpr int NGRP = BF3X ? 1 : CONV_GROUPS;
        constexpr int GT = 32 * NUM_CONV_WARPS / NGRP;              // threads per group
        const int grp = NGRP == 1 ? 0 : (warp - 2) / (NUM_CONV_WARPS / NGRP);
        const int ct = (threadIdx.x - 64) - grp * GT;
        uint32_t it = 0;
        for (int64_t w = w_first; w < p.work_total; w += w_step) {
            int m_blk, n_blk, kb_beg, nkb;
            decode(w, m_blk, n_blk, kb_beg, nkb);
            for (int i = 0; i < nkb; ++i, ++it) {
                if (NGRP == 2 && (int)(it & 1u) != grp) continue;
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(smem_u32(&bars[s]), ph);
                if (CDG_PROBE(p) & 1) {
                    __syncwarp();
                    if (lane == 0) {
                        if (CTA2 && rank != 0) mbar_arrive_rank0(smem_u32(&bars[STAGES + s]));
                        else mbar_arrive(smem_u32(&bars[STAGES + s]));
                    }
                    continue;
                }
                if (BF3X) {
                    // all converter warps share a stage (the in-place conversion needs every read done before any write)
                    constexpr int NTA = 32 * NUM_CONV_WARPS;
                    const int cta = threadIdx.x - 64;
                    if (C_::A_TMEM) bf3x_a_tmem<A_MN>(a_hi(s), tmem_base + C_::A_COL0 + s * C_::A_STAGE_COLS, warp, lane, p.a16swap);
                    else if (!A_MN) bf3x_kmajor<BM, NTA>(a_hi(s), cta, 1, p.sw32);
                    else bf3x_mnmajor<BM, 128, NTA>(a_hi(s), cta, 1, p.sw32);
                    if (p.b_pre) {
                    } else if (!B_MN) bf3x_kmajor<BN, NTA>(b_hi(s), cta, 2, p.sw32);
                    else bf3x_mnmajor<BN, C_::B_CW, NTA>(b_hi(s), cta, 2, p.sw32);
                } else if (C_::A_TMEM) {
                    const uint32_t ta = tmem_base + C_::A_COL0 + s * 2 * BK;
                    if (NGRP == 2) {
                        convert_a_tmem<PASSES, BK, A_MN>(a_hi(s), a_lo(s), ta, warp, lane, 0);
                        convert_a_tmem<PASSES, BK, A_MN>(a_hi(s), a_lo(s), ta, warp, lane, 1);
                    } else {
                        convert_a_tmem<PASSES, BK, A_MN>(a_hi(s), a_lo(s), ta, warp, lane, (warp - 2) >> 2);
                    }
                } else {
                    if (!A_MN) convert_kmajor<PASSES, GT>(a_hi(s), a_lo(s), A_BYTES, ct, p.rawhi);
                    else convert_mnmajor<PASSES, BM, 128, BK, GT>(a_hi(s), a_lo(s), ct, 1 + grp);
                }
                if (BF3X) {
                } else if (!B_MN) convert_kmajor<PASSES, GT>(b_hi(s), b_lo(s), C_::B_BYTES, ct, p.rawhi);
                else convert_mnmajor<PASSES, BN, C_::B_CW, BK, GT>(b_hi(s), b_lo(s), ct, 3 + grp);
                tc_fence_before();                                   // tcgen05.st (A ring) ordered before the arrive
                // generic-proxy writes -> async proxy (UMMA).  Not needed when this stage's conversion wrote no shared memory
                // at all (bf16x3 with the A operand in tensor memory and a pre-split B): the proxy fence waits out every
                // outstanding shared-memory access of the warp, a few hundred cycles in front of each hand-off to the MMA warp
                if (!(BF3X && C_::A_TMEM && p.b_pre && !CDG_SKIP_FENCE_OFF(p))) fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    // CTA pairs: the MMA thread of rank 0 waits for both CTAs' tiles (landed by TMA and converted)
                    if (CTA2 && rank != 0) mbar_arrive_rank0(smem_u32(&bars[STAGES + s]));
                    else mbar_arrive(smem_u32(&bars[STAGES + s]));
                }
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                                      // TMEM lane quarter this warp may access
        if constexpr (STG) {
            // ---- staged reconstruction head --------------------------------------------------------------------------
            // The row-per-thread epilogue reads the target and writes the gradient as 32-byte pieces of 32 different rows
            // per instruction (measured: 2x the HBM time of those bytes).  Here every warp streams its 32 rows x 32 columns
            // chunks through shared memory with TMA: target chunks arrive EPI_NIN deep (running ahead across tiles),
            // gradient (and xhat) chunks leave through EPI_NOUT buffers as bulk stores; the threads only touch shared
            // memory (swizzled, conflict-free) and tensor memory.
            constexpr int CPT = BN / 32;                             // chunks per tile
            const int ew = warp - (2 + NUM_CONV_WARPS);
            uint8_t* ebuf = epi_base + (size_t)ew * (EPI_NIN + EPI_NOUT) * EPI_CHUNK_BYTES;
            uint64_t* infull = in_full + ew * EPI_NIN;
            auto nvalid = [&](int n_blk) {
                const int64_t nv = (p.N - (int64_t)n_blk * BN + 31) / 32;
                return nv < CPT ? (int)nv : CPT;
            };
            // producer side (lane 0): the next target chunk in this CTA's (tile, chunk) order
            int64_t pw = w_first;
            int pc = 0;
            uint32_t pcount = 0;
            auto issue_next = [&]() {
                if (pw >= p.work_total) return;
                int m_blk, n_blk, kb_beg, nkb;
                decode(pw, m_blk, n_blk, kb_beg, nkb);
                const uint32_t b = pcount % EPI_NIN;
                const uint32_t bar = smem_u32(&infull[b]);
                mbar_expect_tx(bar, EPI_CHUNK_BYTES);
                tma_load_2d(smem_u32(ebuf + b * EPI_CHUNK_BYTES), &tmX, bar, n_blk * BN + pc * 32, m_blk * BM + q * 32);
                ++pcount;
                if (++pc >= nvalid(n_blk)) { pc = 0; pw += w_step; }
            };
            if (lane == 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
#pragma unroll
                for (int i = 0; i < EPI_NIN; ++i) issue_next();
            }
            uint32_t j = 0, ccount = 0, ocount = 0;
            float rloss = 0.f;
            auto put_chunk = [&](const CUtensorMap* tm, const float* o, int col0, int row0) {
                uint8_t* xo = ebuf + (size_t)(EPI_NIN + (ocount % EPI_NOUT)) * EPI_CHUNK_BYTES;
                ++ocount;
                // the bulk store that used this buffer EPI_NOUT chunks ago has read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(EPI_NOUT - 1) : "memory");
                __syncwarp();
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<float4*>(xo + sw_chunk<32>((uint32_t)lane, (uint32_t)ch)) =
                        make_float4(o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                     reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(xo)), "r"(col0), "r"(row0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            };
            for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
                int m_blk, n_blk, kb_beg, nkb;
                decode(w, m_blk, n_blk, kb_beg, nkb);
                const uint32_t buf = j % NACC;
                mbar_wait(smem_u32(&acc_full[buf]), (j / NACC) & 1u);
                tc_fence_after();
                const int row0 = m_blk * BM + q * 32;
                const bool row_ok = (int64_t)row0 + lane < p.M;
                const uint32_t trow = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
                const int nv = nvalid(n_blk);
#pragma unroll 1
                for (int c = 0; c < nv; ++c) {
                    float v[32], xs[32];
                    tmem_ld16(trow + (uint32_t)(c * 32), v);
                    tmem_ld16(trow + (uint32_t)(c * 32 + 16), v + 16);
                    if (c == nv - 1) {                               // accumulator drained: the MMA warp may reuse the buffer
                        tc_fence_before();
                        if (CTA2 && rank != 0) mbar_arrive_rank0(smem_u32(&acc_empty[buf]));
                        else mbar_arrive(smem_u32(&acc_empty[buf]));
                    }
                    const uint32_t b = ccount % EPI_NIN;
                    mbar_wait(smem_u32(&infull[b]), (ccount / EPI_NIN) & 1u);
                    ++ccount;
                    const uint8_t* xin = ebuf + b * EPI_CHUNK_BYTES;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 t = *reinterpret_cast<const float4*>(xin + sw_chunk<32>((uint32_t)lane, (uint32_t)ch));
                        xs[4 * ch] = t.x; xs[4 * ch + 1] = t.y; xs[4 * ch + 2] = t.z; xs[4 * ch + 3] = t.w;
                    }
                    const int col0 = n_blk * BN + c * 32;
                    const bool full = (int64_t)col0 + 32 <= p.N;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        float bb[4];
                        if (full) {
                            const float4 t = *reinterpret_cast<const float4*>(p.bias + col0 + 4 * g);
                            bb[0] = t.x; bb[1] = t.y; bb[2] = t.z; bb[3] = t.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) bb[e] = (int64_t)col0 + 4 * g + e < p.N ? p.bias[col0 + 4 * g + e] : 0.f;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = 4 * g + e;
                            const float t = tanh_fast(v[i] + bb[e]);
                            const float df = t - xs[i];
                            const bool ok = row_ok && (full || (int64_t)col0 + i < p.N);
                            rloss += ok ? 0.5f * df * df : 0.f;
                            xs[i] = df * (1.f - t * t) * p.inv_batch;
                            v[i] = t;
                        }
                    }
                    put_chunk(&tmC, xs, col0, row0);
                    if (p.rxhat) put_chunk(&tmXh, v, col0, row0);
                    if (lane == 0) issue_next();                     // refill the target buffer (its values have all been used)
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            if (p.racc) {
                const float s = warp_sum(rloss);
                if (lane == 0) atomicAdd(p.racc, (double)s);
            }
        } else {
        // column range of this warp: the whole tile (4 epilogue warps) or one half (8)
        constexpr int CSPLIT = ((BN / 2 + 15) / 16) * 16;
        const int ehalf = NUM_EPI_WARPS == 8 ? (warp - (2 + NUM_CONV_WARPS)) >> 2 : 0;
        const int c_beg = ehalf == 0 ? 0 : CSPLIT;
        const int c_end = (NUM_EPI_WARPS == 8 && ehalf == 0) ? CSPLIT : BN;
        uint32_t j = 0;
        float rloss = 0.f;
        for (int64_t w = w_first; w < p.work_total; w += w_step, ++j) {
            int m_blk, n_blk, kb_beg, nkb;
            decode(w, m_blk, n_blk, kb_beg, nkb);
            const uint32_t buf = j % NACC;
            const int64_t m = (int64_t)m_blk * BM + q * 32 + lane;
            const uint32_t trow = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
            const bool row_ok = m < p.M;
            // side operand of the vector path, read one 16-column chunk ahead of its use so that its DRAM
            // latency overlaps the previous chunk's math (one warp per SM sub-partition has no other cover)
            const float* side = nullptr;
            if (p.vec && !p.atomic && row_ok) {
                if (p.epi == EPI_RECON) side = p.rx + m * p.rx_ld + (int64_t)n_blk * BN;
                else if (p.epi == EPI_MUL_DACT) side = p.aux + m * p.aux_sm + (int64_t)n_blk * BN;
            }
            // Fused reconstruction head on full 256-wide tiles: TRANSPOSING epilogue.  In the accumulator's natural layout a
            // thread owns one output row, so the target reads and gradient stores of a warp touch 32 rows x 32 bytes per
            // instruction (49 KB apart): measured 2x their HBM time.  Here each 32 x 32 accumulator chunk goes through a
            // conflict-free shared-memory scratch (pitch 33) and comes back with lane = column, so every global access of the
            // warp is one contiguous 128-byte row segment; tanh / loss / gradient are element-wise and do not care.
            const bool tr_tile = C_::TR_BYTES > 0 && p.tr && p.epi == EPI_RECON && p.vec && !p.atomic && !p.extra &&
                                 (int64_t)(n_blk + 1) * BN <= p.N && (int64_t)m_blk * BM + q * 32 + 32 <= p.M &&
                                 p.rx_ld < (1 << 26) && p.sc_m < (1 << 26);
            float xs[C_::TR_BYTES > 0 ? 32 : 1];
            const int64_t trow0 = (int64_t)m_blk * BM + q * 32;
            const float* xg = p.rx + trow0 * p.rx_ld + (int64_t)n_blk * BN + lane;
            if (tr_tile) {
                side = nullptr;
#pragma unroll
                for (int i = 0; i < 32; ++i) xs[i] = __ldg(xg + (int64_t)i * p.rx_ld);   // first chunk's targets: before the accumulator wait
            }
            float4 r0[4], r1[4], r2[4];                                // chunks c0, c0+16, c0+32 in flight
            auto fetch = [&](int c0, float4* dst) {
                const int64_t n0 = (int64_t)n_blk * BN + c0;
                if (side && c0 < c_end && n0 + 16 <= p.N) {
                    if (p.vec8) {
                        ld_nc_v8(side + c0, dst[0], dst[1]);
                        ld_nc_v8(side + c0 + 8, dst[2], dst[3]);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) dst[g] = __ldg(reinterpret_cast<const float4*>(side + c0 + 4 * g));
                    }
                }
            };
            // everything that does not need the accumulator goes BEFORE the wait for it: the first target chunks and the
            // tile's bias values (staged in shared memory: per-chunk global loads of the bias were an exposed L1/L2 latency
            // in front of every tanh, ncu source view: 28 % of the epilogue warps' samples)
            fetch(c_beg, r0); fetch(c_beg + 16, r1); fetch(c_beg + 32, r2);
            const float* bs = bias_s + (j & 1u) * C_::BIAS_LD;
            const bool bias_tile = !STG && !p.bias_on_m && (p.epi == EPI_BIAS || p.epi == EPI_BIAS_ACT || p.epi == EPI_RECON);
            if (bias_tile) {
                const int et = threadIdx.x - 32 * (2 + NUM_CONV_WARPS);
                for (int c = et; c < BN; c += 32 * NUM_EPI_WARPS) {
                    const int64_t n = (int64_t)n_blk * BN + c;
                    bias_s[(j & 1u) * C_::BIAS_LD + c] = n < p.N ? p.bias[n] : 0.f;
                }
                asm volatile("bar.sync 5, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
            }
            mbar_wait(smem_u32(&acc_full[buf]), (j / NACC) & 1u);
            tc_fence_after();
            if (C_::TR_BYTES > 0 && tr_tile) {
                float* sc = tr_s + (warp - (2 + NUM_CONV_WARPS)) * (32 * C_::TR_LD);
                float* cg = p.C + trow0 * p.sc_m + (int64_t)n_blk * BN + lane;
                float* hg = p.rxhat ? p.rxhat + trow0 * p.sc_m + (int64_t)n_blk * BN + lane : nullptr;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                    // 16 columns at a time: 16, not 32, accumulator registers live
                        float v[16];
                        tmem_ld16(trow + (uint32_t)(c * 32 + 16 * h), v);
                        if (h == 1 && c == BN / 32 - 1) {            // accumulator drained: the MMA warp may reuse the buffer
                            tc_fence_before();
                            if (CTA2 && rank != 0) mbar_arrive_rank0(smem_u32(&acc_empty[buf]));
                            else mbar_arrive(smem_u32(&acc_empty[buf]));
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) sc[lane * C_::TR_LD + 16 * h + i] = v[i];
                    }
                    __syncwarp();
                    const float bb = bs[c * 32 + lane];
                    const bool more = c + 1 < BN / 32;
                    // row strides re-read per chunk behind an empty asm: otherwise the 32 + 32 row offsets are hoisted out of
                    // the chunk loop as loop invariants and spill (ptxas: 308 bytes of spill stores at the 128-register cap)
                    uint32_t ldx = (uint32_t)p.rx_ld, ldc = (uint32_t)p.sc_m;
                    asm volatile("" : "+r"(ldx), "+r"(ldc));
                    const float* xr = xg + (c + 1) * 32;
                    float* cr = cg + c * 32;
                    float* hr = hg ? hg + c * 32 : nullptr;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float t = tanh_fast(sc[i * C_::TR_LD + lane] + bb);
                        const float df = t - xs[i];
                        // the register is refilled at once with the next chunk's target of the same row: a full chunk of
                        // work (32 rows) covers its DRAM latency, and no second array of targets is live
                        if (more) xs[i] = __ldg(xr + (uint32_t)i * ldx);
                        rloss += 0.5f * df * df;
                        cr[(uint32_t)i * ldc] = df * (1.f - t * t) * p.inv_batch;
                        if (hr) hr[(uint32_t)i * ldc] = t;
                    }
                    __syncwarp();
                }
                continue;
            }
            uint32_t vraw[16];
            tmem_ld16_issue(trow + (uint32_t)c_beg, vraw);           // accumulator chunks are read one ahead of their use
#pragma unroll 1
            for (int c0 = c_beg; c0 < c_end; c0 += 16) {
                float v[16];
                tmem_ld16_wait(vraw);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(vraw[i]);
                if (c0 + 16 < c_end) tmem_ld16_issue(trow + (uint32_t)(c0 + 16), vraw);
                float4 cur[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) { cur[g] = r0[g]; r0[g] = r1[g]; r1[g] = r2[g]; }
                fetch(c0 + 48, r2);
                const int64_t n0 = (int64_t)n_blk * BN + c0;
                if (!row_ok || n0 >= p.N) continue;
                if (p.vec && !p.atomic && n0 + 16 <= p.N && !(p.extra && n0 + 16 >= p.N)) {
                    float* crow = p.C + m * p.sc_m + n0;
                    // results leave in halves of 8 columns (one 256-bit store each): keeps two float4 of outputs live, not four
                    if (p.epi == EPI_RECON) {
                        float* xh = p.rxhat ? p.rxhat + m * p.sc_m + n0 : nullptr;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float4 og[2], tg[2];
#pragma unroll
                            for (int gg = 0; gg < 2; ++gg) {
                                const int g = 2 * h + gg;
                                const float4 bb = *reinterpret_cast<const float4*>(bs + c0 + 4 * g);
                                const float4 xx = cur[g];
                                float t[4] = {tanh_fast(v[4 * g] + bb.x), tanh_fast(v[4 * g + 1] + bb.y), tanh_fast(v[4 * g + 2] + bb.z),
                                              tanh_fast(v[4 * g + 3] + bb.w)};
                                const float xs[4] = {xx.x, xx.y, xx.z, xx.w};
                                float o[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float df = t[e] - xs[e];
                                    rloss += 0.5f * df * df;
                                    o[e] = df * (1.f - t[e] * t[e]) * p.inv_batch;
                                }
                                og[gg] = make_float4(o[0], o[1], o[2], o[3]);
                                tg[gg] = make_float4(t[0], t[1], t[2], t[3]);
                            }
                            if (p.vec8) {
                                st_v8(crow + 8 * h, og[0], og[1]);
                                if (xh) st_v8(xh + 8 * h, tg[0], tg[1]);
                            } else {
#pragma unroll
                                for (int gg = 0; gg < 2; ++gg) {
                                    *reinterpret_cast<float4*>(crow + 8 * h + 4 * gg) = og[gg];
                                    if (xh) *reinterpret_cast<float4*>(xh + 8 * h + 4 * gg) = tg[gg];
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float4 og[2];
#pragma unroll
                            for (int gg = 0; gg < 2; ++gg) {
                                const int g = 2 * h + gg;
                                float o[4] = {v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]};
                                if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_ACT) {
                                    if (p.bias_on_m) {
                                        const float bm = p.bias[m];
                                        o[0] += bm; o[1] += bm; o[2] += bm; o[3] += bm;
                                    } else {
                                        const float4 bb = *reinterpret_cast<const float4*>(bs + c0 + 4 * g);
                                        o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w;
                                    }
                                }
                                if (p.epi == EPI_BIAS_ACT) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) o[e] = act_fwd(o[e], p.act);
                                }
                                if (p.epi == EPI_MUL_DACT) {
                                    const float4 hh = cur[g];
                                    o[0] *= act_bwd_from_out(hh.x, p.act); o[1] *= act_bwd_from_out(hh.y, p.act);
                                    o[2] *= act_bwd_from_out(hh.z, p.act); o[3] *= act_bwd_from_out(hh.w, p.act);
                                }
                                if (p.accumulate) {
                                    const float4 cc = *reinterpret_cast<const float4*>(crow + 4 * g);
                                    o[0] += cc.x; o[1] += cc.y; o[2] += cc.z; o[3] += cc.w;
                                }
                                og[gg] = make_float4(o[0], o[1], o[2], o[3]);
                            }
                            if (p.vec8) {
                                st_v8(crow + 8 * h, og[0], og[1]);
                            } else {
#pragma unroll
                                for (int gg = 0; gg < 2; ++gg) *reinterpret_cast<float4*>(crow + 8 * h + 4 * gg) = og[gg];
                            }
                        }
                    }
                } else if (p.accumulate && !p.atomic && p.epi == EPI_NONE && !p.extra) {
                    // accumulating stores (weight gradients written in the swapped orientation at small batches: 160
                    // read-modify-writes per thread): four old values are fetched before the first of their stores -- one at a
                    // time every load has to stay behind the previous store (same pointer type), which serialised 160 DRAM round
                    // trips per tile (207 us for the 14.7 MB encoder weight gradient at batch 128)
#pragma unroll
                    for (int e4 = 0; e4 < 16; e4 += 4) {
                        float old[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) old[e] = n0 + e4 + e < p.N ? p.C[m * p.sc_m + (n0 + e4 + e) * p.sc_n] : 0.f;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n0 + e4 + e < p.N) p.C[m * p.sc_m + (n0 + e4 + e) * p.sc_n] = v[e4 + e] + old[e];
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int64_t n = n0 + e;
                        if (n < p.N) {
                            float* cptr = p.C + m * p.sc_m + n * p.sc_n;
                            if (p.extra && n == p.N - 1) cptr = p.extra + m;
                            if (p.atomic) {
                                atomicAdd(cptr, v[e]);
                            } else if (p.epi == EPI_RECON) {
                                const float t = tanh_fast(v[e] + p.bias[n]);
                                const float df = t - p.rx[m * p.rx_ld + n];
                                rloss += 0.5f * df * df;
                                *cptr = df * (1.f - t * t) * p.inv_batch;
                                if (p.rxhat) p.rxhat[m * p.sc_m + n * p.sc_n] = t;
                            } else {
                                *cptr = epi_scalar(p, v[e], m, n, cptr);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (CTA2 && rank != 0) mbar_arrive_rank0(smem_u32(&acc_empty[buf]));
            else mbar_arrive(smem_u32(&acc_empty[buf]));             // TMEM buffer may be overwritten
        }
        if (p.epi == EPI_RECON && p.racc) {
            const float s = warp_sum(rloss);
            if (lane == 0) atomicAdd(p.racc, (double)s);
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();      // both CTAs are done with each other's shared / tensor memory
    if (warp == 1) {
        tc_fence_after();
        if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_::TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_::TMEM_COLS));
    }
}

// ---- host side ----------------------------------------------------------------------------------
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

struct MapKey {
    const void* ptr; int64_t rows, k, row_stride, k_stride; int box0, box1, mn;   // box0/box1 include BK
    bool operator<(const MapKey& o) const {
        return std::tie(ptr, rows, k, row_stride, k_stride, box0, box1, mn) <
               std::tie(o.ptr, o.rows, o.k, o.row_stride, o.k_stride, o.box0, o.box1, o.mn);
    }
};
static std::map<MapKey, CUtensorMap> g_maps;
static std::mutex g_maps_mu;

// Operand X(r,k) = X[r*s_r + k*s_k], r < rows, k < K.  K-major (s_k == 1): dims {K, rows}, box {32, box_rows},
// SWIZZLE_128B.  MN-major (s_r == 1): dims {rows, K}, box {cw, 32}, no swizzle.
static int make_map(const float* X, int64_t rows, int64_t K, int64_t s_r, int64_t s_k, bool mn_major, int box_rows,
                    int BK, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    MapKey key{X, rows, K, s_r, s_k, mn_major ? box_rows : BK, mn_major ? BK : box_rows, mn_major ? 1 : 0};
    {
        std::lock_guard<std::mutex> g(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    cuuint64_t dims[2], strides[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    CUtensorMapSwizzle sw;
    if (!mn_major) {
        dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; strides[0] = (cuuint64_t)s_r * 4;
        box[0] = (cuuint32_t)BK; box[1] = (cuuint32_t)box_rows; sw = BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    } else {
        dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; strides[0] = (cuuint64_t)s_k * 4;
        box[0] = (cuuint32_t)box_rows; box[1] = (cuuint32_t)BK; sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}

// NHWC activation [B, H, W, C] as a 4-D map {C, W, H, B}; box = BK channels x (Wt x Ht x Bt = 128 pixels in output-row order)
static int make_conv_map(const float* X, int64_t B, int H, int W, int C, int BK, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    MapKey key{X, B, (int64_t)H * 100000 + W, C, -4, BK, 128, 2};
    {
        std::lock_guard<std::mutex> g(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    const int Wt = W < 128 ? W : 128, Ht = (128 / Wt) < H ? 128 / Wt : H, Bt = 128 / (Wt * Ht);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)Wt, (cuuint32_t)Ht, (cuuint32_t)Bt}, estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (4-D) failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}
// pre-split bf16 operand [rows][K] (row stride ld16 elements): box {32 bf16 = 64 B, box_rows}, SWIZZLE_64B
static int make_map_bf16(const void* X, int64_t rows, int64_t K, int64_t ld16, int box_rows, int sw32, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CDG_ERR_CUDA; }
    MapKey key{X, rows, K, ld16, -16, sw32 ? 16 : 32, box_rows, 3};
    {
        std::lock_guard<std::mutex> g(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return CDG_OK; }
    }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld16 * 2};
    cuuint32_t box[2] = {sw32 ? 16u : 32u, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r); return CDG_ERR_CUDA; }
    std::lock_guard<std::mutex> g(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return CDG_OK;
}
// spatial extents whose 128-pixel tiles are boxes: powers of two (W >= 128: multiples of 128)
static bool conv_ok(const GemmDesc& g, int BK) {
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (g.conv_C % BK != 0 || ((uintptr_t)g.A & 15) != 0 || g.conv_k < 1 || g.conv_k % 2 == 0) return false;
    if (g.conv_W >= 128) { if (g.conv_W % 128 != 0) return false; }
    else if (!pow2(g.conv_W) || !pow2(g.conv_H)) return false;
    return g.M == g.conv_B * g.conv_H * g.conv_W && g.K == (int64_t)g.conv_k * g.conv_k * g.conv_C;
}

// an operand is usable when it is K-major or MN-major with TMA-legal strides and alignment
static bool operand_ok(const float* X, int64_t s_r, int64_t s_k, int64_t rows, int64_t K, bool* mn_major) {
    if (((uintptr_t)X & 15) != 0) return false;
    if (s_k == 1 && (s_r % 4 == 0) && s_r >= K) { *mn_major = false; return true; }
    if (s_r == 1 && (s_k % 4 == 0) && s_k >= rows) { *mn_major = true; return true; }
    return false;
}

template <int BN, int BK, int PASSES, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2, const Params& p, dim3 grid, cudaStream_t s) {
    auto kern = gemm_tc_kernel<BN, BK, PASSES, A_MN, B_MN>;
    static bool attr_done = false;
    if (!attr_done) {
        CDG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, BK, PASSES>::SMEM));
        attr_done = true;
    }
    kern<<<grid, THREADS, Cfg<BN, BK, PASSES>::SMEM, s>>>(ta, tb, tb2, ta, ta, ta, p);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

// CTA-pair launch: clusters of 2, pre-split bf16 B (tb / tb2: first MMA's half rows, tb3 / tb4: second MMA's), A converted in the loop
template <int BN, bool A_MN, bool STG = false, bool TR = false>
static int launch_cta2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2, const CUtensorMap& tb3,
                       const CUtensorMap& tb4, const Params& p, unsigned clusters, cudaStream_t s, const CUtensorMap* txh = nullptr) {
    auto kern = gemm_tc_kernel<BN, 32, 2, A_MN, false, STG, true, TR>;
    using C_ = Cfg<BN, 32, 2, STG, true, TR>;
    static bool attr_done = false;
    if (!attr_done) {
        CDG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = C_::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CDG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tb2, tb3, tb4, txh ? *txh : ta, p));
    ++g_launches;
    return CDG_OK;
}

// the fused reconstruction head with the staged (TMA in / TMA out) epilogue: 256-wide bf16x3 tile, K-major operands
static int launch_recon_staged(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2, const CUtensorMap& tx,
                               const CUtensorMap& tc_, const CUtensorMap& txh, const Params& p, dim3 grid, cudaStream_t s) {
    auto kern = gemm_tc_kernel<256, 32, 2, false, false, true>;
    using C_ = Cfg<256, 32, 2, true>;
    static bool attr_done = false;
    if (!attr_done) {
        CDG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
        attr_done = true;
    }
    kern<<<grid, THREADS, C_::SMEM, s>>>(ta, tb, tb2, tx, tc_, txh, p);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}

template <int BN, int BK, int PASSES>
static int launch_layout(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2,
                         const Params& p, dim3 grid, cudaStream_t s) {
    if (!a_mn && !b_mn) return launch<BN, BK, PASSES, false, false>(ta, tb, tb2, p, grid, s);
    if (!a_mn && b_mn) return launch<BN, BK, PASSES, false, true>(ta, tb, tb2, p, grid, s);
    if (a_mn && !b_mn) return launch<BN, BK, PASSES, true, false>(ta, tb, tb2, p, grid, s);
    return launch<BN, BK, PASSES, true, true>(ta, tb, tb2, p, grid, s);
}

}  // namespace tc

struct Plan {
    GemmDesc g;
    int64_t sc_m, sc_n, aux_sm, aux_sn;
    int bias_on_m, BN, splits, kb_total, kb_per;
    int64_t tm, tn;
    bool a_mn, b_mn;
};

static bool no160() {
    static const int v = exp_switch("CDG_TC_NO160", 0);
    return v != 0;
}
// shape / layout analysis shared by gemm_tc and gemm_tc_can
static thread_local int g_sm_budget = kNumSMs;
void gemm_tc_set_sm_budget(int sms) { g_sm_budget = sms > 0 && sms < kNumSMs ? sms : kNumSMs; }

static int plan_gemm(const GemmDesc& g0, int BK, Plan* pl) {
    using namespace tc;
    GemmDesc g = g0;
    int64_t sc_m = g.ldc, sc_n = 1, aux_sm = g.ld_aux, aux_sn = 1;
    int bias_on_m = 0;
    // Tensor-core tiles need real extents: tiny N (encoder head 2d = 8) / tiny K (decoder inputs 1-2) stay on SIMT.
    if (g.K < 16 || (g.M < 32 && g.N < 32) || g.M * g.N < 4096) return CDG_ERR_UNSUPPORTED;
    // put the long output dimension on M (128-row tiles) when the other fits one N tile
    if (g.conv_C > 0 && !conv_ok(g, BK)) return CDG_ERR_UNSUPPORTED;
    if (g.N > 304 && g.M <= 304 && g.epi != EPI_RECON && g.conv_C == 0 && !g.b_hi16) {
        std::swap(g.A, g.B); std::swap(g.sa_m, g.sb_n); std::swap(g.sa_k, g.sb_k); std::swap(g.M, g.N);
        std::swap(sc_m, sc_n); std::swap(aux_sm, aux_sn);
        std::swap(g.a_hi16, g.b_hi16); std::swap(g.a_lo16, g.b_lo16); std::swap(g.ld_a16, g.ld_b16);
        bias_on_m = 1;
    }
    if (g.a_hi16) return CDG_ERR_UNSUPPORTED;       // a pre-split A is only usable once swapped into the B slot
    if (g.extra_col && (bias_on_m || g.epi != EPI_NONE)) return CDG_ERR_UNSUPPORTED;   // plain (accumulating) results only
    if (g.N < 16 && g.conv_C == 0) return CDG_ERR_UNSUPPORTED;   // (conv mode: a 3-channel toRGB still beats im2col + rowdot)
    bool a_mn, b_mn;
    if (g.conv_C > 0) a_mn = false;
    else if (!operand_ok(g.A, g.sa_m, g.sa_k, g.M, g.K, &a_mn)) return CDG_ERR_UNSUPPORTED;
    if (g.b_hi16) {
        if (BK != 32 || g.ld_b16 % 8 != 0 || (((uintptr_t)g.b_hi16 | (uintptr_t)g.b_lo16) & 15) != 0) return CDG_ERR_UNSUPPORTED;
        b_mn = false;
    } else if (!operand_ok(g.B, g.sb_n, g.sb_k, g.N, g.K, &b_mn)) return CDG_ERR_UNSUPPORTED;
    int BN;
    if (g.N <= 64 && BK == 32) BN = 64;
    else if (g.N <= 128) BN = 128;
    else if (g.N <= 160) BN = 160;
    else if (g.N <= 256) BN = 256;
    else if (g.N <= 304) BN = (g.K <= 1024 && g.M >= 128 * 64 && !no160()) ? 160 : 304;   // short K: two 160-wide tiles with
    else BN = 256;                                                            // double-buffered accumulators
    const int64_t tm = (g.M + BM - 1) / BM, tn = (g.N + BN - 1) / BN;
    if (tn > (1 << 30) || tm > (1 << 30)) return CDG_ERR_UNSUPPORTED;
    const int kb_total = (int)((g.K + BK - 1) / BK);
    // split-K, for two reasons:
    //  (1) accuracy: the tensor core accumulates in fp32 with truncation, whose error grows linearly with the
    //      number of accumulations (measured on B200: ~3.5e-9 * K relative).  Capping one TMEM accumulation at
    //      KB_CAP K-blocks (K = 2048) keeps a partial sum at ~7e-6; partials are combined with fp32
    //      round-to-nearest adds (red.global.add.f32).
    //  (2) occupancy: fill the chip when the tile grid is small.
    const int KB_CAP = 2048 / BK;
    int splits = (kb_total + KB_CAP - 1) / KB_CAP;
    const int64_t tiles = tm * tn;
    const int64_t sms = g_sm_budget;       // 148, or the share of the chip a caller running several chains side by side wants per GEMM
    if (tiles * splits < sms && kb_total >= 8 && g.epi != EPI_RECON) {
        int want = (int)imin64((sms + tiles - 1) / tiles, kb_total / 4);
        splits = (int)imax64(splits, imin64(want, 64));
    }
    if (splits < 1) splits = 1;
    int kb_per = (kb_total + splits - 1) / splits;
    splits = (kb_total + kb_per - 1) / kb_per;
    if (g.epi == EPI_RECON && splits > 1) return CDG_ERR_UNSUPPORTED;
    pl->g = g; pl->sc_m = sc_m; pl->sc_n = sc_n; pl->aux_sm = aux_sm; pl->aux_sn = aux_sn; pl->bias_on_m = bias_on_m;
    pl->BN = BN; pl->splits = splits; pl->kb_total = kb_total; pl->kb_per = kb_per; pl->tm = tm; pl->tn = tn;
    pl->a_mn = a_mn; pl->b_mn = b_mn;
    return CDG_OK;
}

// K-block of the 3xTF32 kernels: 16 (SWIZZLE_64B, 4 stages) unless CDG_TC_BK=32 asks for the 2-stage variant
static int default_bk() {
    static const int bk = exp_switch("CDG_TC_BK", 16) == 32 ? 32 : 16;
    return bk;
}

bool gemm_tc_can(const GemmDesc& g) {
    Plan pl;
    return plan_gemm(g, default_bk(), &pl) == CDG_OK;
}

static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int gemm_tc(const GemmDesc& g0, int passes, void*, int64_t, cudaStream_t s) {
    using namespace tc;
    Plan pl;
    // narrow outputs (N <= 64) run the 64-wide tile with 32-float K-blocks: the main loop of this kernel has a fixed cost per
    // K-block (TMA -> convert -> MMA hand-offs), so halving the block count matters more than ring depth there
    const bool narrow = g0.N <= 64 && passes != 1 && g0.K >= 32 && (g0.conv_C == 0 || g0.conv_C % 32 == 0);
    const int BK = (passes == 1 || passes == 2 || narrow) ? 32 : default_bk();
    const int pr = plan_gemm(g0, BK, &pl);
    if (pr != CDG_OK) return pr;
    const GemmDesc& g = pl.g;
    const int BN = pl.BN, splits = pl.splits;

    Params p;
    p.C = g.C; p.sc_m = pl.sc_m; p.sc_n = pl.sc_n; p.M = g.M; p.N = g.N;
    p.epi = g.epi; p.act = g.act; p.accumulate = g.accumulate; p.atomic = splits > 1;
    p.bias = g.bias; p.bias_on_m = pl.bias_on_m; p.aux = g.aux; p.aux_sm = pl.aux_sm; p.aux_sn = pl.aux_sn;
    p.rx = g.recon_x; p.rx_ld = g.ld_x; p.rxhat = g.recon_xhat; p.racc = g.recon_acc; p.inv_batch = g.inv_batch;
    p.kb_total = pl.kb_total; p.kb_per_split = pl.kb_per;
    p.tiles_n = (int)pl.tn; p.splits = splits; p.work_total = pl.tm * pl.tn * splits;
    p.extra = g.extra_col;
    {
        static const int sw = exp_switch("CDG_TC_A16SWAP", 0) != 0, sw32 = exp_switch("CDG_TC_SW32", 0) != 0;
        p.a16swap = sw;
        p.sw32 = sw32;
    }
    p.conv_cb = g.conv_C > 0 ? g.conv_C / BK : 0; p.conv_W = g.conv_W; p.conv_H = g.conv_H; p.conv_k = g.conv_k;
    {
        // default on: measured on B200, tcgen05 kind::tf32 ignores the 13 low mantissa bits (results with the raw
        // tile as `hi` match the explicitly rounded split to 2e-7)
        static const int rawhi = exp_switch("CDG_TC_RAWHI", 1) != 0, probe = exp_switch("CDG_TC_PROBE", 0);
        p.rawhi = rawhi;
        p.probe = probe;
    }
    // 128-bit epilogue path: row-major output whose rows, bias, aux and recon operands are 16-byte aligned
    p.vec = (p.sc_n == 1 && p.sc_m % 4 == 0 && al16(p.C)) ? 1 : 0;
    if ((g.epi == EPI_BIAS || g.epi == EPI_BIAS_ACT || g.epi == EPI_RECON) && !pl.bias_on_m && !al16(g.bias)) p.vec = 0;
    if (g.epi == EPI_MUL_DACT && !(p.aux_sn == 1 && p.aux_sm % 4 == 0 && al16(g.aux))) p.vec = 0;
    if (g.epi == EPI_RECON) {
        if (!g.recon_x) { set_error("EPI_RECON without target"); return CDG_ERR_INVALID; }
        if (!(g.ld_x % 4 == 0 && al16(g.recon_x) && (!g.recon_xhat || al16(g.recon_xhat)))) p.vec = 0;
    }
    auto al32 = [](const void* q) { return ((uintptr_t)q & 31) == 0; };
    p.vec8 = (p.vec && p.sc_m % 8 == 0 && al32(p.C)) ? 1 : 0;
    if (g.epi == EPI_MUL_DACT && !(p.aux_sm % 8 == 0 && al32(g.aux))) p.vec8 = 0;
    if (g.epi == EPI_RECON && !(g.ld_x % 8 == 0 && al32(g.recon_x) && (!g.recon_xhat || al32(g.recon_xhat)))) p.vec8 = 0;

    if (p.atomic && !g.accumulate) {
        // zero C (in the caller's orientation) before the partial sums are added
        if (g0.ldc == g0.N) CDG_CHECK_CUDA(cudaMemsetAsync(g0.C, 0, sizeof(float) * g0.M * g0.N, s));
        else CDG_CHECK_CUDA(cudaMemset2DAsync(g0.C, sizeof(float) * g0.ldc, 0, sizeof(float) * g0.N, g0.M, s));
    }

    CUtensorMap ta, tb;
    const int b_box = pl.b_mn ? (BN <= 128 ? BN : (BN % 128 == 0 ? 128 : (BN == 160 ? 80 : BN / 4))) : (BN <= 256 ? BN : BN / 2);
    if (g.conv_C > 0) CDG_TRY(make_conv_map(g.A, g.conv_B, g.conv_H, g.conv_W, g.conv_C, BK, &ta));
    else CDG_TRY(make_map(g.A, g.M, g.K, g.sa_m, g.sa_k, pl.a_mn, 128, BK, &ta));
    CUtensorMap tb2;
    if (g.b_hi16 && passes == 2) {
        const int pb = BN <= 256 ? BN : BN / 2;
        CDG_TRY(make_map_bf16(g.b_hi16, g.N, g.K, g.ld_b16, pb, p.sw32, &tb));
        CDG_TRY(make_map_bf16(g.b_lo16, g.N, g.K, g.ld_b16, pb, p.sw32, &tb2));
        p.b_pre = 1;
    } else {
        CDG_TRY(make_map(g.B, g.N, g.K, g.sb_n, g.sb_k, pl.b_mn, b_box, BK, &tb));
        tb2 = tb;
        p.b_pre = 0;
    }
    dim3 grid((unsigned)imin64(p.work_total, kNumSMs));
    int r;
    // single-buffered staged epilogue, measured on B200: 7.8 vs 7.1 ms per step for the decoder-output GEMMs -> experiments only
    static const int no_stg = exp_switch("CDG_TC_STG", 0) == 0;
    static const int cta2 = exp_switch("CDG_TC_CTA2", 1) != 0;
    // transposing epilogue of the fused reconstruction head (row-coalesced 128-byte accesses through a shared-memory
    // scratch).  Measured on B200: dec2 forward 13.3 vs 6.5 ms per step -- four 32-bit memory instructions per element
    // (STS, LDS, LDG, STG) instead of half a 256-bit one: the single epilogue warp per scheduler is bound by the number of
    // memory instructions it can issue, not by how the sectors coalesce.  Off unless CDG_TC_EPI_TR=1.
    static const int epi_tr = exp_switch("CDG_TC_EPI_TR", 0) != 0;
    p.tr = epi_tr;
    if (passes == 2 && p.b_pre && cta2 && (BN == 304 || BN == 256 || BN == 160) && g.M >= 1024 && g.conv_C == 0) {
        // CTA pairs: a work item is two 128-row blocks; each CTA stages half of the B rows of each MMA
        const int N0 = BN <= 256 ? BN : 160, N1 = BN - N0;
        CDG_TRY(make_map_bf16(g.b_hi16, g.N, g.K, g.ld_b16, N0 / 2, 0, &tb));
        CDG_TRY(make_map_bf16(g.b_lo16, g.N, g.K, g.ld_b16, N0 / 2, 0, &tb2));
        CUtensorMap tb3 = tb, tb4 = tb2;
        if (N1 > 0) {
            CDG_TRY(make_map_bf16(g.b_hi16, g.N, g.K, g.ld_b16, N1 / 2, 0, &tb3));
            CDG_TRY(make_map_bf16(g.b_lo16, g.N, g.K, g.ld_b16, N1 / 2, 0, &tb4));
        }
        p.sw32 = 0;
        p.work_total = ((pl.tm + 1) / 2) * pl.tn * splits;
        const unsigned clusters = (unsigned)imin64(p.work_total, kNumSMs / 2);
        if (BN == 256 && g.epi == EPI_RECON && p.vec && !pl.a_mn && !no_stg && g.M < (1ll << 31) && g.N < (1ll << 31)) {
            // fused reconstruction head with the staged epilogue (N1 = 0: the tmX / tmC slots carry the target / gradient maps)
            CUtensorMap tx, tcm, txh;
            CDG_TRY(make_map(g.recon_x, g.M, g.N, g.ld_x, 1, false, 32, 32, &tx));
            CDG_TRY(make_map(g.C, g.M, g.N, pl.sc_m, 1, false, 32, 32, &tcm));
            if (g.recon_xhat) CDG_TRY(make_map(g.recon_xhat, g.M, g.N, pl.sc_m, 1, false, 32, 32, &txh));
            else txh = tcm;
            r = launch_cta2<256, false, true>(ta, tb, tb2, tx, tcm, p, clusters, s, &txh);
        } else if (BN == 304) r = pl.a_mn ? launch_cta2<304, true>(ta, tb, tb2, tb3, tb4, p, clusters, s) : launch_cta2<304, false>(ta, tb, tb2, tb3, tb4, p, clusters, s);
        else if (BN == 256 && g.epi == EPI_RECON && p.tr && p.vec && !pl.a_mn)       // fused head: transposing epilogue
            r = launch_cta2<256, false, false, true>(ta, tb, tb2, tb3, tb4, p, clusters, s);
        else if (BN == 256) r = pl.a_mn ? launch_cta2<256, true>(ta, tb, tb2, tb3, tb4, p, clusters, s) : launch_cta2<256, false>(ta, tb, tb2, tb3, tb4, p, clusters, s);
        else r = pl.a_mn ? launch_cta2<160, true>(ta, tb, tb2, tb3, tb4, p, clusters, s) : launch_cta2<160, false>(ta, tb, tb2, tb3, tb4, p, clusters, s);
    } else if (passes == 2 && BN == 256 && g.epi == EPI_RECON && p.vec && !pl.a_mn && !pl.b_mn && !no_stg && g.M < (1ll << 31) &&
        g.N < (1ll << 31)) {
        // staged epilogue: target / gradient / xhat as [M rows][N columns] maps with 32 x 32 boxes (SWIZZLE_128B)
        CUtensorMap tx, tcm, txh;
        CDG_TRY(make_map(g.recon_x, g.M, g.N, g.ld_x, 1, false, 32, 32, &tx));
        CDG_TRY(make_map(g.C, g.M, g.N, pl.sc_m, 1, false, 32, 32, &tcm));
        if (g.recon_xhat) CDG_TRY(make_map(g.recon_xhat, g.M, g.N, pl.sc_m, 1, false, 32, 32, &txh));
        else txh = tcm;
        r = launch_recon_staged(ta, tb, tb2, tx, tcm, txh, p, grid, s);
    } else if (passes == 2) {
        if (BN == 64) r = launch_layout<64, 32, 2>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 128) r = launch_layout<128, 32, 2>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 160) r = launch_layout<160, 32, 2>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 256) r = launch_layout<256, 32, 2>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else r = launch_layout<304, 32, 2>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
    } else if (passes == 1) {
        if (BN == 64) r = launch_layout<64, 32, 1>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 128) r = launch_layout<128, 32, 1>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 160) r = launch_layout<160, 32, 1>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 256) r = launch_layout<256, 32, 1>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else r = launch_layout<304, 32, 1>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
    } else if (BK == 32) {
        if (BN == 64) r = launch_layout<64, 32, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 128) r = launch_layout<128, 32, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 160) r = launch_layout<160, 32, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 256) r = launch_layout<256, 32, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else r = launch_layout<304, 32, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
    } else {
        if (BN == 128) r = launch_layout<128, 16, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 160) r = launch_layout<160, 16, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else if (BN == 256) r = launch_layout<256, 16, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
        else r = launch_layout<304, 16, 3>(pl.a_mn, pl.b_mn, ta, tb, tb2, p, grid, s);
    }
    CDG_TRY(r);
    if (p.atomic && g0.epi != EPI_NONE)
        CDG_TRY(launch_bias_act(g0.C, g0.ldc, g0.M, g0.N, g0.bias, g0.epi, g0.act, g0.aux, g0.ld_aux, s, g0.out_hi16, g0.out_lo16,
                                g0.ld_out16, g0.out_ones));
    return CDG_OK;
}

}  // namespace cdg
