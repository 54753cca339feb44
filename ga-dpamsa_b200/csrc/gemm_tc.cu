// l1 of the Q-network on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   C[M][N] = leaky_relu( (A . W^T) * 2^-s + bias ),   A = Z (encoder output), W = l1.weight (dqn.py:57,73)
//
// Precision: the north star needs greedy-action parity with the fp32 reference, which bf16/fp16 inputs
// (2^-9 / 2^-11 per operand) cannot give. Both operands are therefore split into two fp16 terms,
// x = hi + lo (W pre-scaled by 2^s so that lo stays normal), and the product is evaluated as
// hi.hi + lo.hi + hi.lo with fp32 accumulation in TMEM - 2^-22 per operand, i.e. fp32-class - as ONE
// k-loop over three (A-segment, W-segment) pairs of the same tiles:
//     A' = [Zh | Zl]  (M x 2K fp16)       W' = [Wh | Wl]  (Npad x 2K fp16)
//     segments: (Zh,Wh) (Zl,Wh) (Zh,Wl)
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles):
//     warp 0   TMA producer   cp.async.bulk.tensor.2d -> 128B-swizzled smem stages, mbarrier expect_tx
//     warp 1   MMA issuer     one elected lane issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16), commits
//                             to the stage's "empty" barrier and, per tile, to the accumulator's "full" barrier
//     warps 2-9 epilogue      tcgen05.ld 32x32b from their TMEM lane quarter and column half into fp32 registers
//                             per chunk; after the last chunk scale + bias + LeakyReLU and store
// Tail: tiles are dealt round-robin; a last partial round that fits twice on the grid is dealt as column halves
// (see `split` in the kernel).
// Accumulation: tcgen05 adds each K=16 product block into the fp32 TMEM accumulator with truncation, a
// bias that grows linearly with the number of accumulations (measured: 1e-4 relative after the 2232
// MMAs of one 6x30 tile, 50x worse than fp32). The k-loop is therefore cut into chunks of CHUNK_KB
// k-blocks (32 MMAs): each chunk accumulates into one of two TMEM buffers (2 x 256 columns) from zero
// and the epilogue warps add the finished chunk into fp32 registers with round-to-nearest while the
// next chunk runs in the other buffer - the same promotion DeepGEMM applies to FP8 on Hopper.
// Canonical summation order and split-K (fused layers = l1): the k-loop is cut into K_SLICES fixed slices of
// whole chunks; a slice's chunks are folded left to right into a partial p_s and the result is
// ((p_0 + p_1) + p_2) + p_3 - for EVERY launch shape, so a state's Q-values do not depend on the batch it is
// evaluated in (duplicate elimination and sharding change the batches; greedy actions must not move).
// Large batches: one unit runs all slices and keeps the running total in an L2-resident scratch tile between
// slices. Small batches (fewer tiles than SMs: the lock-step episodes after duplicate elimination) are latency-bound by
// the 372 k-blocks of one tile: there ksplit = 2 or 4 units per column half each run K_SLICES / ksplit slices
// and write their partials to a workspace that splitk_reduce_kernel folds in the same order (+ bias,
// LeakyReLU, fp16 split).
#include "gad_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                      // 64 fp16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int ACC_COLS = 256;               // TMEM columns per accumulator buffer
constexpr int TMEM_COLS = 512;
constexpr int CHUNK_KB = 8;                 // k-blocks (of 64) per TMEM accumulation chunk = 32 MMAs
constexpr int EPI_WARPS = 8;                // 2 per TMEM lane quarter, each owning half of the tile's columns

// FUSED: one stage holds the hi AND lo tiles of both operands for a 32-wide k-block (64-byte swizzle rows), so
// the three split products (lo.hi, hi.lo, hi.hi) are issued from one load of four tiles instead of three loads
// of two - a third less L2 -> shared-memory traffic for the same MMAs.
constexpr int BKF = 32;
constexpr int CHUNK_KB_F = 6;               // fused k-blocks per TMEM accumulation chunk = 36 MMAs
constexpr int K_SLICES = 4;                 // canonical k-slices of a fused layer (see the header comment)

template <int BN, bool FUSED = false> struct Cfg {
    static constexpr int A_BYTES = FUSED ? BM * BKF * 2 : BM * BK * 2;          // one tile (hi or lo)
    static constexpr int B_BYTES = FUSED ? BN * BKF * 2 : BN * BK * 2;
    static constexpr int STAGE_BYTES = FUSED ? 2 * (A_BYTES + B_BYTES) : A_BYTES + B_BYTES;
    static constexpr int STAGES = (220 * 1024) / STAGE_BYTES > 6 ? 6 : (220 * 1024) / STAGE_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(B_BYTES % (FUSED ? 512 : 1024) == 0, "tiles must keep the swizzle-atom alignment");
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");
};

using namespace tcptx;

// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN, bool SPLIT_OUT, bool FUSED = false>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
gemm_split_f16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_bh, const float *__restrict__ bias, float out_scale, void *__restrict__ Cout, int ldc, int lo_offset,
                      const int32_t *__restrict__ m_ptr, long long m_fixed, int N, int K, int leaky, int tail_split,
                      float *__restrict__ ws, long long ws_rows, int ws_ld, float *__restrict__ scratch, int ksplit)
{
    using C_ = Cfg<BN, FUSED>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C_::STAGES * C_::STAGE_BYTES);
    // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * C_::STAGES;
    const uint32_t tfull0 = empty0 + 8 * C_::STAGES, tempty0 = tfull0 + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * C_::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long M = m_ptr ? (long long)*m_ptr : m_fixed;
    const int tiles_n = (N + BN - 1) / BN;
    const long long tiles = ((M + BM - 1) / BM) * tiles_n;
    const int kblocks = K / BK;                          // per segment (unfused)
    // k-blocks and chunk length of one tile in this kernel's own units
    const int total_kb = FUSED ? K / BKF : 3 * kblocks;
    constexpr int CHUNK = FUSED ? CHUNK_KB_F : CHUNK_KB;
    // Wave quantisation: tiles are dealt round-robin to gridDim.x persistent CTAs, so a last, partial round of
    // t_tail tiles costs a whole tile time on a few CTAs while the rest idle (l1 at 16 384 rows: 640 tiles on 148 CTAs
    // = 4.32 rounds). When the partial round fits twice, its tiles are cut into two column halves and dealt to twice as
    // many CTAs: a half unit loads the A tile and only the W rows [0, HN) or [BN - HN, BN) of its tile (second tensor
    // map, box of HN rows), issues MMAs of N = HN and its epilogue keeps BN/2 of those columns. work unit u < t_full: tile u; otherwise tile
    // t_full + (u - t_full) / 2, half (u - t_full) & 1.
    constexpr bool CAN_SPLIT = FUSED && (BN / 2) % 8 == 0 && ((BN / 2 + 15) / 16 * 16) % 8 == 0;
    constexpr int HN = (BN / 2 + 15) / 16 * 16;                   // UMMA N of a half unit (208 -> 112)
    const long long t_full = (tiles / gridDim.x) * gridDim.x, t_tail = tiles - t_full;
    const bool split = CAN_SPLIT && tail_split && t_tail > 0 && 2 * t_tail <= (long long)gridDim.x;
    // split-K launches (ksplit > 1, fused only): every tile is dealt as 2 column halves x ksplit k-slice groups
    const int ks = FUSED ? ksplit : 1;
    constexpr int HS = CAN_SPLIT ? 2 : 1;
    const long long units = ks > 1 ? tiles * HS * ks : (split ? t_full + 2 * t_tail : tiles);
    const int nchunks = (total_kb + CHUNK - 1) / CHUNK;
    auto slice_chunk0 = [&](int sl) { return (int)(((long long)nchunks * sl) / K_SLICES); };      // first chunk of slice sl
    // returns -1 for a whole tile, else the column half; [s0, s1) = the canonical k-slices the unit covers
    auto decode = [&](long long u, int &m_blk, int &n_blk, int &s0, int &s1) {
        long long t = u;
        int hsel = -1;
        s0 = 0;
        s1 = FUSED ? K_SLICES : 1;
        if (ks > 1) {
            t = u / (HS * ks);
            const int rem = (int)(u % (HS * ks));
            if (CAN_SPLIT) hsel = rem / ks;
            const int per = K_SLICES / ks;
            s0 = (rem % ks) * per;
            s1 = s0 + per;
        } else if (split && u >= t_full) {
            t = t_full + ((u - t_full) >> 1);
            hsel = (int)((u - t_full) & 1);
        }
        m_blk = (int)(t / tiles_n);
        n_blk = (int)(t % tiles_n);
        return hsel;
    };
    // k-block range of slices [s0, s1) (fused) or of the whole tile
    auto kb_range = [&](int s0, int s1, int &kb_lo, int &kb_hi) {
        if (!FUSED) {
            kb_lo = 0;
            kb_hi = total_kb;
            return;
        }
        kb_lo = slice_chunk0(s0) * CHUNK;
        kb_hi = slice_chunk0(s1) * CHUNK;
        if (kb_hi > total_kb) kb_hi = total_kb;
        if (kb_lo > kb_hi) kb_lo = kb_hi;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < C_::STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                     // TMEM allocation (whole warp), 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < units; t += gridDim.x) {
                int m_blk, n_blk, s0, s1, kb_lo, kb_hi;
                const int hsel = decode(t, m_blk, n_blk, s0, s1);
                kb_range(s0, s1, kb_lo, kb_hi);
                if constexpr (FUSED) {
                    // a half unit loads only its HN rows of W (box of map_bh), starting at row 0 or BN - HN of the tile
                    const CUtensorMap *mb = hsel < 0 ? &map_b : &map_bh;
                    const int b_row = n_blk * BN + (hsel > 0 ? BN - HN : 0);
                    const uint32_t bytes = hsel < 0 ? (uint32_t)C_::STAGE_BYTES : (uint32_t)(2 * C_::A_BYTES + 2 * HN * BKF * 2);
                    for (int kb = kb_lo; kb < kb_hi; ++kb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_u32(smem + stage * C_::STAGE_BYTES);
                        const uint32_t sal = sa + C_::A_BYTES, sb = sal + C_::A_BYTES, sbl = sb + C_::B_BYTES;
                        mbar_expect_tx(full0 + 8 * stage, bytes);
                        tma_load_2d(sa, &map_a, full0 + 8 * stage, kb * BKF, m_blk * BM);          // Zh
                        tma_load_2d(sal, &map_a, full0 + 8 * stage, K + kb * BKF, m_blk * BM);     // Zl
                        tma_load_2d(sb, mb, full0 + 8 * stage, kb * BKF, b_row);                   // Wh
                        tma_load_2d(sbl, mb, full0 + 8 * stage, K + kb * BKF, b_row);              // Wl
                        if (++stage == C_::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                } else {
                    for (int seg = 0; seg < 3; ++seg) {
                        const int a_off = seg == 1 ? K : 0, b_off = seg == 2 ? K : 0;
                        for (int kb = 0; kb < kblocks; ++kb) {
                            mbar_wait(empty0 + 8 * stage, phase ^ 1);
                            const uint32_t sa = smem_u32(smem + stage * C_::STAGE_BYTES), sb = sa + C_::A_BYTES;
                            mbar_expect_tx(full0 + 8 * stage, C_::STAGE_BYTES);
                            tma_load_2d(sa, &map_a, full0 + 8 * stage, a_off + kb * BK, m_blk * BM);
                            tma_load_2d(sb, &map_b, full0 + 8 * stage, b_off + kb * BK, n_blk * BN);
                            if (++stage == C_::STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0;
            uint32_t phase = 0;
            long long chunk = 0;                                                  // running chunk counter (buffer = chunk & 1)
            for (long long t = blockIdx.x; t < units; t += gridDim.x) {
                int m_blk, n_blk, s0, s1, kb_lo, kb_hi;
                const int hsel = decode(t, m_blk, n_blk, s0, s1);
                kb_range(s0, s1, kb_lo, kb_hi);
                const uint32_t idesc_u = hsel < 0 ? idesc : make_idesc(HN);
                (void)n_blk;
                for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += CHUNK, ++chunk) {
                    const int as = (int)(chunk & 1);
                    mbar_wait(tempty0 + 8 * as, (uint32_t)((chunk >> 1) & 1) ^ 1);    // epilogue has drained this buffer
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + as * ACC_COLS;
                    const int kb1 = kb0 + CHUNK < kb_hi ? kb0 + CHUNK : kb_hi;
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(full0 + 8 * stage, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa = smem_u32(smem + stage * C_::STAGE_BYTES);
                        if constexpr (FUSED) {
                            const uint32_t sal = sa + C_::A_BYTES, sb = sal + C_::A_BYTES, sbl = sb + C_::B_BYTES;
                            const uint64_t dah = make_smem_desc(sa, true), dal = make_smem_desc(sal, true);
                            const uint64_t dbh = make_smem_desc(sb, true), dbl = make_smem_desc(sbl, true);
#pragma unroll
                            for (int k = 0; k < BKF / UMMA_K; ++k) {     // small terms first, hi.hi last
                                umma_f16(tmem_d, dal + 2 * k, dbh + 2 * k, idesc_u, (kb > kb0) || k != 0);
                                umma_f16(tmem_d, dah + 2 * k, dbl + 2 * k, idesc_u, 1u);
                                umma_f16(tmem_d, dah + 2 * k, dbh + 2 * k, idesc_u, 1u);
                            }
                        } else {
                            const uint32_t sb = sa + C_::A_BYTES;
                            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the (>>4) address field
                                umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0) || k != 0);
                            }
                        }
                        umma_commit(empty0 + 8 * stage);                          // frees the smem stage when the MMAs retire
                        if (++stage == C_::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(tfull0 + 8 * as);                                 // chunk ready for the epilogue
                }
            }
        }
    } else {
        // ===================== epilogue: warps 2..9; TMEM lane quarter = warp & 3, column half = (warp-2) >> 2 =====
        constexpr int HALF = BN / 2;                       // columns per epilogue warp
        constexpr int NLD = HALF / 8;                      // x8 loads per chunk
        static_assert(HALF % 8 == 0, "BN/2 must be a multiple of 8");
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        long long chunk = 0;
        for (long long t = blockIdx.x; t < units; t += gridDim.x) {
            int m_blk, n_blk, s0, s1;
            const int hsel = decode(t, m_blk, n_blk, s0, s1);
            // columns of this warp: TMEM columns [col_t, col_t + ncols) of the accumulator = tile columns [col_o, ...)
            int col_t = half * HALF, col_o = half * HALF, ncols = HALF;
            if (hsel >= 0) {                               // half unit: BN/2 useful columns, dealt 8-aligned to the two warps
                constexpr int FIRST = (HALF / 2 + 7) / 8 * 8;
                const int sub = half ? FIRST : 0;
                ncols = half ? HALF - FIRST : FIRST;
                col_o = hsel * HALF + sub;                 // tile column
                col_t = col_o - (hsel ? BN - HN : 0);      // accumulator column: the unit's B rows start at BN - HN
            }
            float acc[HALF];
#pragma unroll
            for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
            const long long row = (long long)m_blk * BM + quarter * 32 + lane;
            const int n0 = n_blk * BN + col_o;
            // Where this thread parks slice partials, as float4 units laid out [column / 4][tile row] so that the 32
            // lanes of a warp (32 consecutive rows) store 512 contiguous bytes: the split-K workspace
            // ws[slice][m tile][column / 4][tile row] (ksplit > 1), or this CTA's running-total tile (whole-K units).
            float4 *park = nullptr;
            size_t park_slice = 0;
            if constexpr (FUSED) {
                if (ks > 1) {
                    park = reinterpret_cast<float4 *>(ws) + ((size_t)m_blk * (ws_ld >> 2) + (n0 >> 2)) * BM + quarter * 32 + lane;
                    park_slice = (size_t)(ws_rows / BM) * (ws_ld >> 2) * BM;
                } else {
                    park = reinterpret_cast<float4 *>(scratch) + ((size_t)blockIdx.x * (BN >> 2) + (col_o >> 2)) * BM + quarter * 32 + lane;
                }
            }
            for (int sl = s0; sl < s1; ++sl) {
                int c_lo = 0, c_hi = nchunks;
                if constexpr (FUSED) {
                    c_lo = slice_chunk0(sl);
                    c_hi = slice_chunk0(sl + 1);
                }
                for (int c = c_lo; c < c_hi; ++c, ++chunk) {
                    const int as = (int)(chunk & 1);
                    mbar_wait(tfull0 + 8 * as, (uint32_t)((chunk >> 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * ACC_COLS + col_t;
#pragma unroll
                    for (int g = 0; g < NLD; ++g) {
                        if (8 * g >= ncols) break;
                        uint32_t r[8];
                        tmem_ld8(taddr + 8 * g, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int u = 0; u < 8; ++u) acc[8 * g + u] += __uint_as_float(r[u]);       // fp32 RN promotion
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * as);
                }
                if constexpr (FUSED) {
                    if (ks > 1) {                                  // p_sl goes to the workspace; splitk_reduce_kernel folds
                        float4 *dst = park + (size_t)sl * park_slice;
#pragma unroll
                        for (int j = 0; j < HALF; j += 4) {
                            if (j >= ncols) break;
                            dst[(size_t)(j >> 2) * BM] = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                        }
#pragma unroll
                        for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
                    } else if (sl > 0) {                           // running total ((p_0 + p_1) + p_2) + p_3, in this order
#pragma unroll
                        for (int j = 0; j < HALF; j += 4) {
                            if (j >= ncols) break;
                            const float4 t = park[(size_t)(j >> 2) * BM];         // this thread's own earlier store
                            acc[j] = t.x + acc[j];
                            acc[j + 1] = t.y + acc[j + 1];
                            acc[j + 2] = t.z + acc[j + 2];
                            acc[j + 3] = t.w + acc[j + 3];
                        }
                    }
                    if (ks == 1 && sl + 1 < K_SLICES) {            // park the running total, start the next slice from zero
#pragma unroll
                        for (int j = 0; j < HALF; j += 4) {
                            if (j >= ncols) break;
                            park[(size_t)(j >> 2) * BM] = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                        }
#pragma unroll
                        for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
                    }
                }
            }
            if constexpr (FUSED) {
                if (ks > 1) continue;
            }
            if (row < M) {
#pragma unroll
                for (int j = 0; j < HALF; j += 8) {
                    if (j >= ncols) break;
                    float x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        x[u] = 0.f;
                        if (n0 + j + u < N) {
                            x[u] = fmaf(acc[j + u], out_scale, bias[n0 + j + u]);
                            if (leaky) x[u] = x[u] > 0.f ? x[u] : 0.01f * x[u];
                        }
                    }
                    if (n0 + j >= N) continue;
                    if constexpr (SPLIT_OUT) {           // (hi | lo) fp16 activations of the next layer
                        __half *dh = reinterpret_cast<__half *>(Cout) + (size_t)row * ldc + n0 + j;
                        __half hi[8], lo[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            hi[u] = __float2half_rn(x[u]);
                            lo[u] = __float2half_rn(x[u] - __half2float(hi[u]));
                        }
                        if (n0 + j + 7 < N) {
                            *reinterpret_cast<uint4 *>(dh) = *reinterpret_cast<const uint4 *>(hi);
                            *reinterpret_cast<uint4 *>(dh + lo_offset) = *reinterpret_cast<const uint4 *>(lo);
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (n0 + j + u < N) {
                                    dh[u] = hi[u];
                                    dh[lo_offset + u] = lo[u];
                                }
                        }
                    } else {
                        float *dst = reinterpret_cast<float *>(Cout) + (size_t)row * ldc + n0 + j;
                        if (n0 + j + 7 < N && (ldc & 3) == 0) {
                            *reinterpret_cast<float4 *>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                            *reinterpret_cast<float4 *>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
                        } else {
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (n0 + j + u < N) dst[u] = x[u];
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// Second half of a split-K launch: out = act((((p_0 + p_1) + p_2) + p_3) * 2^-s + bias) as (hi | lo) fp16, the
// same fold the whole-K units do through their running total. One thread per (row, 8 columns); partials are
// float4 units laid out [slice][m tile][column / 4][tile row].
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float *__restrict__ ws, long long ws_rows, int ws_ld, const int32_t *__restrict__ m_ptr,
                     long long m_fixed, int N, const float *__restrict__ bias, float out_scale, int leaky,
                     __half *__restrict__ out, int ldc, int lo_offset)
{
    const long long M = m_ptr ? (long long)*m_ptr : m_fixed;
    const int groups = (N + 7) / 8;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r_in = (int)(i % BM);                       // consecutive threads = consecutive rows of one tile: coalesced reads
    const long long rest = i / BM;
    const int n0 = (int)(rest % groups) * 8;
    const long long m_blk = rest / groups;
    const long long row = m_blk * BM + r_in;
    if (row >= M) return;
    const size_t slice = (size_t)(ws_rows / BM) * (ws_ld >> 2) * BM;
    const float4 *src = reinterpret_cast<const float4 *>(ws) + ((size_t)m_blk * (ws_ld >> 2) + (n0 >> 2)) * BM + r_in;
    float t[8];
#pragma unroll
    for (int sl = 0; sl < K_SLICES; ++sl) {
        const float4 a = src[sl * slice];
        const float4 b = src[sl * slice + BM];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = sl == 0 ? v[u] : t[u] + v[u];
    }
    __half hi[8], lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float x = 0.f;
        if (n0 + u < N) {
            x = fmaf(t[u], out_scale, bias[n0 + u]);
            if (leaky) x = x > 0.f ? x : 0.01f * x;
        }
        hi[u] = __float2half_rn(x);
        lo[u] = __float2half_rn(x - __half2float(hi[u]));
    }
    __half *dh = out + (size_t)row * ldc + n0;
    if (n0 + 7 < N) {
        *reinterpret_cast<uint4 *>(dh) = *reinterpret_cast<const uint4 *>(hi);
        *reinterpret_cast<uint4 *>(dh + lo_offset) = *reinterpret_cast<const uint4 *>(lo);
    } else {
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (n0 + u < N) {
                dh[u] = hi[u];
                dh[lo_offset + u] = lo[u];
            }
    }
}

// fp32 weights [N][K] -> fp16 [Npad][2*Kpad] = (hi | lo) of w * 2^s; rows >= N and columns >= K are zero
__global__ void split_weights_kernel(const float *__restrict__ W, int N, int Npad, int K, int Kpad, float scale,
                                     __half *__restrict__ out)
{
    const long long total = (long long)Npad * Kpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / Kpad;
        const int k = (int)(i - n * Kpad);
        const float w = (n < N && k < K) ? W[(size_t)n * K + k] * scale : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        out[(size_t)n * 2 * Kpad + k] = hi;
        out[(size_t)n * 2 * Kpad + Kpad + k] = lo;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map(CUtensorMap *map, const void *base, long long rows, long long cols, int box_rows, bool fused = false)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GAD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        GAD_REQUIRE(p && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
    const cuuint32_t box[2] = {(cuuint32_t)(fused ? BKF : BK), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, fused ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        gad_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld box=%dx%d)", (int)rc, rows, cols,
                      BK, box_rows);
        return GAD_ECUDA;
    }
    return GAD_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host interface used by qnet.cu
// ---------------------------------------------------------------------------------------------
template <int BN>
int set_smem_attr()
{
    GAD_CUDA(cudaFuncSetAttribute(gemm_split_f16_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    GAD_CUDA(cudaFuncSetAttribute(gemm_split_f16_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM));
    GAD_CUDA(cudaFuncSetAttribute(gemm_split_f16_kernel<BN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<BN, true>::SMEM));
    GAD_CUDA(cudaFuncSetAttribute(gemm_split_f16_kernel<BN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<BN, true>::SMEM));
    return GAD_OK;
}

int tc_layer_init(TcLayer *L, const float *d_W, const float *h_W, int N, int K, int BN, int fused)
{
    L->fused = fused;
    GAD_REQUIRE(BN == TC_BN_L1 || BN == TC_BN_L2 || BN == TC_BN_L3, "tcgen05 layer: unsupported BN=%d", BN);
    L->N = N;
    L->K = K;
    L->Kpad = (K + BK - 1) / BK * BK;
    L->BN = BN;
    L->Npad = (N + BN - 1) / BN * BN;
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)N * K; ++i) mx = fmaxf(mx, fabsf(h_W[i]));
    int s = 0;
    if (mx > 0.f) s = (int)floorf(log2f(32768.f / mx));
    if (s > 60) s = 60;
    if (s < -60) s = -60;
    L->scale = ldexpf(1.f, s);
    L->inv_scale = ldexpf(1.f, -s);
    GAD_CUDA(cudaMalloc((void **)&L->W2, (size_t)L->Npad * 2 * L->Kpad * sizeof(__half)));
    split_weights_kernel<<<GAD_NUM_SMS * 8, 256>>>(d_W, N, L->Npad, K, L->Kpad, L->scale, (__half *)L->W2);
    GAD_LAUNCH_CHECK();
    GAD_CUDA(cudaDeviceSynchronize());
    int rc = encode_map((CUtensorMap *)L->map_b, L->W2, L->Npad, 2ll * L->Kpad, BN, L->fused != 0);
    if (!rc && L->fused)                                   // box of half a tile's rows, rounded up to a valid UMMA N
        rc = encode_map((CUtensorMap *)L->map_bh, L->W2, L->Npad, 2ll * L->Kpad, (BN / 2 + 15) / 16 * 16, true);
    if (rc) return rc;
    if (L->fused) {
        // split-K workspace (K_SLICES partial tiles for up to TC_SPLITK_MAX_ROWS rows) and the whole-K units'
        // scratch (one running-total tile of 128 x BN per CTA); both stay L2-resident
        GAD_CUDA(cudaMalloc((void **)&L->ws, (size_t)K_SLICES * TC_SPLITK_MAX_ROWS * L->Npad * sizeof(float)));
        GAD_CUDA(cudaMalloc((void **)&L->scratch, (size_t)GAD_NUM_SMS * BM * BN * sizeof(float)));
    }
    if (BN == TC_BN_L1) return set_smem_attr<TC_BN_L1>();
    if (BN == TC_BN_L2) return set_smem_attr<TC_BN_L2>();
    return set_smem_attr<TC_BN_L3>();
}

void tc_layer_free(TcLayer *L)
{
    if (L->W2) cudaFree(L->W2);
    if (L->ws) cudaFree(L->ws);
    if (L->scratch) cudaFree(L->scratch);
    L->W2 = nullptr;
    L->ws = L->scratch = nullptr;
}

int tc_layer_bind_a(TcLayer *L, const void *d_A2, long long rows)
{
    return encode_map((CUtensorMap *)L->map_a, d_A2, rows, 2ll * L->Kpad, BM, L->fused != 0);
}

template <int BN>
int run_bn(const TcLayer *L, const float *d_bias, void *d_out, int ldc, int lo_offset, int split_out, const int32_t *d_m,
           long long m_max, int leaky, cudaStream_t st)
{
    const long long tiles = ((m_max + BM - 1) / BM) * ((L->N + BN - 1) / BN);
    if (tiles == 0) return GAD_OK;
    static const int tail_split = [] {                            // A/B switch: GAD_L1_TAILSPLIT=0 keeps whole tiles only
        const char *e = getenv("GAD_L1_TAILSPLIT");
        return e ? atoi(e) : 1;
    }();
    static const int ksplit_max = [] {                            // A/B switch: GAD_L1_KSPLIT=1 never splits K
        const char *e = getenv("GAD_L1_KSPLIT");
        return e ? atoi(e) : K_SLICES;
    }();
    // small launches of a fused layer: 2 column halves x ksplit k-slice groups per tile, until about two rounds of units
    // Cost model in tile-times: whole tiles take ceil(tiles / SMs) rounds (a partial last round that fits twice
    // runs as column halves: half a round); split units are 1 / (2 ksplit) of a tile each. Split when that saves
    // at least 15 % (the fold kernel and the workspace round trip are not free).
    int ksplit = 1;
    if (L->fused && split_out && L->ws && (m_max + BM - 1) / BM * BM <= TC_SPLITK_MAX_ROWS) {
        const long long full = tiles / GAD_NUM_SMS, tail = tiles % GAD_NUM_SMS;
        double best = (double)full + (tail == 0 ? 0.0 : (tail_split && 2 * tail <= GAD_NUM_SMS ? 0.5 : 1.0));
        best *= 0.85;
        for (int s = 2; s <= K_SLICES && s <= ksplit_max; s *= 2) {
            const long long units = tiles * 2 * s;
            const double cost = (double)((units + GAD_NUM_SMS - 1) / GAD_NUM_SMS) / (2.0 * s);
            if (cost < best) {
                best = cost;
                ksplit = s;
            }
        }
    }
    const long long ws_rows = (m_max + BM - 1) / BM * BM;
    long long want = L->fused && tail_split ? 2 * tiles : tiles;       // half units need up to two CTAs per tile
    if (ksplit > 1) want = 2 * tiles * ksplit;
    const unsigned grid = (unsigned)(want < GAD_NUM_SMS ? want : GAD_NUM_SMS);
    const CUtensorMap &ma = *(const CUtensorMap *)L->map_a, &mb = *(const CUtensorMap *)L->map_b;
    const CUtensorMap &mh = *(const CUtensorMap *)(L->fused ? L->map_bh : L->map_b);
    if (L->fused && split_out)
        gemm_split_f16_kernel<BN, true, true><<<grid, 64 + 32 * EPI_WARPS, Cfg<BN, true>::SMEM, st>>>(
            ma, mb, mh, d_bias, L->inv_scale, d_out, ldc, lo_offset, d_m, m_max, L->N, L->Kpad, leaky, tail_split, L->ws, ws_rows, L->Npad, L->scratch, ksplit);
    else if (L->fused)
        gemm_split_f16_kernel<BN, false, true><<<grid, 64 + 32 * EPI_WARPS, Cfg<BN, true>::SMEM, st>>>(
            ma, mb, mh, d_bias, L->inv_scale, d_out, ldc, lo_offset, d_m, m_max, L->N, L->Kpad, leaky, tail_split, L->ws, ws_rows, L->Npad, L->scratch, ksplit);
    else if (split_out)
        gemm_split_f16_kernel<BN, true><<<grid, 64 + 32 * EPI_WARPS, Cfg<BN>::SMEM, st>>>(
            ma, mb, mh, d_bias, L->inv_scale, d_out, ldc, lo_offset, d_m, m_max, L->N, L->Kpad, leaky, tail_split, L->ws, ws_rows, L->Npad, L->scratch, ksplit);
    else
        gemm_split_f16_kernel<BN, false><<<grid, 64 + 32 * EPI_WARPS, Cfg<BN>::SMEM, st>>>(
            ma, mb, mh, d_bias, L->inv_scale, d_out, ldc, lo_offset, d_m, m_max, L->N, L->Kpad, leaky, tail_split, L->ws, ws_rows, L->Npad, L->scratch, ksplit);
    GAD_LAUNCH_CHECK();
    if (ksplit > 1) {
        const long long threads = ws_rows * ((L->N + 7) / 8);
        splitk_reduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
            L->ws, ws_rows, L->Npad, d_m, m_max, L->N, d_bias, L->inv_scale, leaky, (__half *)d_out, ldc, lo_offset);
        GAD_LAUNCH_CHECK();
    }
    return GAD_OK;
}

int tc_layer_run(const TcLayer *L, const float *d_bias, void *d_out, int ldc, int lo_offset, int split_out,
                 const int32_t *d_m, long long m_max, int leaky, cudaStream_t st)
{
    if (L->BN == TC_BN_L1) return run_bn<TC_BN_L1>(L, d_bias, d_out, ldc, lo_offset, split_out, d_m, m_max, leaky, st);
    if (L->BN == TC_BN_L2) return run_bn<TC_BN_L2>(L, d_bias, d_out, ldc, lo_offset, split_out, d_m, m_max, leaky, st);
    return run_bn<TC_BN_L3>(L, d_bias, d_out, ldc, lo_offset, split_out, d_m, m_max, leaky, st);
}
