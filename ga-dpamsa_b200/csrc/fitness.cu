// Fitness pass: pad + clean + sum-of-pairs + exact-match columns for every board of the
// population in one launch. Replaces the per-board Python loop of GA.calculate_fitness_score
// (reference GA.py:156-178, utils.py:27-128, utils.py:390-428).
//
// Mapping: a group of GS lanes (4/8/16/32, chosen from the stride) owns one board at a time;
// lane l of the group owns the 32-bit words l, l+GS, ... of every row (4 columns per word), so
// one row is read by the group as contiguous 16*GS-byte segments. Per-column symbol counts live
// in byte lanes of registers (gad_common.cuh); nothing is staged through shared memory because
// every cell is used exactly once. HBM-bound: R*W bytes in (+R*W' out only for the boards that
// actually change), 12 bytes of scores out.
#include "gad_common.cuh"
#include "hof_decide.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <atomic>

namespace {

template <int GS>
__device__ __forceinline__ uint32_t group_mask()
{
    if constexpr (GS == 32) {
        return 0xFFFFFFFFu;
    } else {
        const unsigned lane = threadIdx.x & 31u;
        return ((1u << GS) - 1u) << (lane & ~(unsigned)(GS - 1));
    }
}

template <int GS>
__device__ __forceinline__ uint32_t group_sum(uint32_t v, uint32_t mask)
{
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

template <int GS>
__device__ __forceinline__ int group_max(int v, uint32_t mask)
{
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// counts of the 4 columns of word `j` over rows [0, R)
__device__ __forceinline__ void count_word(const uint8_t *__restrict__ board, int R, int stride, int j, ColCounts &cc)
{
    const uint32_t *col = reinterpret_cast<const uint32_t *>(board) + j;
    const int wstride = stride >> 2;
    colcounts_clear(cc);
    const uint32_t first = col[0];
    int r = 0;
    for (; r + 4 <= R; r += 4) {
        const uint32_t x0 = col[(size_t)(r + 0) * wstride];
        const uint32_t x1 = col[(size_t)(r + 1) * wstride];
        const uint32_t x2 = col[(size_t)(r + 2) * wstride];
        const uint32_t x3 = col[(size_t)(r + 3) * wstride];
        colcounts_add(cc, x0, first);
        colcounts_add(cc, x1, first);
        colcounts_add(cc, x2, first);
        colcounts_add(cc, x3, first);
    }
    for (; r < R; ++r) colcounts_add(cc, col[(size_t)r * wstride], first);
}

// counts of the 16 columns of unit `u` (one uint4 = 4 words) over rows [0, R); words at or beyond
// `nwords` are outside the board (their bytes may be stale) and get an empty valid mask
template <bool PAIRED>
__device__ __forceinline__ void count_unit(const uint8_t *__restrict__ board, int R, int stride, int u, int nwords,
                                           ColCounts (&cc)[4], uint32_t (&valid)[4])
{
    const uint4 *col = reinterpret_cast<const uint4 *>(board) + u;
    const int ustride = stride >> 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        colcounts_clear(cc[k]);
        valid[k] = 4 * u + k < nwords ? GAD_MSB4 : 0u;
    }
    const uint4 first = col[0];
    if constexpr (!PAIRED) {                              // few rows: plain byte counters, fewer registers
        int r = 0;
        for (; r + 2 <= R; r += 2) {
            const uint4 x0 = col[(size_t)r * ustride], x1 = col[(size_t)(r + 1) * ustride];
            colcounts_add(cc[0], x0.x, first.x);
            colcounts_add(cc[1], x0.y, first.y);
            colcounts_add(cc[2], x0.z, first.z);
            colcounts_add(cc[3], x0.w, first.w);
            colcounts_add(cc[0], x1.x, first.x);
            colcounts_add(cc[1], x1.y, first.y);
            colcounts_add(cc[2], x1.z, first.z);
            colcounts_add(cc[3], x1.w, first.w);
        }
        if (r < R) {
            const uint4 x0 = col[(size_t)r * ustride];
            colcounts_add(cc[0], x0.x, first.x);
            colcounts_add(cc[1], x0.y, first.y);
            colcounts_add(cc[2], x0.z, first.z);
            colcounts_add(cc[3], x0.w, first.w);
        }
        return;
    }
    const uint32_t f2[4] = {first.x * 17u, first.y * 17u, first.z * 17u, first.w * 17u};   // first row in both nibbles
    ColCounts nn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) colcounts_clear(nn[k]);
    int r = 0, pairs = 0;
    for (; r + 4 <= R; r += 4) {                          // 4 independent 16-byte loads in flight per lane
        uint4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = col[(size_t)(r + i) * ustride];
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            colcounts_add2(nn[0], x[i].x, x[i + 1].x, f2[0]);
            colcounts_add2(nn[1], x[i].y, x[i + 1].y, f2[1]);
            colcounts_add2(nn[2], x[i].z, x[i + 1].z, f2[2]);
            colcounts_add2(nn[3], x[i].w, x[i + 1].w, f2[3]);
        }
        pairs += 2;
        if (pairs >= 14) {                                // 4-bit counters hold at most 15
#pragma unroll
            for (int k = 0; k < 4; ++k) colcounts_flush2(cc[k], nn[k]);
            pairs = 0;
        }
    }
    if (r + 2 <= R) {
        const uint4 x0 = col[(size_t)r * ustride], x1 = col[(size_t)(r + 1) * ustride];
        colcounts_add2(nn[0], x0.x, x1.x, f2[0]);
        colcounts_add2(nn[1], x0.y, x1.y, f2[1]);
        colcounts_add2(nn[2], x0.z, x1.z, f2[2]);
        colcounts_add2(nn[3], x0.w, x1.w, f2[3]);
        r += 2;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) colcounts_flush2(cc[k], nn[k]);
    if (r < R) {
        const uint4 x0 = col[(size_t)r * ustride];
        colcounts_add(cc[0], x0.x, first.x);
        colcounts_add(cc[1], x0.y, first.y);
        colcounts_add(cc[2], x0.z, first.z);
        colcounts_add(cc[3], x0.w, first.w);
    }
}

// One board by a group of GS lanes, cells read from global memory: width/raggedness from the row lengths (w, wmin
// already reduced over the group), column counts, scores, and - when all-gap columns exist - the in-place
// compaction of utils.py:408-428. `pre` (PRE only) holds this lane's first unit of rows 0..6, loaded before the row
// lengths were known. Returns the cleaned width.
template <int GS, bool PAIRED, bool PRE>
__device__ __forceinline__ int board_pass_global(uint8_t *__restrict__ board, int32_t *__restrict__ rl, int R, int stride,
                                                 int gl, uint32_t gmask, int w, int wmin, const uint4 (&pre)[7], bool have_pre,
                                                 int match, int mismatch, int gap, int32_t *__restrict__ sp_p,
                                                 int32_t *__restrict__ em_p, int32_t *__restrict__ al_p)
{
    const int nwords = (w + 3) >> 2;
    const int nunits = (nwords + 3) >> 2;

    BoardSums bs;
    boardsums_clear(bs);
    for (int u = gl; u < nunits; u += GS) {
        ColCounts cc[4];
        uint32_t valid[4];
        if (PRE && have_pre && u == gl) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                colcounts_clear(cc[k]);
                valid[k] = 4 * u + k < nwords ? GAD_MSB4 : 0u;
            }
#pragma unroll
            for (int r = 0; r < 7; ++r)
                if (r < R) {
                    colcounts_add(cc[0], pre[r].x, pre[0].x);
                    colcounts_add(cc[1], pre[r].y, pre[0].y);
                    colcounts_add(cc[2], pre[r].z, pre[0].z);
                    colcounts_add(cc[3], pre[r].w, pre[0].w);
                }
        } else {
            count_unit<PAIRED>(board, R, stride, u, nwords, cc, valid);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) boardsums_fold_word(bs, cc[k], valid[k] != 0u);
    }
    bs.q2 = group_sum<GS>(bs.q2, gmask);
    bs.s1 = group_sum<GS>(bs.s1, gmask);
    bs.s2 = group_sum<GS>(bs.s2, gmask);
    bs.kept = group_sum<GS>(bs.kept, gmask);
    bs.exact = group_sum<GS>(bs.exact, gmask);
    const int nk = (int)bs.kept;

    if (nk != w) {
        // utils.py:408-428: drop every all-gap column, keep the order. In place: chunk by chunk, each
        // row's units are read by the whole group before any lane stores, and stores only land at or
        // left of the columns just read.
        int base = 0;
        const int nchunks = (nunits + GS - 1) / GS;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int u = ch * GS + gl;
            uint32_t keep[4] = {0u, 0u, 0u, 0u};
            if (u < nunits) {
                ColCounts cc[4];
                uint32_t valid[4];
                count_unit<PAIRED>(board, R, stride, u, nwords, cc, valid);
#pragma unroll
                for (int k = 0; k < 4; ++k) keep[k] = nonzero_lanes(cc[k].c1 + cc[k].c2 + cc[k].c3 + cc[k].c4) & valid[k];
            }
            const int mine = __popc(keep[0]) + __popc(keep[1]) + __popc(keep[2]) + __popc(keep[3]);
            int incl = mine;                                  // inclusive scan inside the group
#pragma unroll
            for (int o = 1; o < GS; o <<= 1) {
                const int t = __shfl_up_sync(gmask, incl, o, GS);
                if (gl >= o) incl += t;
            }
            const int total = __shfl_sync(gmask, incl, GS - 1, GS);
            const int dst0 = base + incl - mine;
            for (int r = 0; r < R; ++r) {
                uint8_t *row = board + (size_t)r * stride;
                uint4 x = make_uint4(0u, 0u, 0u, 0u);
                if (u < nunits) x = reinterpret_cast<const uint4 *>(row)[u];
                __syncwarp(gmask);
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
                int d = dst0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (keep[k] & (0x80u << (8 * b))) row[d++] = (uint8_t)(xs[k] >> (8 * b));
                __syncwarp(gmask);
            }
            base += total;
        }
        // restore the padding invariant for the shrunken board
        const int pad_to = (nk + 3) & ~3;
        for (int i = gl; i < R * 4; i += GS) {
            const int r = i >> 2, c = nk + (i & 3);
            if (c < pad_to) board[(size_t)r * stride + c] = GAD_GAP;
        }
    }
    if (nk != w || wmin != w)
        for (int r = gl; r < R; r += GS) rl[r] = nk;

    if (gl == 0) {
        *sp_p = (int32_t)sum_of_pairs_closed_form(bs, R, match, mismatch, gap);
        *em_p = (int32_t)bs.exact;
        *al_p = nk;
    }
    return nk;
}

// GS lanes own one board; lane l owns the 16-byte units l, l+GS, ... of every row.
template <int GS, bool PAIRED>
__global__ void __launch_bounds__(256, 3)
fitness_kernel(uint8_t *__restrict__ boards, int32_t *__restrict__ rowlen, long long P, int R, int stride,
               int match, int mismatch, int gap, int32_t *__restrict__ sp, int32_t *__restrict__ em,
               int32_t *__restrict__ al, unsigned long long *__restrict__ max_al, unsigned long long max_tag)
{
    const uint32_t gmask = group_mask<GS>();
    const int gl = threadIdx.x & (GS - 1);                       // lane inside the group
    const long long groups_per_block = blockDim.x / GS;
    const long long group0 = (long long)blockIdx.x * groups_per_block + threadIdx.x / GS;
    const long long ngroups = (long long)gridDim.x * groups_per_block;
    int local_max = 0;

    for (long long p = group0; p < P; p += ngroups) {
        uint8_t *board = boards + (size_t)p * R * stride;
        int32_t *rl = rowlen + (size_t)p * R;

        // Few rows (R < 8): issue the loads of this lane's first unit before the row lengths are known - the
        // unit lies inside the stride, so the loads are always legal - and overlap the two DRAM round trips.
        uint4 pre[7];
        bool have_pre = false;
        if constexpr (!PAIRED) {
            if ((gl + 1) * 16 <= stride) {
                have_pre = true;
                const uint4 *col = reinterpret_cast<const uint4 *>(board) + gl;
                const int ustride = stride >> 4;
#pragma unroll
                for (int r = 0; r < 7; ++r)
                    if (r < R) pre[r] = col[(size_t)r * ustride];
            }
        }

        // width = longest row (GA.py:158); ragged = some row is shorter
        int w = 0, wmin = 0x7fffffff;
        for (int r = gl; r < R; r += GS) {
            const int v = rl[r];
            w = max(w, v);
            wmin = min(wmin, v);
        }
        w = group_max<GS>(w, gmask);
        wmin = -group_max<GS>(-wmin, gmask);
        const int nk = board_pass_global<GS, PAIRED, !PAIRED>(board, rl, R, stride, gl, gmask, w, wmin, pre, have_pre, match,
                                                             mismatch, gap, sp + p, em + p, al + p);
        local_max = max(local_max, nk);
    }
    if (max_al != nullptr) {
        local_max = group_max<32>(local_max, 0xFFFFFFFFu);
        if ((threadIdx.x & 31) == 0) atomicMax(max_al, max_tag | (unsigned long long)local_max);
    }
}

// One board, arbitrary half-open window: utils.get_sum_of_pairs / get_column_score called
// directly on a list-of-lists board (all-gap columns count, nothing is cleaned).
__global__ void __launch_bounds__(32)
window_score_kernel(const uint8_t *__restrict__ board, int stride, int fr, int tr, int fc, int tc,
                    int match, int mismatch, int gap, long long *__restrict__ sp, int32_t *__restrict__ em)
{
    const int lane = threadIdx.x;
    const int rows = tr - fr;
    BoardSums bs;
    boardsums_clear(bs);
    const int j0 = fc >> 2, j1 = (tc + 3) >> 2;
    for (int j = j0 + lane; j < j1; j += 32) {
        uint32_t valid = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = 4 * j + b;
            if (c >= fc && c < tc) valid |= 0x80u << (8 * b);
        }
        ColCounts cc;
        if (rows > 0) {
            count_word(board + (size_t)fr * stride, rows, stride, j, cc);
            boardsums_fold(bs, cc, valid, true);
        }
    }
    bs.q2 = group_sum<32>(bs.q2, 0xFFFFFFFFu);
    bs.s1 = group_sum<32>(bs.s1, 0xFFFFFFFFu);
    bs.s2 = group_sum<32>(bs.s2, 0xFFFFFFFFu);
    bs.kept = group_sum<32>(bs.kept, 0xFFFFFFFFu);
    bs.exact = group_sum<32>(bs.exact, 0xFFFFFFFFu);
    if (lane == 0) {
        *sp = rows > 0 ? sum_of_pairs_closed_form(bs, rows, match, mismatch, gap) : 0;
        *em = rows > 0 ? (int32_t)bs.exact : (tc > fc ? tc - fc : 0);
    }
}

// =============================================================================================
// TMA-staged variant (the default for boards up to 256 columns): the board cells never pass through registers
// on their way from HBM. One persistent CTA per SM; a producer thread streams boxes of NB boards x R rows x BW
// bytes (BW = the live width rounded to 16, not the stride) through a 3-D tensor map into a ring of shared-memory
// stages, all stages in flight from the first cycle (c3: the 14 boxes of an SM's share are requested at once =
// 170 KB in flight per SM, which a register-staged kernel cannot hold). Teams of NB x GS consumer threads take the
// boxes in turn; GS lanes own one board and read its rows from shared memory (16 bytes per lane and row).
//   NIB (R <= 15): two 4-column words are packed into one register of eight 4-bit lanes (x_hi * 16 + x_lo), so
//   one LOP3 per symbol class covers 8 columns of a row and the per-column counters are nibbles that cannot
//   overflow - half the integer work per cell of the byte-lane counters.
//   R >= 16: byte-lane counters with two rows per word (colcounts_add2), flushed every 14 row pairs.
// All-gap columns (utils.py:408-428) are compacted from the shared-memory copy straight to global memory (no
// in-place hazard); boards wider than the box, or with more units than lanes when they need compaction, take the
// global-memory path of the register kernel (board_pass_global) inside the same launch.
// =============================================================================================
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t ft_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ft_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ft_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(uint32_t bar, uint32_t parity)
{
    uint64_t t0 = 0;
    for (;;) {
#pragma unroll 1
        for (int spin = 0; spin < 4096; ++spin) {
            uint32_t done;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity), "r"(10000u)
                : "memory");
            if (done) return;
        }
        uint64_t now;                               // a lost arrival must fail loudly, never hang the GPU
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) __trap();
    }
}
__device__ __forceinline__ void ft_tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// nibble-lane counters of 8 columns (two packed words) over the rows of one unit half
struct NibCounts {
    uint32_t n1, n2, n3, n4, diff;
};
__device__ __forceinline__ void nib_clear(NibCounts &n) { n.n1 = n.n2 = n.n3 = n.n4 = n.diff = 0u; }
__device__ __forceinline__ void nib_add(NibCounts &n, uint32_t z, uint32_t z_first)
{
    const uint32_t b0 = z & GAD_LSN8;
    const uint32_t b1 = (z >> 1) & GAD_LSN8;
    const uint32_t b2 = (z >> 2) & GAD_LSN8;
    n.n1 += b0 & ~b1 & ~b2;
    n.n2 += b1 & ~b0;
    n.n3 += b0 & b1;
    n.n4 += b2 & ~b0;
    n.diff |= z ^ z_first;
}
// low (hi = 0) or high (hi = 1) nibbles of the counters as byte-lane ColCounts of one 4-column word
__device__ __forceinline__ void nib_unpack(const NibCounts &n, int hi, ColCounts &c)
{
    const int sh = hi * 4;
    c.c1 = (n.n1 >> sh) & GAD_NIB4;
    c.c2 = (n.n2 >> sh) & GAD_NIB4;
    c.c3 = (n.n3 >> sh) & GAD_NIB4;
    c.c4 = (n.n4 >> sh) & GAD_NIB4;
    c.diff = (n.diff >> sh) & GAD_NIB4;
}

// Column counts of unit u (16 columns = words 4u..4u+3) of a board in shared memory: `sb` = shared address of the
// board's first row, `pitch` = bytes between rows. Same outputs as count_unit.
template <bool NIB>
__device__ __forceinline__ void count_unit_smem(uint32_t sb, int R, int pitch, int u, int nwords, ColCounts (&cc)[4],
                                                uint32_t (&valid)[4])
{
    const uint32_t a0 = sb + 16u * u;
#pragma unroll
    for (int k = 0; k < 4; ++k) valid[k] = 4 * u + k < nwords ? GAD_MSB4 : 0u;
    const uint4 first = lds128(a0);
    if constexpr (NIB) {
        NibCounts na, nb;
        nib_clear(na);
        nib_clear(nb);
        // words at or beyond nwords hold stale bytes of any value: a stale high word must not carry into the valid
        // low word it is packed with, so its multiplier is 0 instead of 16 (a stale low word implies a stale high one)
        const uint32_t my = valid[1] ? 16u : 0u, mw = valid[3] ? 16u : 0u;
        const uint32_t fa = first.y * my + first.x, fb = first.w * mw + first.z;
#pragma unroll 3
        for (int r = 0; r < R; ++r) {
            const uint4 x = lds128(a0 + (uint32_t)(r * pitch));
            nib_add(na, x.y * my + x.x, fa);
            nib_add(nb, x.w * mw + x.z, fb);
        }
        nib_unpack(na, 0, cc[0]);
        nib_unpack(na, 1, cc[1]);
        nib_unpack(nb, 0, cc[2]);
        nib_unpack(nb, 1, cc[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) colcounts_clear(cc[k]);
        const uint32_t f2[4] = {first.x * 17u, first.y * 17u, first.z * 17u, first.w * 17u};
        ColCounts nn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) colcounts_clear(nn[k]);
        int r = 0;
        const int rp = R & ~1;
        while (r < rp) {                                      // epochs of at most 14 row pairs: 4-bit counters hold 15
            const int rend = min(rp, r + 28);
#pragma unroll 2
            for (; r < rend; r += 2) {
                const uint4 x0 = lds128(a0 + (uint32_t)(r * pitch)), x1 = lds128(a0 + (uint32_t)((r + 1) * pitch));
                colcounts_add2(nn[0], x0.x, x1.x, f2[0]);
                colcounts_add2(nn[1], x0.y, x1.y, f2[1]);
                colcounts_add2(nn[2], x0.z, x1.z, f2[2]);
                colcounts_add2(nn[3], x0.w, x1.w, f2[3]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) colcounts_flush2(cc[k], nn[k]);
        }
        if (r < R) {
            const uint4 x0 = lds128(a0 + (uint32_t)(r * pitch));
            colcounts_add(cc[0], x0.x, first.x);
            colcounts_add(cc[1], x0.y, first.y);
            colcounts_add(cc[2], x0.z, first.z);
            colcounts_add(cc[3], x0.w, first.w);
        }
    }
}

// consumer threads per CTA (+ one producer warp): the nibble path fits 64 registers, the row-paired byte path needs more
template <bool NIB> struct FtThreads { static constexpr int CONSUMERS = NIB ? 896 : 480; };
constexpr int FT_MAX_STAGES = 32;

// Launch geometry, computed on the host so that the kernel does no division (a thread scores only two or three boards
// per launch at c3: a 64-bit divide in its prologue costs as much as a board).
struct FtGeom {
    int BW;            // bytes of a row inside a box (multiple of 16, <= 256, <= stride)
    int NB;            // boards per box (power of two)
    int stages;        // ring depth, a multiple of nteams
    int stage_bytes;   // NB * R * BW rounded up to 128
    int nteams;        // teams of NB * GS consumer threads
    int spt;           // stages per team = stages / nteams
    int team_shift;    // log2(NB * GS)
    int nboxes;        // ceil(P / NB)
    int rev;           // odd boards walk their rows backwards (see the kernel: shared-memory bank conflicts)
};

// Fold the nibble counters of 8 columns (words w and w + 1 of the board; w even) into the board sums.
// hi_valid = word w + 1 is inside the board. Returns nothing; keep masks are recomputed when a board needs compaction.
__device__ __forceinline__ void nib_fold(BoardSums &b, const NibCounts &n, bool hi_valid)
{
    // nibble-level: non-gap cells per column (<= 15), kept = columns with any, exact = kept and no differing row
    const uint32_t s = n.n1 + n.n2 + n.n3 + n.n4;
    const uint32_t vmask = hi_valid ? 0x88888888u : 0x08080808u;
    const uint32_t kept = (((s & 0x77777777u) + 0x77777777u) | s) & vmask;
    const uint32_t differ = (((n.diff & 0x77777777u) + 0x77777777u) | n.diff) & 0x88888888u;
    b.kept += __popc(kept);
    b.exact += __popc(kept & ~differ);
    // squares need byte lanes: low nibbles = word w, high nibbles = word w + 1 (zero counts when that word is invalid
    // because its multiplier was 0 - except codes of word w + 1 read as 0 never count as a symbol)
    const uint32_t m = hi_valid ? 0xFFFFFFFFu : 0u;
    const uint32_t a1 = n.n1 & GAD_NIB4, a2 = n.n2 & GAD_NIB4, a3 = n.n3 & GAD_NIB4, a4 = n.n4 & GAD_NIB4;
    const uint32_t h1 = (n.n1 >> 4) & GAD_NIB4 & m, h2 = (n.n2 >> 4) & GAD_NIB4 & m, h3 = (n.n3 >> 4) & GAD_NIB4 & m,
                   h4 = (n.n4 >> 4) & GAD_NIB4 & m;
    b.q2 = __dp4a(a1, a1, b.q2);
    b.q2 = __dp4a(a2, a2, b.q2);
    b.q2 = __dp4a(a3, a3, b.q2);
    b.q2 = __dp4a(a4, a4, b.q2);
    b.q2 = __dp4a(h1, h1, b.q2);
    b.q2 = __dp4a(h2, h2, b.q2);
    b.q2 = __dp4a(h3, h3, b.q2);
    b.q2 = __dp4a(h4, h4, b.q2);
    const uint32_t sl = a1 + a2 + a3 + a4, sh = h1 + h2 + h3 + h4;
    b.s1 = __dp4a(sl, GAD_LSB4, b.s1);
    b.s1 = __dp4a(sh, GAD_LSB4, b.s1);
    b.s2 = __dp4a(sl, sl, b.s2);
    b.s2 = __dp4a(sh, sh, b.s2);
}

// keep mask (bit 7 per byte lane) of one 4-column word from nibble counters; hi selects the high nibbles
__device__ __forceinline__ uint32_t nib_keep_word(const NibCounts &n, int hi)
{
    const int sh = hi * 4;
    const uint32_t s = ((n.n1 >> sh) & GAD_NIB4) + ((n.n2 >> sh) & GAD_NIB4) + ((n.n3 >> sh) & GAD_NIB4) + ((n.n4 >> sh) & GAD_NIB4);
    return nonzero_lanes(s);
}

// GA.update_hall_of_fame in the tail of the fitness launch (gad_fitness_pass): hof == nullptr = scores only
struct FtHof {
    int *hof;                    // record {valid, index, sp, em, al, barrier counter, ticket, -}: words 5 and 6 are zero
    uint8_t *hof_board;          // between launches
    int mode, hof_valid;
    GadDecideParts parts;
    unsigned long long *stamps;  // measurement hook (gad_debug_stamps): where the tail leaves its timestamps, or nullptr
};

// timestamps of the decision tail, by thread 0 of CTA 0 (slots 0..7) and of the last CTA (8..11): globaltimer ns
__device__ __forceinline__ void ft_stamp(const FtHof &fh, int slot, bool who)
{
    if (fh.stamps != nullptr && who) fh.stamps[slot] = global_ns();
}

// What a CTA remembers of the boards it scored, for the decision in its tail: the score tuples and column scores
// (when its share fits) and, in 'mo' mode, their range - kept with shared-memory atomics by the lane that writes a
// board's scores, so the tail starts with the CTA's range in hand and never goes back to L2 for a score.
constexpr int FT_TAIL_CAP = 2048;  // c3: 443 boards per CTA, c4: 1 771
struct FtTail {
    int lo, hi;
    u64 clo, chi;                // bit patterns of non-negative doubles (order-preserving)
    int from_global;             // a board took the global-memory path, or the share exceeds FT_TAIL_CAP: re-read scores
    int rec[3];                  // CTA 0: the record's (sp, em, al), fetched while the boards stream in
    int win_sp, win_em, win_al;
    u64 cs[FT_TAIL_CAP];         // em / al as python computes it ('cs' and 'mo')
    int sp[FT_TAIL_CAP];
    int ea[FT_TAIL_CAP];         // em | al << 16 (boards of this kernel are at most 256 columns wide)
};

__device__ __forceinline__ void ft_tail_note(FtTail &t, int mode, int j, int s, int e, int a)
{
    const bool fits = j < FT_TAIL_CAP;
    if (fits) {
        t.sp[j] = s;
        t.ea[j] = e | (a << 16);
    }
    if (mode != GAD_MODE_SP) {
        const u64 c = (u64)__double_as_longlong(column_score(e, a));
        if (fits) t.cs[j] = c;
        if (mode == GAD_MODE_MO) {
            atomicMin(&t.lo, s);
            atomicMax(&t.hi, s);
            atomicMin(&t.clo, c);
            atomicMax(&t.chi, c);
        }
    }
}

// make_key with the column score already divided (the same IEEE operations in the same order)
__device__ __forceinline__ double ft_key(int mode, int s, u64 cbits, const Range &g)
{
    if (mode == GAD_MODE_SP) return (double)s;
    const double c = __longlong_as_double((long long)cbits);
    if (mode == GAD_MODE_CS) return c;
    const double clo = __longlong_as_double((long long)g.clo), chi = __longlong_as_double((long long)g.chi);
    const double nsp = g.hi > g.lo ? __ddiv_rn((double)((long long)s - g.lo), (double)((long long)g.hi - g.lo)) : 0.5;
    const double ncs = chi > clo ? __ddiv_rn(__dsub_rn(c, clo), __dsub_rn(chi, clo)) : 0.5;
    return __dadd_rn(nsp, ncs);
}

__device__ __forceinline__ void ft_red_release(int *p)
{
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ int ft_ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ft_ticket_acq_rel(int *p)
{
    int v;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The hall-of-fame decision (GA.py:110-136) by the consumer threads of every CTA once their boards are scored.
// Every step that goes to L2 and back costs about a microsecond, so the chain is kept short (measured with
// gad_debug_stamps, scripts/pass_timing.py): the CTA's score range is already in shared memory (ft_tail_note);
// 'mo' publishes it and arrives at ONE grid barrier with a release (the grid is one persistent CTA per SM, all
// resident; thread 0 speaks for the CTA after a CTA barrier, like a cooperative grid sync), polls it with acquire
// loads, and every WARP folds the partial ranges for itself (no block-wide fold); the keys of the CTA's own boards
// come from the shared-memory tuples with the arithmetic of make_key; the best per CTA is published with its
// tuple under an acq_rel ticket, and the last CTA reads every CTA's best and tuple in one trip, settles the first
// maximum and copies the winner's board. The record takes part as element P (CTA 0). Words 5 and 6 of the record
// are the barrier counter and the ticket: zero before the launch, left zero by it.
__device__ __noinline__ void ft_decide_tail(const FtHof &fh, FtTail &ts, int NB, int nk, int P, int R, int stride,
                                            const uint8_t *boards, const int32_t *sp, const int32_t *em, const int32_t *al,
                                            int tid, int n_consumers)
{
    __shared__ DecideShared sh;
    const BlockCtx cx{tid, n_consumers, 1};
    const int mode = fh.mode, lane = tid & 31, grid = (int)gridDim.x;
    const bool has_rec = fh.hof_valid != 0 && blockIdx.x == 0;          // the record is element P of CTA 0's share
    int *sync = fh.hof + 5;
    const int nown = nk * NB;
    cx.sync();                           // every board of this CTA is scored
    const bool st0 = blockIdx.x == 0 && tid == 0;
    ft_stamp(fh, 1, st0);
    const bool from_global = nown > FT_TAIL_CAP || *reinterpret_cast<volatile int *>(&ts.from_global) != 0;
    Range g;
    range_clear(g);
    if (mode == GAD_MODE_MO) {
        if (from_global) {               // (own writes: visible to the CTA after the barrier)
            for (int j = tid; j < nown; j += n_consumers) {
                const int k = j / NB, p = (blockIdx.x + k * grid) * NB + (j - k * NB);
                if (p < P) {
                    const int s = __ldcg(sp + p);
                    const u64 c = (u64)__double_as_longlong(column_score(__ldcg(em + p), __ldcg(al + p)));
                    range_merge(g, s, s, c, c);
                }
            }
            g = range_fold_block(g, sh, cx);
        } else {
            g.lo = ts.lo, g.hi = ts.hi, g.clo = ts.clo, g.chi = ts.chi;
        }
        if (has_rec) {
            const u64 c = (u64)__double_as_longlong(column_score(ts.rec[1], ts.rec[2]));
            range_merge(g, ts.rec[0], ts.rec[0], c, c);
        }
        if (grid > 1) {
            uint4 *rec = fh.parts.rec;   // 64 bytes per CTA: {lo, hi, clo, chi} | {key, idx, sp, em, al}
            if (tid == 0) {
                rec[4 * blockIdx.x] = make_uint4((unsigned)g.lo, (unsigned)g.hi, (unsigned)g.clo, (unsigned)(g.clo >> 32));
                rec[4 * blockIdx.x + 1] = make_uint4((unsigned)g.chi, (unsigned)(g.chi >> 32), 0u, 0u);
                ft_stamp(fh, 2, st0);
                ft_red_release(&sync[0]);                // the CTA's scores, compacted cells and partials, then the arrival
                const unsigned long long t0 = global_ns();
                unsigned spins = 0;
                while (ft_ld_acquire(&sync[0]) < grid)   // all CTAs are resident; a lost arrival fails loudly
                    if ((++spins & 1023u) == 0u && global_ns() - t0 > 2000000000ull) __trap();
                ft_stamp(fh, 3, st0);
            }
            cx.sync();
            // one warp reads every CTA's range - all loads in flight at once, 32 bytes per CTA (148 CTAs reading the
            // same few lines: 28 warps each doing so queue up in L2 for microseconds) - and hands the fold to the rest
            if (tid < 32) {
                range_clear(g);
                uint4 ra[5], rb[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int b = min(lane + 32 * q, grid - 1);
                    ra[q] = __ldcg(&rec[4 * b]);
                    rb[q] = __ldcg(&rec[4 * b + 1]);
                }
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    range_merge(g, (int)ra[q].x, (int)ra[q].y, (u64)ra[q].z | ((u64)ra[q].w << 32), (u64)rb[q].x | ((u64)rb[q].y << 32));
                for (int b = lane + 160; b < grid; b += 32) {        // more than one CTA per SM (GAD_FT_CTAS)
                    const uint4 a4 = __ldcg(&rec[4 * b]), b4 = __ldcg(&rec[4 * b + 1]);
                    range_merge(g, (int)a4.x, (int)a4.y, (u64)a4.z | ((u64)a4.w << 32), (u64)b4.x | ((u64)b4.y << 32));
                }
                range_fold_warp(g);
                if (lane == 0) sh.range[0] = g;
            }
            cx.sync();
            g = sh.range[0];
        }
    }
    ft_stamp(fh, 5, st0);
    double bk = 0.0;
    int bi = 0x7fffffff;
    for (int j = tid; j < nown; j += n_consumers) {
        const int k = j / NB, p = (blockIdx.x + k * grid) * NB + (j - k * NB);
        if (p < P) {
            double key;
            if (from_global) key = make_key(mode, __ldcg(sp + p), __ldcg(em + p), __ldcg(al + p), g.lo, g.hi, g.clo, g.chi);
            else key = ft_key(mode, ts.sp[j], mode != GAD_MODE_SP ? ts.cs[j] : 0ull, g);
            best_merge(bk, bi, key, p);
        }
    }
    if (has_rec && tid == 0) best_merge(bk, bi, make_key(mode, ts.rec[0], ts.rec[1], ts.rec[2], g.lo, g.hi, g.clo, g.chi), P);
    best_fold_block(bk, bi, sh, cx);     // the best of the CTA, in warp 0
    ft_stamp(fh, 6, st0);
    if (tid == 0) {
        sh.last = true;
        sh.winner = bi;
        if (bi == P) {
            ts.win_sp = ts.rec[0], ts.win_em = ts.rec[1], ts.win_al = ts.rec[2];
        } else if (bi != 0x7fffffff) {
            const int box = bi / NB, j = (box - (int)blockIdx.x) / grid * NB + (bi - box * NB);
            if (from_global) ts.win_sp = __ldcg(sp + bi), ts.win_em = __ldcg(em + bi), ts.win_al = __ldcg(al + bi);
            else ts.win_sp = ts.sp[j], ts.win_em = ts.ea[j] & 0xFFFF, ts.win_al = ts.ea[j] >> 16;
        }
        if (grid > 1) {
            const u64 kb = (u64)__double_as_longlong(bk);
            fh.parts.rec[4 * blockIdx.x + 2] = make_uint4((unsigned)kb, (unsigned)(kb >> 32), (unsigned)bi, (unsigned)ts.win_sp);
            fh.parts.rec[4 * blockIdx.x + 3] = make_uint4((unsigned)ts.win_em, (unsigned)ts.win_al, 0u, 0u);
            // release: for the whole CTA (the barrier above) - cells compacted by it, its partials; acquire: the others'
            sh.last = ft_ticket_acq_rel(&sync[1]) == grid - 1;
        }
    }
    ft_stamp(fh, 7, st0);
    cx.sync();
    if (!sh.last) return;
    ft_stamp(fh, 8, tid == 0);
    if (grid > 1) {
        bk = 0.0;
        bi = 0x7fffffff;
        int ws = 0, we = 0, wa = 0;
        for (int b = tid; b < grid; b += n_consumers) {
            const uint4 a4 = __ldcg(&fh.parts.rec[4 * b + 2]), b4 = __ldcg(&fh.parts.rec[4 * b + 3]);
            const double key = __longlong_as_double((long long)((u64)a4.x | ((u64)a4.y << 32)));
            const int i = (int)a4.z;
            if (i != 0x7fffffff && (bi == 0x7fffffff || better(key, i, bk, bi)))
                bk = key, bi = i, ws = (int)a4.w, we = (int)b4.x, wa = (int)b4.y;
        }
        const int cand = bi;
        best_fold_block(bk, bi, sh, cx);
        if (tid == 0) sh.winner = bi;
        cx.sync();
        if (cand == sh.winner && cand != 0x7fffffff) ts.win_sp = ws, ts.win_em = we, ts.win_al = wa;     // indices are distinct
    }
    ft_stamp(fh, 10, tid == 0);
    if (tid == 0) {
        sync[0] = 0;
        sync[1] = 0;
        *fh.parts.winner = sh.winner;
    }
    cx.sync();
    // GA.py:110-136: the winner's board and score tuple become the hall of fame; winner == P: the record keeps its
    // place under the fresh index len(scores) (GA.py:127-132)
    const int win = sh.winner;
    if (win >= P) {
        if (tid == 0) fh.hof[1] = P;
        return;
    }
    const int width = ts.win_al;
    const int n16 = min(stride >> 4, (width + 15) >> 4);
    for (int u = tid; u < R * n16; u += n_consumers) {
        const int r = u / n16, q = u - r * n16;
        reinterpret_cast<uint4 *>(fh.hof_board + (size_t)r * stride)[q] =
            __ldcg(reinterpret_cast<const uint4 *>(boards + ((size_t)win * R + r) * stride) + q);
    }
    if (tid == 0) {
        fh.hof[0] = 1;
        fh.hof[1] = win;
        fh.hof[2] = ts.win_sp;
        fh.hof[3] = ts.win_em;
        fh.hof[4] = width;
    }
    ft_stamp(fh, 11, tid == 0);
}

// RT > 0: the row count is RT (loops fully unrolled, row lengths read as vectors); RT = 0: any R
template <int GS, bool NIB, int RT>
__global__ void __launch_bounds__(FtThreads<NIB>::CONSUMERS + 32, 1)
fitness_tma_kernel(const __grid_constant__ CUtensorMap map, uint8_t *__restrict__ boards, int32_t *__restrict__ rowlen,
                   int P, int R_rt, int stride, FtGeom geo, int match, int mismatch, int gap,
                   int32_t *__restrict__ sp, int32_t *__restrict__ em, int32_t *__restrict__ al,
                   unsigned long long *__restrict__ max_al, unsigned long long max_tag, const __grid_constant__ FtHof fh)
{
    extern __shared__ uint8_t ft_smem_raw[];
    const int R = RT > 0 ? RT : R_rt;
    const uint32_t smem0 = (ft_smem_u32(ft_smem_raw) + 127u) & ~127u;
    const uint32_t full0 = smem0 + (uint32_t)(geo.stages * geo.stage_bytes), empty0 = full0 + 8u * geo.stages;
    const int tid = threadIdx.x;
    const int n_consumers = geo.nteams << geo.team_shift;
    // boxes of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ... (32-bit: P < 2^31)
    const int nk = geo.nboxes > (int)blockIdx.x ? (int)((unsigned)(geo.nboxes - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    // Programmatic dependent launch: the next launch in the stream may start its prologue (CTA launch, barrier
    // set-up) while this grid is still running; it blocks at its own griddepcontrol.wait until this grid has
    // completed and its writes are visible. Without the launch attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ FtTail ts;                                         // used only when the launch takes the decision (fh.hof)
    if (tid == 0) {
        ts.lo = 0x7fffffff;
        ts.hi = (int)0x80000000;
        ts.clo = ~0ull;
        ts.chi = 0ull;
        ts.from_global = 0;
        ts.win_sp = ts.win_em = ts.win_al = 0;
        for (int s2 = 0; s2 < geo.stages; ++s2) {
            ft_mbar_init(full0 + 8u * s2, 1);
            ft_mbar_init(empty0 + 8u * s2, 1u << (geo.team_shift - 5));          // one arrival per warp of the owning team
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");            // nothing before this line touched global memory
    if (fh.hof != nullptr) {
        ft_stamp(fh, 0, blockIdx.x == 0 && tid == 0);
        if (blockIdx.x == 0 && tid < 3 && fh.hof_valid) ts.rec[tid] = fh.hof[2 + tid];      // read long before the tail needs it
    }

    if (tid >= n_consumers) {
        // ===================== producer: one thread streams this CTA's boxes =====================
        if (tid == n_consumers) {
            const uint32_t box_bytes = (uint32_t)(geo.NB * R * geo.BW);
            int stage = 0, round = 0;
            for (int k = 0; k < nk; ++k) {
                if (round > 0) ft_mbar_wait(empty0 + 8u * stage, (uint32_t)((round - 1) & 1));
                const int box = blockIdx.x + k * gridDim.x;
                ft_mbar_expect_tx(full0 + 8u * stage, box_bytes);
                ft_tma_load_3d(smem0 + (uint32_t)(stage * geo.stage_bytes), &map, full0 + 8u * stage, 0, 0, box * geo.NB);
                if (++stage == geo.stages) {
                    stage = 0;
                    ++round;
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const uint32_t gmask = group_mask<GS>();
    const int team = tid >> geo.team_shift, tt = tid & ((1 << geo.team_shift) - 1);
    const int b = tt / GS, gl = tt & (GS - 1);          // GS is a power of two: a shift
    const int lane = tid & 31;
    const int pitch = geo.BW;
    int local_max = 0;
    // Stage s always belongs to team s % nteams (stages is a multiple of nteams), and a team walks its stages round
    // by round: nobody ever waits on a barrier phase that is more than one completion away (an mbarrier parity
    // wait cannot tell phase n from phase n + 2).
    // (width, shortest row) of this thread's board in the team's box k, from the row lengths in global memory.
    // Called one box ahead, so the DRAM round trip overlaps the previous box's arithmetic (and, for the first box,
    // the flight of the box itself).
    auto load_widths = [&](int k, int &w, int &wmin) {
        w = 0;
        wmin = 0x7fffffff;
        if (k >= nk) return;
        const int p = (blockIdx.x + k * gridDim.x) * geo.NB + b;
        if (p >= P) return;
        const int32_t *rl = rowlen + (size_t)p * R;
        if constexpr (RT > 0 && RT % 2 == 0) {                    // every lane reads all of them: no shuffles
            const int2 *rl2 = reinterpret_cast<const int2 *>(rl);
#pragma unroll
            for (int r = 0; r < RT / 2; ++r) {
                const int2 v = rl2[r];
                w = max(w, max(v.x, v.y));
                wmin = min(wmin, min(v.x, v.y));
            }
        } else if constexpr (RT > 0) {
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                const int v = rl[r];
                w = max(w, v);
                wmin = min(wmin, v);
            }
        } else {
            for (int r = gl; r < R; r += GS) {
                const int v = rl[r];
                w = max(w, v);
                wmin = min(wmin, v);
            }
            w = group_max<GS>(w, gmask);
            wmin = -group_max<GS>(-wmin, gmask);
        }
    };
    int idx = 0, round = 0;
    int w_next, wmin_next;
    load_widths(team, w_next, wmin_next);
    for (;;) {
        const int stage = team + idx * geo.nteams;
        const int k = round * geo.stages + stage;
        if (k >= nk) break;                                       // k grows monotonically: nothing left for this team
        const uint32_t parity = (uint32_t)(round & 1);
        if (++idx == geo.spt) {
            idx = 0;
            ++round;
        }
        const int box = blockIdx.x + k * gridDim.x;
        const int p = box * geo.NB + b;
        const bool live = p < P;
        int32_t *rl = rowlen + (size_t)(live ? p : 0) * R;
        const int w = w_next, wmin = wmin_next;
        load_widths(round * geo.stages + team + idx * geo.nteams, w_next, wmin_next);      // the team's next box
        ft_mbar_wait(full0 + 8u * stage, parity);
        if (live) {
            uint8_t *board = boards + (size_t)p * R * stride;
            const uint32_t sb = smem0 + (uint32_t)(stage * geo.stage_bytes) + (uint32_t)(b * R * pitch);
            const int nwords = (w + 3) >> 2;
            const int nunits = (nwords + 3) >> 2;
            int nk_cols;
            if (w <= geo.BW) {
                BoardSums bs;
                boardsums_clear(bs);
                if constexpr (NIB) {
                    // 64-byte rows, an even row count, two boards per quarter-warp (c3): the same unit of two neighbouring
                    // boards lies a multiple of 128 bytes apart, so every lds128 would be a 2-way bank conflict (ncu: half
                    // of the shared-load wavefronts). Odd boards walk their rows backwards: rows r and R-1-r are an odd
                    // multiple of 64 bytes apart, the pair covers all 32 banks. The counts do not depend on the row
                    // order, and any fixed row serves as the reference of the "all rows equal" test.
                    const bool back = geo.rev != 0 && (b & 1) != 0;
                    const int rstep = back ? -pitch : pitch;
                    const uint32_t sb0 = back ? sb + (uint32_t)((R - 1) * pitch) : sb;
                    for (int u = gl; u < nunits; u += GS) {
                        const uint32_t a0 = sb0 + 16u * u;
                        // words at or beyond nwords hold stale bytes of any value: a stale high word must not carry into
                        // the valid low word it is packed with, so its multiplier is 0 instead of 16
                        const bool v1 = 4 * u + 1 < nwords, v2 = 4 * u + 2 < nwords, v3 = 4 * u + 3 < nwords;
                        const uint32_t my = v1 ? 16u : 0u, mw = v3 ? 16u : 0u;
                        const uint4 first = lds128(a0);
                        const uint32_t fa = first.y * my + first.x, fb = first.w * mw + first.z;
                        NibCounts na, nb;
                        nib_clear(na);
                        nib_clear(nb);
                        if constexpr (RT > 0) {
#pragma unroll
                            for (int r = 0; r < RT; ++r) {
                                const uint4 x = lds128(a0 + (uint32_t)(r * rstep));
                                nib_add(na, x.y * my + x.x, fa);
                                nib_add(nb, x.w * mw + x.z, fb);
                            }
                        } else {
#pragma unroll 2
                            for (int r = 0; r < R; ++r) {
                                const uint4 x = lds128(a0 + (uint32_t)(r * rstep));
                                nib_add(na, x.y * my + x.x, fa);
                                nib_add(nb, x.w * mw + x.z, fb);
                            }
                        }
                        nib_fold(bs, na, v1);                     // word 4u is valid because u < nunits
                        if (v2) nib_fold(bs, nb, v3);
                    }
                } else {
                    for (int u = gl; u < nunits; u += GS) {
                        ColCounts cc[4];
                        uint32_t valid[4];
                        count_unit_smem<false>(sb, R, pitch, u, nwords, cc, valid);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) boardsums_fold_word(bs, cc[kk], valid[kk] != 0u);
                    }
                }
                // group reduction of the five sums, packed into two words while they fit 16 bits (narrow boards, few rows)
                if (NIB) {
                    uint32_t pa = bs.q2 | (bs.s2 << 16), pb = bs.s1 | (bs.kept << 12) | (bs.exact << 22);
                    pa = group_sum<GS>(pa, gmask);
                    pb = group_sum<GS>(pb, gmask);
                    bs.q2 = pa & 0xFFFFu;
                    bs.s2 = pa >> 16;
                    bs.s1 = pb & 0xFFFu;
                    bs.kept = (pb >> 12) & 0x3FFu;
                    bs.exact = pb >> 22;
                } else {
                    bs.q2 = group_sum<GS>(bs.q2, gmask);
                    bs.s1 = group_sum<GS>(bs.s1, gmask);
                    bs.s2 = group_sum<GS>(bs.s2, gmask);
                    bs.kept = group_sum<GS>(bs.kept, gmask);
                    bs.exact = group_sum<GS>(bs.exact, gmask);
                }
                nk_cols = (int)bs.kept;
                if (nk_cols != w && nunits > GS) {
                    // more units than lanes and columns to drop: the in-place path on global memory (rare, wide boards)
                    const uint4 none[7] = {};
                    nk_cols = board_pass_global<GS, !NIB, false>(board, rl, R, stride, gl, gmask, w, wmin, none, false, match,
                                                                mismatch, gap, sp + p, em + p, al + p);
                    ts.from_global = 1;
                } else {
                    if (nk_cols != w) {
                        // utils.py:408-428 from the shared-memory copy: every lane owns one unit; kept bytes go to
                        // their final place in global memory (no in-place hazard)
                        const int u = gl;
                        uint32_t keep[4] = {0u, 0u, 0u, 0u};
                        if (u < nunits) {
                            ColCounts cc[4];
                            uint32_t valid[4];
                            count_unit_smem<NIB>(sb, R, pitch, u, nwords, cc, valid);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                keep[kk] = nonzero_lanes(cc[kk].c1 + cc[kk].c2 + cc[kk].c3 + cc[kk].c4) & valid[kk];
                        }
                        const int mine = __popc(keep[0]) + __popc(keep[1]) + __popc(keep[2]) + __popc(keep[3]);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < GS; o <<= 1) {
                            const int t = __shfl_up_sync(gmask, incl, o, GS);
                            if (gl >= o) incl += t;
                        }
                        const int dst0 = incl - mine;
                        if (u < nunits) {
                            for (int r = 0; r < R; ++r) {
                                const uint4 x = lds128(sb + (uint32_t)(r * pitch) + 16u * u);
                                uint8_t *row = board + (size_t)r * stride;
                                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
                                int d = dst0;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                                    for (int bb = 0; bb < 4; ++bb)
                                        if (keep[kk] & (0x80u << (8 * bb))) row[d++] = (uint8_t)(xs[kk] >> (8 * bb));
                            }
                        }
                        __syncwarp(gmask);
                        const int pad_to = (nk_cols + 3) & ~3;      // the padding invariant of the shrunken board
                        for (int i = gl; i < R * 4; i += GS) {
                            const int r = i >> 2, c = nk_cols + (i & 3);
                            if (c < pad_to) board[(size_t)r * stride + c] = GAD_GAP;
                        }
                    }
                    if (nk_cols != w || wmin != w)
                        for (int r = gl; r < R; r += GS) rl[r] = nk_cols;
                    if (gl == 0) {
                        const int32_t spv = (int32_t)sum_of_pairs_closed_form(bs, R, match, mismatch, gap);
                        sp[p] = spv;
                        em[p] = (int32_t)bs.exact;
                        al[p] = nk_cols;
                        if (fh.hof != nullptr) ft_tail_note(ts, fh.mode, k * geo.NB + b, spv, (int32_t)bs.exact, nk_cols);
                    }
                }
            } else {
                // wider than the box (the width hint was too small): everything from global memory
                const uint4 none[7] = {};
                nk_cols = board_pass_global<GS, !NIB, false>(board, rl, R, stride, gl, gmask, w, wmin, none, false, match, mismatch,
                                                            gap, sp + p, em + p, al + p);
                ts.from_global = 1;
            }
            local_max = max(local_max, nk_cols);
        }
        __syncwarp();
        if (lane == 0) ft_mbar_arrive(empty0 + 8u * stage);          // this warp is done with the stage
    }
    if (max_al != nullptr) {
        local_max = group_max<32>(local_max, 0xFFFFFFFFu);
        if (lane == 0) atomicMax(max_al, max_tag | (unsigned long long)local_max);
    }
    if (fh.hof != nullptr) ft_decide_tail(fh, ts, geo.NB, nk, P, R, stride, boards, sp, em, al, tid, n_consumers);
}

typedef CUresult (*FtEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Returns GAD_OK after launching, or 1 when the shape is outside the TMA path (the caller launches the register kernel).
template <int GS, bool NIB>
int launch_fitness_tma(uint8_t *boards, int32_t *rowlen, long long P, int R, int stride, int width, int match, int mismatch,
                       int gap, int32_t *sp, int32_t *em, int32_t *al, unsigned long long *max_al, unsigned long long max_tag,
                       const FtHof &fh, cudaStream_t st)
{
    static FtEncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || !fp ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return 1;
        }
        encode = (FtEncodeTiledFn)fp;
    }
    if (P >= (1ll << 30)) return 1;                                  // the kernel indexes boards with 32-bit integers
    FtGeom g;
    g.BW = (width + 15) / 16 * 16;
    if (g.BW > stride) g.BW = stride;
    const int board_bytes = R * g.BW;
    int nb = 32 / GS > 1 ? 32 / GS : 1;                              // a team is whole warps
    static const int box_target = [] { const char *e = getenv("GAD_FT_BOX_BYTES"); return e && atoi(e) > 0 ? atoi(e) : 6144; }();
    while (nb * 2 * board_bytes <= box_target && nb * 2 <= 256 && nb * 2 * GS <= 256) nb *= 2;
    g.NB = nb;
    g.stage_bytes = (nb * board_bytes + 127) / 128 * 128;
    // tuning knobs (A/B): CTAs per SM, shared memory and consumer threads per CTA
    static const int ctas_env = [] { const char *e = getenv("GAD_FT_CTAS"); return e && atoi(e) > 0 ? atoi(e) : 1; }();
    // a launch that takes the hall-of-fame decision meets at a grid barrier: every CTA must be resident, and with the
    // tail's static shared memory only one CTA fits an SM
    const int ctas_per_sm = fh.hof != nullptr ? 1 : ctas_env;
    static const int smem_kb = [] {                                  // + 34 KB of static shared memory (FtTail, DecideShared)
        const char *e = getenv("GAD_FT_SMEM_KB");
        return e && atoi(e) > 0 ? (atoi(e) < 190 ? atoi(e) : 190) : 190;
    }();
    static const int max_cons = [] { const char *e = getenv("GAD_FT_CONSUMERS"); return e && atoi(e) > 0 ? atoi(e) : 1 << 20; }();
    g.stages = (smem_kb * 1024 / ctas_per_sm) / g.stage_bytes;
    if (g.stages > FT_MAX_STAGES) g.stages = FT_MAX_STAGES;
    int cons = FtThreads<NIB>::CONSUMERS / ctas_per_sm / 32 * 32;
    if (cons > max_cons) cons = max_cons;
    g.nteams = cons / (nb * GS);
    if (g.stages < 2 || g.nteams < 1 || (nb * GS) % 32 != 0) return 1;
    g.nboxes = (int)((P + nb - 1) / nb);
    // fewer boxes than teams x SMs: shrink the team count so that the CTAs stay balanced
    const int max_grid = GAD_NUM_SMS * ctas_per_sm;
    const int grid = g.nboxes < max_grid ? g.nboxes : max_grid;
    const int per_cta = (g.nboxes + grid - 1) / grid;
    if (per_cta < g.nteams) g.nteams = per_cta;
    if (g.nteams > g.stages) g.nteams = g.stages;
    g.stages = g.stages / g.nteams * g.nteams;                     // every stage has one owning team (see the kernel)
    g.spt = g.stages / g.nteams;
    g.team_shift = 0;
    while ((1 << g.team_shift) < nb * GS) ++g.team_shift;
    static const int allow_rev = [] { const char *e = getenv("GAD_FT_REVERSE"); return e ? atoi(e) : 1; }();      // A/B switch
    g.rev = allow_rev && NIB && GS == 4 && g.BW == 64 && R % 2 == 0;

    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)stride, (cuuint64_t)R, (cuuint64_t)P};
    const cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)stride * R};
    const cuuint32_t box[3] = {(cuuint32_t)g.BW, (cuuint32_t)R, (cuuint32_t)nb};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, boards, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return 1;
    const int smem = g.stages * g.stage_bytes + 16 * g.stages + 128;
    const int threads = g.nteams * nb * GS + 32;
    static const int use_pdl = [] { const char *e = getenv("GAD_FT_PDL"); return e ? atoi(e) : 1; }();      // A/B switch
#define GAD_FT_LAUNCH(RT)                                                                                                  \
    do {                                                                                                                   \
        GAD_CUDA(cudaFuncSetAttribute(fitness_tma_kernel<GS, NIB, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      227 * 1024 - 36 * 1024));                                                                 \
        cudaLaunchConfig_t cfg = {};                                                                                       \
        cfg.gridDim = dim3((unsigned)grid);                                                                                \
        cfg.blockDim = dim3((unsigned)threads);                                                                            \
        cfg.dynamicSmemBytes = (size_t)smem;                                                                               \
        cfg.stream = st;                                                                                                   \
        cudaLaunchAttribute attr[1];                                                                                       \
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                   \
        attr[0].val.programmaticStreamSerializationAllowed = 1;                                                            \
        cfg.attrs = attr;                                                                                                  \
        cfg.numAttrs = use_pdl ? 1 : 0;                                                                                    \
        GAD_CUDA(cudaLaunchKernelEx(&cfg, fitness_tma_kernel<GS, NIB, RT>, map, boards, rowlen, (int)P, R, stride, g,      \
                                    match, mismatch, gap, sp, em, al, max_al, max_tag, fh));                               \
    } while (0)
    if (NIB && R == 6) GAD_FT_LAUNCH(NIB ? 6 : 0);
    else if (NIB && R == 4) GAD_FT_LAUNCH(NIB ? 4 : 0);
    else if (NIB && R == 3) GAD_FT_LAUNCH(NIB ? 3 : 0);
    else GAD_FT_LAUNCH(0);
#undef GAD_FT_LAUNCH
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

template <int GS>
int launch_fitness(uint8_t *boards, int32_t *rowlen, long long P, int R, int stride, int match, int mismatch,
                   int gap, int32_t *sp, int32_t *em, int32_t *al, unsigned long long *max_al, unsigned long long max_tag, cudaStream_t st)
{
    const int threads = 256;
    const long long groups_per_block = threads / GS;
    long long blocks = (P + groups_per_block - 1) / groups_per_block;
    const long long cap = (long long)GAD_NUM_SMS * 8 * 4;          // 8 resident CTAs/SM, <=4 passes each
    if (blocks > cap) blocks = cap;
    static const int force_paired = [] {                            // tuning override: GAD_FITNESS_PAIRED=0|1
        const char *e = getenv("GAD_FITNESS_PAIRED");
        return e ? atoi(e) : -1;
    }();
    if (force_paired < 0 ? R >= 8 : force_paired != 0)
        fitness_kernel<GS, true><<<(unsigned)blocks, threads, 0, st>>>(boards, rowlen, P, R, stride, match, mismatch,
                                                                        gap, sp, em, al, max_al, max_tag);
    else
        fitness_kernel<GS, false><<<(unsigned)blocks, threads, 0, st>>>(boards, rowlen, P, R, stride, match, mismatch,
                                                                         gap, sp, em, al, max_al, max_tag);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

}  // namespace

// gad_fitness, and - with a decision to take (fh.hof != nullptr) - the whole GA.calculate_fitness_score: *fused says
// whether the launch took the hall-of-fame decision in its tail (the TMA-staged kernel) or the caller still has to
static int fitness_dispatch(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride, int width_hint,
                            int match, int mismatch, int gap, int32_t *d_sp, int32_t *d_em, int32_t *d_al,
                            int64_t *d_max_al, const FtHof &fh, bool *fused, void *stream)
{
    *fused = false;
    GAD_REQUIRE(P >= 0 && R >= 1 && R <= 255, "gad_fitness: need P >= 0 and 1 <= R <= 255 (got P=%lld R=%d)",
                (long long)P, R);
    GAD_REQUIRE(stride >= 16 && stride % 16 == 0, "gad_fitness: stride must be a positive multiple of 16 (got %d)",
                stride);
    GAD_REQUIRE((long long)R * R * stride < (1ll << 31), "gad_fitness: R*R*stride overflows the 32-bit column sums");
    GAD_REQUIRE(d_boards && d_rowlen && d_sp && d_em && d_al, "gad_fitness: null device pointer");
    if (P == 0) return GAD_OK;
    cudaStream_t st = gad_stream(stream);
    // max(al) without a memset per pass: every launch ORs a fresh, growing tag into the bits above the width, so the
    // atomicMax of this pass beats whatever an earlier pass left behind (44-bit tag: never wraps in practice)
    static std::atomic<unsigned long long> launch_tag{0};
    unsigned long long *max_al = reinterpret_cast<unsigned long long *>(d_max_al);
    const unsigned long long max_tag = max_al ? (++launch_tag) << GAD_MAX_AL_BITS : 0ull;
    GAD_REQUIRE(!max_al || stride < (1 << GAD_MAX_AL_BITS), "gad_fitness: stride %d does not fit the max_al encoding", stride);
    // lanes per board from the expected width (a performance hint only: wider boards take more trips)
    int width = width_hint > 0 && width_hint < stride ? width_hint : stride;
    int units = (width + 15) / 16;
    // wide boards in a large population: two 16-byte units per lane (more loads in flight per lane, half the
    // per-board reduction overhead) - measured on c4: 56.5 % -> 62.6 % of the HBM peak
    if (units >= 8 && P >= 4096) units = (units + 1) / 2;
    if (const char *g = getenv("GAD_FITNESS_GS")) units = atoi(g) > 0 ? atoi(g) : units;      // tuning override: lanes per board
    // TMA-staged kernel (default): live width up to 256 bytes, 16-byte aligned board base, enough boards to fill the SMs.
    // GAD_FITNESS_TMA=0 keeps the register-staged kernel (A/B).
    const char *tma_env = getenv("GAD_FITNESS_TMA");
    const int use_tma = tma_env ? atoi(tma_env) : 1;
    if (use_tma && width <= 256 && ((uintptr_t)d_boards & 15) == 0 && P >= 2048) {
        int rc = 1;
#define GAD_FIT_TMA(GS)                                                                                                   \
    rc = R <= 15 ? launch_fitness_tma<GS, true>(d_boards, d_rowlen, P, R, stride, width, match, mismatch, gap, d_sp, d_em, \
                                                d_al, max_al, max_tag, fh, st)                                             \
                 : launch_fitness_tma<GS, false>(d_boards, d_rowlen, P, R, stride, width, match, mismatch, gap, d_sp,      \
                                                 d_em, d_al, max_al, max_tag, fh, st)
        // 16+ rows (byte-lane counters, 128 registers, 480 consumer threads): one unit per lane - boxes of two boards
        // instead of four, so that every one of the 15 consumer warps owns a stage (c4: 11 teams -> 15, 0.66 -> 0.69
        // of the HBM peak); GAD_FITNESS_GS still overrides
        int tu = units;
        if (R >= 16 && !getenv("GAD_FITNESS_GS")) tu = (width + 15) / 16;
        if (tu <= 1) GAD_FIT_TMA(1);
        else if (tu <= 2) GAD_FIT_TMA(2);
        else if (tu <= 4) GAD_FIT_TMA(4);
        else if (tu <= 8) GAD_FIT_TMA(8);
        else GAD_FIT_TMA(16);
#undef GAD_FIT_TMA
        if (rc == 0) *fused = fh.hof != nullptr;
        if (rc <= 0) return rc;                            // launched (0) or failed (< 0); 1 = shape outside the TMA path
    }
#define GAD_FIT(GS) launch_fitness<GS>(d_boards, d_rowlen, P, R, stride, match, mismatch, gap, d_sp, d_em, d_al, max_al, max_tag, st)
    if (units <= 1) return GAD_FIT(1);
    if (units <= 2) return GAD_FIT(2);
    if (units <= 4) return GAD_FIT(4);
    if (units <= 8) return GAD_FIT(8);
    if (units <= 16) return GAD_FIT(16);
    return GAD_FIT(32);
#undef GAD_FIT
}

extern "C" int gad_fitness(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride, int width_hint,
                           int match, int mismatch, int gap, int32_t *d_sp, int32_t *d_em, int32_t *d_al,
                           int64_t *d_max_al, void *stream)
{
    bool fused;
    return fitness_dispatch(d_boards, d_rowlen, P, R, stride, width_hint, match, mismatch, gap, d_sp, d_em, d_al, d_max_al,
                            FtHof{}, &fused, stream);
}

// GA.calculate_fitness_score (GA.py:138-181) = gad_fitness + gad_hof_update. With the TMA-staged kernel (boards up
// to 256 columns, 2 048+ individuals) both are ONE launch: the CTAs that scored the boards take the decision in
// their tail (ft_decide_tail). Other shapes: the fitness launch followed by the decision kernel. Same result
// either way; GAD_FITNESS_FUSED_HOF=0 always takes the two launches (A/B).
// Measurement hook: a device buffer of 16 uint64 where the decision tail of the next gad_fitness_pass launches leaves
// globaltimer stamps (nullptr switches it off again). Slots: 0 kernel entry, 1 tail entry, 2 before the arrival,
// 3 barrier passed, 4 fenced, 5 ranges folded, 6 CTA's best known, 7 ticket taken (all CTA 0); 8-11 the last CTA:
// entry, fenced, winner known, board copied.
static unsigned long long *g_tail_stamps = nullptr;
extern "C" int gad_debug_stamps(void *d_buf16)
{
    g_tail_stamps = reinterpret_cast<unsigned long long *>(d_buf16);
    return GAD_OK;
}

extern "C" int gad_fitness_pass(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride, int width_hint,
                                int match, int mismatch, int gap, int32_t *d_sp, int32_t *d_em, int32_t *d_al,
                                int64_t *d_max_al, int mode, uint8_t *d_hof_board, int32_t *d_hof, int hof_valid,
                                void *d_ws, size_t ws_bytes, void *stream)
{
    GAD_REQUIRE(d_hof_board && d_hof, "gad_fitness_pass: null device pointer");
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_fitness_pass: unknown mode %d", mode);
    GAD_REQUIRE(P >= 1, "gad_fitness_pass: empty population");
    static const bool allow = !(getenv("GAD_FITNESS_FUSED_HOF") && atoi(getenv("GAD_FITNESS_FUSED_HOF")) == 0);
    FtHof fh{};
    if (allow) {
        int rc = gad_decide_parts(d_ws, ws_bytes, P + 1, "gad_fitness_pass", &fh.parts);
        if (rc) return rc;
        fh.hof = d_hof;
        fh.hof_board = d_hof_board;
        fh.mode = mode;
        fh.hof_valid = hof_valid ? 1 : 0;
        fh.stamps = g_tail_stamps;
    }
    bool fused = false;
    int rc = fitness_dispatch(d_boards, d_rowlen, P, R, stride, width_hint, match, mismatch, gap, d_sp, d_em, d_al, d_max_al,
                              fh, &fused, stream);
    if (rc || fused) return rc;
    return gad_hof_update(d_boards, d_sp, d_em, d_al, P, R, stride, mode, d_hof_board, d_hof, hof_valid, d_ws, ws_bytes, stream);
}

// ---------------------------------------------------------------------------------------------
// Measurement hook (bench.py): what a launch that only READS the live cells of the population costs.
// One uint4 (16 columns) per thread and trip, consecutive threads on consecutive units of a row, every
// SM full, no arithmetic beyond an XOR, one 16-byte result per block - the speed of light of a single
// pass over these bytes at this launch size. The fitness kernel is reported against it next to the
// HBM peak: at c3 a launch carries only 25 MB, which the memory system can stream in about 4 us, so
// launch ramp and drain (several microseconds per launch whatever the kernel does) set the floor.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256, 8)
probe_read_kernel(const uint8_t *__restrict__ boards, long long n_items, int units, int row_units, uint4 *__restrict__ sink)
{
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += step) {
        const long long row = i / units;
        const int u = (int)(i - row * units);
        const uint4 x = reinterpret_cast<const uint4 *>(boards)[row * row_units + u];
        acc.x ^= x.x;
        acc.y ^= x.y;
        acc.z ^= x.z;
        acc.w ^= x.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc.x ^= __shfl_xor_sync(0xFFFFFFFFu, acc.x, o);
        acc.y ^= __shfl_xor_sync(0xFFFFFFFFu, acc.y, o);
        acc.z ^= __shfl_xor_sync(0xFFFFFFFFu, acc.z, o);
        acc.w ^= __shfl_xor_sync(0xFFFFFFFFu, acc.w, o);
    }
    if ((threadIdx.x & 31) == 0 && sink) atomicXor(reinterpret_cast<unsigned int *>(sink) + ((threadIdx.x >> 5) & 3), acc.x ^ acc.y ^ acc.z ^ acc.w);
}
}  // namespace

extern "C" int gad_probe_read(const uint8_t *d_boards, int64_t P, int R, int stride, int width, void *d_sink16, void *stream)
{
    GAD_REQUIRE(d_boards && P >= 0 && R >= 1 && stride >= 16 && stride % 16 == 0 && width >= 1 && width <= stride,
                "gad_probe_read: bad argument");
    if (P == 0) return GAD_OK;
    const int units = (width + 15) / 16;
    const long long n_items = (long long)P * R * units;
    long long blocks = (n_items + 255) / 256;
    const long long cap = (long long)GAD_NUM_SMS * 8;
    if (blocks > cap) blocks = cap;
    probe_read_kernel<<<(unsigned)blocks, 256, 0, gad_stream(stream)>>>(d_boards, n_items, units, stride >> 4, (uint4 *)d_sink16);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------
namespace {
struct HostStage {
    void *d_boards = nullptr, *d_rowlen = nullptr, *d_scores = nullptr;
    size_t cap_boards = 0, cap_rowlen = 0, cap_scores = 0;
    cudaStream_t stream = nullptr;
};
thread_local HostStage g_stage;

int stage_reserve(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return GAD_OK;
    if (*p) GAD_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    GAD_CUDA(cudaMalloc(p, need));
    *cap = need;
    return GAD_OK;
}
}  // namespace

extern "C" int gad_fitness_host(uint8_t *h_boards, int32_t *h_rowlen, int64_t P, int R, int stride, int match,
                                int mismatch, int gap, int32_t *h_sp, int32_t *h_em, int32_t *h_al, int copy_back)
{
    GAD_REQUIRE(h_boards && h_rowlen && h_sp && h_em && h_al, "gad_fitness_host: null host pointer");
    GAD_REQUIRE(P >= 0 && R >= 1 && stride >= 16, "gad_fitness_host: bad shape");
    if (P == 0) return GAD_OK;
    HostStage &s = g_stage;
    if (!s.stream) GAD_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    const size_t nb = (size_t)P * R * stride, nr = (size_t)P * R * sizeof(int32_t), ns = (size_t)P * sizeof(int32_t);
    int rc;
    if ((rc = stage_reserve(&s.d_boards, &s.cap_boards, nb))) return rc;
    if ((rc = stage_reserve(&s.d_rowlen, &s.cap_rowlen, nr))) return rc;
    if ((rc = stage_reserve(&s.d_scores, &s.cap_scores, 3 * ns))) return rc;
    int32_t *d_sp = (int32_t *)s.d_scores, *d_em = d_sp + P, *d_al = d_em + P;
    GAD_CUDA(cudaMemcpyAsync(s.d_boards, h_boards, nb, cudaMemcpyHostToDevice, s.stream));
    GAD_CUDA(cudaMemcpyAsync(s.d_rowlen, h_rowlen, nr, cudaMemcpyHostToDevice, s.stream));
    rc = gad_fitness((uint8_t *)s.d_boards, (int32_t *)s.d_rowlen, P, R, stride, 0, match, mismatch, gap, d_sp, d_em,
                     d_al, nullptr, s.stream);
    if (rc) return rc;
    GAD_CUDA(cudaMemcpyAsync(h_sp, d_sp, ns, cudaMemcpyDeviceToHost, s.stream));
    GAD_CUDA(cudaMemcpyAsync(h_em, d_em, ns, cudaMemcpyDeviceToHost, s.stream));
    GAD_CUDA(cudaMemcpyAsync(h_al, d_al, ns, cudaMemcpyDeviceToHost, s.stream));
    if (copy_back) {
        GAD_CUDA(cudaMemcpyAsync(h_boards, s.d_boards, nb, cudaMemcpyDeviceToHost, s.stream));
        GAD_CUDA(cudaMemcpyAsync(h_rowlen, s.d_rowlen, nr, cudaMemcpyDeviceToHost, s.stream));
    }
    GAD_CUDA(cudaStreamSynchronize(s.stream));
    return GAD_OK;
}

extern "C" int gad_window_score_host(const uint8_t *h_board, int R, int stride, int from_row, int to_row,
                                     int from_col, int to_col, int match, int mismatch, int gap, int64_t *sp,
                                     int32_t *em)
{
    GAD_REQUIRE(h_board && sp && em, "gad_window_score_host: null pointer");
    GAD_REQUIRE(R >= 1 && R <= 255 && stride >= 4 && stride % 4 == 0, "gad_window_score_host: bad shape R=%d stride=%d",
                R, stride);
    GAD_REQUIRE(0 <= from_row && from_row <= to_row && to_row <= R, "gad_window_score_host: row window [%d,%d) outside [0,%d)",
                from_row, to_row, R);
    GAD_REQUIRE((long long)R * R * stride < (1ll << 31), "gad_window_score_host: R*R*stride overflows the 32-bit column sums");
    GAD_REQUIRE(0 <= from_col && from_col <= to_col && to_col <= stride,
                "gad_window_score_host: column window [%d,%d) outside [0,%d)", from_col, to_col, stride);
    HostStage &s = g_stage;
    if (!s.stream) GAD_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    const size_t nb = (size_t)R * stride;
    int rc;
    if ((rc = stage_reserve(&s.d_boards, &s.cap_boards, nb))) return rc;
    if ((rc = stage_reserve(&s.d_scores, &s.cap_scores, 16))) return rc;
    GAD_CUDA(cudaMemcpyAsync(s.d_boards, h_board, nb, cudaMemcpyHostToDevice, s.stream));
    long long *d_sp = (long long *)s.d_scores;
    int32_t *d_em = (int32_t *)(d_sp + 1);
    window_score_kernel<<<1, 32, 0, s.stream>>>((const uint8_t *)s.d_boards, stride, from_row, to_row, from_col, to_col,
                                                match, mismatch, gap, d_sp, d_em);
    GAD_LAUNCH_CHECK();
    long long out[2];
    GAD_CUDA(cudaMemcpyAsync(out, s.d_scores, 16, cudaMemcpyDeviceToHost, s.stream));
    GAD_CUDA(cudaStreamSynchronize(s.stream));
    *sp = out[0];
    *em = (int32_t)(out[1] & 0xffffffffll);
    return GAD_OK;
}
