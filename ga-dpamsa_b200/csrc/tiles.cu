// Worst-fitted sub-board search for every mutating individual in one launch.
// Replaces utils.calculate_worst_fitted_sub_board + get_all_different_sub_range + check_overlap
// (reference utils.py:198-313). The reference's overlap scan accepts exactly the full tiles whose
// origins lie on the (rw, cw) lattice, row-major, so the tile list is implicit.
// Per tile SP and exact-match columns are computed from packed column counts (gad_common.cuh); the
// 'mo' ranking is the reference's fp64 min/max normalisation, first minimum wins.
#include "gad_common.cuh"

namespace {

// (SP, EM) of tile rows [fr, fr+rw) x columns [fc, fc+cw), computed by ONE lane: a tile of a trained window is
// a few words wide (30 columns = 8 or 9 words), so a lane per tile keeps all 32 lanes busy on boards with many
// tiles (c4: 42, c5: 735) and needs no cross-lane reduction per tile.
__device__ __forceinline__ void tile_score(const uint8_t *board, int stride, int fr, int rw, int fc, int cw,
                                           int match, int mismatch, int gap, long long &sp, int &em)
{
    BoardSums bs;
    boardsums_clear(bs);
    const int tc = fc + cw;
    const int j0 = fc >> 2, j1 = (tc + 3) >> 2;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(board + (size_t)fr * stride);
    const int wstride = stride >> 2;
    for (int j = j0; j < j1; ++j) {
        uint32_t valid = GAD_MSB4;
        if (4 * j < fc) valid &= GAD_MSB4 << (8 * (fc - 4 * j));                 // first word: columns before the tile
        if (4 * j + 4 > tc) valid &= GAD_MSB4 >> (8 * (4 * j + 4 - tc));         // last word: columns after it
        ColCounts cc;
        colcounts_clear(cc);
        const uint32_t first = base[j];
        for (int r = 0; r < rw; ++r) colcounts_add(cc, base[(size_t)r * wstride + j], first);
        boardsums_fold(bs, cc, valid, true);
    }
    sp = sum_of_pairs_closed_form(bs, rw, match, mismatch, gap);
    em = (int)bs.exact;
}

// one warp per mutating individual, one lane per tile (tiles in row-major order: t = lane, lane + 32, ...)
__global__ void __launch_bounds__(128)
worst_tile_kernel(const uint8_t *__restrict__ boards, const int32_t *__restrict__ rowlen, int R, int stride,
                  const int32_t *__restrict__ idx, long long M, int rw, int cw, int mode, int match, int mismatch,
                  int gap, int32_t *__restrict__ tile)
{
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const int lane = threadIdx.x & 31;
    const long long p = idx ? idx[m] : m;
    const uint8_t *board = boards + (size_t)p * R * stride;
    int n = 0x7fffffff;                                       // utils.py:231: shortest row
    for (int r = lane; r < R; r += 32) n = min(n, rowlen[p * R + r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n = min(n, __shfl_xor_sync(0xFFFFFFFFu, n, o));
    const int tr = R / rw, tcn = n / cw;
    if (tr == 0 || tcn == 0) {                                // no tile fits: the reference's min() raises
        if (lane == 0) {
            tile[2 * m] = -1;
            tile[2 * m + 1] = -1;
        }
        return;
    }
    const int ntiles = tr * tcn;
    long long lo_s = 0x7fffffffffffffffll, hi_s = -lo_s - 1;
    int lo_e = 0x7fffffff, hi_e = -1;
    if (mode == GAD_MODE_MO) {                                // utils.py:296-301 needs min/max over all tiles first
        for (int t = lane; t < ntiles; t += 32) {
            long long s;
            int e;
            tile_score(board, stride, (t / tcn) * rw, rw, (t % tcn) * cw, cw, match, mismatch, gap, s, e);
            lo_s = min(lo_s, s);
            hi_s = max(hi_s, s);
            lo_e = min(lo_e, e);
            hi_e = max(hi_e, e);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo_s = min(lo_s, __shfl_xor_sync(0xFFFFFFFFu, lo_s, o));
            hi_s = max(hi_s, __shfl_xor_sync(0xFFFFFFFFu, hi_s, o));
            lo_e = min(lo_e, __shfl_xor_sync(0xFFFFFFFFu, lo_e, o));
            hi_e = max(hi_e, __shfl_xor_sync(0xFFFFFFFFu, hi_e, o));
        }
    }
    // cs = exact / cw is monotone in exact for a fixed cw, so min/max over cs are those over exact
    const double lo_c = __ddiv_rn((double)lo_e, (double)cw), hi_c = __ddiv_rn((double)hi_e, (double)cw);
    double best = 0.0;
    int best_t = 0x7fffffff;
    for (int t = lane; t < ntiles; t += 32) {
        long long s;
        int e;
        tile_score(board, stride, (t / tcn) * rw, rw, (t % tcn) * cw, cw, match, mismatch, gap, s, e);
        double k;
        if (mode == GAD_MODE_SP) {
            k = (double)s;
        } else if (mode == GAD_MODE_CS) {
            k = __ddiv_rn((double)e, (double)cw);
        } else {
            const double nsp = hi_s > lo_s ? __ddiv_rn((double)(s - lo_s), (double)(hi_s - lo_s)) : 0.5;
            const double c = __ddiv_rn((double)e, (double)cw);
            const double ncs = hi_c > lo_c ? __ddiv_rn(__dsub_rn(c, lo_c), __dsub_rn(hi_c, lo_c)) : 0.5;
            k = __dadd_rn(nsp, ncs);
        }
        if (best_t == 0x7fffffff || k < best) {               // first minimum of this lane's tiles (ascending t)
            best = k;
            best_t = t;
        }
    }
    // first minimum over all tiles (python min): smallest key, ties to the smallest tile number
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int ot = __shfl_xor_sync(0xFFFFFFFFu, best_t, o);
        if (ot != 0x7fffffff && (best_t == 0x7fffffff || ok < best || (ok == best && ot < best_t))) {
            best = ok;
            best_t = ot;
        }
    }
    if (lane == 0) {
        tile[2 * m] = (best_t / tcn) * rw;
        tile[2 * m + 1] = (best_t % tcn) * cw;
    }
}

}  // namespace

extern "C" int gad_worst_tiles(const uint8_t *d_boards, const int32_t *d_rowlen, int R, int stride,
                               const int32_t *d_idx, int64_t M, int rw, int cw, int mode, int match, int mismatch,
                               int gap, int32_t *d_tile, void *stream)
{
    GAD_REQUIRE(d_boards && d_rowlen && d_tile, "gad_worst_tiles: null device pointer");
    GAD_REQUIRE(M >= 0 && R >= 1 && R <= 255 && stride >= 16 && stride % 16 == 0, "gad_worst_tiles: bad shape");
    GAD_REQUIRE(rw >= 1 && cw >= 1 && rw <= R, "gad_worst_tiles: window %dx%d does not fit %d rows", rw, cw, R);
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_worst_tiles: unknown mode %d", mode);
    if (M == 0) return GAD_OK;
    const unsigned blocks = (unsigned)((M + 3) / 4);
    worst_tile_kernel<<<blocks, 128, 0, gad_stream(stream)>>>(d_boards, d_rowlen, R, stride, d_idx, M, rw, cw, mode,
                                                             match, mismatch, gap, d_tile);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
