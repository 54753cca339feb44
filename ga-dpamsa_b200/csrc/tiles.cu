// Worst-fitted sub-board search for every mutating individual in one launch.
// Replaces utils.calculate_worst_fitted_sub_board + get_all_different_sub_range + check_overlap
// (reference utils.py:198-313). The reference's overlap scan accepts exactly the full tiles whose
// origins lie on the (rw, cw) lattice, row-major, so the tile list is implicit.
// Per tile SP and exact-match columns are computed from packed column counts (gad_common.cuh); the
// 'mo' ranking is the reference's fp64 min/max normalisation, first minimum wins.
#include "gad_common.cuh"

namespace {

// (SP, EM) of tile rows [fr, fr+rw) x columns [fc, fc+cw), computed by one warp
__device__ __forceinline__ void tile_score(const uint8_t *board, int stride, int fr, int rw, int fc, int cw,
                                           int match, int mismatch, int gap, long long &sp, int &em)
{
    const int lane = threadIdx.x & 31;
    BoardSums bs;
    boardsums_clear(bs);
    const int tc = fc + cw;
    const int j0 = fc >> 2, j1 = (tc + 3) >> 2;
    const uint32_t *base = reinterpret_cast<const uint32_t *>(board + (size_t)fr * stride);
    const int wstride = stride >> 2;
    for (int j = j0 + lane; j < j1; j += 32) {
        uint32_t valid = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int c = 4 * j + b;
            if (c >= fc && c < tc) valid |= 0x80u << (8 * b);
        }
        ColCounts cc;
        colcounts_clear(cc);
        const uint32_t first = base[j];
        for (int r = 0; r < rw; ++r) colcounts_add(cc, base[(size_t)r * wstride + j], first);
        boardsums_fold(bs, cc, valid, true);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bs.q2 += __shfl_xor_sync(0xFFFFFFFFu, bs.q2, o);
        bs.s1 += __shfl_xor_sync(0xFFFFFFFFu, bs.s1, o);
        bs.s2 += __shfl_xor_sync(0xFFFFFFFFu, bs.s2, o);
        bs.kept += __shfl_xor_sync(0xFFFFFFFFu, bs.kept, o);
        bs.exact += __shfl_xor_sync(0xFFFFFFFFu, bs.exact, o);
    }
    sp = sum_of_pairs_closed_form(bs, rw, match, mismatch, gap);
    em = (int)bs.exact;
}

// one warp per mutating individual
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
    for (int r = 0; r < R; ++r) n = min(n, rowlen[p * R + r]);
    const int tr = R / rw, tcn = n / cw;
    if (tr == 0 || tcn == 0) {                                // no tile fits: the reference's min() raises
        if (lane == 0) {
            tile[2 * m] = -1;
            tile[2 * m + 1] = -1;
        }
        return;
    }
    long long lo_s = 0x7fffffffffffffffll, hi_s = -lo_s - 1;
    int lo_e = 0x7fffffff, hi_e = -1;
    if (mode == GAD_MODE_MO) {                                // utils.py:296-301 needs min/max first
        for (int t = 0; t < tr * tcn; ++t) {
            long long s;
            int e;
            tile_score(board, stride, (t / tcn) * rw, rw, (t % tcn) * cw, cw, match, mismatch, gap, s, e);
            lo_s = min(lo_s, s);
            hi_s = max(hi_s, s);
            lo_e = min(lo_e, e);
            hi_e = max(hi_e, e);
        }
    }
    // cs = exact / cw is monotone in exact for a fixed cw, so min/max over cs are those over exact
    const double lo_c = __ddiv_rn((double)lo_e, (double)cw), hi_c = __ddiv_rn((double)hi_e, (double)cw);
    double best = 0.0;
    int best_t = -1;
    for (int t = 0; t < tr * tcn; ++t) {
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
        if (best_t < 0 || k < best) {                         // first minimum wins (python min)
            best = k;
            best_t = t;
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
