/*
 * oracle/ga_oracle.c - CPU restatement of the GA-DPAMSA scoring path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library this
 * file builds (oracle/_build/libga_oracle.so).  The product path
 * (ga-dpamsa_b200/) never links, imports or calls it.
 *
 * Parity pin: every function here is checked against the reference's own
 * python functions (run in-process through oracle/ref_shim.py) and against
 * the 1340 golden (alignment -> SP, EM, AL) records of results/reports
 * (tests/test_oracle_pins.py, tests/golden/report_vectors.json).
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference checkout).  Cells are uint8 codes A=1 T=2 C=3 G=4 gap=5
 * (GA.py:35).  A board is R rows at a fixed byte stride; rowlen[r] is the
 * live length of row r (rows are ragged between GA phases).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GAPC 5

/* utils.py:62-76 - sum over columns, over row pairs j<k:
 *   either cell is a gap -> gap penalty; equal -> match; else mismatch. */
long orc_sum_of_pairs(const uint8_t *b, int stride, int from_row, int to_row,
                      int from_col, int to_col, int match, int mismatch, int gap)
{
    long score = 0;
    for (int c = from_col; c < to_col; ++c)
        for (int j = from_row; j < to_row; ++j)
            for (int k = j + 1; k < to_row; ++k) {
                uint8_t x = b[(size_t)j * stride + c], y = b[(size_t)k * stride + c];
                if (x == GAPC || y == GAPC) score += gap;
                else if (x == y) score += match;
                else score += mismatch;
            }
    return score;
}

/* utils.py:122-125 - number of columns whose cells all equal the cell of
 * from_row (an all-gap column counts).  The caller divides by the number of
 * columns (utils.py:128) to get the python float. */
int orc_exact_columns(const uint8_t *b, int stride, int from_row, int to_row,
                      int from_col, int to_col)
{
    int n = 0;
    for (int c = from_col; c < to_col; ++c) {
        int same = 1;
        for (int r = from_row; r < to_row; ++r)
            if (b[(size_t)r * stride + c] != b[(size_t)from_row * stride + c]) { same = 0; break; }
        n += same;
    }
    return n;
}

/* GA.py:158-162 (right-pad every row with gaps to the longest row) followed by
 * utils.py:404-428 (cut the common all-gap tail, then delete every remaining
 * all-gap column).  In place; returns the new width and sets every rowlen. */
int orc_pad_clean(uint8_t *b, int R, int stride, int32_t *rowlen)
{
    int width = 0;
    for (int r = 0; r < R; ++r) if (rowlen[r] > width) width = rowlen[r];
    for (int r = 0; r < R; ++r) {
        for (int c = rowlen[r]; c < width; ++c) b[(size_t)r * stride + c] = GAPC;
        rowlen[r] = width;
    }
    if (R == 0 || width == 0) return width;            /* utils.py:404-405 */
    /* step 1: utils.py:408-419 */
    int common = 0;
    for (int r = 0; r < R; ++r) {
        int i = width;
        while (i > 0 && b[(size_t)r * stride + i - 1] == GAPC) --i;
        if (i > common) common = i;
    }
    width = common;
    /* step 2: utils.py:422-428 */
    int out = 0;
    for (int c = 0; c < width; ++c) {
        int allgap = 1;
        for (int r = 0; r < R; ++r) if (b[(size_t)r * stride + c] != GAPC) { allgap = 0; break; }
        if (allgap) continue;
        if (out != c) for (int r = 0; r < R; ++r) b[(size_t)r * stride + out] = b[(size_t)r * stride + c];
        ++out;
    }
    for (int r = 0; r < R; ++r) rowlen[r] = out;
    return out;
}

/* GA.py:156-178 - one fitness pass over a population stored as
 * boards[P][R][stride] + rowlen[P][R]: pad, clean, SP over the whole board,
 * exact-match columns (EM) and alignment length (AL); CS = EM/AL is formed by
 * the caller as a python float.  Boards are rewritten in place, as the
 * reference mutates its lists in place.  nthreads<=1 -> serial. */
void orc_fitness_population(uint8_t *boards, int32_t *rowlen, long P, int R, int stride,
                            int match, int mismatch, int gap,
                            int32_t *sp, int32_t *em, int32_t *al, int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (long p = 0; p < P; ++p) {
        uint8_t *b = boards + (size_t)p * R * stride;
        int32_t *rl = rowlen + (size_t)p * R;
        int w = orc_pad_clean(b, R, stride, rl);
        sp[p] = (int32_t)orc_sum_of_pairs(b, stride, 0, R, 0, w, match, mismatch, gap);
        em[p] = orc_exact_columns(b, stride, 0, R, 0, w);
        al[p] = w;
    }
}

/* utils.py:276-289 - per-tile SP and exact-column count over the tile grid that
 * get_all_different_sub_range (utils.py:230-252) yields: origins
 * (i*Rw, j*Cw) for every full tile inside R x n (n = shortest row).
 * Tiles are emitted row-major; returns the tile count. */
int orc_tile_scores(const uint8_t *b, int R, int stride, int n, int Rw, int Cw,
                    int match, int mismatch, int gap, int32_t *tile_sp, int32_t *tile_em)
{
    int t = 0;
    for (int i = 0; i + Rw <= R; i += Rw)
        for (int j = 0; j + Cw <= n; j += Cw) {
            tile_sp[t] = (int32_t)orc_sum_of_pairs(b, stride, i, i + Rw, j, j + Cw, match, mismatch, gap);
            tile_em[t] = orc_exact_columns(b, stride, i, i + Rw, j, j + Cw);
            ++t;
        }
    return t;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
