// Shared helpers for libgadpamsa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gadpamsa.h"

#define GAD_GAP 5
#define GAD_NUM_SMS 148

void gad_set_error(const char *fmt, ...);

#define GAD_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            gad_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return GAD_ECUDA;                                                            \
        }                                                                                \
    } while (0)

#define GAD_REQUIRE(cond, ...)                                                           \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            gad_set_error(__VA_ARGS__);                                                  \
            return GAD_EINVAL;                                                           \
        }                                                                                \
    } while (0)

#define GAD_LAUNCH_CHECK()                                                               \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            gad_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GAD_ECUDA;                                                            \
        }                                                                                \
    } while (0)

static inline cudaStream_t gad_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------
// Packed-byte column arithmetic. One 32-bit word = 4 adjacent columns of one row; cell
// codes are 1..5 (3 bits), so the per-symbol indicator of every byte is one LOP3 over the
// three bit planes of the word, and per-column symbol counts over <=255 rows accumulate
// as four independent byte lanes with plain integer adds.
// ---------------------------------------------------------------------------------------
#define GAD_LSB4 0x01010101u
#define GAD_MSB4 0x80808080u

struct ColCounts {
    uint32_t c1, c2, c3, c4;   // per byte lane: rows holding A / T / C / G in that column
    uint32_t diff;             // per byte lane: OR over rows of (cell XOR first row's cell)
};

__device__ __forceinline__ void colcounts_clear(ColCounts &c) { c.c1 = c.c2 = c.c3 = c.c4 = 0u; c.diff = 0u; }

__device__ __forceinline__ void colcounts_add(ColCounts &c, uint32_t x, uint32_t x_first)
{
    const uint32_t b0 = x & GAD_LSB4;
    const uint32_t b1 = (x >> 1) & GAD_LSB4;
    const uint32_t b2 = (x >> 2) & GAD_LSB4;
    c.c1 += b0 & ~b1 & ~b2;    // 001
    c.c2 += b1 & ~b0;          // 010 (b2 is 0 whenever b1 is set for codes 1..5)
    c.c3 += b0 & b1;           // 011
    c.c4 += b2 & ~b0;          // 100
    c.diff |= x ^ x_first;
}

// Two rows at once: z = xa | xb << 4 puts the 3-bit codes of row a in the low and of row b in the high
// nibble of every byte, so one LOP3 per symbol yields the indicators of both rows (bits 0 and 4) and
// the accumulators become pairs of 4-bit counters (at most 15 row pairs between two flushes).
#define GAD_LSN8 0x11111111u
#define GAD_NIB4 0x0F0F0F0Fu

__device__ __forceinline__ void colcounts_add2(ColCounts &n, uint32_t xa, uint32_t xb, uint32_t first2)
{
    const uint32_t z = xb * 16u + xa;
    const uint32_t b0 = z & GAD_LSN8;
    const uint32_t b1 = (z >> 1) & GAD_LSN8;
    const uint32_t b2 = (z >> 2) & GAD_LSN8;
    n.c1 += b0 & ~b1 & ~b2;
    n.c2 += b1 & ~b0;
    n.c3 += b0 & b1;
    n.c4 += b2 & ~b0;
    n.diff |= z ^ first2;
}

// fold the nibble-pair counters of an epoch into the byte counters
__device__ __forceinline__ void colcounts_flush2(ColCounts &c, ColCounts &n)
{
    c.c1 += (n.c1 & GAD_NIB4) + ((n.c1 >> 4) & GAD_NIB4);
    c.c2 += (n.c2 & GAD_NIB4) + ((n.c2 >> 4) & GAD_NIB4);
    c.c3 += (n.c3 & GAD_NIB4) + ((n.c3 >> 4) & GAD_NIB4);
    c.c4 += (n.c4 & GAD_NIB4) + ((n.c4 >> 4) & GAD_NIB4);
    c.diff |= (n.diff | (n.diff >> 4)) & GAD_NIB4;
    n.c1 = n.c2 = n.c3 = n.c4 = 0u;
    n.diff = 0u;
}

// bit 7 of every byte lane that is non-zero (exact per lane, lanes may hold up to 255)
__device__ __forceinline__ uint32_t nonzero_lanes(uint32_t v)
{
    return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & GAD_MSB4;
}

// Per-board accumulators over the kept (not all-gap) columns.
struct BoardSums {
    uint32_t q2;     // sum over columns of c1^2+c2^2+c3^2+c4^2
    uint32_t s1;     // sum over columns of non-gap cells
    uint32_t s2;     // sum over columns of (non-gap cells)^2
    uint32_t kept;   // columns that are not all-gap
    uint32_t exact;  // kept columns whose cells are all equal
};

__device__ __forceinline__ void boardsums_clear(BoardSums &b) { b.q2 = b.s1 = b.s2 = b.kept = b.exact = 0u; }

// Fold one word (4 columns) of finished column counts into the board sums; returns the
// keep mask (bit 7 of each lane whose column has at least one non-gap cell).
// `lane_valid` has bit 7 set for the lanes that are real columns of the window.
__device__ __forceinline__ uint32_t boardsums_fold(BoardSums &b, const ColCounts &c, uint32_t lane_valid,
                                                   bool count_allgap_columns)
{
    const uint32_t s = c.c1 + c.c2 + c.c3 + c.c4;          // non-gap cells per column
    uint32_t keep = nonzero_lanes(s) & lane_valid;
    const uint32_t uniform = ~nonzero_lanes(c.diff) & GAD_MSB4 & lane_valid;
    if (count_allgap_columns) {
        // windowed utils.get_column_score counts an all-gap column as exact and
        // utils.get_sum_of_pairs scores its pairs as gaps: keep every valid lane.
        keep = lane_valid;
    }
    // lanes outside `keep` must not contribute: all-gap lanes have zero counts already;
    // invalid lanes are masked here (0xFF per kept lane).
    const uint32_t m = (keep >> 7) * 0xFFu;
    const uint32_t a1 = c.c1 & m, a2 = c.c2 & m, a3 = c.c3 & m, a4 = c.c4 & m, as = s & m;
    b.q2 = __dp4a(a1, a1, b.q2);
    b.q2 = __dp4a(a2, a2, b.q2);
    b.q2 = __dp4a(a3, a3, b.q2);
    b.q2 = __dp4a(a4, a4, b.q2);
    b.s1 = __dp4a(as, GAD_LSB4, b.s1);
    b.s2 = __dp4a(as, as, b.s2);
    b.kept += __popc(keep);
    b.exact += __popc(uniform & keep);
    return keep;
}

// Fast path of boardsums_fold for the fitness pass: the word is either wholly inside the board or wholly
// outside it (`inside`), all-gap columns have zero counts already, so no per-lane masking is needed.
__device__ __forceinline__ uint32_t boardsums_fold_word(BoardSums &b, const ColCounts &c, bool inside)
{
    if (!inside) return 0u;
    const uint32_t s = c.c1 + c.c2 + c.c3 + c.c4;          // non-gap cells per column
    const uint32_t keep = nonzero_lanes(s);
    b.q2 = __dp4a(c.c1, c.c1, b.q2);
    b.q2 = __dp4a(c.c2, c.c2, b.q2);
    b.q2 = __dp4a(c.c3, c.c3, b.q2);
    b.q2 = __dp4a(c.c4, c.c4, b.q2);
    b.s1 = __dp4a(s, GAD_LSB4, b.s1);
    b.s2 = __dp4a(s, s, b.s2);
    b.kept += __popc(keep);
    b.exact += __popc(~nonzero_lanes(c.diff) & keep);
    return keep;
}

// Closed form of utils.py:62-76 from the board sums over `kept` columns of `rows` cells:
//   match pairs    M  = sum c(c-1)/2                      = (q2 - s1)/2
//   gap pairs      Gp = sum g(g-1)/2 + g(rows-g), g=rows-s = ((2 rows-1) G1 - G2)/2
//   mismatch pairs    = kept*rows(rows-1)/2 - M - Gp
__device__ __forceinline__ long long sum_of_pairs_closed_form(const BoardSums &b, int rows, int match, int mismatch,
                                                              int gap)
{
    // every term is below 2 * rows^2 * columns, which the entry points keep under 2^32: 32-bit unsigned
    // arithmetic, 64 bits only for the weighted sum
    const uint32_t R = (uint32_t)rows, nk = b.kept, s1 = b.s1, s2 = b.s2, q2 = b.q2;
    const uint32_t M = (q2 - s1) >> 1;
    const uint32_t G1 = R * nk - s1;
    const uint32_t G2 = R * R * nk - 2u * R * s1 + s2;
    const uint32_t Gp = ((2u * R - 1u) * G1 - G2) >> 1;
    const uint32_t T = (nk * R * (R - 1u)) >> 1;
    return (long long)match * M + (long long)gap * Gp + (long long)mismatch * (long long)(T - M - Gp);
}
