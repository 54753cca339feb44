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

#include <stdlib.h>

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

// GS lanes own one board; lane l owns the 16-byte units l, l+GS, ... of every row.
template <int GS, bool PAIRED>
__global__ void __launch_bounds__(256, 3)
fitness_kernel(uint8_t *__restrict__ boards, int32_t *__restrict__ rowlen, long long P, int R, int stride,
               int match, int mismatch, int gap, int32_t *__restrict__ sp, int32_t *__restrict__ em,
               int32_t *__restrict__ al, int32_t *__restrict__ max_al)
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
        const int nwords = (w + 3) >> 2;
        const int nunits = (nwords + 3) >> 2;

        BoardSums bs;
        boardsums_clear(bs);
        for (int u = gl; u < nunits; u += GS) {
            ColCounts cc[4];
            uint32_t valid[4];
            if (!PAIRED && have_pre && u == gl) {
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
            sp[p] = (int32_t)sum_of_pairs_closed_form(bs, R, match, mismatch, gap);
            em[p] = (int32_t)bs.exact;
            al[p] = nk;
        }
        local_max = max(local_max, nk);
    }
    if (max_al != nullptr) {
        local_max = group_max<32>(local_max, 0xFFFFFFFFu);
        if ((threadIdx.x & 31) == 0 && local_max > 0) atomicMax(max_al, local_max);
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

template <int GS>
int launch_fitness(uint8_t *boards, int32_t *rowlen, long long P, int R, int stride, int match, int mismatch,
                   int gap, int32_t *sp, int32_t *em, int32_t *al, int32_t *max_al, cudaStream_t st)
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
                                                                        gap, sp, em, al, max_al);
    else
        fitness_kernel<GS, false><<<(unsigned)blocks, threads, 0, st>>>(boards, rowlen, P, R, stride, match, mismatch,
                                                                         gap, sp, em, al, max_al);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

}  // namespace

extern "C" int gad_fitness(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride, int width_hint,
                           int match, int mismatch, int gap, int32_t *d_sp, int32_t *d_em, int32_t *d_al,
                           int32_t *d_max_al, void *stream)
{
    GAD_REQUIRE(P >= 0 && R >= 1 && R <= 255, "gad_fitness: need P >= 0 and 1 <= R <= 255 (got P=%lld R=%d)",
                (long long)P, R);
    GAD_REQUIRE(stride >= 16 && stride % 16 == 0, "gad_fitness: stride must be a positive multiple of 16 (got %d)",
                stride);
    GAD_REQUIRE((long long)R * R * stride < (1ll << 31), "gad_fitness: R*R*stride overflows the 32-bit column sums");
    GAD_REQUIRE(d_boards && d_rowlen && d_sp && d_em && d_al, "gad_fitness: null device pointer");
    if (P == 0) return GAD_OK;
    cudaStream_t st = gad_stream(stream);
    if (d_max_al) GAD_CUDA(cudaMemsetAsync(d_max_al, 0, sizeof(int32_t), st));
    // lanes per board from the expected width (a performance hint only: wider boards take more trips)
    int width = width_hint > 0 && width_hint < stride ? width_hint : stride;
    int units = (width + 15) / 16;
    // wide boards in a large population: two 16-byte units per lane (more loads in flight per lane, half the
    // per-board reduction overhead) - measured on c4: 56.5 % -> 62.6 % of the HBM peak
    if (units >= 8 && P >= 4096) units = (units + 1) / 2;
    if (const char *g = getenv("GAD_FITNESS_GS")) units = atoi(g) > 0 ? atoi(g) : units;      // tuning override: lanes per board
#define GAD_FIT(GS) launch_fitness<GS>(d_boards, d_rowlen, P, R, stride, match, mismatch, gap, d_sp, d_em, d_al, d_max_al, st)
    if (units <= 1) return GAD_FIT(1);
    if (units <= 2) return GAD_FIT(2);
    if (units <= 4) return GAD_FIT(4);
    if (units <= 8) return GAD_FIT(8);
    if (units <= 16) return GAD_FIT(16);
    return GAD_FIT(32);
#undef GAD_FIT
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
