// GA.update_hall_of_fame (reference GA.py:96-136) as device code shared by the stand-alone decision kernel
// (rank_select.cu: hof_decide_kernel) and by the tail of the fitness kernel (fitness.cu: the whole
// GA.calculate_fitness_score in one launch). Keys are the fp64 values python computes
// (utils.py:316-355: int/int true division, float subtract / divide / add, no contraction); the first maximum
// wins, like the reference's stable descending sort.
#pragma once
#include "gad_common.cuh"

// where the blocks of one decision leave their partial results (the head of the ranking workspace); plain data
// shared by rank_select.cu (which owns the workspace layout) and fitness.cu
struct GadDecideParts {
    double *key;
    int *idx;
    int *lo, *hi;
    unsigned long long *clo, *chi;
    uint4 *rec;          // 64 bytes per block: {lo, hi, clo, chi} | {key, idx, sp, em, al}, each read with one 32-byte trip
    int *winner;         // the decision, for callers that read it from the workspace
};

// rank_select.cu: the partial-result arrays inside a ranking workspace sized for n elements (gad_workspace_bytes)
int gad_decide_parts(void *d_ws, size_t ws_bytes, long long n, const char *who, GadDecideParts *out);

namespace {

typedef unsigned long long u64;

// the threads that take part in a block-wide step: all of a block (bar 0) or the consumer threads of the
// fitness kernel (a named barrier, the producer warp has left)
struct BlockCtx {
    int tid, nthreads, bar;
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory"); }
};

// python: exact / num_columns (utils.py:128), 0 when there are no columns
__device__ __forceinline__ double column_score(int em, int al) { return al > 0 ? __ddiv_rn((double)em, (double)al) : 0.0; }

// element i < P comes from the population, element P (when n == P + 1) from the hall of fame
__device__ __forceinline__ void load_score(const int *sp, const int *em, const int *al, const int *hof, long long P,
                                           long long i, int &s, int &e, int &a)
{
    if (i < P) {
        s = sp[i];
        e = em[i];
        a = al[i];
    } else {
        s = hof[2];
        e = hof[3];
        a = hof[4];
    }
}

// utils.py:332-351: single metric -> the metric; MO -> (sp-min)/(max-min) + (cs-min)/(max-min), 0.5 when flat
__device__ __forceinline__ double make_key(int mode, int s, int e, int a, long long lo, long long hi, u64 cs_min, u64 cs_max)
{
    if (mode == GAD_MODE_SP) return (double)s;
    if (mode == GAD_MODE_CS) return column_score(e, a);
    const double clo = __longlong_as_double((long long)cs_min), chi = __longlong_as_double((long long)cs_max);
    const double nsp = hi > lo ? __ddiv_rn((double)((long long)s - lo), (double)(hi - lo)) : 0.5;
    const double c = column_score(e, a);
    const double ncs = chi > clo ? __ddiv_rn(__dsub_rn(c, clo), __dsub_rn(chi, clo)) : 0.5;
    return __dadd_rn(nsp, ncs);
}

__device__ __forceinline__ bool better(double ka, int ia, double kb, int ib)
{
    return ka > kb || (ka == kb && ia < ib);      // first maximum wins (stable descending sort, utils.py:340,351)
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct Range {
    int lo, hi;
    u64 clo, chi;        // bit patterns of non-negative doubles (order-preserving)
};

__device__ __forceinline__ void range_clear(Range &g)
{
    g.lo = 0x7fffffff;
    g.hi = (int)0x80000000;
    g.clo = ~0ull;
    g.chi = 0ull;
}

__device__ __forceinline__ void range_merge(Range &g, int lo, int hi, u64 clo, u64 chi)
{
    g.lo = min(g.lo, lo);
    g.hi = max(g.hi, hi);
    g.clo = clo < g.clo ? clo : g.clo;
    g.chi = chi > g.chi ? chi : g.chi;
}

__device__ __forceinline__ void range_fold_warp(Range &g)
{
    for (int o = 16; o > 0; o >>= 1)
        range_merge(g, __shfl_xor_sync(0xFFFFFFFFu, g.lo, o), __shfl_xor_sync(0xFFFFFFFFu, g.hi, o),
                    __shfl_xor_sync(0xFFFFFFFFu, g.clo, o), __shfl_xor_sync(0xFFFFFFFFu, g.chi, o));
}

struct DecideShared {                // up to 32 warps
    Range range[32];
    double key[32];
    int idx[32];
    int winner;
    bool last;
};

// score range of the block's elements; result in every thread
__device__ __forceinline__ Range range_fold_block(Range g, DecideShared &sh, const BlockCtx &cx)
{
    const int lane = cx.tid & 31, wid = cx.tid >> 5, nw = (cx.nthreads + 31) >> 5;
    range_fold_warp(g);
    if (lane == 0) sh.range[wid] = g;
    cx.sync();
    range_clear(g);
    if (lane < nw) g = sh.range[lane];
    range_fold_warp(g);
    cx.sync();
    return g;
}

__device__ __forceinline__ void best_merge(double &bk, int &bi, double k, int i)
{
    if (i != 0x7fffffff && (bi == 0x7fffffff || better(k, i, bk, bi))) {
        bk = k;
        bi = i;
    }
}

__device__ __forceinline__ void best_fold_warp(double &bk, int &bi)
{
    for (int o = 16; o > 0; o >>= 1) best_merge(bk, bi, __shfl_xor_sync(0xFFFFFFFFu, bk, o), __shfl_xor_sync(0xFFFFFFFFu, bi, o));
}

// best of the block; result in warp 0
__device__ __forceinline__ void best_fold_block(double &bk, int &bi, DecideShared &sh, const BlockCtx &cx)
{
    const int lane = cx.tid & 31, wid = cx.tid >> 5, nw = (cx.nthreads + 31) >> 5;
    best_fold_warp(bk, bi);
    if (lane == 0) {
        sh.key[wid] = bk;
        sh.idx[wid] = bi;
    }
    cx.sync();
    if (cx.tid < 32) {
        bk = lane < nw ? sh.key[lane] : 0.0;
        bi = lane < nw ? sh.idx[lane] : 0x7fffffff;
        best_fold_warp(bk, bi);
    }
}

// GA.py:110-136 on one GPU by one whole block: the winner's board and score tuple become the hall of fame.
// Winner == P: the record keeps its place under the fresh index len(scores) (GA.py:127-132).
__device__ __forceinline__ void hof_commit_local(int win, long long P, const int *sp, const int *em, const int *al,
                                                 const uint8_t *boards, int R, int stride, uint8_t *hof_board, int *hof,
                                                 const BlockCtx &cx)
{
    if (win >= P) {
        if (cx.tid == 0) hof[1] = (int)P;
        return;
    }
    const int width = __ldcg(al + win);
    const int n16 = min(stride >> 4, (width + 15) >> 4);
    for (int u = cx.tid; u < R * n16; u += cx.nthreads) {
        const int r = u / n16, j = u - r * n16;
        reinterpret_cast<uint4 *>(hof_board + (size_t)r * stride)[j] =
            __ldcg(reinterpret_cast<const uint4 *>(boards + ((size_t)win * R + r) * stride) + j);
    }
    if (cx.tid == 0) {
        hof[0] = 1;
        hof[1] = win;
        hof[2] = __ldcg(sp + win);
        hof[3] = __ldcg(em + win);
        hof[4] = width;
    }
}

}  // namespace
