// Ranking, selection, survivor compaction and the hall of fame, all on device.
// Replaces utils.get_index_of_the_best_fitted_individuals (reference utils.py:316-355),
// GA._find_pareto_front (GA.py:183-211), GA.selection (GA.py:213-260) and
// GA.update_hall_of_fame (GA.py:96-136).
//
// Every ordering rule of the reference is a stable descending python sort, so ties keep
// ascending index. Keys are fp64 computed with the same IEEE operations python performs
// (int/int true division and float subtract/divide/add, no contraction), which makes the
// device order bit-identical to the reference's. Device-wide sort and scan are CUB primitives
// (toolkit headers); the kernels that move boards are hand-written.
#include "gad_common.cuh"
#include "hof_decide.cuh"

#include <cub/cub.cuh>
#include <stdlib.h>

namespace {

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Stats {            // lives at the head of the workspace
    int sp_min, sp_max;
    u64 cs_min, cs_max;   // bit patterns of non-negative doubles (order-preserving)
    int winner;           // argbest of the last hof pass
    int n_front;          // size of the Pareto front of the last selection
    int n_keep;
    unsigned int ticket;  // last-block-done counter
};

struct Workspace {
    Stats *stats;
    u64 *k0, *k1;         // sort keys (double buffer)
    int *v0, *v1;         // sort values (double buffer)
    int *ord_a, *ord_b, *ord_c;   // extra orders kept by the MO selection
    u64 *scan_u64;
    int *gstart, *flags, *scan;
    double *part_key;     // per-block argmax partials
    int *part_idx;
    int *part_lo, *part_hi;       // per-block score ranges of the one-kernel hall-of-fame decision
    u64 *part_clo, *part_chi;
    uint4 *part_rec;              // 64 bytes per block: its score range and its best element with the score tuple, one
                                  // 32-byte record each (the decision in the tail of the fitness kernel)
    void *cub_tmp;
    size_t cub_bytes;
};

const int kArgBlocks = 1024;

size_t cub_temp_bytes(long long n)
{
    size_t a = 0, b = 0, c = 0, d = 0;
    cub::DoubleBuffer<u64> keys((u64 *)nullptr, (u64 *)nullptr);
    cub::DoubleBuffer<int> vals((int *)nullptr, (int *)nullptr);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, a, keys, vals, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int *)nullptr, (int *)nullptr, (int)n);
    cub::DeviceScan::InclusiveScan(nullptr, c, (u64 *)nullptr, (u64 *)nullptr, cub::Max(), (int)n);
    cub::DeviceScan::InclusiveScan(nullptr, d, (int *)nullptr, (int *)nullptr, cub::Max(), (int)n);
    size_t m = a > b ? a : b;
    m = m > c ? m : c;
    m = m > d ? m : d;
    return align256(m + 256);
}

size_t carve(Workspace &w, void *base, long long n, size_t cub_bytes)
{
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { void *q = p ? p + off : nullptr; off += align256(bytes); return q; };
    w.stats = (Stats *)take(sizeof(Stats));
    w.k0 = (u64 *)take(sizeof(u64) * n);
    w.k1 = (u64 *)take(sizeof(u64) * n);
    w.v0 = (int *)take(sizeof(int) * n);
    w.v1 = (int *)take(sizeof(int) * n);
    w.ord_a = (int *)take(sizeof(int) * n);
    w.ord_b = (int *)take(sizeof(int) * n);
    w.ord_c = (int *)take(sizeof(int) * n);
    w.scan_u64 = (u64 *)take(sizeof(u64) * n);
    w.gstart = (int *)take(sizeof(int) * n);
    w.flags = (int *)take(sizeof(int) * n);
    w.scan = (int *)take(sizeof(int) * n);
    w.part_key = (double *)take(sizeof(double) * kArgBlocks);
    w.part_idx = (int *)take(sizeof(int) * kArgBlocks);
    w.part_lo = (int *)take(sizeof(int) * kArgBlocks);
    w.part_hi = (int *)take(sizeof(int) * kArgBlocks);
    w.part_clo = (u64 *)take(sizeof(u64) * kArgBlocks);
    w.part_chi = (u64 *)take(sizeof(u64) * kArgBlocks);
    w.part_rec = (uint4 *)take(64 * (size_t)kArgBlocks);
    w.cub_tmp = take(cub_bytes);
    w.cub_bytes = cub_bytes;
    return off;
}

int open_workspace(Workspace &w, void *ws, size_t ws_bytes, long long n, const char *who)
{
    GAD_REQUIRE(ws != nullptr, "%s: null workspace", who);
    GAD_REQUIRE(n < (1ll << 31) - 2, "%s: population too large for 32-bit indices", who);
    const size_t cb = cub_temp_bytes(n);
    const size_t need = carve(w, ws, n, cb);
    GAD_REQUIRE(need <= ws_bytes, "%s: workspace of %zu bytes is too small, need %zu (gad_workspace_bytes)", who,
                ws_bytes, need);
    return GAD_OK;
}

__device__ __forceinline__ u64 sortable(double v)
{
    const u64 b = (u64)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------------
__global__ void stats_init_kernel(Stats *s)
{
    s->sp_min = 0x7fffffff;
    s->sp_max = (int)0x80000000;
    s->cs_min = ~0ull;
    s->cs_max = 0ull;
}

__global__ void stats_kernel(const int *__restrict__ sp, const int *__restrict__ em, const int *__restrict__ al,
                             const int *__restrict__ hof, long long P, long long n, Stats *st)
{
    int lo = 0x7fffffff, hi = (int)0x80000000;
    u64 clo = ~0ull, chi = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int s, e, a;
        load_score(sp, em, al, hof, P, i, s, e, a);
        lo = min(lo, s);
        hi = max(hi, s);
        const u64 c = (u64)__double_as_longlong(column_score(e, a));
        clo = c < clo ? c : clo;
        chi = c > chi ? c : chi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, o));
        const u64 a = __shfl_xor_sync(0xFFFFFFFFu, clo, o), b = __shfl_xor_sync(0xFFFFFFFFu, chi, o);
        clo = a < clo ? a : clo;
        chi = b > chi ? b : chi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&st->sp_min, lo);
        atomicMax(&st->sp_max, hi);
        atomicMin(&st->cs_min, clo);
        atomicMax(&st->cs_max, chi);
    }
}

__global__ void keys_kernel(const int *__restrict__ sp, const int *__restrict__ em, const int *__restrict__ al,
                            const int *__restrict__ hof, long long P, long long n, int mode, const Stats *st,
                            double *__restrict__ key)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s, e, a;
    load_score(sp, em, al, hof, P, i, s, e, a);
    key[i] = make_key(mode, s, e, a, st->sp_min, st->sp_max, st->cs_min, st->cs_max);
}

int compute_keys(const int *sp, const int *em, const int *al, const int *hof, long long P, int mode, double *key,
                 Workspace &w, cudaStream_t st)
{
    const long long n = P + (hof ? 1 : 0);
    if (n == 0) return GAD_OK;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (mode == GAD_MODE_MO) {
        stats_init_kernel<<<1, 1, 0, st>>>(w.stats);
        stats_kernel<<<blocks < 592u ? blocks : 592u, threads, 0, st>>>(sp, em, al, hof, P, n, w.stats);
    }
    keys_kernel<<<blocks, threads, 0, st>>>(sp, em, al, hof, P, n, mode, w.stats, key);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// ---------------------------------------------------------------------------------------------
// stable descending order of fp64 keys (ties keep ascending index)
// ---------------------------------------------------------------------------------------------
__global__ void sort_prepare_kernel(const double *__restrict__ key, long long n, u64 *__restrict__ k, int *__restrict__ v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    k[i] = sortable(key[i]);
    v[i] = (int)i;
}

// sorts (w.k0, w.v0) descending, result copied to `order`
int sort_desc(Workspace &w, long long n, int *order, cudaStream_t st)
{
    cub::DoubleBuffer<u64> keys(w.k0, w.k1);
    cub::DoubleBuffer<int> vals(w.v0, w.v1);
    size_t bytes = w.cub_bytes;
    GAD_CUDA(cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp, bytes, keys, vals, (int)n, 0, 64, st));
    GAD_CUDA(cudaMemcpyAsync(order, vals.Current(), sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
    return GAD_OK;
}

// ---------------------------------------------------------------------------------------------
// selection
// ---------------------------------------------------------------------------------------------
__global__ void fill_int_kernel(int *p, long long n, int v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void mark_prefix_kernel(const int *__restrict__ order, long long top_n, int *__restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < top_n) flags[order[t]] = 1;
}

// keys for the MO sorts: which = 0 -> cs of element i, 1 -> sp of element i, 2 -> sp of element perm[i]
__global__ void mo_sort_keys_kernel(const int *__restrict__ sp, const int *__restrict__ em, const int *__restrict__ al,
                                    const int *__restrict__ perm, long long n, int which, u64 *__restrict__ k,
                                    int *__restrict__ v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int src = which == 2 ? perm[i] : (int)i;
    k[i] = which == 0 ? sortable(column_score(em[src], al[src])) : sortable((double)sp[src]);
    v[i] = src;
}

// Along the (sp desc, cs desc, index asc) order: group heads and the cs sequence for the max-scan.
__global__ void front_prepare_kernel(const int *__restrict__ order, const int *__restrict__ sp,
                                     const int *__restrict__ em, const int *__restrict__ al, long long n,
                                     u64 *__restrict__ cs_seq, int *__restrict__ head_pos)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = order[t];
    cs_seq[t] = (u64)__double_as_longlong(column_score(em[i], al[i]));
    head_pos[t] = (t == 0 || sp[order[t - 1]] != sp[i]) ? (int)t : 0;
}

// GA.py:199-209 in sorted order: i is dominated iff a strictly-higher-sp point has cs >= cs_i (prefix max
// over earlier groups) or a same-sp point has cs > cs_i (the group's first element holds the group max).
__global__ void front_flags_kernel(const u64 *__restrict__ cs_seq, const u64 *__restrict__ cs_prefmax,
                                   const int *__restrict__ gstart, long long n, int *__restrict__ front)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int g = gstart[t];
    const u64 c = cs_seq[t];
    bool keep = c == cs_seq[g];
    if (g > 0) keep = keep && c > cs_prefmax[g - 1];
    front[t] = keep ? 1 : 0;
}

// GA.py:246-254. front_rank = exclusive scan of `front` along the sorted order.
__global__ void mo_decide_kernel(const int *__restrict__ order, const int *__restrict__ front,
                                 const int *__restrict__ front_rank, const int *__restrict__ ord_sp,
                                 const int *__restrict__ ord_cs, long long n, long long top_n, Stats *st,
                                 int *__restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int n_front = front_rank[n - 1] + front[n - 1];
    if (t == 0) st->n_front = n_front;
    if (top_n < 0) {                                       // the Pareto front itself (GA._find_pareto_front)
        if (front[t]) flags[order[t]] = 1;
    } else if (n_front < top_n) {
        if (front[t]) flags[order[t]] = 1;
        if (t < top_n) {
            flags[ord_sp[t]] = 1;
            flags[ord_cs[t]] = 1;
        }
    } else if (front[t] && front_rank[t] < top_n) {
        flags[order[t]] = 1;
    }
}

__global__ void keep_finish_kernel(const int *__restrict__ flags, const int *__restrict__ scan, long long n,
                                   uint8_t *__restrict__ keep, int *__restrict__ dst, int *__restrict__ count, Stats *st)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keep[i] = (uint8_t)flags[i];
    dst[i] = scan[i];
    if (i == n - 1) {
        const int c = scan[i] + flags[i];
        *count = c;
        st->n_keep = c;
    }
}

// One warp per (board,row) unit: survivors move to their compacted slot, 16 bytes per lane per trip.
__global__ void __launch_bounds__(256)
compact_kernel(const uint8_t *__restrict__ src, const int *__restrict__ src_rowlen, const int *__restrict__ src_sp,
               const int *__restrict__ src_em, const int *__restrict__ src_al, const uint8_t *__restrict__ keep,
               const int *__restrict__ dst_of, long long P, int R, int stride, uint8_t *__restrict__ dst,
               int *__restrict__ dst_rowlen, int *__restrict__ dst_sp, int *__restrict__ dst_em,
               int *__restrict__ dst_al)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long units = P * R;
    for (long long u = warp0; u < units; u += nwarps) {
        const long long p = u / R;
        const int r = (int)(u - p * R);
        if (!keep[p]) continue;
        const long long q = dst_of[p];
        int w = 0;
        for (int i = 0; i < R; ++i) w = max(w, src_rowlen[p * R + i]);
        const int n16 = min(stride >> 4, (w + 15) >> 4);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + ((size_t)p * R + r) * stride);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + ((size_t)q * R + r) * stride);
        for (int j = lane; j < n16; j += 32) d4[j] = s4[j];
        if (lane == 0) {
            dst_rowlen[q * R + r] = src_rowlen[p * R + r];
            if (r == 0) {
                dst_sp[q] = src_sp[p];
                dst_em[q] = src_em[p];
                dst_al[q] = src_al[p];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// hall of fame
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
argbest_kernel(const double *__restrict__ key, long long n, double *__restrict__ part_key,
               int *__restrict__ part_idx, Stats *st)
{
    __shared__ double sk[8];
    __shared__ int si[8];
    __shared__ bool last;
    double bk = 0.0;
    int bi = 0x7fffffff;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double k = key[i];
        if (bi == 0x7fffffff || better(k, (int)i, bk, bi)) {
            bk = k;
            bi = (int)i;
        }
    }
    auto warp_reduce = [&]() {
        for (int o = 16; o > 0; o >>= 1) {
            const double ok = __shfl_xor_sync(0xFFFFFFFFu, bk, o);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ok, oi, bk, bi))) {
                bk = ok;
                bi = oi;
            }
        }
    };
    warp_reduce();
    if ((threadIdx.x & 31) == 0) {
        sk[threadIdx.x >> 5] = bk;
        si[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        bk = threadIdx.x < 8 ? sk[threadIdx.x] : 0.0;
        bi = threadIdx.x < 8 ? si[threadIdx.x] : 0x7fffffff;
        warp_reduce();
        if (threadIdx.x == 0) {
            part_key[blockIdx.x] = bk;
            part_idx[blockIdx.x] = bi;
            __threadfence();
            last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    bk = 0.0;
    bi = 0x7fffffff;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        const double k = part_key[b];
        const int i = part_idx[b];
        if (i != 0x7fffffff && (bi == 0x7fffffff || better(k, i, bk, bi))) {
            bk = k;
            bi = i;
        }
    }
    warp_reduce();
    if ((threadIdx.x & 31) == 0) {
        sk[threadIdx.x >> 5] = bk;
        si[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        bk = threadIdx.x < 8 ? sk[threadIdx.x] : 0.0;
        bi = threadIdx.x < 8 ? si[threadIdx.x] : 0x7fffffff;
        warp_reduce();
        if (threadIdx.x == 0) {
            st->winner = bi;
            st->ticket = 0u;
        }
    }
}

// GA.py:110-136: the winner's board and score tuple become the hall of fame. Winner == P is the old
// hall of fame itself, re-labelled with the fresh index len(scores) (GA.py:127-132).
__global__ void hof_commit_kernel(const uint8_t *__restrict__ boards, const int *__restrict__ sp,
                                  const int *__restrict__ em, const int *__restrict__ al, long long P, int R,
                                  int stride, const Stats *st, uint8_t *__restrict__ hof_board, int *__restrict__ hof)
{
    const int win = st->winner;
    const int r = blockIdx.x;
    if (win >= P) {
        if (r == 0 && threadIdx.x == 0) hof[1] = (int)P;
        return;
    }
    const int w = al[win];
    const int n16 = min(stride >> 4, (w + 15) >> 4);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(boards + ((size_t)win * R + r) * stride);
    uint4 *d4 = reinterpret_cast<uint4 *>(hof_board + (size_t)r * stride);
    for (int j = threadIdx.x; j < n16; j += blockDim.x) d4[j] = s4[j];
    if (r == 0 && threadIdx.x == 0) {
        hof[0] = 1;
        hof[1] = win;
        hof[2] = sp[win];
        hof[3] = em[win];
        hof[4] = w;
    }
}

// ---------------------------------------------------------------------------------------------
// GA.update_hall_of_fame (GA.py:96-136) as ONE launch: score ranges ('mo'), keys, first maximum, and the
// winner's board + record - what stats_init / stats / keys / argbest / hof_commit do in five.
// The grid is at most one block per SM, so all blocks are resident and may meet at a counter: phase A leaves
// every block's score range in the workspace, the barrier makes them visible, every block folds all of them
// (a few hundred values); phase B forms each key with the arithmetic of keys_kernel and keeps the best per
// block; the last block to finish settles the winner and commits. sync[0] = barrier counter, sync[1] = ticket:
// zero before the first launch, left zero by every launch.
//
// Sharded populations (gad_pass_decide): sp/em/al are the rows of this rank's score table - filled by every
// rank's gad_pass_scatter over NVLink - and the kernel first waits for the arrival flag of every rank. Every
// rank takes the same decision; the rank that owns the winner stores its board into every rank's hall of fame
// and raises their `board` flag, the others wait for that flag. No collective call, nothing read by the host.
// ---------------------------------------------------------------------------------------------
struct PassPeer {                    // all null / zero for a population that lives on one GPU
    const long long *peer_base;      // device array [world]: every rank's pass block (symmetric memory)
    long long flags_off, table_off, hof_off;   // byte offsets inside a pass block
    long long cap1;                  // table row length (int32 [2][5][cap1])
    int rank, world;
    int hof_pitch;                   // row pitch of the hall-of-fame board inside the pass block
    long long stage_off, cap_s;      // staging area (0 = none): int32 [2][world][4 + 4 * cap_s], see gad_pass_scatter
};

// one sender's block of a pass in a receiver's staging area: header {count, -, -, -} | sp | em | al | logical index
__device__ __forceinline__ int *stage_block(char *base, const PassPeer &pp, int parity, int sender)
{
    return reinterpret_cast<int *>(base + pp.stage_off) + ((size_t)parity * pp.world + sender) * (size_t)(4 + 4 * pp.cap_s);
}

// words of the flags area of a pass block (int32 [64])
enum { PF_ARRIVE = 0, PF_BOARD = 16, PF_SEQ = 32, PF_SCATTER_TICKET = 33, PF_DECIDE_SYNC = 34, PF_ERROR = 36 };

__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// wait until *p >= want (a peer's store); a peer that never arrives must not hang the GPU: trap after 20 s
__device__ __forceinline__ void wait_flag(const int *p, int want, int *err)
{
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(p) - want < 0) {
        if ((++spins & 1023u) == 0u && global_ns() - t0 > 20000000000ull) {
            *err = 1;
            __threadfence_system();
            __trap();
        }
    }
}

// ---- pieces shared by the two shapes of the decision kernel ------------------------------------------------
__device__ __forceinline__ Range range_scan(const int *sp, const int *em, const int *al, const int *hofp, long long P,
                                            long long n, long long first, long long step)
{
    Range g;
    range_clear(g);
    for (long long i = first; i < n; i += step) {
        int s, e, a;
        load_score(sp, em, al, hofp, P, i, s, e, a);
        const u64 c = (u64)__double_as_longlong(column_score(e, a));
        range_merge(g, s, s, c, c);
    }
    return g;
}

__device__ __forceinline__ void best_scan(const int *sp, const int *em, const int *al, const int *hofp, long long P,
                                          long long n, long long first, long long step, int mode, const Range &g,
                                          double &bk, int &bi)
{
    bk = 0.0;
    bi = 0x7fffffff;
    for (long long i = first; i < n; i += step) {
        int s, e, a;
        load_score(sp, em, al, hofp, P, i, s, e, a);
        best_merge(bk, bi, make_key(mode, s, e, a, g.lo, g.hi, g.clo, g.chi), (int)i);
    }
}

struct DecideView {                  // where this launch finds its scores, flags and sync words
    const int *sp, *em, *al, *owner, *slot;
    int *flags, *sync;
    int seq;
    int *table;                      // sharded: this pass's table (written here when the scores arrive staged)
    char *mine;                      // sharded: this rank's pass block
};

// The same two scans over STAGED scores (gad_pass_scatter with a staging area): every sender left its scores as
// contiguous arrays with the logical index beside them. The second scan also files every tuple under its logical
// index in this pass's table - the arrays the ranking and the selection read next - with its owner and slot.
__device__ __forceinline__ Range range_scan_staged(const DecideView &v, const PassPeer &pp, const int *hofp,
                                                   long long first, long long step)
{
    Range g;
    range_clear(g);
    for (int r = 0; r < pp.world; ++r) {
        const int *blk = stage_block(v.mine, pp, v.seq & 1, r);
        const int n_r = blk[0];
        const int *s_sp = blk + 4, *s_em = s_sp + pp.cap_s, *s_al = s_em + pp.cap_s;
        for (long long i = first; i < n_r; i += step) {
            const u64 c = (u64)__double_as_longlong(column_score(s_em[i], s_al[i]));
            range_merge(g, s_sp[i], s_sp[i], c, c);
        }
    }
    if (hofp && first == 0) {
        const u64 c = (u64)__double_as_longlong(column_score(hofp[3], hofp[4]));
        range_merge(g, hofp[2], hofp[2], c, c);
    }
    return g;
}

__device__ __forceinline__ void best_scan_staged(const DecideView &v, const PassPeer &pp, const int *hofp, long long P,
                                                 long long first, long long step, int mode, const Range &g, double &bk, int &bi)
{
    bk = 0.0;
    bi = 0x7fffffff;
    int *t_sp = v.table, *t_em = t_sp + pp.cap1, *t_al = t_em + pp.cap1, *t_owner = t_al + pp.cap1, *t_slot = t_owner + pp.cap1;
    for (int r = 0; r < pp.world; ++r) {
        const int *blk = stage_block(v.mine, pp, v.seq & 1, r);
        const int n_r = blk[0];
        const int *s_sp = blk + 4, *s_em = s_sp + pp.cap_s, *s_al = s_em + pp.cap_s, *s_gi = s_al + pp.cap_s;
        for (long long i = first; i < n_r; i += step) {
            const int s2 = s_sp[i], e2 = s_em[i], a2 = s_al[i], gi = s_gi[i];
            t_sp[gi] = s2;
            t_em[gi] = e2;
            t_al[gi] = a2;
            t_owner[gi] = r;
            t_slot[gi] = (int)i;
            best_merge(bk, bi, make_key(mode, s2, e2, a2, g.lo, g.hi, g.clo, g.chi), gi);
        }
    }
    if (hofp && first == 0) best_merge(bk, bi, make_key(mode, hofp[2], hofp[3], hofp[4], g.lo, g.hi, g.clo, g.chi), (int)P);
}

// Sharded: wait for every rank's arrival flag of this pass and take the pass's table. Whole block.
__device__ __forceinline__ void decide_enter(DecideView &v, const PassPeer &pp)
{
    v.owner = v.slot = nullptr;
    v.flags = nullptr;
    v.seq = 0;
    v.table = nullptr;
    v.mine = nullptr;
    if (!pp.peer_base) return;
    char *mine = reinterpret_cast<char *>(pp.peer_base[pp.rank]);
    v.mine = mine;
    v.flags = reinterpret_cast<int *>(mine + pp.flags_off);
    v.sync = v.flags + PF_DECIDE_SYNC;
    v.seq = *reinterpret_cast<volatile int *>(v.flags + PF_SEQ);              // set by this pass's scatter kernel
    if ((int)threadIdx.x < pp.world) wait_flag(v.flags + PF_ARRIVE + threadIdx.x, v.seq, v.flags + PF_ERROR);
    __syncthreads();
    int *table = reinterpret_cast<int *>(mine + pp.table_off) + (size_t)(v.seq & 1) * 5 * pp.cap1;
    v.table = table;
    v.sp = table;
    v.em = table + pp.cap1;
    v.al = table + 2 * pp.cap1;
    v.owner = table + 3 * pp.cap1;
    v.slot = table + 4 * pp.cap1;
}

// GA.py:110-136 by one whole block. Winner == P: the record keeps its place under the fresh index len(scores).
__device__ __forceinline__ void decide_commit(int win, long long P, const DecideView &v, const uint8_t *boards, int R,
                                              int stride, uint8_t *hof_board, int *hof, const PassPeer &pp)
{
    const int tid = threadIdx.x;
    if (win >= P) {
        if (tid == 0) hof[1] = (int)P;
        return;
    }
    const int width = v.al[win];
    const int n16 = min(stride >> 4, (width + 15) >> 4);
    if (!pp.peer_base) {
        for (int u = tid; u < R * n16; u += blockDim.x) {
            const int r = u / n16, j = u - r * n16;
            reinterpret_cast<uint4 *>(hof_board + (size_t)r * stride)[j] =
                reinterpret_cast<const uint4 *>(boards + ((size_t)win * R + r) * stride)[j];
        }
    } else if (v.owner[win] == pp.rank) {
        const size_t src = (size_t)v.slot[win] * R;
        for (int q = 0; q < pp.world; ++q) {
            uint8_t *dst = reinterpret_cast<uint8_t *>(pp.peer_base[q]) + pp.hof_off;
            for (int u = tid; u < R * n16; u += blockDim.x) {
                const int r = u / n16, j = u - r * n16;
                reinterpret_cast<uint4 *>(dst + (size_t)r * pp.hof_pitch)[j] =
                    reinterpret_cast<const uint4 *>(boards + (src + r) * stride)[j];
            }
        }
        __threadfence_system();
        __syncthreads();
        if (tid < pp.world)
            st_release_sys(reinterpret_cast<int *>(reinterpret_cast<char *>(pp.peer_base[tid]) + pp.flags_off) + PF_BOARD, v.seq);
    } else if (tid == 0) {
        wait_flag(v.flags + PF_BOARD, v.seq, v.flags + PF_ERROR);          // the owner's copy of the board has landed here
    }
    if (tid == 0) {
        hof[0] = 1;
        hof[1] = win;
        hof[2] = v.sp[win];
        hof[3] = v.em[win];
        hof[4] = width;
    }
}

// ---- the kernel: any population size; one block per SM at most, blocks meet at a counter ---------------------------
__global__ void __launch_bounds__(256)
hof_decide_kernel(const int *__restrict__ sp_in, const int *__restrict__ em_in, const int *__restrict__ al_in,
                  long long P, int mode, int hof_valid, const uint8_t *__restrict__ boards, int R, int stride,
                  Workspace w, int *sync, uint8_t *__restrict__ hof_board, int *hof, PassPeer pp)
{
    __shared__ DecideShared sh;
    const int tid = threadIdx.x;
    const BlockCtx cx{tid, (int)blockDim.x, 0};
    DecideView v;
    v.sp = sp_in, v.em = em_in, v.al = al_in, v.sync = sync;
    decide_enter(v, pp);
    const int *hofp = hof_valid ? hof : nullptr;
    const long long n = P + (hof_valid ? 1 : 0);
    const long long first = (long long)blockIdx.x * blockDim.x + tid, step = (long long)gridDim.x * blockDim.x;

    // ---- phase A ('mo' only): score ranges over the population and the record
    Range g;
    range_clear(g);
    if (mode == GAD_MODE_MO) {
        const bool staged = pp.peer_base != nullptr && pp.stage_off != 0;
        g = range_fold_block(staged ? range_scan_staged(v, pp, hofp, first, step)
                                    : range_scan(v.sp, v.em, v.al, hofp, P, n, first, step), sh, cx);
        if (gridDim.x > 1) {
            if (tid == 0) {
                w.part_lo[blockIdx.x] = g.lo;
                w.part_hi[blockIdx.x] = g.hi;
                w.part_clo[blockIdx.x] = g.clo;
                w.part_chi[blockIdx.x] = g.chi;
                __threadfence();
                atomicAdd(&v.sync[0], 1);
                while (*reinterpret_cast<volatile int *>(&v.sync[0]) < (int)gridDim.x) {}
                __threadfence();
            }
            __syncthreads();
            range_clear(g);
            for (int b = tid; b < (int)gridDim.x; b += blockDim.x)
                range_merge(g, __ldcg(&w.part_lo[b]), __ldcg(&w.part_hi[b]), __ldcg(&w.part_clo[b]), __ldcg(&w.part_chi[b]));
            g = range_fold_block(g, sh, cx);
        }
    }

    // ---- phase B: keys and the first maximum
    double bk;
    int bi;
    if (pp.peer_base != nullptr && pp.stage_off != 0) best_scan_staged(v, pp, hofp, P, first, step, mode, g, bk, bi);
    else best_scan(v.sp, v.em, v.al, hofp, P, n, first, step, mode, g, bk, bi);
    best_fold_block(bk, bi, sh, cx);
    if (tid == 0) {
        sh.last = true;
        if (gridDim.x > 1) {
            w.part_key[blockIdx.x] = bk;
            w.part_idx[blockIdx.x] = bi;
            __threadfence();
            sh.last = atomicAdd(&v.sync[1], 1) == (int)gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!sh.last) return;
    if (gridDim.x > 1) {
        __threadfence();
        bk = 0.0;
        bi = 0x7fffffff;
        for (int b = tid; b < (int)gridDim.x; b += blockDim.x) best_merge(bk, bi, __ldcg(&w.part_key[b]), __ldcg(&w.part_idx[b]));
        __syncthreads();
        best_fold_block(bk, bi, sh, cx);
    }
    if (tid == 0) {
        sh.winner = bi;
        v.sync[0] = 0;
        v.sync[1] = 0;
        w.stats->winner = bi;
    }
    __syncthreads();
    decide_commit(sh.winner, P, v, boards, R, stride, hof_board, hof, pp);
}

inline unsigned decide_blocks(long long n)
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1)
            v = GAD_NUM_SMS;
        sms = v;
    }
    long long b = (n + 511) / 512;                   // two elements per thread before another block pays off
    if (b > sms) b = sms;
    if (b > kArgBlocks) b = kArgBlocks;
    return (unsigned)(b < 1 ? 1 : b);
}

int launch_decide(const int *sp, const int *em, const int *al, long long P, int mode, int hof_valid, const uint8_t *boards,
                  int R, int stride, const Workspace &w, int *sync, uint8_t *hof_board, int *hof, const PassPeer &pp,
                  cudaStream_t st)
{
    const long long n = P + (hof_valid ? 1 : 0);
    hof_decide_kernel<<<decide_blocks(n), 256, 0, st>>>(sp, em, al, P, mode, hof_valid, boards, R, stride, w, sync, hof_board,
                                                        hof, pp);
    return GAD_OK;
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

__global__ void copy_winner_kernel(const Stats *st, int32_t *out) { *out = st->winner; }

}  // namespace


// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int gad_workspace_bytes(int64_t P, size_t *bytes)
{
    GAD_REQUIRE(bytes && P >= 0, "gad_workspace_bytes: bad argument");
    Workspace w;
    const long long n = P + 1;
    *bytes = carve(w, nullptr, n, cub_temp_bytes(n));
    return GAD_OK;
}

extern "C" int gad_rank_keys(const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, int64_t P, int mode,
                             const int32_t *d_hof, double *d_key, void *d_ws, size_t ws_bytes, void *stream)
{
    GAD_REQUIRE(d_sp && d_em && d_al && d_key, "gad_rank_keys: null device pointer");
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_rank_keys: unknown mode %d", mode);
    Workspace w;
    int rc = open_workspace(w, d_ws, ws_bytes, P + 1, "gad_rank_keys");
    if (rc) return rc;
    return compute_keys(d_sp, d_em, d_al, d_hof, P, mode, d_key, w, gad_stream(stream));
}

extern "C" int gad_rank_order(const double *d_key, int64_t n, int32_t *d_order, void *d_ws, size_t ws_bytes,
                              void *stream)
{
    GAD_REQUIRE(d_key && d_order && n >= 0, "gad_rank_order: bad argument");
    if (n == 0) return GAD_OK;
    Workspace w;
    int rc = open_workspace(w, d_ws, ws_bytes, n, "gad_rank_order");
    if (rc) return rc;
    cudaStream_t st = gad_stream(stream);
    sort_prepare_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_key, n, w.k0, w.v0);
    GAD_LAUNCH_CHECK();
    return sort_desc(w, n, d_order, st);
}

extern "C" int gad_select(const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, int64_t P, int mode,
                          int64_t top_n, uint8_t *d_keep, int32_t *d_dst, int32_t *d_count, void *d_ws,
                          size_t ws_bytes, void *stream)
{
    GAD_REQUIRE(d_sp && d_em && d_al && d_keep && d_dst && d_count, "gad_select: null device pointer");
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_select: unknown mode %d", mode);
    GAD_REQUIRE(P >= 1 && top_n >= -1, "gad_select: need P >= 1 and top_n >= 0 (or -1: Pareto front only)");
    GAD_REQUIRE(top_n >= 0 || mode == GAD_MODE_MO, "gad_select: the Pareto front needs mode 'mo'");
    if (top_n > P) top_n = P;
    Workspace w;
    int rc = open_workspace(w, d_ws, ws_bytes, P + 1, "gad_select");
    if (rc) return rc;
    cudaStream_t st = gad_stream(stream);
    const unsigned nb = blocks_for(P, 256);
    fill_int_kernel<<<nb, 256, 0, st>>>(w.flags, P, 0);
    if (mode != GAD_MODE_MO) {
        // GA.py:234-238: stable descending sort on the single metric, first top_n positions
        mo_sort_keys_kernel<<<nb, 256, 0, st>>>(d_sp, d_em, d_al, nullptr, P, mode == GAD_MODE_SP ? 1 : 0, w.k0, w.v0);
        GAD_LAUNCH_CHECK();
        if ((rc = sort_desc(w, P, w.ord_a, st))) return rc;
        if (top_n > 0) mark_prefix_kernel<<<blocks_for(top_n, 256), 256, 0, st>>>(w.ord_a, top_n, w.flags);
    } else {
        // order by cs (also "top_n by CS"), by sp ("top_n by SP"), and by (sp, cs) for the front
        mo_sort_keys_kernel<<<nb, 256, 0, st>>>(d_sp, d_em, d_al, nullptr, P, 0, w.k0, w.v0);
        GAD_LAUNCH_CHECK();
        if ((rc = sort_desc(w, P, w.ord_b, st))) return rc;                       // ord_b = by cs
        mo_sort_keys_kernel<<<nb, 256, 0, st>>>(d_sp, d_em, d_al, nullptr, P, 1, w.k0, w.v0);
        if ((rc = sort_desc(w, P, w.ord_a, st))) return rc;                       // ord_a = by sp
        mo_sort_keys_kernel<<<nb, 256, 0, st>>>(d_sp, d_em, d_al, w.ord_b, P, 2, w.k0, w.v0);
        if ((rc = sort_desc(w, P, w.ord_c, st))) return rc;                       // ord_c = by (sp, cs)
        const int *order = w.ord_c;
        front_prepare_kernel<<<nb, 256, 0, st>>>(order, d_sp, d_em, d_al, P, w.k0, w.v1);   // k0 = cs seq, v1 = head pos
        GAD_LAUNCH_CHECK();
        size_t bytes = w.cub_bytes;
        GAD_CUDA(cub::DeviceScan::InclusiveScan(w.cub_tmp, bytes, w.k0, w.scan_u64, cub::Max(), (int)P, st));
        bytes = w.cub_bytes;
        GAD_CUDA(cub::DeviceScan::InclusiveScan(w.cub_tmp, bytes, w.v1, w.gstart, cub::Max(), (int)P, st));
        front_flags_kernel<<<nb, 256, 0, st>>>(w.k0, w.scan_u64, w.gstart, P, w.v1);        // v1 = front flag
        bytes = w.cub_bytes;
        GAD_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, w.v1, w.scan, (int)P, st));
        mo_decide_kernel<<<nb, 256, 0, st>>>(order, w.v1, w.scan, w.ord_a, w.ord_b, P, top_n, w.stats, w.flags);
    }
    GAD_LAUNCH_CHECK();
    size_t bytes = w.cub_bytes;
    GAD_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, w.flags, w.scan, (int)P, st));
    keep_finish_kernel<<<nb, 256, 0, st>>>(w.flags, w.scan, P, d_keep, d_dst, d_count, w.stats);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

extern "C" int gad_compact(const uint8_t *d_src_boards, const int32_t *d_src_rowlen, const int32_t *d_src_sp,
                           const int32_t *d_src_em, const int32_t *d_src_al, const uint8_t *d_keep,
                           const int32_t *d_dst, int64_t P, int R, int stride, uint8_t *d_dst_boards,
                           int32_t *d_dst_rowlen, int32_t *d_dst_sp, int32_t *d_dst_em, int32_t *d_dst_al,
                           void *stream)
{
    GAD_REQUIRE(d_src_boards && d_src_rowlen && d_src_sp && d_src_em && d_src_al && d_keep && d_dst && d_dst_boards &&
                    d_dst_rowlen && d_dst_sp && d_dst_em && d_dst_al,
                "gad_compact: null device pointer");
    GAD_REQUIRE(P >= 0 && R >= 1 && stride >= 16 && stride % 16 == 0, "gad_compact: bad shape");
    GAD_REQUIRE(d_src_boards != d_dst_boards, "gad_compact: source and destination must be different buffers");
    if (P == 0) return GAD_OK;
    const long long units = (long long)P * R;
    long long blocks = (units + 7) / 8;
    const long long cap = (long long)GAD_NUM_SMS * 8 * 4;
    if (blocks > cap) blocks = cap;
    compact_kernel<<<(unsigned)blocks, 256, 0, gad_stream(stream)>>>(d_src_boards, d_src_rowlen, d_src_sp, d_src_em,
                                                                      d_src_al, d_keep, d_dst, P, R, stride,
                                                                      d_dst_boards, d_dst_rowlen, d_dst_sp, d_dst_em,
                                                                      d_dst_al);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// Index of the best element of population (+ hall-of-fame record as element P) - the decision of
// GA.update_hall_of_fame without the board copy (used when the boards are sharded over several GPUs).
extern "C" int gad_hof_argbest(const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, int64_t P, int mode,
                               const int32_t *d_hof, int32_t *d_winner, void *d_ws, size_t ws_bytes, void *stream)
{
    GAD_REQUIRE(d_sp && d_em && d_al && d_winner, "gad_hof_argbest: null device pointer");
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_hof_argbest: unknown mode %d", mode);
    GAD_REQUIRE(P >= 1, "gad_hof_argbest: empty population");
    Workspace w;
    int rc = open_workspace(w, d_ws, ws_bytes, P + 1, "gad_hof_argbest");
    if (rc) return rc;
    cudaStream_t st = gad_stream(stream);
    double *key = reinterpret_cast<double *>(w.k0);
    const long long n = P + (d_hof ? 1 : 0);
    if ((rc = compute_keys(d_sp, d_em, d_al, d_hof, P, mode, key, w, st))) return rc;
    GAD_CUDA(cudaMemsetAsync(&w.stats->ticket, 0, sizeof(unsigned), st));
    unsigned blocks = blocks_for(n, 256 * 4);
    if (blocks > (unsigned)kArgBlocks) blocks = kArgBlocks;
    if (blocks < 1) blocks = 1;
    argbest_kernel<<<blocks, 256, 0, st>>>(key, n, w.part_key, w.part_idx, w.stats);
    copy_winner_kernel<<<1, 1, 0, st>>>(w.stats, d_winner);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// fitness.cu (gad_fitness_pass): where a decision taken in the tail of the fitness kernel leaves its partials
int gad_decide_parts(void *d_ws, size_t ws_bytes, long long n, const char *who, GadDecideParts *out)
{
    Workspace w;
    int rc = open_workspace(w, d_ws, ws_bytes, n, who);
    if (rc) return rc;
    out->key = w.part_key;
    out->idx = w.part_idx;
    out->lo = w.part_lo;
    out->hi = w.part_hi;
    out->clo = w.part_clo;
    out->chi = w.part_chi;
    out->rec = w.part_rec;
    out->winner = &w.stats->winner;
    return GAD_OK;
}

extern "C" int gad_hof_update(const uint8_t *d_boards, const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al,
                              int64_t P, int R, int stride, int mode, uint8_t *d_hof_board, int32_t *d_hof,
                              int hof_valid, void *d_ws, size_t ws_bytes, void *stream)
{
    GAD_REQUIRE(d_boards && d_sp && d_em && d_al && d_hof_board && d_hof, "gad_hof_update: null device pointer");
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_hof_update: unknown mode %d", mode);
    GAD_REQUIRE(P >= 1 && R >= 1 && stride >= 16 && stride % 16 == 0, "gad_hof_update: bad shape");
    Workspace w;
    int rc = open_workspace(w, d_ws, ws_bytes, P + 1, "gad_hof_update");
    if (rc) return rc;
    cudaStream_t st = gad_stream(stream);
    const long long n = P + (hof_valid ? 1 : 0);
    static const bool unfused = getenv("GAD_HOF_UNFUSED") && atoi(getenv("GAD_HOF_UNFUSED")) != 0;
    if (!unfused) {                                  // one launch; d_hof[5..6] = its barrier counter and ticket
        if ((rc = launch_decide(d_sp, d_em, d_al, P, mode, hof_valid ? 1 : 0, d_boards, R, stride, w, d_hof + 5, d_hof_board,
                                d_hof, PassPeer{}, st)))
            return rc;
        GAD_LAUNCH_CHECK();
        return GAD_OK;
    }
    double *key = reinterpret_cast<double *>(w.k0);
    if ((rc = compute_keys(d_sp, d_em, d_al, hof_valid ? d_hof : nullptr, P, mode, key, w, st))) return rc;
    GAD_CUDA(cudaMemsetAsync(&w.stats->ticket, 0, sizeof(unsigned), st));
    unsigned blocks = blocks_for(n, 256 * 4);
    if (blocks > (unsigned)kArgBlocks) blocks = kArgBlocks;
    if (blocks < 1) blocks = 1;
    argbest_kernel<<<blocks, 256, 0, st>>>(key, n, w.part_key, w.part_idx, w.stats);
    hof_commit_kernel<<<R, 128, 0, st>>>(d_boards, d_sp, d_em, d_al, P, R, stride, w.stats, d_hof_board, d_hof);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// =============================================================================================
// The fitness pass of a SHARDED population without a collective call (sharded.py): gad_pass_scatter stores the
// scores of the local individuals into every rank's score table over NVLink and raises this rank's arrival flag
// on every rank; gad_pass_decide waits for all flags, takes GA.update_hall_of_fame's decision on the complete
// table (the same on every rank) and moves the winner's board from its owner into every rank's hall of fame.
// A pass block (symmetric memory, one per rank, zeroed once): int32 flags[64] | int32 table[2][5][cap1] | board.
// Tables alternate with the pass number: a rank that runs ahead writes the other table, and cannot start the
// pass after that before every rank has raised its flag for this one.
// =============================================================================================
namespace {

__global__ void __launch_bounds__(256)
pass_scatter_kernel(PassPeer pp, const int32_t *__restrict__ sp, const int32_t *__restrict__ em,
                    const int32_t *__restrict__ al, const long long *__restrict__ gidx, long long n_local)
{
    __shared__ bool last;
    int *flags = reinterpret_cast<int *>(reinterpret_cast<char *>(pp.peer_base[pp.rank]) + pp.flags_off);
    const int seq = *reinterpret_cast<volatile int *>(flags + PF_SEQ) + 1;       // this pass
    const size_t toff = (size_t)pp.table_off + (size_t)(seq & 1) * 5 * pp.cap1 * sizeof(int32_t);
    if (pp.stage_off != 0) {
        // staged: this rank's block in every receiver's staging area, as contiguous arrays - 16 bytes per store, whole
        // lines over NVLink instead of 4-byte stores strided by the logical index; the receiver files them itself
        const long long quads = (n_local + 3) >> 2;
        const long long total = quads * 4 * pp.world;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
            const long long q = t % quads;                 // consecutive threads: consecutive quads of one array
            const int k = (int)((t / quads) & 3), r = (int)(t / (quads * 4));
            int *blk = stage_block(reinterpret_cast<char *>(pp.peer_base[r]), pp, seq & 1, pp.rank);
            int *dst = blk + 4 + (size_t)k * pp.cap_s + 4 * q;
            int v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long i = 4 * q + u;
                v4[u] = i >= n_local ? 0 : k == 0 ? sp[i] : k == 1 ? em[i] : k == 2 ? al[i] : (int)gidx[i];
            }
            *reinterpret_cast<int4 *>(dst) = make_int4(v4[0], v4[1], v4[2], v4[3]);
            if (q == 0 && k == 0) blk[0] = (int)n_local;
        }
        if (n_local == 0 && blockIdx.x == 0 && (int)threadIdx.x < pp.world)
            stage_block(reinterpret_cast<char *>(pp.peer_base[threadIdx.x]), pp, seq & 1, pp.rank)[0] = 0;
    }
    const long long total = pp.stage_off != 0 ? 0 : n_local * pp.world;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / pp.world;              // consecutive threads: the peers of one individual
        const int r = (int)(t - i * pp.world);
        int32_t *table = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(pp.peer_base[r]) + toff);
        const long long g = gidx[i];
        table[g] = sp[i];
        table[pp.cap1 + g] = em[i];
        table[2 * pp.cap1 + g] = al[i];
        table[3 * pp.cap1 + g] = pp.rank;
        table[4 * pp.cap1 + g] = (int32_t)i;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&flags[PF_SCATTER_TICKET], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x < pp.world)
        st_release_sys(reinterpret_cast<int *>(reinterpret_cast<char *>(pp.peer_base[threadIdx.x]) + pp.flags_off) + PF_ARRIVE + pp.rank, seq);
    if (threadIdx.x == 0) {
        flags[PF_SEQ] = seq;
        flags[PF_SCATTER_TICKET] = 0;
    }
}

int check_pass_block(const char *who, const void *d_peer_base, int64_t flags_off, int64_t table_off, int64_t hof_off,
                     int64_t cap1, int rank, int world)
{
    GAD_REQUIRE(d_peer_base && cap1 >= 1 && world >= 1 && world <= 16 && rank >= 0 && rank < world, "%s: bad argument", who);
    GAD_REQUIRE(flags_off >= 0 && flags_off % 16 == 0 && table_off >= 0 && table_off % 16 == 0 && hof_off >= 0 &&
                hof_off % 16 == 0, "%s: pass-block offsets must be multiples of 16", who);
    return GAD_OK;
}

}  // namespace

extern "C" int gad_pass_scatter(const void *d_peer_base, int64_t flags_off, int64_t table_off, int64_t cap1,
                                int64_t stage_off, int64_t cap_s, const int32_t *d_sp, const int32_t *d_em,
                                const int32_t *d_al, const int64_t *d_gidx, int64_t n_local, int rank, int world, void *stream)
{
    GAD_REQUIRE(stage_off >= 0 && stage_off % 16 == 0 && (stage_off == 0 || (cap_s >= n_local && cap_s % 4 == 0)),
                "gad_pass_scatter: bad staging area (offset %lld, capacity %lld, %lld local individuals)", (long long)stage_off,
                (long long)cap_s, (long long)n_local);
    int rc = check_pass_block("gad_pass_scatter", d_peer_base, flags_off, table_off, 0, cap1, rank, world);
    if (rc) return rc;
    GAD_REQUIRE(n_local >= 0 && (n_local == 0 || (d_sp && d_em && d_al && d_gidx)), "gad_pass_scatter: null pointer");
    PassPeer pp{};
    pp.peer_base = reinterpret_cast<const long long *>(d_peer_base);
    pp.flags_off = flags_off;
    pp.table_off = table_off;
    pp.cap1 = cap1;
    pp.rank = rank;
    pp.world = world;
    pp.stage_off = stage_off;
    pp.cap_s = cap_s;
    long long blocks = (n_local * world + 255) / 256;       // a rank without individuals still raises its flag
    if (blocks < 1) blocks = 1;
    if (blocks > 4 * GAD_NUM_SMS) blocks = 4 * GAD_NUM_SMS;
    pass_scatter_kernel<<<(unsigned)blocks, 256, 0, gad_stream(stream)>>>(pp, d_sp, d_em, d_al,
                                                                          reinterpret_cast<const long long *>(d_gidx), n_local);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

extern "C" int gad_pass_decide(const void *d_peer_base, int64_t flags_off, int64_t table_off, int64_t hof_off,
                               int64_t cap1, int64_t stage_off, int64_t cap_s, int hof_pitch, int64_t n, int mode,
                               int hof_valid, const uint8_t *d_boards, int R, int stride, int32_t *d_hof, int rank,
                               int world, void *d_ws, size_t ws_bytes, void *stream)
{
    GAD_REQUIRE(stage_off >= 0 && stage_off % 16 == 0 && (stage_off == 0 || (cap_s >= 1 && cap_s % 4 == 0)),
                "gad_pass_decide: bad staging area");
    int rc = check_pass_block("gad_pass_decide", d_peer_base, flags_off, table_off, hof_off, cap1, rank, world);
    if (rc) return rc;
    GAD_REQUIRE(d_boards && d_hof, "gad_pass_decide: null device pointer");
    GAD_REQUIRE(mode >= GAD_MODE_SP && mode <= GAD_MODE_MO, "gad_pass_decide: unknown mode %d", mode);
    GAD_REQUIRE(n >= 1 && n < cap1 && R >= 1 && stride >= 16 && stride % 16 == 0 && hof_pitch >= stride && hof_pitch % 16 == 0,
                "gad_pass_decide: bad shape");
    Workspace w;
    if ((rc = open_workspace(w, d_ws, ws_bytes, n + 1, "gad_pass_decide"))) return rc;
    PassPeer pp{};
    pp.peer_base = reinterpret_cast<const long long *>(d_peer_base);
    pp.flags_off = flags_off;
    pp.table_off = table_off;
    pp.hof_off = hof_off;
    pp.cap1 = cap1;
    pp.rank = rank;
    pp.world = world;
    pp.hof_pitch = hof_pitch;
    pp.stage_off = stage_off;
    pp.cap_s = cap_s;
    // the kernel finds its table (pass parity), its flags and the sync words inside the pass block
    if ((rc = launch_decide(nullptr, nullptr, nullptr, n, mode, hof_valid ? 1 : 0, d_boards, R, stride, w, nullptr, nullptr,
                            d_hof, pp, gad_stream(stream))))
        return rc;
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
