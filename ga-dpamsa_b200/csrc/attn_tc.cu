// The encoder's attention on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only: the default encoder
// for agent windows of 16..192 positions (launch_forward in qnet.cu; GAD_QNET_ENCODER=fa selects the mma.sync one).
// Same mathematics as encode_fa_kernel - S = Q'K'^T from split fp16 row tables, softmax in the log2 domain,
// O = P V with the output projection folded into V, residual, LayerNorm - with both products as tcgen05.mma and
// the accumulators in tensor memory.
//
//   work item = (state, tile of TR query rows); TR = 96 when the state's rows split evenly into tiles of <= 96
//   (186 positions -> two tiles of 96, 93 -> one), else 128. Persistent CTAs, two per SM, so that one CTA's staging
//   and softmax overlap the other's MMAs. Warps of a CTA (TR = 96: seven):
//     row warps    two threads per query row: tensor-memory lanes 32q..32q+31 are only reachable from warps with
//                  (warp & 3) == q, so a row's threads sit in warps q and q+4; thread h of a row owns every other
//                  32-key block of S and channels 32h..32h+31 of O, and trades max / sum / LayerNorm partials with
//                  its partner through shared memory (named barrier 1)
//     MMA warp     the first warp that owns no rows: lane 0 issues every tcgen05.mma; while the row warps work on
//                  item n the whole warp builds the tokens and the key list of item n+1 (double-buffered); it also
//                  owns the 256 TMEM columns
//   shared memory: K-major 128-byte swizzled operand tiles, two of them used twice per item
//     QP  hi|lo  [128 rows][64]      32 KB   Q' (A operand of S), then - once S is complete - two buffers of P
//                                            blocks, hi|lo [128][32 keys] each (64-byte swizzled rows), so that the
//                                            softmax of block b+1 runs while the tensor core consumes block b
//     KV  hi|lo  [<=192][64]         48 KB   K' rows (B operand of S), then the V rows of the same keys, which the
//                                            second product reads as an MN-major B operand: no transpose anywhere
//     XR  fp32   [96][64]            24 KB   the tile's residual rows (TR = 96 only; two CTAs still fit an SM)
//   tensor memory: S at columns 0..191 (fp32, one column per key), O at 192..255.
//
//   item: cp.async Q', K', residual rows -> S = 12 MMAs (N = keys) -> commit
//         row threads: wait S; cp.async the V rows over K'; row max from TMEM while they land; per 32-key block:
//                      p = ex2(s - max) as hi|lo fp16 into the block's P buffer, fence.proxy.async, arrive
//         MMA thread : per block: wait P -> 6 MMAs O += P V (N = 64) -> commit (frees the buffer / completes O)
//         row threads: wait O; /sum, residual, LayerNorm over the row's 64 channels, (hi|lo) store
// Every product is the 3-term fp16 split (lo.hi + hi.lo + hi.hi): fp32-class like the other encoders.
#include "qnet.cuh"
#include "tc_ptx.cuh"

#include <math.h>
#include <stdlib.h>

#include <type_traits>

using namespace tcptx;

namespace {

constexpr int AT_M = 128;                    // UMMA M (rows of the accumulators; a tile uses the first TR of them)
constexpr int AT_MAXK = 192;                 // keys
constexpr int AT_TILE_Q = AT_M * 128;        // bytes of one [128][64] fp16 tile
constexpr int AT_TILE_K = AT_MAXK * 128;
constexpr int OFF_QH = 0;                    // Q' hi, later P hi
constexpr int OFF_QL = OFF_QH + AT_TILE_Q;
constexpr int OFF_KH = OFF_QL + AT_TILE_Q;   // K' hi, later V^T hi (3 atoms)
constexpr int OFF_KL = OFF_KH + AT_TILE_K;
constexpr int OFF_XR = OFF_KL + AT_TILE_K;   // 81 920: the tile's residual rows (fp32, 16-byte chunks swizzled), 96-row tiles
constexpr int COL_S = 0, COL_O = 192;
constexpr int TMEM_COLS = 256;

// K-major tiles with 128-byte rows and 128B swizzle keep the 16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4);
// the same for 64-byte rows with 64B swizzle (the P tiles: 32 keys per row):
__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ float fast_ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One warp builds the tokens of a work item's state (env.py:147-153, same rule as build_tokens in qnet.cuh) and
// the ordered list of its unmasked keys. *nk_out = 0 marks a fully masked state (kj is then the identity).
__device__ __forceinline__ void warp_item_tokens(const gad_qnet &net, const uint8_t *__restrict__ tok_in,
                                                 const int32_t *__restrict__ active, long long row, int L, int lane,
                                                 uint8_t *tok, int *kj, int *nk_out)
{
    if (tok_in) {
        for (int i = lane; i < L; i += 32) tok[i] = tok_in[row * L + i];
    } else {
        const int e = active[row];
        const int rows = net.rows, cols = net.cols;
        const uint8_t *sub = net.sub + (size_t)e * rows * cols;
        const int h = lane < rows ? (int)net.heads[(size_t)e * rows + lane] : cols + 1;
        const int len = cols - h + 1;                     // a row's remaining tokens + its gap code; 0 beyond the rows
        int end = len;                                    // inclusive prefix sum over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, end, d);
            if (lane >= d) end += v;
        }
        const int nvalid = __shfl_sync(0xFFFFFFFFu, end, 31);
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            int r = 0;
            for (int q = 0; q < rows; ++q) r += i >= __shfl_sync(0xFFFFFFFFu, end, q);
            r = min(r, rows - 1);
            const int hr = __shfl_sync(0xFFFFFFFFu, h, r);
            const int off = __shfl_sync(0xFFFFFFFFu, end - len, r);
            uint8_t t = 0;
            if (i < nvalid) {
                const int k = i - off;
                t = k < cols - hr ? sub[r * cols + hr + k] : (uint8_t)GAD_GAP;
            }
            if (i < L) tok[i] = t;
        }
    }
    __syncwarp();
    int base = 0;
    for (int j0 = 0; j0 < L; j0 += 32) {
        const int j = j0 + lane;
        const bool v = j < L && tok[j] != 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, v);
        if (v) kj[base + __popc(m & ((1u << lane) - 1u))] = j;
        base += __popc(m);
    }
    if (base == 0)                                        // reference: softmax over all -1e9 is uniform over every key
        for (int j = lane; j < L; j += 32) kj[j] = j;
    if (lane == 0) *nk_out = base;
    __syncwarp();
}

// Thread roles of a CTA. Tensor-memory lanes 32q..32q+31 are only reachable from warps with (warp & 3) == q, so the two
// threads of a row sit in warps q and q + 4; the first warp that owns no rows issues the MMAs.
template <int TR> struct AtCfg {
    static constexpr int RW = TR / 32;                             // row warps per half
    static constexpr int MMA_WARP = TR < 128 ? RW : 8;
    static constexpr int THREADS = TR < 128 ? (4 + RW) * 32 : 288;
    static constexpr int NROW = 2 * TR;                            // row threads: two per query row
    static constexpr bool XSTAGE = TR < 128;                       // residual rows staged in shared memory (two CTAs still fit)
    static constexpr int OFF_BAR = OFF_XR + (XSTAGE ? TR * 256 : 0);
    static constexpr int SMEM = OFF_BAR + 128 + 1024;
};

template <int TR>
__global__ void __launch_bounds__(AtCfg<TR>::THREADS, 2)
encode_tc_kernel(const gad_qnet net, const uint8_t *__restrict__ tok_in, const int32_t *__restrict__ active,
                 const int32_t *__restrict__ n_rows_ptr, long long n_rows_fixed, const int n_tiles,
                 float *__restrict__ Zout, __half *__restrict__ Z2out)
{
    using Cfg = AtCfg<TR>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + Cfg::OFF_BAR);
    // bars: s_full, o_full, p_ready[2], p_free[2] (one pair per P buffer), then the TMEM base address
    const uint32_t b_sfull = smem_u32(bars), b_ofull = b_sfull + 8, b_pready = b_sfull + 16, b_pfree = b_sfull + 32;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    __shared__ int kj_s[2][AT_MAXK];
    __shared__ uint8_t tok_s[2][AT_MAXK + 64];
    __shared__ int nk_s[2];
    __shared__ float ln_g[QN_D], ln_b[QN_D];
    __shared__ float xch[2][4][AT_M];                     // what the two threads of a row tell each other

    const int L = net.L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, h = warp >> 2;                // TMEM lane quarter, half of the columns
    const bool is_row = q < Cfg::RW && h < 2, is_mma = warp == Cfg::MMA_WARP;
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const long long n_items = n_rows * n_tiles;

    if (threadIdx.x == 0) {
        mbar_init(b_sfull, 1);
        mbar_init(b_ofull, 1);
        mbar_init(b_pready, Cfg::NROW);
        mbar_init(b_pready + 8, Cfg::NROW);
        mbar_init(b_pfree, 1);
        mbar_init(b_pfree + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < QN_D) {
        ln_g[threadIdx.x] = net.ln2_g[threadIdx.x];
        ln_b[threadIdx.x] = net.ln2_b[threadIdx.x];
    }
    // rows TR..127 of the A tiles are never staged: zero them once (their accumulator rows are never read)
    for (int idx = threadIdx.x; idx < (AT_M - TR) * 8; idx += blockDim.x) {
        *reinterpret_cast<uint4 *>(sm + OFF_QH + TR * 128 + idx * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(sm + OFF_QL + TR * 128 + idx * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (is_mma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        if ((long long)blockIdx.x < n_items)
            warp_item_tokens(net, tok_in, active, blockIdx.x / n_tiles, L, lane, tok_s[0], kj_s[0], &nk_s[0]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // s_full and o_full complete once per item; p_ready[b & 1] and p_free[b & 1] once per 32-key block b, so their
    // phase is the running count of blocks of that parity
    uint32_t ph_s = 0, ph_o = 0, done0 = 0, done1 = 0;
    int buf = 0;

    for (long long item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
        const long long row = item / n_tiles;
        const int t = (int)(item - row * n_tiles);
        __syncthreads();                                  // previous item fully consumed; this item's tokens published
        const uint8_t *tok = tok_s[buf];
        const int *kj = kj_s[buf];
        const bool all_masked = nk_s[buf] == 0;
        const int nk = all_masked ? L : nk_s[buf];
        const int nk16 = (nk + 15) & ~15;
        const int n_blocks = (nk16 + 31) >> 5;           // 32-key blocks of P, alternating between two buffers

        // ---- stage Q' (this tile's rows) and K' (the valid keys): cp.async 16-byte chunks into swizzled positions,
        // rows outside the state zero-filled through the src-size operand. Row thread rt owns chunk rt & 7 of rows
        // rt/8, rt/8 + NROW/8, ... : NROW/8 is a multiple of 8, so its swizzled chunk position is one constant and a
        // row's table offset, computed once, serves the K' copy now and the V copy of the same key later.
        constexpr int KIT = (AT_MAXK * 8 + Cfg::NROW - 1) / Cfg::NROW, RSTEP = Cfg::NROW / 8;
        uint32_t koff[KIT], qrow[TR * 8 / Cfg::NROW];     // table offsets (elements) of this thread's key / query rows
        const int st_c = (h * TR + q * 32 + lane) & 7, st_r0 = (h * TR + q * 32 + lane) >> 3;
        const uint32_t st_swz = (uint32_t)((st_c ^ (st_r0 & 7)) << 4);
        if (is_row) {
#pragma unroll
            for (int n = 0; n < TR * 8 / Cfg::NROW; ++n) {
                const int r = st_r0 + RSTEP * n;
                const int i = t * TR + r;
                const int ii = i < L ? i : 0;
                qrow[n] = ((uint32_t)tok[ii] * L + ii) * QN_D;
                const uint32_t off = qrow[n] + st_c * 8;
                const uint32_t nbytes = i < L ? 16u : 0u;
                const uint32_t dst = smem_u32(sm + OFF_QH) + r * 128 + st_swz;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(net.Qh + off), "r"(nbytes) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + AT_TILE_Q), "l"(net.Ql + off), "r"(nbytes)
                             : "memory");
            }
#pragma unroll
            for (int n = 0; n < KIT; ++n) {
                const int k = st_r0 + RSTEP * n;
                koff[n] = 0u;
                if (k < nk16) {
                    const int j = kj[k < nk ? k : 0];
                    koff[n] = ((uint32_t)tok[j] * L + j) * QN_D + st_c * 8;
                    const uint32_t nbytes = k < nk ? 16u : 0u;
                    const uint32_t dst = smem_u32(sm + OFF_KH) + k * 128 + st_swz;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(net.Kh + koff[n]), "r"(nbytes)
                                 : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + AT_TILE_K), "l"(net.Kl + koff[n]),
                                 "r"(nbytes) : "memory");
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        fence_async_smem();                               // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncthreads();

        if (is_mma) {
            // ======================= MMA issuer + next item's tokens =======================
            const uint32_t idesc_s = make_idesc_f16(AT_M, nk16);
            // O's B operand: the V rows as staged, [key][64 ch], read as an MN-major operand (instruction descriptor bit 16):
            // 8-key groups 1024 bytes apart, 16 keys = 2048 bytes per k-step - no transpose anywhere
            const uint32_t idesc_o = make_idesc_f16(AT_M, QN_D) | (1u << 16);
            const uint32_t ds = tmem_base + COL_S, dout = tmem_base + COL_O;
            if (lane == 0) {
                tc_fence_after();
                const uint64_t dkh = make_smem_desc(smem_u32(sm + OFF_KH)), dkl = make_smem_desc(smem_u32(sm + OFF_KL));
                const uint64_t dqh = make_smem_desc(smem_u32(sm + OFF_QH)), dql = make_smem_desc(smem_u32(sm + OFF_QL));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {          // S = Q' K'^T : 4 k-steps x (lo.hi, hi.lo, hi.hi)
                    umma_f16(ds, dql + 2 * ks, dkh + 2 * ks, idesc_s, ks != 0);
                    umma_f16(ds, dqh + 2 * ks, dkl + 2 * ks, idesc_s, 1u);
                    umma_f16(ds, dqh + 2 * ks, dkh + 2 * ks, idesc_s, 1u);
                }
                umma_commit(b_sfull);
            }
            __syncwarp();
            const long long next = item + gridDim.x;      // the row warps are busy for a while: prepare the next item
            if (next < n_items)
                warp_item_tokens(net, tok_in, active, next / n_tiles, L, lane, tok_s[buf ^ 1], kj_s[buf ^ 1],
                                 &nk_s[buf ^ 1]);
            if (lane == 0) {
                for (int b = 0; b < n_blocks; ++b) {      // O += P(block) V(block): P in the halves of the Q' buffer (64-byte
                    const int pb = b & 1;                 // swizzled rows), V rows where K' was
                    const int ksteps = min(2, (nk16 - 32 * b) >> 4);
                    mbar_wait(b_pready + 8 * pb, ((pb ? done1 : done0) + (b >> 1)) & 1);
                    tc_fence_after();
                    const uint64_t dph = make_smem_desc(smem_u32(sm + OFF_QH + pb * AT_TILE_Q), true);
                    const uint64_t dpl = make_smem_desc(smem_u32(sm + OFF_QH + pb * AT_TILE_Q + AT_TILE_Q / 2), true);
                    const uint64_t dvh = make_smem_desc(smem_u32(sm + OFF_KH + b * 32 * 128));
                    const uint64_t dvl = make_smem_desc(smem_u32(sm + OFF_KL + b * 32 * 128));
                    for (int ks = 0; ks < ksteps; ++ks) {                       // descriptor addresses count 16-byte units
                        umma_f16(dout, dpl + 2 * ks, dvh + 128 * ks, idesc_o, (b | ks) != 0);
                        umma_f16(dout, dph + 2 * ks, dvl + 128 * ks, idesc_o, 1u);
                        umma_f16(dout, dph + 2 * ks, dvh + 128 * ks, idesc_o, 1u);
                    }
                    umma_commit(b_pfree + 8 * pb);                              // this P buffer may be overwritten
                    if (b + 1 == n_blocks) umma_commit(b_ofull);                // O complete
                }
            }
        } else if (is_row) {
            // ============ row threads: two per query row, each owns every other 32-column block ============
            const int r = q * 32 + lane;                  // row inside the tile = TMEM lane
            const int i = t * TR + r;                     // query position
            const bool warp_live = t * TR + q * 32 < L;   // warps past the state's last row only stage and arrive
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const float sc = all_masked ? 0.f : net.s_inv_scale;           // fully masked: every p = ex2(0) = 1
            auto row_barrier = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(Cfg::NROW) : "memory"); };

            if (Cfg::XSTAGE) {                            // the residual rows: off the path to S, waited for with the V rows.
#pragma unroll
                for (int n = 0; n < TR * 8 / Cfg::NROW; ++n) {             // Same rows as this thread's Q' rows: chunks c, c + 8
                    const int xr = st_r0 + RSTEP * n;
                    const uint32_t nbytes = t * TR + xr < L ? 16u : 0u;
                    const uint32_t dst = smem_u32(sm + OFF_XR) + xr * 256 + st_swz;
                    const float *src = net.X + qrow[n] + st_c * 4;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 128), "l"(src + 32), "r"(nbytes)
                                 : "memory");
                }
            }
            mbar_wait(b_sfull, ph_s);                     // S complete: Q' and K' are dead, their buffers are free
            tc_fence_after();
#pragma unroll
            for (int n = 0; n < KIT; ++n) {               // V rows over the K' rows, same thread -> (key, chunk) mapping and
                const int k = st_r0 + RSTEP * n;          // table offsets; they land while the row maxima are computed
                if (k < nk16) {
                    const uint32_t nbytes = k < nk ? 16u : 0u;
                    const uint32_t dst = smem_u32(sm + OFF_KH) + k * 128 + st_swz;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(net.VFh + koff[n]), "r"(nbytes)
                                 : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + AT_TILE_K), "l"(net.VFl + koff[n]),
                                 "r"(nbytes) : "memory");
                }
            }

            // pass 1: row maximum over the valid keys; this thread reads the 32-column blocks h, h + 2, h + 4
            float mx = -INFINITY;
            if (warp_live) {
                for (int c0 = 32 * h; c0 < nk16; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + COL_S + c0, v);
                    tmem_ld_wait();
                    if (c0 + 32 <= nk) {
#pragma unroll
                        for (int u = 0; u < 32; ++u) mx = fmaxf(mx, __uint_as_float(v[u]));
                    } else {
#pragma unroll
                        for (int u = 0; u < 32; ++u)
                            if (c0 + u < nk) mx = fmaxf(mx, __uint_as_float(v[u]));
                    }
                }
            }
            xch[h][0][r] = mx;
            row_barrier();
            mx = fmaxf(mx, xch[h ^ 1][0][r]);
            const float mxs = mx * sc;
            asm volatile("cp.async.wait_all;" ::: "memory");             // V rows: published by the fence before the first arrive
            float lsum = 0.f;
            for (int b = 0; b < n_blocks; ++b) {
                const int pb = b & 1;
                if (b >= 2)                               // block b-2's MMAs must have read this P buffer
                    mbar_wait(b_pfree + 8 * pb, ((pb ? done1 : done0) + (b >> 1) - 1) & 1);
                const int c0 = 32 * b + 16 * h;           // this thread's 16 keys of the block: chunks 2h, 2h+1
                if (warp_live && c0 < nk16) {
                    uint32_t v[16];
                    tmem_ld16(lane_addr + COL_S + c0, v);
                    tmem_ld_wait();
                    uint8_t *pbase = sm + OFF_QH + pb * AT_TILE_Q;
                    auto chunk = [&](int c, auto masked) {    // one 16-byte chunk = 8 keys
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float x0 = fast_ex2(fmaf(__uint_as_float(v[8 * c + 2 * u]), sc, -mxs));
                            float x1 = fast_ex2(fmaf(__uint_as_float(v[8 * c + 2 * u + 1]), sc, -mxs));
                            if (decltype(masked)::value) {            // padding keys of the last 16-key step
                                const int key = c0 + 8 * c + 2 * u;
                                if (key >= nk) x0 = 0.f;
                                if (key + 1 >= nk) x1 = 0.f;
                            }
                            lsum += x0 + x1;
                            const __half2 hh = __floats2half2_rn(x0, x1);
                            const float2 hf = __half22float2(hh);
                            const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                            hi[u] = *reinterpret_cast<const uint32_t *>(&hh);
                            lo[u] = *reinterpret_cast<const uint32_t *>(&ll);
                        }
                        const uint32_t off = sw64(r, 2 * h + c);
                        *reinterpret_cast<uint4 *>(pbase + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(pbase + AT_TILE_Q / 2 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    };
                    if (c0 + 16 <= nk) {                  // the common case: all 16 keys are real
                        chunk(0, std::false_type{});
                        chunk(1, std::false_type{});
                    } else {
                        chunk(0, std::true_type{});
                        chunk(1, std::true_type{});
                    }
                }
                fence_async_smem();                       // also publishes this thread's share of the V rows (first arrive)
                tc_fence_before();
                mbar_arrive(b_pready + 8 * pb);
            }
            // ---- epilogue: O / sum * 2^-sv + residual, LayerNorm over the row's 64 channels; this thread owns
            // channels 32h .. 32h+31 and trades the partial sums with its partner
            float y[32];
            xch[h][1][r] = lsum;
            row_barrier();                                // also: every thread's share of the residual rows has landed
            const float inv = net.vf_inv_scale / (lsum + xch[h ^ 1][1][r]);
            if (warp_live) {
                const int ii = i < L ? i : 0;
                const float4 *x4 = reinterpret_cast<const float4 *>(net.X + ((size_t)tok[ii] * L + ii) * QN_D + 32 * h);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 x = Cfg::XSTAGE ? *reinterpret_cast<const float4 *>(sm + OFF_XR + r * 256 +
                                                                                     (((8 * h + c) ^ (r & 7)) << 4))
                                                 : __ldg(x4 + c);
                    y[4 * c + 0] = x.x; y[4 * c + 1] = x.y; y[4 * c + 2] = x.z; y[4 * c + 3] = x.w;
                }
            }
            float s1 = 0.f;
            if (warp_live) {
                mbar_wait(b_ofull, ph_o);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(lane_addr + COL_O + 32 * h, v);
                tmem_ld_wait();
                tc_fence_before();
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    y[u] = fmaf(__uint_as_float(v[u]), inv, y[u]);
                    s1 += y[u];
                }
            }
            xch[h][2][r] = s1;
            row_barrier();
            const float mu = (s1 + xch[h ^ 1][2][r]) * (1.f / QN_D);
            float s2 = 0.f;
            if (warp_live) {
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    y[u] -= mu;
                    s2 = fmaf(y[u], y[u], s2);
                }
            }
            xch[h][3][r] = s2;
            row_barrier();
            s2 += xch[h ^ 1][3][r];
            const float rstd = rsqrtf(s2 * (1.f / QN_D) + 1e-6f);              // LayerNorm(eps=1e-6), models.py:166
            if (warp_live && i < L) {
                if (Z2out) {
                    __half *zh = Z2out + (size_t)row * 2 * net.K + (size_t)i * QN_D + 32 * h, *zl = zh + net.K;
#pragma unroll
                    for (int c = 0; c < 32; c += 8) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int ch = 32 * h + c + 2 * u;
                            const float z0 = y[c + 2 * u] * rstd * ln_g[ch] + ln_b[ch];
                            const float z1 = y[c + 2 * u + 1] * rstd * ln_g[ch + 1] + ln_b[ch + 1];
                            const __half2 hh = __floats2half2_rn(z0, z1);
                            const float2 hf = __half22float2(hh);
                            const __half2 ll = __floats2half2_rn(z0 - hf.x, z1 - hf.y);
                            hi[u] = *reinterpret_cast<const uint32_t *>(&hh);
                            lo[u] = *reinterpret_cast<const uint32_t *>(&ll);
                        }
                        *reinterpret_cast<uint4 *>(zh + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(zl + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                } else {
                    float *z = Zout + (size_t)row * net.K + (size_t)i * QN_D + 32 * h;
#pragma unroll
                    for (int c = 0; c < 32; ++c) z[c] = y[c] * rstd * ln_g[32 * h + c] + ln_b[32 * h + c];
                }
            }
        }
        // per-item phase flips for the barriers a thread did not wait on itself
        ph_s ^= 1;
        ph_o ^= 1;
        done0 += (n_blocks + 1) >> 1;
        done1 += n_blocks >> 1;
    }
    __syncthreads();
    if (is_mma) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

template <int TR>
int launch_tc(gad_qnet *q, const uint8_t *tok_in, const int32_t *active, const int32_t *n_ptr, long long n_max,
              int n_tiles, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        GAD_CUDA(cudaFuncSetAttribute(encode_tc_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<TR>::SMEM));
        attr_set = true;
    }
    const long long items = n_max * n_tiles;
    const long long grid = items < 2ll * GAD_NUM_SMS ? items : 2ll * GAD_NUM_SMS;
    encode_tc_kernel<TR><<<(unsigned)grid, AtCfg<TR>::THREADS, AtCfg<TR>::SMEM, st>>>(*q, tok_in, active, n_ptr, n_max, n_tiles,
                                                                            q->Z, q->Z2);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

}  // namespace

int attn_tc_supported(const gad_qnet *q) { return q->L <= AT_MAXK && q->L >= 16; }

int launch_encode_tc(gad_qnet *q, const uint8_t *tok_in, const int32_t *active, const int32_t *n_ptr, long long n_max,
                     cudaStream_t st)
{
    const int n_tiles = (q->L + AT_M - 1) / AT_M;                     // even tiles: 186 positions -> 2 x 96 rows
    const int per_tile = (q->L + n_tiles - 1) / n_tiles;
    return per_tile <= 96 ? launch_tc<96>(q, tok_in, active, n_ptr, n_max, n_tiles, st)
                          : launch_tc<128>(q, tok_in, active, n_ptr, n_max, n_tiles, st);
}
