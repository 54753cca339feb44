// Crossover: children are gathered from pairs of survivors by a (parent1, parent2, cut) plan.
// Replaces the per-child deepcopy loop of GA.horizontal_crossover (reference GA.py:281-297).
// The plan itself comes from a CPython-compatible MT19937 replay on the host (mt_host.cu), so the
// children are bit-identical to the reference's for the same `random` state.
//
// HBM-bound gather: 2*R*W bytes per child (R rows read, R rows written), one warp per child row,
// 16 bytes per lane per trip.
#include "gad_common.cuh"

namespace {

// kind 0 (horizontal, GA.py:296): rows [0,cut) from parent1, rows [cut,R) from parent2 - whole rows.
// kind 1 (vertical, not in the reference code - supplement table A5): every row keeps parent1's
// first `cut` cells and continues with parent2's cells from `cut` on.
__global__ void __launch_bounds__(256)
crossover_kernel(const uint8_t *src_boards, const int32_t *src_rowlen, uint8_t *boards, int32_t *rowlen,
                 long long n_parents, const int32_t *__restrict__ plan, long long n_children, int R, int stride, int kind)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long units = n_children * R;
    for (long long u = warp0; u < units; u += nwarps) {
        const long long c = u / R;
        const int r = (int)(u - c * R);
        const int p1 = plan[3 * c + 0], p2 = plan[3 * c + 1], cut = plan[3 * c + 2];
        // parents are rectangular here (selection's fitness pass has just cleaned them)
        const int w1 = src_rowlen[(size_t)p1 * R + r], w2 = src_rowlen[(size_t)p2 * R + r];
        const int wmax = max(w1, w2);
        const int n16 = min(stride >> 4, (wmax + 15) >> 4);
        const long long child = n_parents + c;
        uint4 *d4 = reinterpret_cast<uint4 *>(boards + ((size_t)child * R + r) * stride);
        const uint4 *a4 = reinterpret_cast<const uint4 *>(src_boards + ((size_t)p1 * R + r) * stride);
        const uint4 *b4 = reinterpret_cast<const uint4 *>(src_boards + ((size_t)p2 * R + r) * stride);
        int len;
        if (kind == 0) {
            const bool first = r < cut;
            const uint4 *s4 = first ? a4 : b4;
            len = first ? w1 : w2;
            for (int j = lane; j < n16; j += 32) {
                uint4 v = s4[j];
                // cells at or beyond the source row's length become the gap code (layout invariant)
                uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int base = 16 * j + 4 * k;
                    if (base + 4 > len) {
                        uint32_t x = w[k];
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (base + b >= len) x = (x & ~(0xFFu << (8 * b))) | ((uint32_t)GAD_GAP << (8 * b));
                        w[k] = x;
                    }
                }
                d4[j] = v;
            }
        } else {
            len = min(cut, w1) + max(0, w2 - cut);      // p1[r][:cut] + p2[r][cut:]; the plan keeps cut < min(w1, w2)
            for (int j = lane; j < n16; j += 32) {
                const uint4 va = a4[j], vb = b4[j];
                uint4 v;
                const uint32_t *wa = reinterpret_cast<const uint32_t *>(&va);
                const uint32_t *wb = reinterpret_cast<const uint32_t *>(&vb);
                uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t x = 0u;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int col = 16 * j + 4 * k + b;
                        uint32_t cell;
                        if (col < cut) cell = col < w1 ? (wa[k] >> (8 * b)) & 0xFFu : (uint32_t)GAD_GAP;
                        else cell = col < w2 ? (wb[k] >> (8 * b)) & 0xFFu : (uint32_t)GAD_GAP;
                        x |= cell << (8 * b);
                    }
                    w[k] = x;
                }
                d4[j] = v;
            }
        }
        if (lane == 0) rowlen[(size_t)child * R + r] = len;
    }
}

// Fused exchange + crossover: parents are read straight from the owning rank's population buffer over
// NVLink (peer pointers of a symmetric allocation), children are written locally. A parent is named by
// pos = owner_rank * peer_slots + slot. No staging copy, no all-gather of boards: each rank moves exactly
// the R rows per child it needs.
__global__ void __launch_bounds__(256)
crossover_peer_kernel(const uint8_t *const *__restrict__ peer_boards, const int32_t *const *__restrict__ peer_rowlen,
                      long long peer_slots, uint8_t *boards, int32_t *rowlen, long long dst_first,
                      const int32_t *__restrict__ plan, long long n_children, int R, int stride, int kind)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long units = n_children * R;
    for (long long u = warp0; u < units; u += nwarps) {
        const long long c = u / R;
        const int r = (int)(u - c * R);
        const long long p1 = plan[3 * c + 0], p2 = plan[3 * c + 1];
        const int cut = plan[3 * c + 2];
        const int o1 = (int)(p1 / peer_slots), o2 = (int)(p2 / peer_slots);
        const long long s1 = p1 - o1 * peer_slots, s2 = p2 - o2 * peer_slots;
        const int w1 = peer_rowlen[o1][(size_t)s1 * R + r], w2 = peer_rowlen[o2][(size_t)s2 * R + r];
        const int n16 = min(stride >> 4, (max(w1, w2) + 15) >> 4);
        const long long child = dst_first + c;
        uint4 *d4 = reinterpret_cast<uint4 *>(boards + ((size_t)child * R + r) * stride);
        const uint4 *a4 = reinterpret_cast<const uint4 *>(peer_boards[o1] + ((size_t)s1 * R + r) * stride);
        const uint4 *b4 = reinterpret_cast<const uint4 *>(peer_boards[o2] + ((size_t)s2 * R + r) * stride);
        int len;
        if (kind == 0) {
            const bool first = r < cut;
            const uint4 *s4 = first ? a4 : b4;
            len = first ? w1 : w2;
            for (int j = lane; j < n16; j += 32) {
                uint4 v = s4[j];
                uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int base = 16 * j + 4 * k;
                    if (base + 4 > len) {
                        uint32_t x = w[k];
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (base + b >= len) x = (x & ~(0xFFu << (8 * b))) | ((uint32_t)GAD_GAP << (8 * b));
                        w[k] = x;
                    }
                }
                d4[j] = v;
            }
        } else {
            len = min(cut, w1) + max(0, w2 - cut);
            for (int j = lane; j < n16; j += 32) {
                const uint4 va = a4[j], vb = b4[j];
                uint4 v;
                const uint32_t *wa = reinterpret_cast<const uint32_t *>(&va);
                const uint32_t *wb = reinterpret_cast<const uint32_t *>(&vb);
                uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t x = 0u;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int col = 16 * j + 4 * k + b;
                        uint32_t cell;
                        if (col < cut) cell = col < w1 ? (wa[k] >> (8 * b)) & 0xFFu : (uint32_t)GAD_GAP;
                        else cell = col < w2 ? (wb[k] >> (8 * b)) & 0xFFu : (uint32_t)GAD_GAP;
                        x |= cell << (8 * b);
                    }
                    w[k] = x;
                }
                d4[j] = v;
            }
        }
        if (lane == 0) rowlen[(size_t)child * R + r] = len;
    }
}

}  // namespace

extern "C" int gad_crossover_peer(const uint8_t *const *d_peer_boards, const int32_t *const *d_peer_rowlen,
                                  int64_t peer_slots, uint8_t *d_dst_boards, int32_t *d_dst_rowlen, int64_t dst_first,
                                  const int32_t *d_plan, int64_t n_children, int R, int stride, int kind, void *stream)
{
    GAD_REQUIRE(d_peer_boards && d_peer_rowlen && d_dst_boards && d_dst_rowlen && (d_plan || n_children == 0),
                "gad_crossover_peer: null device pointer");
    GAD_REQUIRE(peer_slots >= 1 && dst_first >= 0 && n_children >= 0 && R >= 1 && stride >= 16 && stride % 16 == 0,
                "gad_crossover_peer: bad shape");
    GAD_REQUIRE(kind == 0 || kind == 1, "gad_crossover_peer: kind must be 0 (horizontal) or 1 (vertical)");
    if (n_children == 0) return GAD_OK;
    const long long units = (long long)n_children * R;
    long long blocks = (units + 7) / 8;
    const long long cap = (long long)GAD_NUM_SMS * 8 * 4;
    if (blocks > cap) blocks = cap;
    crossover_peer_kernel<<<(unsigned)blocks, 256, 0, gad_stream(stream)>>>(d_peer_boards, d_peer_rowlen, peer_slots,
                                                                             d_dst_boards, d_dst_rowlen, dst_first,
                                                                             d_plan, n_children, R, stride, kind);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

extern "C" int gad_crossover(uint8_t *d_boards, int32_t *d_rowlen, int64_t n_parents, const int32_t *d_plan,
                             int64_t n_children, int R, int stride, int kind, void *stream)
{
    GAD_REQUIRE(d_boards && d_rowlen && (d_plan || n_children == 0), "gad_crossover: null device pointer");
    GAD_REQUIRE(n_parents >= 0 && n_children >= 0 && R >= 1 && stride >= 16 && stride % 16 == 0,
                "gad_crossover: bad shape");
    GAD_REQUIRE(kind == 0 || kind == 1, "gad_crossover: kind must be 0 (horizontal) or 1 (vertical)");
    if (n_children == 0) return GAD_OK;
    const long long units = (long long)n_children * R;
    long long blocks = (units + 7) / 8;
    const long long cap = (long long)GAD_NUM_SMS * 8 * 4;
    if (blocks > cap) blocks = cap;
    crossover_kernel<<<(unsigned)blocks, 256, 0, gad_stream(stream)>>>(d_boards, d_rowlen, d_boards, d_rowlen, n_parents,
                                                                        d_plan, n_children, R, stride, kind);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// Sharded form: parents come from a separate (all-gathered) survivor buffer, children are written to
// slots [dst_first, dst_first + n_children) of this rank's population buffer.
extern "C" int gad_crossover_from(const uint8_t *d_src_boards, const int32_t *d_src_rowlen, uint8_t *d_dst_boards,
                                  int32_t *d_dst_rowlen, int64_t dst_first, const int32_t *d_plan, int64_t n_children,
                                  int R, int stride, int kind, void *stream)
{
    GAD_REQUIRE(d_src_boards && d_src_rowlen && d_dst_boards && d_dst_rowlen && (d_plan || n_children == 0),
                "gad_crossover_from: null device pointer");
    GAD_REQUIRE(dst_first >= 0 && n_children >= 0 && R >= 1 && stride >= 16 && stride % 16 == 0,
                "gad_crossover_from: bad shape");
    GAD_REQUIRE(kind == 0 || kind == 1, "gad_crossover_from: kind must be 0 (horizontal) or 1 (vertical)");
    if (n_children == 0) return GAD_OK;
    const long long units = (long long)n_children * R;
    long long blocks = (units + 7) / 8;
    const long long cap = (long long)GAD_NUM_SMS * 8 * 4;
    if (blocks > cap) blocks = cap;
    crossover_kernel<<<(unsigned)blocks, 256, 0, gad_stream(stream)>>>(d_src_boards, d_src_rowlen, d_dst_boards,
                                                                        d_dst_rowlen, dst_first, d_plan, n_children, R,
                                                                        stride, kind);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}
