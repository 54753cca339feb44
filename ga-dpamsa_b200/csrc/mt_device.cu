// GA.generate_population (reference GA.py:71-94) on the GPU, draw for draw on CPython's MT19937 stream.
//
// The reference inserts max(1, round(len * GAP_RATE)) gaps into every row of every non-clone individual, one
// at a time, at random.randint(0, len_now - 1) = Random._randbelow_with_getrandbits(len_now): k = bit_length(n),
// r = genrand_uint32() >> (32 - k), redrawn while r >= n. At BASELINE c5 that is 2.5 G draws out of ~3.8 G
// generator words - 50 s of host time per rank in round 1. Three sequential-looking parts, made parallel:
//
//  1. the raw generator. One CTA regenerates MT19937 blocks of 624 words in three phases (0..226, 227..453,
//     454..623 - inside a phase every word depends only on values of earlier phases), tempers them and streams
//     them to HBM: one SM, several G words/s, segment by segment (the stream is consumed while it is produced,
//     so memory stays bounded).
//  2. the rejection sampling. Whether word i is accepted depends on which draw it is offered to, i.e. on how many
//     earlier words were accepted - a finite-state scan. The (k, n) sequence of the draws is periodic with the
//     individual (period D' <= sum of the gap counts, usually the gap count of one row), so a chunk of 256 words
//     is a function "residue at entry -> words accepted", tabulated for every residue (W * D' cheap integer steps
//     over all SMs), composed over groups of 256 chunks, chased sequentially over the few hundred groups of a
//     segment, expanded back, and replayed per chunk with the now known entry state to emit the accepted values
//     of the individuals this rank owns.
//  3. the boards. One warp per row turns the insert positions (each relative to the row as it was then) into the
//     final gap cells - lane q holds gap q; inserting at a moves every earlier gap at or right of a one cell to the
//     right - ranks them, and streams residues and gap codes into the row.
// Output is bit-identical to gad_mt_population_host / CPython (tests/test_population_device_gpu.py), including the
// generator state handed back.
#include "gad_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <vector>

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr int CHUNK = 256;              // words per level-0 chunk
constexpr int GROUP = 256;              // chunks per level-1 group

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b)
{
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// state[0..623] + state[624] = position (CPython's layout). Emits n_words tempered outputs (out may be NULL: only
// advance the state) and leaves the state ready for the next call. One CTA of 256 threads. A block of 624 words
// is regenerated in three phases between two copies of the state (read the old copy, write the new one: one
// barrier per phase), every thread tempering and storing the word it has just made.
__global__ void __launch_bounds__(256)
mt_stream_kernel(uint32_t *__restrict__ state, long long n_words, uint32_t *__restrict__ out)
{
    __shared__ uint32_t buf[2][MT_N];
    const int t = threadIdx.x;
    int cur = 0;
    for (int i = t; i < MT_N; i += 256) buf[0][i] = state[i];
    int pos = (int)state[MT_N];
    __syncthreads();
    long long done = 0;
    // what is left of the current block
    if (pos < MT_N && n_words > 0) {
        long long take = MT_N - pos;
        if (take > n_words) take = n_words;
        if (out)
            for (int i = t; i < take; i += 256) out[i] = mt_temper(buf[0][pos + i]);
        pos += (int)take;
        done = take;
    }
    constexpr int H = MT_N - MT_M;                       // 227
    while (done < n_words) {                             // pos == 624 here: regenerate a whole block
        const uint32_t *o = buf[cur];
        uint32_t *n = buf[cur ^ 1];
        const long long left = n_words - done;           // words of this block that are wanted (all 624 but for the last)
        if (t < H) {                                     // words 0..226: old values only
            const uint32_t v = o[t + MT_M] ^ mt_twist(o[t], o[t + 1]);
            n[t] = v;
            if (out && t < left) out[done + t] = mt_temper(v);
        }
        __syncthreads();
        if (t < H) {                                     // words 227..453: new words 0..226
            const int i = t + H;
            const uint32_t v = n[t] ^ mt_twist(o[i], o[i + 1]);
            n[i] = v;
            if (out && i < left) out[done + i] = mt_temper(v);
        }
        __syncthreads();
        if (t < MT_N - 2 * H) {                          // words 454..623: new words 227..396; 623 wraps to the new word 0
            const int i = t + 2 * H;
            const uint32_t v = n[i - H] ^ mt_twist(o[i], i + 1 < MT_N ? o[i + 1] : n[0]);
            n[i] = v;
            if (out && i < left) out[done + i] = mt_temper(v);
        }
        __syncthreads();
        cur ^= 1;
        pos = left < MT_N ? (int)left : MT_N;
        done += pos;
    }
    for (int i = t; i < MT_N; i += 256) state[i] = buf[cur][i];
    if (t == 0) state[MT_N] = (uint32_t)pos;
}

struct Pattern {
    const uint8_t *shift;       // [period] 32 - bit_length(n)
    const uint32_t *limit;      // [period] n
    int period;                 // D'
};

__device__ __forceinline__ bool accepted(const Pattern &pt, int residue, uint32_t word, uint32_t &value)
{
    value = word >> pt.shift[residue];
    return value < pt.limit[residue];
}

// level 0: adv0[chunk][s] = words of the chunk accepted when its first word is offered to a draw of residue s
__global__ void __launch_bounds__(64)
scan_tabulate_kernel(const uint32_t *__restrict__ words, long long n_words, Pattern pt, uint16_t *__restrict__ adv0)
{
    __shared__ uint32_t w[CHUNK];
    extern __shared__ uint8_t pat_raw[];
    uint32_t *lim = reinterpret_cast<uint32_t *>(pat_raw);
    uint8_t *sh = pat_raw + 4 * (size_t)pt.period;
    for (int i = threadIdx.x; i < pt.period; i += blockDim.x) {
        lim[i] = pt.limit[i];
        sh[i] = pt.shift[i];
    }
    const long long chunk = blockIdx.x;
    const long long base = chunk * CHUNK;
    const int n = (int)(n_words - base < CHUNK ? n_words - base : CHUNK);
    for (int i = threadIdx.x; i < CHUNK; i += blockDim.x) w[i] = i < n ? words[base + i] : 0u;
    __syncthreads();
    for (int s0 = threadIdx.x; s0 < pt.period; s0 += blockDim.x) {
        int s = s0, acc = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t v = w[i] >> sh[s];
            if (v < lim[s]) {
                ++acc;
                if (++s == pt.period) s = 0;
            }
        }
        adv0[chunk * pt.period + s0] = (uint16_t)acc;
    }
}

// level 1: adv1[group][s] = words accepted by the group's chunks when it is entered at residue s
__global__ void __launch_bounds__(256)
scan_compose_kernel(const uint16_t *__restrict__ adv0, long long n_chunks, int period, uint32_t *__restrict__ adv1)
{
    const long long group = blockIdx.x;
    const long long c0 = group * GROUP;
    const int nc = (int)(n_chunks - c0 < GROUP ? n_chunks - c0 : GROUP);
    for (int s0 = threadIdx.x; s0 < period; s0 += blockDim.x) {
        int s = s0;
        uint32_t total = 0;
        for (int c = 0; c < nc; ++c) {
            const int a = adv0[(c0 + c) * period + s];
            total += a;
            s = (s + a) % period;
        }
        adv1[group * period + s0] = total;
    }
}

// level 2 (one thread): draw ordinal at the entry of every group; progress[0] = ordinal after the segment
__global__ void scan_chase_kernel(const uint32_t *__restrict__ adv1, long long n_groups, int period,
                                  unsigned long long *__restrict__ progress, unsigned long long *__restrict__ group_entry)
{
    unsigned long long q = progress[0];
    for (long long g = 0; g < n_groups; ++g) {
        group_entry[g] = q;
        q += adv1[g * period + (int)(q % (unsigned long long)period)];
    }
    progress[0] = q;
}

// level 1 expand: draw ordinal at the entry of every chunk
__global__ void __launch_bounds__(256)
scan_expand_kernel(const uint16_t *__restrict__ adv0, long long n_chunks, int period,
                   const unsigned long long *__restrict__ group_entry, unsigned long long *__restrict__ chunk_entry)
{
    const long long group = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c0 = group * GROUP;
    if (c0 >= n_chunks) return;
    const int nc = (int)(n_chunks - c0 < GROUP ? n_chunks - c0 : GROUP);
    unsigned long long q = group_entry[group];
    for (int c = 0; c < nc; ++c) {
        chunk_entry[c0 + c] = q;
        q += adv0[(c0 + c) * period + (int)(q % (unsigned long long)period)];
    }
}

// level 0 replay: with the entry ordinal known, emit the accepted values of the owned individuals.
// Draw ordinal q belongs to individual n_clone + q / D (D draws per individual), draw q % D of it.
// at[] is indexed [owned individual][draw]; last_word receives the stream index of the word that served draw
// n_draws - 1 (the generator stops there).
template <typename T>
__global__ void __launch_bounds__(128)
scan_emit_kernel(const uint32_t *__restrict__ words, long long n_words, long long word_base, Pattern pt,
                 const unsigned long long *__restrict__ chunk_entry, long long n_chunks, unsigned long long n_draws, int D,
                 long long n_clone, long long own_first, long long own_step, T *__restrict__ at,
                 unsigned long long *__restrict__ last_word)
{
    const long long chunk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= n_chunks) return;
    unsigned long long q = chunk_entry[chunk];
    if (q >= n_draws) return;
    const long long base = chunk * CHUNK;
    const int n = (int)(n_words - base < CHUNK ? n_words - base : CHUNK);
    int s = (int)(q % (unsigned long long)pt.period);
    unsigned long long ind = q / (unsigned long long)D;             // non-clone individual number
    int j = (int)(q - ind * (unsigned long long)D);
    for (int i = 0; i < n; ++i) {
        uint32_t v;
        if (!accepted(pt, s, words[base + i], v)) continue;
        const long long p = n_clone + (long long)ind;
        if (p % own_step == own_first) at[(size_t)((p - own_first) / own_step) * D + j] = (T)v;
        if (++q == n_draws) {
            last_word[0] = (unsigned long long)(word_base + base + i);
            return;
        }
        if (++s == pt.period) s = 0;
        if (++j == D) {
            j = 0;
            ++ind;
        }
    }
}

// One warp per (owned individual, row): boards from insert positions. Gap q sits in lane q % 32, register q / 32.
template <typename T, int MAXG>
__global__ void __launch_bounds__(256)
build_rows_kernel(const T *__restrict__ at, const uint8_t *__restrict__ seqs, const int32_t *__restrict__ seqlen, int seq_stride,
                  const int32_t *__restrict__ ngaps, const int32_t *__restrict__ first_draw, int R, int D, long long n_own,
                  long long n_clone, long long own_first, long long own_step, uint8_t *__restrict__ boards,
                  int32_t *__restrict__ rowlen, int stride)
{
    constexpr int NREG = MAXG / 32;
    __shared__ int raw_s[8][MAXG], sorted_s[8][MAXG];
    const long long unit = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (unit >= n_own * R) return;
    const int lane = threadIdx.x & 31, wslot = threadIdx.x >> 5;
    const long long o = unit / R;
    const int r = (int)(unit - o * R);
    const long long p = own_first + o * own_step;                      // individual number in the whole population
    const int len0 = seqlen[r];
    const uint8_t *seq = seqs + (size_t)r * seq_stride;
    uint8_t *row = boards + ((size_t)o * R + r) * stride;
    const int g = p >= n_clone ? ngaps[r] : 0;
    if (lane == 0) rowlen[o * R + r] = len0 + g;
    if (g == 0) {
        for (int c = lane; c < stride; c += 32) row[c] = c < len0 ? seq[c] : (uint8_t)GAD_GAP;
        return;
    }
    const T *pos = at + (size_t)o * D + first_draw[r];
    int mine[NREG], gaps[NREG];                         // mine[e] = insert position of draw e * 32 + lane
#pragma unroll
    for (int e = 0; e < NREG; ++e) {
        mine[e] = e * 32 + lane < g ? (int)pos[e * 32 + lane] : 0;
        gaps[e] = 0x7fffffff;
    }
    for (int k = 0; k < g; ++k) {                       // GA.py:87-91: insert at a -> earlier gaps at or right of a move right
        int src = mine[0];
        if (NREG > 1 && k >= 32) src = mine[NREG - 1];  // k is warp-uniform
        const int a = __shfl_sync(0xFFFFFFFFu, src, k & 31);
#pragma unroll
        for (int e = 0; e < NREG; ++e) {
            const int qidx = e * 32 + lane;
            if (qidx < k && gaps[e] >= a) ++gaps[e];
            if (qidx == k) gaps[e] = a;
        }
    }
    // the final gap cells are distinct: rank them into a sorted list
    int *raw = raw_s[wslot], *srt = sorted_s[wslot];
#pragma unroll
    for (int e = 0; e < NREG; ++e)
        if (e * 32 + lane < g) raw[e * 32 + lane] = gaps[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < NREG; ++e)
        if (e * 32 + lane < g) {
            int rank = 0;
            for (int k = 0; k < g; ++k) rank += raw[k] < gaps[e];
            srt[rank] = gaps[e];
        }
    __syncwarp();
    const int len = len0 + g;
    for (int c = lane; c < stride; c += 32) {
        uint8_t v = GAD_GAP;
        if (c < len) {
            int lo = 0, hi = g;                         // number of gap cells < c, and whether c is one
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (srt[mid] < c) lo = mid + 1;
                else hi = mid;
            }
            if (!(lo < g && srt[lo] == c)) v = seq[c - lo];
        }
        row[c] = v;
    }
}

int bit_length(uint32_t n)
{
    int k = 0;
    for (; n; n >>= 1) ++k;
    return k;
}

}  // namespace

// GA.generate_population on the device. mt_state: CPython's 624 words + position (host, in/out). Boards p with
// p % own_step == own_first are written densely to d_boards / d_rowlen (device; ceil((P - own_first) / own_step)
// boards of `stride` bytes per row). Returns GAD_EINVAL for shapes it does not cover (more than 64 gaps per row,
// a draw pattern longer than 4096): the caller then uses gad_mt_population_host. Synchronises.
extern "C" int gad_population_device(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen, int seq_stride,
                                     const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone, int64_t own_first,
                                     int64_t own_step, uint8_t *d_boards, int32_t *d_rowlen, int stride, void *stream)
{
    GAD_REQUIRE(mt_state && h_seqs && h_seqlen && h_ngaps && d_boards && d_rowlen, "gad_population_device: null pointer");
    GAD_REQUIRE(R >= 1 && P >= 0 && n_clone >= 0 && n_clone <= P && stride >= 1 && own_step >= 1 && own_first >= 0 &&
                    own_first < own_step && seq_stride >= 1,
                "gad_population_device: bad sizes");
    GAD_REQUIRE(mt_state[MT_N] <= (uint32_t)MT_N, "gad_population_device: corrupt MT19937 position");
    int D = 0, max_g = 0;
    std::vector<int32_t> first(R);
    const bool grows = n_clone < P;
    for (int r = 0; r < R; ++r) {
        GAD_REQUIRE(h_seqlen[r] >= 0 && h_seqlen[r] <= seq_stride, "gad_population_device: bad sequence length");
        GAD_REQUIRE(h_ngaps[r] >= 0, "gad_population_device: negative gap count");
        GAD_REQUIRE(h_seqlen[r] + (grows ? h_ngaps[r] : 0) <= stride, "gad_population_device: row %d does not fit the stride", r);
        GAD_REQUIRE(!grows || h_seqlen[r] >= 1 || h_ngaps[r] == 0, "empty sequence: the reference raises ValueError in randint(0, -1)");
        first[r] = D;
        D += h_ngaps[r];
        max_g = h_ngaps[r] > max_g ? h_ngaps[r] : max_g;
    }
    GAD_REQUIRE(max_g <= 64, "gad_population_device: more than 64 gaps per row (use gad_mt_population_host)");
    const long long n_own = own_first < P ? (P - own_first + own_step - 1) / own_step : 0;
    cudaStream_t st = gad_stream(stream);

    // draw pattern of one individual and its minimal period
    std::vector<uint8_t> shift(D > 0 ? D : 1);
    std::vector<uint32_t> limit(D > 0 ? D : 1);
    uint32_t max_n = 1;
    for (int r = 0, j = 0; r < R; ++r)
        for (int k = 0; k < h_ngaps[r]; ++k, ++j) {
            const uint32_t n = (uint32_t)(h_seqlen[r] + k);
            limit[j] = n;
            shift[j] = (uint8_t)(32 - bit_length(n));
            max_n = n > max_n ? n : max_n;
        }
    int period = D;
    for (int cand = 1; cand < D; ++cand) {
        if (D % cand) continue;
        bool ok = true;
        for (int j = cand; j < D && ok; ++j) ok = limit[j] == limit[j - cand];
        if (ok) {
            period = cand;
            break;
        }
    }
    GAD_REQUIRE(period <= 4096, "gad_population_device: draw pattern of period %d (use gad_mt_population_host)", period);
    const bool wide = max_n > 65535u;

    // device copies of the small inputs
    uint8_t *d_seqs = nullptr, *d_shift = nullptr;
    int32_t *d_seqlen = nullptr, *d_ngaps = nullptr, *d_first = nullptr;
    uint32_t *d_limit = nullptr, *d_state = nullptr, *d_ckpt = nullptr, *d_words = nullptr, *d_adv1 = nullptr;
    uint16_t *d_adv0 = nullptr;
    unsigned long long *d_progress = nullptr, *d_gentry = nullptr, *d_centry = nullptr;
    void *d_at = nullptr;
    int rc = GAD_OK;
    bool seg_used[2] = {false, false};
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == GAD_OK) {
            gad_set_error("gad_population_device: %s -> %s", what, cudaGetErrorString(e));
            rc = GAD_ECUDA;
        }
        return e != cudaSuccess;
    };
    const unsigned long long n_draws = (unsigned long long)(P - n_clone) * (unsigned long long)D;
    // expected words per draw from the acceptance probabilities; a segment holds at most SEG words
    long long SEG = 16ll << 20;
    if (n_draws * 2ull + 4096ull < (unsigned long long)SEG) SEG = (long long)(n_draws * 2ull + 4096ull);   // small populations
    if (const char *e = getenv("GAD_POPGEN_SEGMENT")) SEG = atoll(e) > 4096 ? atoll(e) : SEG;       // test hook: tiny segments
    SEG = (SEG + CHUNK - 1) / CHUNK * CHUNK;
    const long long max_chunks = SEG / CHUNK, max_groups = (max_chunks + GROUP - 1) / GROUP;
    const size_t at_elems = (size_t)(n_own > 0 ? n_own : 1) * (size_t)(D > 0 ? D : 1);
#define PD_ALLOC(ptr, bytes) fail(cudaMalloc((void **)&(ptr), (bytes) ? (bytes) : 1), #ptr)
    PD_ALLOC(d_seqs, (size_t)R * seq_stride);
    PD_ALLOC(d_seqlen, (size_t)R * 4);
    PD_ALLOC(d_ngaps, (size_t)R * 4);
    PD_ALLOC(d_first, (size_t)R * 4);
    PD_ALLOC(d_shift, (size_t)(D > 0 ? D : 1));
    PD_ALLOC(d_limit, (size_t)(D > 0 ? D : 1) * 4);
    PD_ALLOC(d_state, (size_t)(MT_N + 1) * 4);
    PD_ALLOC(d_ckpt, (size_t)(MT_N + 1) * 4 * 2);
    PD_ALLOC(d_progress, 16);
    if (n_draws > 0) {
        PD_ALLOC(d_words, (size_t)SEG * 4 * 2);
        PD_ALLOC(d_adv0, (size_t)max_chunks * period * 2);
        PD_ALLOC(d_adv1, (size_t)max_groups * period * 4);
        PD_ALLOC(d_gentry, (size_t)max_groups * 8);
        PD_ALLOC(d_centry, (size_t)max_chunks * 8);
        PD_ALLOC(d_at, at_elems * (wide ? 4 : 2));
    }
#undef PD_ALLOC
    auto cleanup = [&]() {
        cudaFree(d_seqs); cudaFree(d_seqlen); cudaFree(d_ngaps); cudaFree(d_first); cudaFree(d_shift); cudaFree(d_limit);
        cudaFree(d_state); cudaFree(d_ckpt); cudaFree(d_progress); cudaFree(d_words); cudaFree(d_adv0); cudaFree(d_adv1);
        cudaFree(d_gentry); cudaFree(d_centry); cudaFree(d_at);
    };
    if (rc) {
        cleanup();
        return rc;
    }
    fail(cudaMemcpyAsync(d_seqs, h_seqs, (size_t)R * seq_stride, cudaMemcpyHostToDevice, st), "copy");
    fail(cudaMemcpyAsync(d_seqlen, h_seqlen, (size_t)R * 4, cudaMemcpyHostToDevice, st), "copy");
    fail(cudaMemcpyAsync(d_ngaps, h_ngaps, (size_t)R * 4, cudaMemcpyHostToDevice, st), "copy");
    fail(cudaMemcpyAsync(d_first, first.data(), (size_t)R * 4, cudaMemcpyHostToDevice, st), "copy");
    fail(cudaMemcpyAsync(d_state, mt_state, (size_t)(MT_N + 1) * 4, cudaMemcpyHostToDevice, st), "copy");
    if (D > 0) {
        fail(cudaMemcpyAsync(d_shift, shift.data(), (size_t)D, cudaMemcpyHostToDevice, st), "copy");
        fail(cudaMemcpyAsync(d_limit, limit.data(), (size_t)D * 4, cudaMemcpyHostToDevice, st), "copy");
    }
    fail(cudaMemsetAsync(d_progress, 0, 16, st), "memset");

    Pattern pt{d_shift, d_limit, period};
    long long word_base = 0;                        // stream index of the current segment's first word
    unsigned long long h_progress[2] = {0, 0};      // [0] draws served so far, [1] stream index of the last word used
    const size_t pat_smem = (size_t)period * 5 + 16;
    // The generator (one SM) runs one segment ahead of the scan (all other SMs) on its own stream.
    cudaStream_t gen = nullptr;
    cudaEvent_t made[2] = {nullptr, nullptr}, scanned[2] = {nullptr, nullptr}, ready = nullptr;
    if (n_draws > 0) {
        fail(cudaStreamCreateWithFlags(&gen, cudaStreamNonBlocking), "stream");
        for (int i = 0; i < 2; ++i) {
            fail(cudaEventCreateWithFlags(&made[i], cudaEventDisableTiming), "event");
            fail(cudaEventCreateWithFlags(&scanned[i], cudaEventDisableTiming), "event");
        }
        fail(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming), "event");
        fail(cudaEventRecord(ready, st), "event");                    // inputs uploaded
        fail(cudaStreamWaitEvent(gen, ready, 0), "event");
    }
    auto launch_generator = [&](int slot) {          // next segment into words[slot], state checkpointed before it
        if (scanned[slot] && seg_used[slot]) fail(cudaStreamWaitEvent(gen, scanned[slot], 0), "event");
        fail(cudaMemcpyAsync(d_ckpt + (size_t)slot * (MT_N + 1), d_state, (size_t)(MT_N + 1) * 4, cudaMemcpyDeviceToDevice, gen),
             "checkpoint");
        mt_stream_kernel<<<1, 256, 0, gen>>>(d_state, SEG, d_words + (size_t)slot * SEG);
        fail(cudaEventRecord(made[slot], gen), "event");
        seg_used[slot] = true;
    };
    int slot = 0;
    if (!rc && n_draws > 0) launch_generator(0);
    while (!rc && h_progress[0] < n_draws) {
        // acceptance is at least 1/2, so the remaining draws need at most about two words each: only then is
        // another segment worth generating ahead
        const unsigned long long remaining = n_draws - h_progress[0];
        const bool more_likely = remaining > (unsigned long long)SEG / 2;
        if (more_likely) launch_generator(slot ^ 1);
        const long long n_words = SEG, n_chunks = n_words / CHUNK, n_groups = (n_chunks + GROUP - 1) / GROUP;
        const uint32_t *words = d_words + (size_t)slot * SEG;
        fail(cudaStreamWaitEvent(st, made[slot], 0), "event");
        scan_tabulate_kernel<<<(unsigned)n_chunks, 64, pat_smem, st>>>(words, n_words, pt, d_adv0);
        scan_compose_kernel<<<(unsigned)n_groups, 64, 0, st>>>(d_adv0, n_chunks, period, d_adv1);
        scan_chase_kernel<<<1, 1, 0, st>>>(d_adv1, n_groups, period, d_progress, d_gentry);
        scan_expand_kernel<<<(unsigned)((n_groups + 127) / 128), 128, 0, st>>>(d_adv0, n_chunks, period, d_gentry, d_centry);
        if (wide)
            scan_emit_kernel<uint32_t><<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(
                words, n_words, word_base, pt, d_centry, n_chunks, n_draws, D, n_clone, own_first, own_step, (uint32_t *)d_at,
                d_progress + 1);
        else
            scan_emit_kernel<uint16_t><<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(
                words, n_words, word_base, pt, d_centry, n_chunks, n_draws, D, n_clone, own_first, own_step, (uint16_t *)d_at,
                d_progress + 1);
        fail(cudaGetLastError(), "kernel launch");
        fail(cudaEventRecord(scanned[slot], st), "event");
        fail(cudaMemcpyAsync(h_progress, d_progress, 16, cudaMemcpyDeviceToHost, st), "progress");
        fail(cudaStreamSynchronize(st), "sync");
        if (h_progress[0] < n_draws) {
            word_base += n_words;
            if (!more_likely) launch_generator(slot ^ 1);          // the estimate was short: generate now
            slot ^= 1;
        }
    }
    if (!rc && n_draws > 0) {
        // the generator stops right after the word that served the last draw: rewind to the checkpoint taken before the
        // segment that word lies in and advance by the words actually used of it
        fail(cudaStreamSynchronize(gen), "sync");
        const long long used = (long long)h_progress[1] + 1 - word_base;
        fail(cudaMemcpyAsync(d_state, d_ckpt + (size_t)slot * (MT_N + 1), (size_t)(MT_N + 1) * 4, cudaMemcpyDeviceToDevice, st),
             "rewind");
        mt_stream_kernel<<<1, 256, 0, st>>>(d_state, used, nullptr);
        fail(cudaGetLastError(), "kernel launch");
    }
    if (gen) {
        cudaStreamSynchronize(gen);
        for (int i = 0; i < 2; ++i) {
            if (made[i]) cudaEventDestroy(made[i]);
            if (scanned[i]) cudaEventDestroy(scanned[i]);
        }
        if (ready) cudaEventDestroy(ready);
        cudaStreamDestroy(gen);
    }
    if (!rc && n_own > 0) {
        const long long units = n_own * R;
        const unsigned blocks = (unsigned)((units + 7) / 8);
#define PD_BUILD(T, MAXG)                                                                                               \
    build_rows_kernel<T, MAXG><<<blocks, 256, 0, st>>>((const T *)d_at, d_seqs, d_seqlen, seq_stride, d_ngaps, d_first, R, \
                                                       D > 0 ? D : 1, n_own, n_clone, own_first, own_step, d_boards,      \
                                                       d_rowlen, stride)
        if (wide) {
            if (max_g <= 32) PD_BUILD(uint32_t, 32);
            else PD_BUILD(uint32_t, 64);
        } else {
            if (max_g <= 32) PD_BUILD(uint16_t, 32);
            else PD_BUILD(uint16_t, 64);
        }
#undef PD_BUILD
        fail(cudaGetLastError(), "kernel launch");
    }
    if (!rc) {
        fail(cudaMemcpyAsync(mt_state, d_state, (size_t)(MT_N + 1) * 4, cudaMemcpyDeviceToHost, st), "state");
        fail(cudaStreamSynchronize(st), "sync");
    }
    cleanup();
    return rc;
}
