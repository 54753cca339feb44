// CPython-compatible MT19937 replay on the host.
// The reference draws every random number through CPython's `random` module
// (GA.py:90 randint, GA.py:283-284 choice, GA.py:293 randint), i.e. through
// Random._randbelow_with_getrandbits over genrand_uint32 (Lib/random.py, Modules/_randommodule.c).
// These entry points take the 625-word state of random.getstate() (624 words + position), consume
// exactly the draws the reference would and hand the advanced state back for random.setstate(),
// so "identical seeds" gives identical populations and crossover plans at any population size.
#include "gad_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <stdexcept>
#include <thread>
#include <vector>

namespace {

struct MT {
    uint32_t *mt;      // 624 words, caller's buffer
    int pos;

    void reload()
    {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAG = 0x9908b0dfu;
        int kk;
        for (kk = 0; kk < 624 - 397; ++kk) {
            const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        for (; kk < 623; ++kk) {
            const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        pos = 0;
    }

    uint32_t next32()
    {
        if (pos >= 624) reload();
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }

    // Random.random(): 53 bits from two words, (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53
    double random53()
    {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
    }

    // Random._randbelow_with_getrandbits for 0 < n < 2^32: k = n.bit_length(); draw getrandbits(k) until < n
    uint32_t randbelow(uint32_t n)
    {
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;
        uint32_t r = next32() >> (32 - k);
        while (r >= n) r = next32() >> (32 - k);
        return r;
    }
};

// One generated row (GA.py:86-91): `ng` gaps inserted one at a time at positions at[0..ng) (each drawn in the row as
// it was then). Inserting at position a moves every earlier gap at or right of a one cell to the right, so the final
// gap cells are known after O(ng^2) integer steps and the residues stream into the other cells in one pass.
void emit_row(uint8_t *row, const uint8_t *seq, int len0, int ng, const int32_t *at, int *gaps, int stride)
{
    if (ng <= 8 || len0 <= 128) {                          // few gaps or a short row: shifting the tail is cheaper
        memset(row, GAD_GAP, (size_t)stride);
        memcpy(row, seq, (size_t)len0);
        int len = len0;
        for (int k = 0; k < ng; ++k, ++len) {
            memmove(row + at[k] + 1, row + at[k], (size_t)(len - at[k]));
            row[at[k]] = GAD_GAP;
        }
        return;
    }
    for (int k = 0; k < ng; ++k) {
        const int a = at[k];
        for (int q = 0; q < k; ++q)
            if (gaps[q] >= a) ++gaps[q];
        gaps[k] = a;
    }
    std::sort(gaps, gaps + ng);                            // distinct cells of the final row, left to right
    memset(row, GAD_GAP, (size_t)stride);
    int src = 0, dst = 0;
    for (int k = 0; k < ng; ++k) {                         // residues fill the runs between consecutive gap cells
        const int run = gaps[k] - dst;
        memcpy(row + dst, seq + src, (size_t)run);
        src += run;
        dst = gaps[k] + 1;
    }
    memcpy(row + dst, seq + src, (size_t)(len0 - src));
}

}  // namespace

// GA.py:281-293: while children are missing: p1 = choice(pop), p2 = choice(pop); same object -> retry;
// cut = randint(1, rows - 1). Vertical (h_widths != NULL): cut = randint(1, min(width1, width2) - 1).
static int crossover_plan_impl(uint32_t *mt_state, int64_t n_parents, int n_rows, int64_t n_children,
                               const int32_t *h_widths, bool with_probability, double probability, int copy_cut,
                               int32_t *h_plan);

extern "C" int gad_mt_crossover_plan_host(uint32_t *mt_state, int64_t n_parents, int n_rows, int64_t n_children,
                                          const int32_t *h_widths, int32_t *h_plan)
{
    return crossover_plan_impl(mt_state, n_parents, n_rows, n_children, h_widths, false, 1.0, 0, h_plan);
}

// The same plan with a crossover probability (an extension: the paper's supplement names c_prob, the reference has no
// code for it): after the draws of a child one random.random(); at or above `probability` the child's cut becomes
// copy_cut (a cut past the last row / cell: the child is a copy of parent1).
extern "C" int gad_mt_crossover_plan_prob_host(uint32_t *mt_state, int64_t n_parents, int n_rows, int64_t n_children,
                                               const int32_t *h_widths, double probability, int copy_cut, int32_t *h_plan)
{
    return crossover_plan_impl(mt_state, n_parents, n_rows, n_children, h_widths, true, probability, copy_cut, h_plan);
}

static int crossover_plan_impl(uint32_t *mt_state, int64_t n_parents, int n_rows, int64_t n_children,
                               const int32_t *h_widths, bool with_probability, double probability, int copy_cut,
                               int32_t *h_plan)
{
    GAD_REQUIRE(mt_state && (h_plan || n_children == 0), "gad_mt_crossover_plan_host: null pointer");
    GAD_REQUIRE(n_children >= 0 && n_parents >= 0 && n_parents < (1ll << 31), "gad_mt_crossover_plan_host: bad sizes");
    if (n_children == 0) return GAD_OK;
    // the reference never terminates in these cases (GA.py:285-290)
    GAD_REQUIRE(n_parents >= 2, "crossover needs at least 2 survivors (the reference loops forever with %lld)",
                (long long)n_parents);
    GAD_REQUIRE(h_widths || n_rows >= 2, "horizontal crossover needs at least 2 rows (the reference loops forever)");
    MT g{mt_state, (int)mt_state[624]};
    GAD_REQUIRE(g.pos >= 0 && g.pos <= 624, "gad_mt_crossover_plan_host: corrupt MT19937 position %d", g.pos);
    for (int64_t c = 0; c < n_children;) {
        const uint32_t a = g.randbelow((uint32_t)n_parents);
        const uint32_t b = g.randbelow((uint32_t)n_parents);
        if (a == b) continue;
        int cut;
        if (h_widths) {
            const int m = h_widths[a] < h_widths[b] ? h_widths[a] : h_widths[b];
            GAD_REQUIRE(m >= 2, "vertical crossover needs parents at least 2 columns wide");
            cut = 1 + (int)g.randbelow((uint32_t)(m - 1));
        } else {
            cut = 1 + (int)g.randbelow((uint32_t)(n_rows - 1));
        }
        if (with_probability && g.random53() >= probability) cut = copy_cut;
        h_plan[3 * c + 0] = (int32_t)a;
        h_plan[3 * c + 1] = (int32_t)b;
        h_plan[3 * c + 2] = cut;
        ++c;
    }
    mt_state[624] = (uint32_t)g.pos;
    return GAD_OK;
}

// The draws of the mutation variants (extensions, config.MUTATION_PROBABILITY / MUTATION_TILE = 'random': the paper's
// supplement names them, the reference has no code): for every selected individual in ranking order one
// random.random() when probability < 1 - at or above it the individual is skipped - and then, when h_ntiles is not
// NULL, one random.randrange(h_ntiles[i]) for its tile. h_keep[i] = 1 / 0, h_tile[i] = the tile ordinal (row-major
// in the lattice of utils.py:230-252) or -1.
extern "C" int gad_mt_mutation_draws_host(uint32_t *mt_state, int64_t n, double probability, const int32_t *h_ntiles,
                                          uint8_t *h_keep, int32_t *h_tile)
{
    GAD_REQUIRE(mt_state && (n == 0 || (h_keep && h_tile)), "gad_mt_mutation_draws_host: null pointer");
    GAD_REQUIRE(n >= 0, "gad_mt_mutation_draws_host: bad size");
    MT g{mt_state, (int)mt_state[624]};
    GAD_REQUIRE(g.pos >= 0 && g.pos <= 624, "gad_mt_mutation_draws_host: corrupt MT19937 position %d", g.pos);
    for (int64_t i = 0; i < n; ++i) {
        h_keep[i] = 1;
        h_tile[i] = -1;
        if (probability < 1.0 && g.random53() >= probability) {
            h_keep[i] = 0;
            continue;
        }
        if (h_ntiles) {
            GAD_REQUIRE(h_ntiles[i] >= 1, "min() arg is an empty sequence");          // no tile fits (utils.py:291)
            h_tile[i] = (int32_t)g.randbelow((uint32_t)h_ntiles[i]);
        }
    }
    mt_state[624] = (uint32_t)g.pos;
    return GAD_OK;
}

// GA.py:71-94: n_clone exact copies first, then every row of every other individual gets h_ngaps[r] gaps
// inserted one at a time at randint(0, len_now - 1) (the row grows with every insert; never appended).
// Output in the library's layout: uint8 boards[P][R][stride] (gap code beyond each row), int32 rowlen[P][R].
static int population_host_impl(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen,
                                int seq_stride, const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone,
                                uint8_t *h_boards, int32_t *h_rowlen, int stride)
{
    GAD_REQUIRE(mt_state && h_seqs && h_seqlen && h_ngaps && h_boards && h_rowlen, "gad_mt_population_host: null pointer");
    GAD_REQUIRE(R >= 1 && P >= 0 && n_clone >= 0 && n_clone <= P && stride >= 1, "gad_mt_population_host: bad sizes");
    for (int r = 0; r < R; ++r) {
        GAD_REQUIRE(h_seqlen[r] >= 0 && h_seqlen[r] <= seq_stride, "gad_mt_population_host: bad sequence length");
        const bool grows = n_clone < P;
        GAD_REQUIRE(h_seqlen[r] + (grows ? h_ngaps[r] : 0) <= stride, "gad_mt_population_host: row %d does not fit the stride", r);
        GAD_REQUIRE(!grows || h_seqlen[r] >= 1, "empty sequence: the reference raises ValueError in randint(0, -1)");
    }
    MT g{mt_state, (int)mt_state[624]};
    GAD_REQUIRE(g.pos >= 0 && g.pos <= 624, "gad_mt_population_host: corrupt MT19937 position %d", g.pos);
    // The generator is sequential, the boards are not: the calling thread draws the insert positions of one block of
    // individuals while worker threads build the boards of the previous block from its positions.
    int per_board = 0, max_gaps = 0;
    std::vector<int> first(R);
    for (int r = 0; r < R; ++r) {
        first[r] = per_board;
        per_board += h_ngaps[r];
        max_gaps = h_ngaps[r] > max_gaps ? h_ngaps[r] : max_gaps;
    }
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *e = getenv("GAD_HOST_THREADS")) hw = (unsigned)(atoi(e) > 0 ? atoi(e) : 1);    // cap / pin the workers
    const int n_workers = (int)(hw > 16 ? 16 : (hw < 1 ? 1 : hw));
    int64_t block = per_board > 0 ? (4ll << 20) / per_board : 16384;        // <= 16 MB of positions per buffer
    block = block > 16384 ? 16384 : (block < 256 ? 256 : block);
    if (const char *e = getenv("GAD_HOST_BLOCK")) block = atoi(e) > 0 ? atoi(e) : block;          // test hook: small blocks
    std::atomic<int> worker_failed{0};
    std::vector<int32_t> pos[2];
    pos[0].resize((size_t)block * per_board + 1);
    pos[1].resize((size_t)block * per_board + 1);
    auto build = [&](int64_t p0, int64_t p1, const int32_t *at_all) {         // boards [p0, p1) from their positions
        auto work = [&](int w) {
          try {                                            // nothing may escape a std::thread body (std::terminate)
            std::vector<int> gaps((size_t)max_gaps + 1);
            const int64_t share = (p1 - p0 + n_workers - 1) / n_workers;         // contiguous boards per worker
            const int64_t lo = p0 + share * w, hi = lo + share < p1 ? lo + share : p1;
            for (int64_t p = lo; p < hi; ++p)
                for (int r = 0; r < R; ++r) {
                    uint8_t *row = h_boards + ((size_t)p * R + r) * stride;
                    const uint8_t *seq = h_seqs + (size_t)r * seq_stride;
                    const int len0 = h_seqlen[r], ng = p >= n_clone ? h_ngaps[r] : 0;
                    if (ng) emit_row(row, seq, len0, ng, at_all + (size_t)(p - p0) * per_board + first[r], gaps.data(), stride);
                    else {
                        memset(row, GAD_GAP, (size_t)stride);
                        memcpy(row, seq, (size_t)len0);
                    }
                    h_rowlen[(size_t)p * R + r] = len0 + ng;
                }
          } catch (...) {
            worker_failed.store(1);
          }
        };
        if (n_workers == 1 || p1 - p0 < 64) {
            for (int w = 0; w < n_workers; ++w) work(w);
            return;
        }
        std::vector<std::thread> th;
        int started = 1;
        try {
            for (; started < n_workers; ++started) th.emplace_back(work, started);
        } catch (...) {                                    // thread limit reached: the remaining shares run here
        }
        work(0);
        for (int w = started; w < n_workers; ++w) work(w);
        for (auto &t : th) t.join();
    };
    std::thread builder;
    struct Joiner {                                        // an exception on the calling thread must not leave a joinable thread
        std::thread &t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{builder};
    int cur = 0;
    for (int64_t p0 = 0; p0 < P; p0 += block, cur ^= 1) {
        const int64_t p1 = p0 + block < P ? p0 + block : P;
        int32_t *at = pos[cur].data();
        for (int64_t p = p0; p < p1; ++p) {                                   // sequential: the reference's draw order
            if (p < n_clone) continue;
            int32_t *a = at + (size_t)(p - p0) * per_board;
            for (int r = 0; r < R; ++r)
                for (int k = 0; k < h_ngaps[r]; ++k) *a++ = (int32_t)g.randbelow((uint32_t)(h_seqlen[r] + k));
        }
        if (builder.joinable()) builder.join();                               // previous block done: its buffer is free again
        try {
            builder = std::thread(build, p0, p1, at);
        } catch (...) {
            build(p0, p1, at);
        }
    }
    if (builder.joinable()) builder.join();
    if (worker_failed.load()) {
        gad_set_error("gad_mt_population_host: a board-building worker ran out of memory");
        return GAD_ENOMEM;
    }
    mt_state[624] = (uint32_t)g.pos;
    return GAD_OK;
}

// Sharded form of gad_mt_population_host: the whole random stream of GA.py:71-94 is consumed (so every
// rank stays in step with the reference's generator), but only the boards p with p % own_step == own_first
// are materialised, densely, into h_boards / h_rowlen (ceil((P - own_first) / own_step) boards).
// Gap positions are tracked instead of shifting the row at every insert: inserting at position a moves every
// earlier gap at or right of a one cell to the right - O(gaps^2) per row instead of O(gaps * length).
static int population_shard_host_impl(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen,
                                      int seq_stride, const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone,
                                      int64_t own_first, int64_t own_step, uint8_t *h_boards, int32_t *h_rowlen,
                                      int stride)
{
    GAD_REQUIRE(mt_state && h_seqs && h_seqlen && h_ngaps && h_boards && h_rowlen, "gad_mt_population_shard_host: null pointer");
    GAD_REQUIRE(R >= 1 && P >= 0 && n_clone >= 0 && n_clone <= P && stride >= 1 && own_step >= 1 && own_first >= 0 &&
                    own_first < own_step,
                "gad_mt_population_shard_host: bad sizes");
    int max_gaps = 0;
    for (int r = 0; r < R; ++r) {
        GAD_REQUIRE(h_seqlen[r] >= 0 && h_seqlen[r] <= seq_stride, "gad_mt_population_shard_host: bad sequence length");
        const bool grows = n_clone < P;
        GAD_REQUIRE(h_seqlen[r] + (grows ? h_ngaps[r] : 0) <= stride, "gad_mt_population_shard_host: row %d does not fit the stride", r);
        GAD_REQUIRE(!grows || h_seqlen[r] >= 1, "empty sequence: the reference raises ValueError in randint(0, -1)");
        max_gaps = h_ngaps[r] > max_gaps ? h_ngaps[r] : max_gaps;
    }
    MT g{mt_state, (int)mt_state[624]};
    GAD_REQUIRE(g.pos >= 0 && g.pos <= 624, "gad_mt_population_shard_host: corrupt MT19937 position %d", g.pos);
    std::vector<int> gaps((size_t)max_gaps + 1);
    std::vector<int32_t> at((size_t)max_gaps + 1);
    int64_t out = 0;
    for (int64_t p = 0; p < P; ++p) {
        const bool mine = p % own_step == own_first;
        for (int r = 0; r < R; ++r) {
            const int len0 = h_seqlen[r];
            const int ng = p >= n_clone ? h_ngaps[r] : 0;
            if (!mine) {                                   // keep the generator in step only
                for (int k = 0; k < ng; ++k) g.randbelow((uint32_t)(len0 + k));
                continue;
            }
            uint8_t *row = h_boards + ((size_t)out * R + r) * stride;
            const uint8_t *seq = h_seqs + (size_t)r * seq_stride;
            for (int k = 0; k < ng; ++k) at[k] = (int32_t)g.randbelow((uint32_t)(len0 + k));      // randint(0, len - 1)
            if (ng == 0) {
                memset(row, GAD_GAP, (size_t)stride);
                memcpy(row, seq, (size_t)len0);
            } else {
                emit_row(row, seq, len0, ng, at.data(), gaps.data(), stride);
            }
            h_rowlen[(size_t)out * R + r] = len0 + ng;
        }
        if (mine) ++out;
    }
    mt_state[624] = (uint32_t)g.pos;
    return GAD_OK;
}

// C-ABI entry points: no C++ exception may cross them (std::bad_alloc from the position buffers, std::system_error
// from thread creation) - they become GAD_ENOMEM / GAD_ECUDA through gad_last_error().
template <typename F>
static int guarded(const char *who, F &&f)
{
    try {
        return f();
    } catch (const std::bad_alloc &) {
        gad_set_error("%s: out of host memory", who);
        return GAD_ENOMEM;
    } catch (const std::exception &e) {
        gad_set_error("%s: %s", who, e.what());
        return GAD_ECUDA;
    } catch (...) {
        gad_set_error("%s: unknown C++ exception", who);
        return GAD_ECUDA;
    }
}

extern "C" int gad_mt_population_host(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen,
                                      int seq_stride, const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone,
                                      uint8_t *h_boards, int32_t *h_rowlen, int stride)
{
    return guarded("gad_mt_population_host", [&] {
        return population_host_impl(mt_state, h_seqs, h_seqlen, seq_stride, h_ngaps, R, P, n_clone, h_boards, h_rowlen, stride);
    });
}

extern "C" int gad_mt_population_shard_host(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen,
                                            int seq_stride, const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone,
                                            int64_t own_first, int64_t own_step, uint8_t *h_boards, int32_t *h_rowlen,
                                            int stride)
{
    return guarded("gad_mt_population_shard_host", [&] {
        return population_shard_host_impl(mt_state, h_seqs, h_seqlen, seq_stride, h_ngaps, R, P, n_clone, own_first, own_step,
                                          h_boards, h_rowlen, stride);
    });
}
