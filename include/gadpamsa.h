/*
 * gadpamsa.h - C ABI of libgadpamsa: the GA-DPAMSA generation loop on B200 (sm_100a).
 *
 * The reference (FLaTNNBio/GA-DPAMSA) is pure Python and has NO FFI of its own;
 * its drop-in boundary is the module surface of GA.py, utils.py and the DPAMSA package.
 * Every entry point below therefore names the reference function(s) whose
 * per-individual Python loop it replaces for a whole device-resident
 * population; the Python facade in ga-dpamsa_b200/ keeps the reference
 * signatures and calls these through ctypes.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.
 *  - every function returns 0 on success, a negative GAD_E* code otherwise;
 *    gad_last_error() gives the thread-local message of the last failure.
 *  - pointers named d_* are DEVICE pointers (cudaMalloc / torch tensor
 *    data_ptr()), h_* are HOST pointers; `stream` is a cudaStream_t passed as
 *    void* (NULL = legacy default stream). Calls are asynchronous on `stream`
 *    unless the name ends in _host or the comment says "synchronises".
 *  - population layout: uint8 boards[P][R][stride] (row-major, stride % 16 == 0)
 *    + int32 rowlen[P][R]. Cell codes follow GA.py:35: A=1 T=2 C=3 G=4 gap=5.
 *    Invariant kept by every producer in this library: with
 *    w = max_r rowlen[p][r], cells [rowlen[p][r], roundup4(w)) of row r hold 5,
 *    i.e. the right-padding GA.py:158-162 performs is already materialised.
 */
#ifndef GADPAMSA_H
#define GADPAMSA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAD_OK 0
#define GAD_EINVAL (-1)   /* bad argument                          -> ValueError   */
#define GAD_ECUDA (-2)    /* CUDA runtime / launch failure          -> RuntimeError */
#define GAD_ECAPACITY (-3)/* a row would outgrow `stride`           -> facade regrows and retries */
#define GAD_ENOMEM (-4)

#define GAD_MODE_SP 0     /* 'sp' */
#define GAD_MODE_CS 1     /* 'cs' */
#define GAD_MODE_MO 2     /* 'mo' */

int gad_version(void);
const char *gad_last_error(void);

/* ---- device plumbing for hosts without torch (C / cgo / JNI callers) ---- */
int gad_device_count(int *count);
int gad_set_device(int device);
int gad_malloc(void **d_ptr, size_t bytes);
int gad_free(void *d_ptr);
int gad_malloc_host(void **h_ptr, size_t bytes);           /* pinned */
int gad_free_host(void *h_ptr);
int gad_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
int gad_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
int gad_memset(void *d_dst, int value, size_t bytes, void *stream);
int gad_stream_sync(void *stream);
/* Allow kernels on the current device to read `peer_device`'s memory (measurement scripts; see api.cu). */
int gad_enable_peer_access(int peer_device);                         /* synchronises */

/* ---- a4-a7: one fitness pass over the whole population ------------------
 * Replaces GA.calculate_fitness_score's per-board loop (GA.py:156-178):
 * right-pad (GA.py:158-162), utils.clean_unnecessary_gaps (utils.py:390-428),
 * utils.get_sum_of_pairs (utils.py:27-76) and utils.get_column_score
 * (utils.py:79-128) over the full board, for all P boards in one launch.
 * Boards are compacted in place and rowlen[p][*] is set to the new width.
 * Outputs (device, int32[P]): sp = sum-of-pairs, em = exact-match columns,
 * al = alignment length after cleaning; the facade forms CS = em / al as a
 * Python float so it is bit-identical to utils.py:128.
 * d_max_al (optional, device int64[1], zero-initialised once by the caller) receives max(al) in its low
 * GAD_MAX_AL_BITS bits: the bits above hold a per-launch tag that grows with every call, so the pass's
 * atomicMax overrides the previous pass without a memset (read it as value & GAD_MAX_AL_MASK).
 * width_hint (0 = unknown) is the expected board width: it selects how many lanes share a board and how
 * many bytes of a row are staged; any width up to `stride` stays correct. Requires R <= 255. */
#define GAD_MAX_AL_BITS 20
#define GAD_MAX_AL_MASK ((1ll << GAD_MAX_AL_BITS) - 1)
int gad_fitness(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride, int width_hint,
                int match, int mismatch, int gap,
                int32_t *d_sp, int32_t *d_em, int32_t *d_al, int64_t *d_max_al, void *stream);

/* Measurement hook (bench.py): a launch that only reads the live cells (`width` bytes of every row, as
 * 16-byte units) of P boards and folds them into d_sink16 (device, 16 bytes). What a single pass over
 * these bytes costs at this launch size, whatever it computes. Asynchronous on `stream`. */
int gad_probe_read(const uint8_t *d_boards, int64_t P, int R, int stride, int width, void *d_sink16, void *stream);

/* Same pass through HOST buffers (pinned or pageable): H2D of boards+rowlen,
 * the kernel, D2H of sp/em/al and of the cleaned boards+rowlen when
 * `copy_back` != 0. Device staging is cached per calling thread. Synchronises. */
int gad_fitness_host(uint8_t *h_boards, int32_t *h_rowlen, int64_t P, int R, int stride,
                     int match, int mismatch, int gap,
                     int32_t *h_sp, int32_t *h_em, int32_t *h_al, int copy_back);

/* Windowed scoring of ONE rectangular board held in host memory - the direct
 * replacement for calling utils.get_sum_of_pairs(chromosome, fr, tr, fc, tc)
 * and utils.get_column_score(...) (numerator) from Python. Synchronises. */
int gad_window_score_host(const uint8_t *h_board, int R, int stride,
                          int from_row, int to_row, int from_col, int to_col,
                          int match, int mismatch, int gap, int64_t *sp, int32_t *em);

/* ---- a9-a11, a8: ranking, selection, survivor compaction, hall of fame ----
 * All take a caller-provided device workspace of gad_workspace_bytes(P) bytes (valid for every
 * population size <= P). Score inputs are the int32 sp/em/al arrays gad_fitness writes. */
int gad_workspace_bytes(int64_t P, size_t *bytes);

/* The fp64 sort key of utils.get_index_of_the_best_fitted_individuals (utils.py:332-351):
 * 'sp' -> sp, 'cs' -> em/al, 'mo' -> (sp-min)/(max-min) + (cs-min)/(max-min) (0.5 when flat).
 * When d_hof (the int32[8] hall-of-fame record, see gad_hof_update) is not NULL its score is
 * appended as element P, exactly like GA.py:127-132 - d_key then holds P+1 doubles. */
int gad_rank_keys(const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, int64_t P, int mode,
                  const int32_t *d_hof, double *d_key, void *d_ws, size_t ws_bytes, void *stream);

/* Stable descending order of n keys (ties keep ascending index) = python sorted(..., reverse=True):
 * d_order[:k] is the list utils.get_index_of_the_best_fitted_individuals(scores, k) returns. */
int gad_rank_order(const double *d_key, int64_t n, int32_t *d_order, void *d_ws, size_t ws_bytes, void *stream);

/* GA.selection's survivor set (GA.py:232-256; Pareto front GA.py:195-211 in O(P log P)):
 * d_keep[p] = 1 for survivors, d_dst[p] = their slot after compaction in original order
 * (GA.py:259), *d_count = number of survivors (may exceed top_n in 'mo' mode, like the reference). */
int gad_select(const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, int64_t P, int mode, int64_t top_n,
               uint8_t *d_keep, int32_t *d_dst, int32_t *d_count, void *d_ws, size_t ws_bytes, void *stream);

/* Move the survivors (boards, row lengths, scores) to their slots in a second population buffer. */
int gad_compact(const uint8_t *d_src_boards, const int32_t *d_src_rowlen, const int32_t *d_src_sp,
                const int32_t *d_src_em, const int32_t *d_src_al, const uint8_t *d_keep, const int32_t *d_dst,
                int64_t P, int R, int stride, uint8_t *d_dst_boards, int32_t *d_dst_rowlen, int32_t *d_dst_sp,
                int32_t *d_dst_em, int32_t *d_dst_al, void *stream);

/* GA.update_hall_of_fame (GA.py:110-136) without the population deep copy, one launch. d_hof is int32[8]:
 * {valid, index, sp, em, al, 0, 0, 0} - words 5..7 must be zero before the first call and belong to the library
 * (grid barrier and ticket of the decision kernel; zero again after every call); d_hof_board is uint8[R][stride].
 * hof_valid = 0 on the first call (GA.py:110-115). On ties a population member replaces the hall of fame
 * (lowest index). */
int gad_hof_update(const uint8_t *d_boards, const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al,
                   int64_t P, int R, int stride, int mode, uint8_t *d_hof_board, int32_t *d_hof, int hof_valid,
                   void *d_ws, size_t ws_bytes, void *stream);

/* GA.calculate_fitness_score as a whole (GA.py:138-181): gad_fitness followed by gad_hof_update, same arguments,
 * same results. For populations the TMA-staged kernel takes (boards up to 256 columns, 2 048+ individuals, 16-byte
 * aligned) it is ONE launch: the CTAs that scored the boards take the hall-of-fame decision in their tail (per-CTA
 * score ranges, one grid barrier, keys of the CTA's own boards, last CTA commits); otherwise the two launches.
 * GAD_FITNESS_FUSED_HOF=0 forces the two launches (A/B). Workspace: gad_workspace_bytes(P).
 * The one-launch form meets at a grid barrier, so all of its CTAs (one per SM) must become resident: do not run two
 * such launches concurrently on different streams of one device (each would hold some SMs and wait for the rest;
 * the barrier traps after 2 s instead of hanging). Launches on one stream, CUDA graphs and programmatic dependent
 * launch are fine. */
int gad_fitness_pass(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride, int width_hint,
                     int match, int mismatch, int gap, int32_t *d_sp, int32_t *d_em, int32_t *d_al,
                     int64_t *d_max_al, int mode, uint8_t *d_hof_board, int32_t *d_hof, int hof_valid,
                     void *d_ws, size_t ws_bytes, void *stream);

/* Measurement hook (scripts/pass_timing.py): d_buf16 = device uint64[16] where the decision tail of the following
 * gad_fitness_pass launches leaves %globaltimer stamps of its steps (NULL switches it off). Not part of the path. */
int gad_debug_stamps(void *d_buf16);

/* The decision of gad_hof_update alone: *d_winner = index of the best of population + hall of fame
 * (P = the hall of fame keeps its place). Used when the boards are sharded over several GPUs and
 * the winner's board has to be fetched from its owner. */
int gad_hof_argbest(const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, int64_t P, int mode,
                    const int32_t *d_hof, int32_t *d_winner, void *d_ws, size_t ws_bytes, void *stream);

/* ---- a12/a13: crossover --------------------------------------------------
 * Children n_parents + c (c < n_children) of the same buffer are built from d_plan[c] =
 * {parent1, parent2, cut}. kind 0 = horizontal (GA.py:296: rows [0,cut) of parent1, the rest of
 * parent2); kind 1 = vertical (extension, no reference code: every row is parent1[:cut] + parent2[cut:]). */
int gad_crossover(uint8_t *d_boards, int32_t *d_rowlen, int64_t n_parents, const int32_t *d_plan,
                  int64_t n_children, int R, int stride, int kind, void *stream);

/* Sharded form of gad_crossover: parents are rows of a separate (all-gathered) survivor buffer,
 * children go to slots [dst_first, dst_first + n_children) of this rank's population buffer. */
int gad_crossover_from(const uint8_t *d_src_boards, const int32_t *d_src_rowlen, uint8_t *d_dst_boards,
                       int32_t *d_dst_rowlen, int64_t dst_first, const int32_t *d_plan, int64_t n_children, int R,
                       int stride, int kind, void *stream);

/* Fused exchange + crossover for a sharded population: d_peer_boards[rank] / d_peer_rowlen[rank] are the
 * ranks' population buffers mapped into this process (peer pointers of a symmetric allocation, read over
 * NVLink); a parent is named by pos = owner_rank * peer_slots + slot in d_plan[c] = {pos1, pos2, cut}.
 * Children go to slots [dst_first, dst_first + n_children) of the local buffer. The caller orders the launch
 * after every rank has finished writing its survivors (the preceding score all-gather does). */
int gad_crossover_peer(const uint8_t *const *d_peer_boards, const int32_t *const *d_peer_rowlen, int64_t peer_slots,
                       uint8_t *d_dst_boards, int32_t *d_dst_rowlen, int64_t dst_first, const int32_t *d_plan,
                       int64_t n_children, int R, int stride, int kind, void *stream);

/* The fitness pass of a SHARDED population without a collective call (sharded.py; replaces the score all-gather +
 * hall-of-fame broadcast around GA.calculate_fitness_score / GA.update_hall_of_fame, GA.py:96-181).
 * Every rank owns a PASS BLOCK in symmetric memory, zeroed once: int32 flags[64] at flags_off, the score tables
 * int32 [2][5][cap1] = {sp, em, al, owner rank, slot on the owner} indexed by LOGICAL index at table_off, and the
 * hall-of-fame board uint8 [R][hof_pitch] at hof_off. d_peer_base = device array of `world` int64: the address of
 * every rank's pass block as mapped into this process.
 * Staging area (stage_off != 0): int32 [2][world][4 + 4 * cap_s] at stage_off - per pass parity and SENDER a header
 * {count, -, -, -} and the contiguous arrays sp | em | al | logical index (cap_s = capacity per sender, a multiple of
 * 4, the same on every rank).
 * gad_pass_scatter: with a staging area, copies the scores and logical indices (d_gidx) of the n_local local
 * individuals as contiguous 16-byte stores into this rank's block of EVERY rank's staging area over NVLink; without
 * one (stage_off = 0) it stores every tuple at its logical index in every rank's table directly (4-byte stores
 * strided by the placement: several times slower from 4 ranks on). Then it raises this rank's arrival flag on
 * every rank.
 * gad_pass_decide: waits for all arrival flags, decides like gad_hof_update on the n gathered scores (identically
 * on every rank) - reading the staged blocks and filing every tuple under its logical index in this pass's table
 * on the way - writes the record d_hof; the rank that owns the winner stores its board into every rank's pass
 * block, the others wait for it. Nothing is read back by the host; a flag that does not arrive within 20 s traps. */
int gad_pass_scatter(const void *d_peer_base, int64_t flags_off, int64_t table_off, int64_t cap1, int64_t stage_off,
                     int64_t cap_s, const int32_t *d_sp, const int32_t *d_em, const int32_t *d_al, const int64_t *d_gidx,
                     int64_t n_local, int rank, int world, void *stream);
int gad_pass_decide(const void *d_peer_base, int64_t flags_off, int64_t table_off, int64_t hof_off, int64_t cap1,
                    int64_t stage_off, int64_t cap_s, int hof_pitch, int64_t n, int mode, int hof_valid,
                    const uint8_t *d_boards, int R, int stride, int32_t *d_hof, int rank, int world, void *d_ws,
                    size_t ws_bytes, void *stream);

/* CPython-compatible MT19937 replay (random.getstate()[1] = 624 words + position, in/out):
 * the (parent1, parent2, cut) triples of GA.py:281-293 for n_children children. h_widths (one
 * per parent) selects the vertical cut rule randint(1, min(w1,w2)-1); NULL = horizontal. */
int gad_mt_crossover_plan_host(uint32_t *mt_state, int64_t n_parents, int n_rows, int64_t n_children,
                               const int32_t *h_widths, int32_t *h_plan);

/* The same plan with a crossover probability - an extension (config.CROSSOVER_PROBABILITY: the paper's supplement
 * names c_prob, the reference checkout has no code for it): after the reference's draws for a child one
 * random.random() (53 bits from two words); at or above `probability` the cut becomes copy_cut, a cut past the last
 * row (horizontal) or cell (vertical), i.e. the child is a copy of parent1. */
int gad_mt_crossover_plan_prob_host(uint32_t *mt_state, int64_t n_parents, int n_rows, int64_t n_children,
                                    const int32_t *h_widths, double probability, int copy_cut, int32_t *h_plan);

/* The draws of the mutation variants - extensions (config.MUTATION_PROBABILITY, config.MUTATION_TILE = 'random'; the
 * paper's supplement names them, the reference checkout has no code): for each of the n selected individuals in
 * ranking order one random.random() when probability < 1 (at or above it h_keep[i] = 0, nothing else is drawn for
 * it), then - h_ntiles not NULL - h_tile[i] = random.randrange(h_ntiles[i]), the ordinal of its tile in the row-major
 * lattice of utils.get_all_different_sub_range (utils.py:230-252); h_tile[i] = -1 otherwise. */
int gad_mt_mutation_draws_host(uint32_t *mt_state, int64_t n, double probability, const int32_t *h_ntiles,
                               uint8_t *h_keep, int32_t *h_tile);

/* GA.generate_population (GA.py:71-94) with the same replay: n_clone copies, then h_ngaps[r]
 * gaps per row inserted one at a time at randint(0, len-1). Host buffers in the library layout. */
int gad_mt_population_host(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen, int seq_stride,
                           const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone, uint8_t *h_boards,
                           int32_t *h_rowlen, int stride);

/* Sharded form: the whole stream is consumed, only boards p with p % own_step == own_first are
 * written (densely) - each rank of a sharded run generates just its slice, in step with the reference. */
int gad_mt_population_shard_host(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen, int seq_stride,
                                 const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone, int64_t own_first,
                                 int64_t own_step, uint8_t *h_boards, int32_t *h_rowlen, int stride);

/* The same population built ON THE DEVICE, draw for draw on the same MT19937 stream (csrc/mt_device.cu: one CTA
 * streams the generator, a chunked finite-state scan resolves the rejection sampling, one warp per row places the
 * gaps). mt_state is the host copy of CPython's state (in/out); boards p with p % own_step == own_first are
 * written densely into d_boards / d_rowlen (device). Returns GAD_EINVAL for shapes it does not cover (more than 64
 * gaps per row, a draw pattern of period above 4096): use gad_mt_population_host then. Synchronises. */
int gad_population_device(uint32_t *mt_state, const uint8_t *h_seqs, const int32_t *h_seqlen, int seq_stride,
                          const int32_t *h_ngaps, int R, int64_t P, int64_t n_clone, int64_t own_first, int64_t own_step,
                          uint8_t *d_boards, int32_t *d_rowlen, int stride, void *stream);

/* ---- a15: worst-fitted sub-board of every mutating individual -------------
 * utils.calculate_worst_fitted_sub_board (utils.py:276-313) over the implicit (rw, cw) tile lattice
 * of utils.get_all_different_sub_range (utils.py:230-252). d_idx[m] (NULL = identity) names the
 * individual; d_tile[m] = {from_row, from_col} of its first-minimum tile ({-1,-1}: no tile fits). */
int gad_worst_tiles(const uint8_t *d_boards, const int32_t *d_rowlen, int R, int stride, const int32_t *d_idx,
                    int64_t M, int rw, int cw, int mode, int match, int mismatch, int gap, int32_t *d_tile,
                    void *stream);

/* ---- a14, a16-a18: batched Q-network and environment episodes -------------
 * gad_qnet holds the reference state_dict (DPAMSA/dqn.py:187-213; 16 fp32 tensors, order in
 * csrc/qnet.cu) for an agent trained on rows x cols windows. Eval mode (no dropout). */
typedef struct gad_qnet gad_qnet;
int gad_qnet_create(gad_qnet **out, int rows, int cols, int d_k, int d_v, int h1, int h2,
                    const float *const *h_weights);
int gad_qnet_destroy(gad_qnet *q);
int gad_qnet_dims(const gad_qnet *q, int *L, int *A);
/* Net.forward (dqn.py:69-79) + argmax (dqn.py:153-154) for B states of L uint8 tokens each. */
int gad_qnet_forward(gad_qnet *q, const uint8_t *d_states, int64_t B, float max_value, float *d_q,
                     int32_t *d_action, void *stream);
int gad_qnet_forward_host(gad_qnet *q, const uint8_t *h_states, int64_t B, float max_value, float *h_q,
                          int32_t *h_action);
/* Measurement hook (bench.py): the l1 tensor-core GEMM alone over M rows, `reps` launches between
 * CUDA events. *flops_per_launch counts the three fp16 passes actually executed. Synchronises. */
int gad_qnet_time_l1(gad_qnet *q, int64_t M, int reps, float *ms_per_launch, double *flops_per_launch,
                     void *stream);
/* GA.mutation's episode loop + Environment.padding + splice (GA.py:335-373) for M individuals.
 * Requires max row length + (rows-1)*cols <= stride; otherwise bit 0 of *d_status is set and the
 * board is left as it was. *h_forwards (optional) = lock-step forwards executed. Synchronises. */
int gad_mutate(gad_qnet *q, uint8_t *d_boards, int32_t *d_rowlen, int R, int stride, const int32_t *d_idx,
               const int32_t *d_tile, int64_t M, float max_value, int32_t *d_status, int64_t *h_forwards,
               void *stream);
/* Counters of the last gad_mutate on this handle: out4 = {episodes (= M, what GA.py:331-373 would run),
 * episodes actually run, lock-step forwards, states evaluated}. Episodes that start from byte-identical
 * windows are run once (the eval-mode network is a pure function of the window, GA.py:353-366) and share
 * the result; GAD_MUTATE_DEDUP=0 in the environment runs every one. */
int gad_mutate_stats(const gad_qnet *q, int64_t *out4);

/* GA.mutation in separate steps, for a population sharded over several GPUs (sharded.py): the windows of all
 * ranks are gathered, every rank finds the distinct ones, runs its share and the results are gathered back, so
 * that each distinct window runs once in the whole job and the ranks do equal work.
 *  gad_mutate_windows  d_windows[m] = the rows x cols window of individual d_idx[m] at d_tile[m] (GA.py:338-350);
 *                      all zero when no tile fits
 *  gad_windows_rep     d_rep[m] = smallest index of a window byte-identical to window m. Scratch: d_table int32
 *                      [table_slots] (a power of two >= 2 M), d_tmp int32 [M]
 *  gad_episodes_run    greedy episodes (GA.py:353-366; env.py:231-259) of windows d_list[0..U) (NULL: the first
 *                      U; scratch is kept for `reserve` >= U episodes so that later calls do not re-allocate).
 *                      d_aligned uint8 [U][rows][rows*cols], d_heads uint8 [U][rows], d_steps int32 [U].
 *                      Synchronises; gad_mutate_stats reports it like a gad_mutate of U distinct windows
 *  gad_splice_results  Environment.padding + splice (env.py:338-351, GA.py:369-373): individual m takes entry
 *                      d_slot[m] of the result arrays; d_windows = the local windows of gad_mutate_windows */
int gad_mutate_windows(const gad_qnet *q, const uint8_t *d_boards, int R, int stride, const int32_t *d_idx,
                       const int32_t *d_tile, int64_t M, uint8_t *d_windows, void *stream);
int gad_windows_rep(const uint8_t *d_windows, int64_t M, int window_bytes, int32_t *d_rep, int32_t *d_table,
                    int64_t table_slots, int32_t *d_tmp, void *stream);
int gad_episodes_run(gad_qnet *q, const uint8_t *d_windows, const int32_t *d_list, int64_t U, int64_t reserve,
                     float max_value, uint8_t *d_aligned, uint8_t *d_heads, int32_t *d_steps, int64_t *h_forwards,
                     void *stream);
int gad_splice_results(const gad_qnet *q, uint8_t *d_boards, int32_t *d_rowlen, int R, int stride,
                       const int32_t *d_idx, const int32_t *d_tile, int64_t M, const uint8_t *d_windows,
                       const int32_t *d_slot, const uint8_t *d_aligned, const uint8_t *d_heads,
                       const int32_t *d_steps, int32_t *d_status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GADPAMSA_H */
