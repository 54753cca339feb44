/*
 * gadpamsa.h - C ABI of libgadpamsa: the GA-DPAMSA generation loop on B200 (sm_100a).
 *
 * The reference (FLaTNNBio/GA-DPAMSA) is pure Python and has NO FFI of its own;
 * its drop-in boundary is the module surface of GA.py / utils.py / DPAMSA/*.py.
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
int gad_stream_sync(void *stream);                         /* synchronises */

/* ---- a4-a7: one fitness pass over the whole population ------------------
 * Replaces GA.calculate_fitness_score's per-board loop (GA.py:156-178):
 * right-pad (GA.py:158-162), utils.clean_unnecessary_gaps (utils.py:390-428),
 * utils.get_sum_of_pairs (utils.py:27-76) and utils.get_column_score
 * (utils.py:79-128) over the full board, for all P boards in one launch.
 * Boards are compacted in place and rowlen[p][*] is set to the new width.
 * Outputs (device, int32[P]): sp = sum-of-pairs, em = exact-match columns,
 * al = alignment length after cleaning; the facade forms CS = em / al as a
 * Python float so it is bit-identical to utils.py:128.
 * d_max_al (optional, device int32[1], caller zero-initialises) receives
 * max(al) by atomicMax. Requires R <= 255. */
int gad_fitness(uint8_t *d_boards, int32_t *d_rowlen, int64_t P, int R, int stride,
                int match, int mismatch, int gap,
                int32_t *d_sp, int32_t *d_em, int32_t *d_al, int32_t *d_max_al, void *stream);

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

#ifdef __cplusplus
}
#endif
#endif /* GADPAMSA_H */
