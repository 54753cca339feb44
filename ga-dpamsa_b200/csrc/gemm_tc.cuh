// Host interface of the tcgen05 split-fp16 GEMM (gemm_tc.cu) used for the Q-network's dense layers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_BN_L1 208       // l1: 1028 outputs pad to 5 x 208 = 1040 (1.2 % waste)
#define TC_BN_L2 128       // l2: 512 outputs = 4 x 128
#define TC_BN_L3 64        // l3: 2^rows - 1 <= 64 outputs
#define TC_SPLITK_MAX_ROWS 4096   // launches of a fused layer with at most this many rows may split K over CTAs

struct TcLayer {
    int N = 0, K = 0, Kpad = 0, BN = 0, Npad = 0;
    int fused = 0;                           // 1: hi and lo tiles of a 32-wide k-block share a stage (64B swizzle)
    float scale = 1.f, inv_scale = 1.f;      // weights are stored as w * 2^s; the epilogue multiplies by 2^-s
    void *W2 = nullptr;                      // fp16 [Npad][2*Kpad] = (hi | lo), zero beyond N and K
    float *ws = nullptr;                     // fused layers: split-K partials [K_SLICES][rows][Npad]
    float *scratch = nullptr;                // fused layers: parked slice partials of whole-K units, per CTA
    alignas(64) unsigned char map_a[128];    // CUtensorMap of the activations (hi | lo), rebound when they move
    alignas(64) unsigned char map_b[128];    // CUtensorMap of W2
    alignas(64) unsigned char map_bh[128];   // fused layers: the same with a box of half a tile's rows (tail units)
};

// BN must be one of the TC_BN_* values. K is padded to a multiple of 64 (Kpad); the activation buffer
// bound with tc_layer_bind_a is fp16 [rows][2*Kpad] = (hi | lo) with zeros in the padding columns.
int tc_layer_init(TcLayer *L, const float *d_W, const float *h_W, int N, int K, int BN, int fused);
void tc_layer_free(TcLayer *L);
int tc_layer_bind_a(TcLayer *L, const void *d_A2, long long rows);
// out[m][n] = act((A . W^T)[m][n] * 2^-s + bias[n]) for m < *d_m (or m_max when d_m is NULL), n < N.
// split_out == 0: out is fp32 [.][ldc]. split_out != 0: out is fp16, hi at out[m*ldc + n] and lo at
// out[m*ldc + lo_offset + n] - the (hi | lo) activation layout of the next layer.
int tc_layer_run(const TcLayer *L, const float *d_bias, void *d_out, int ldc, int lo_offset, int split_out,
                 const int32_t *d_m, long long m_max, int leaky, cudaStream_t st);
