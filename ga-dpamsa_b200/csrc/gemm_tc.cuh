// Host interface of the tcgen05 split-fp16 GEMM (gemm_tc.cu) used for the Q-network's l1 layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_BN 208          // UMMA N: 1028 outputs pad to 5 x 208 = 1040 (1.2 % waste)

struct TcLayer {
    int N = 0, K = 0, BN = 0, Npad = 0;
    float scale = 1.f, inv_scale = 1.f;      // weights are stored as w * 2^s; the epilogue multiplies by 2^-s
    void *W2 = nullptr;                      // fp16 [Npad][2K] = (hi | lo)
    alignas(64) unsigned char map_a[128];    // CUtensorMap of the activations (hi | lo), rebound when they move
    alignas(64) unsigned char map_b[128];    // CUtensorMap of W2
};

int tc_layer_init(TcLayer *L, const float *d_W, const float *h_W, int N, int K);
void tc_layer_free(TcLayer *L);
int tc_layer_bind_a(TcLayer *L, const void *d_A2, long long rows);
// C[m][n] = act((A . W^T)[m][n] * 2^-s + bias[n]) for m < *d_m (or m_max when d_m is NULL), n < N
int tc_layer_run(const TcLayer *L, const float *d_bias, float *d_C, int ldc, const int32_t *d_m, long long m_max,
                 int leaky, cudaStream_t st);
