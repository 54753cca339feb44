// Q-network handle and the helpers shared by qnet.cu and attn_tc.cu.
#pragma once
#include "gad_common.cuh"
#include "gemm_tc.cuh"

#include <cuda_fp16.h>

#define QN_D 64            // d_model of the reference encoder (dqn.py:54)
#define QN_VOCAB 6         // 0 = padding, 1..5 = A T C G gap

struct gad_qnet {
    int rows, cols, L, A, K, h1, h2;
    // parameters / folded tables (device, fp32)
    float *X = nullptr, *T = nullptr, *VF = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    __half *VFh = nullptr, *VFl = nullptr;   // VF * 2^sv split into fp16 hi + lo (encode_mma_kernel's B operand)
    float vf_inv_scale = 1.f;
    // attention as tensor-core products (encode_fa_kernel): Q' = QM * log2(e) * 2^sq and K' = X * 2^sk, fp16 hi + lo
    __half *Qh = nullptr, *Ql = nullptr, *Kh = nullptr, *Kl = nullptr;
    float s_inv_scale = 1.f;                 // 2^-(sq+sk): raw MMA sum -> logit in the log2 domain
    float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr, *W3 = nullptr, *b3 = nullptr;
    // activation scratch for `cap` states
    long long cap = 0;
    uint8_t *tok = nullptr;
    float *Z = nullptr, *H1 = nullptr, *H2 = nullptr, *Q = nullptr;
    __half *Z2 = nullptr;            // (hi | lo) fp16 split of Z, [cap][2K]: A operand of the tcgen05 l1
    TcLayer *tc1 = nullptr;          // NULL = fp32 SIMT dense layers (GAD_QNET_SIMT=1)
    TcLayer *tc2 = nullptr, *tc3 = nullptr;
    __half *H1s = nullptr, *H2s = nullptr;   // (hi | lo) activations between the tensor-core layers
    float *Qpre = nullptr;           // l3 output before tanh, [cap][64]
    int32_t *act = nullptr;
    // episode state for `ecap` episodes
    long long ecap = 0;
    uint8_t *sub = nullptr, *heads = nullptr, *aligned = nullptr;
    int32_t *steps = nullptr, *list0 = nullptr, *list1 = nullptr, *counts = nullptr;   // counts[2]
    int32_t *h_count = nullptr;   // pinned
    // duplicate-window elimination (gad_mutate): rep[m] = the episode whose result episode m re-uses
    int32_t *rep = nullptr, *uniq = nullptr, *table = nullptr;
    long long tcap = 0;                          // slots of `table` (power of two)
    unsigned long long *d_stats = nullptr;       // [0] states evaluated (sum of the active counts of all forwards)
    long long last_stats[4] = {0, 0, 0, 0};      // episodes, unique episodes, lock-step forwards, states evaluated
    // staging of gad_qnet_forward_host (DQN.predict from python)
    long long hcap = 0;
    uint8_t *hs_states = nullptr;
    float *hs_q = nullptr;
    int32_t *hs_act = nullptr;
};


// tokens of episode e from its sub-board and head pointers (env.py:147-153): every row's remaining
// tokens followed by one gap code, rows concatenated, zero padding up to L
__device__ __forceinline__ int build_tokens(uint8_t *tok, const uint8_t *sub, const uint8_t *heads, int rows, int cols,
                                            int L)
{
    __shared__ int s_off[33];
    if (threadIdx.x == 0) {
        int o = 0;
        for (int r = 0; r < rows; ++r) {
            s_off[r] = o;
            o += cols - heads[r] + 1;
        }
        s_off[rows] = o;
    }
    __syncthreads();
    const int nvalid = s_off[rows];
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        uint8_t t = 0;
        if (i < nvalid) {
            int r = 0;
            while (i >= s_off[r + 1]) ++r;
            const int k = i - s_off[r];
            const int rem = cols - heads[r];
            t = k < rem ? sub[r * cols + heads[r] + k] : (uint8_t)GAD_GAP;
        }
        tok[i] = t;
    }
    __syncthreads();
    return nvalid;
}


// attn_tc.cu: the encoder on tcgen05 (returns GAD_EINVAL when the window does not fit its tiles)
int attn_tc_supported(const gad_qnet *q);
int launch_encode_tc(gad_qnet *q, const uint8_t *tok_in, const int32_t *active, const int32_t *n_ptr, long long n_max,
                     cudaStream_t st);
