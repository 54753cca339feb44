// Batched DPAMSA Q-network and alignment-environment episodes: the mutation phase of the generation
// loop for every mutating individual at once. Replaces the per-individual python loop of GA.mutation
// (reference GA.py:335-373): Environment.reset/step/padding (DPAMSA/env.py:139-153,231-259,338-351),
// DQN.predict (DPAMSA/dqn.py:151-154) and Net.forward (dqn.py:69-79; models.py:241-249,190-207,87-95),
// in eval mode (dropout is the identity).
//
// B200-first restructuring of the encoder (not a translation of the torch graph):
//   * tokens are in {0..5} and the positional encoding is fixed, so everything per-token - embedding,
//     positional add, first LayerNorm and the bias-free Q/K/V projections - depends only on
//     (token, position). At load time the host folds them (in fp64) into three small tables:
//         X [6][L][64]      first-LayerNorm output (also the residual)
//         T [6][L][6][L]    attention logit  (x_i Wq^T / sqrt(dk)) . (x_j Wk^T)  for every pair
//         VF[6][L][64]      (x_j Wv^T) Wfc^T  - the value projection with the output projection folded in
//     so one forward needs no QKV or fc GEMM. The fast encoders recompute the logits on the tensor cores
//     from split-fp16 row tables (Q' = x M log2e, K' = x, V = VF): attn_tc.cu (tcgen05, the default for windows
//     of 16..192 positions) and encode_fa_kernel below (mma.sync); the fallback encoders gather logits from T.
//   * what remains dense is l1 (L*64 -> h1, ~95 % of the weights), l2 and l3: tcgen05 GEMMs over all active
//     episodes (gemm_tc.cu; fp32 SIMT tiles below as the fallback); tanh + argmax + Environment.step are fused
//     in one kernel.
//   * episodes run in lock-step on device with an active list that shrinks as episodes finish; the
//     host only polls the active count.
#include "qnet.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

const long long kMaxBatch = 32768;     // states per lock-step chunk (Z is kMaxBatch * K * 4 bytes)

// ---------------------------------------------------------------------------------------------
// encoder: tokens -> Z[row][L*64]  (attention over table logits, + residual, second LayerNorm)
// ---------------------------------------------------------------------------------------------
#define ENC_QT 4           // query rows per warp trip

__global__ void __launch_bounds__(256)
encode_kernel(const gad_qnet net, const uint8_t *__restrict__ tok_in, const int32_t *__restrict__ active,
              const int32_t *__restrict__ n_rows_ptr, long long n_rows_fixed, float *__restrict__ Zout,
              __half *__restrict__ Z2out)
{
    extern __shared__ __align__(16) float smem[];
    const int L = net.L;
    const int nw = blockDim.x >> 5;
    float *Pw = smem;                                         // [nw][L][ENC_QT]
    uint8_t *tok = reinterpret_cast<uint8_t *>(smem + (size_t)nw * L * ENC_QT);   // [L]
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();
        int nvalid;
        if (tok_in) {
            for (int i = threadIdx.x; i < L; i += blockDim.x) tok[i] = tok_in[row * L + i];
            __syncthreads();
            nvalid = 0;                                        // keys are masked where the token is 0 (dqn.py:67)
            for (int i = 0; i < L; ++i) nvalid += tok[i] != 0;
        } else {
            const int e = active[row];
            nvalid = build_tokens(tok, net.sub + (size_t)e * net.rows * net.cols, net.heads + (size_t)e * net.rows,
                                  net.rows, net.cols, L);
        }
        // Valid keys are not necessarily a prefix when tokens are given explicitly: build the key list.
        // (For environment states they are exactly the first nvalid positions.)
        float *P = Pw + (size_t)warp * L * ENC_QT;
        for (int i0 = warp * ENC_QT; i0 < L; i0 += nw * ENC_QT) {
            float mx[ENC_QT], sum[ENC_QT];
            const float *trow[ENC_QT];
#pragma unroll
            for (int q = 0; q < ENC_QT; ++q) {
                const int i = min(i0 + q, L - 1);
                trow[q] = net.T + ((size_t)tok[i] * L + i) * (QN_VOCAB * L);
                mx[q] = -INFINITY;
                sum[q] = 0.f;
            }
            // logits (masked keys get -1e9 like models.py:90-91)
            for (int j = lane; j < L; j += 32) {
                const int tj = tok[j];
#pragma unroll
                for (int q = 0; q < ENC_QT; ++q) {
                    const float s = tj ? __ldg(trow[q] + tj * L + j) : -1e9f;
                    P[j * ENC_QT + q] = s;
                    mx[q] = fmaxf(mx[q], s);
                }
            }
#pragma unroll
            for (int q = 0; q < ENC_QT; ++q)
                for (int o = 16; o > 0; o >>= 1) mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xFFFFFFFFu, mx[q], o));
            __syncwarp();
            for (int j = lane; j < L; j += 32) {
#pragma unroll
                for (int q = 0; q < ENC_QT; ++q) {
                    const float e = expf(P[j * ENC_QT + q] - mx[q]);
                    P[j * ENC_QT + q] = e;
                    sum[q] += e;
                }
            }
#pragma unroll
            for (int q = 0; q < ENC_QT; ++q)
                for (int o = 16; o > 0; o >>= 1) sum[q] += __shfl_xor_sync(0xFFFFFFFFu, sum[q], o);
            __syncwarp();
            // attention output with the fc projection folded in: acc[q][c] = sum_j e[q][j] * VF[tok_j][j][c]
            float acc0[ENC_QT], acc1[ENC_QT];
#pragma unroll
            for (int q = 0; q < ENC_QT; ++q) acc0[q] = acc1[q] = 0.f;
            for (int j = 0; j < L; ++j) {
                const int tj = tok[j];
                if (tj == 0 && nvalid) continue;               // exp(-1e9 - max) == 0 exactly
                const float *vf = net.VF + ((size_t)tj * L + j) * QN_D;
                const float v0 = __ldg(vf + lane), v1 = __ldg(vf + lane + 32);
                const float4 p = *reinterpret_cast<const float4 *>(P + j * ENC_QT);
                acc0[0] = fmaf(p.x, v0, acc0[0]);
                acc1[0] = fmaf(p.x, v1, acc1[0]);
                acc0[1] = fmaf(p.y, v0, acc0[1]);
                acc1[1] = fmaf(p.y, v1, acc1[1]);
                acc0[2] = fmaf(p.z, v0, acc0[2]);
                acc1[2] = fmaf(p.z, v1, acc1[2]);
                acc0[3] = fmaf(p.w, v0, acc0[3]);
                acc1[3] = fmaf(p.w, v1, acc1[3]);
            }
            const float g0 = net.ln2_g[lane], g1 = net.ln2_g[lane + 32], be0 = net.ln2_b[lane], be1 = net.ln2_b[lane + 32];
#pragma unroll
            for (int q = 0; q < ENC_QT; ++q) {
                const int i = i0 + q;
                if (i >= L) break;
                const float *x = net.X + ((size_t)tok[i] * L + i) * QN_D;
                const float inv = 1.f / sum[q];
                const float y0 = fmaf(acc0[q], inv, x[lane]), y1 = fmaf(acc1[q], inv, x[lane + 32]);   // + residual
                float s = y0 + y1;
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                const float mu = s * (1.f / QN_D);
                const float d0 = y0 - mu, d1 = y1 - mu;
                float v = d0 * d0 + d1 * d1;
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                const float rstd = rsqrtf(v * (1.f / QN_D) + 1e-6f);       // LayerNorm(eps=1e-6), models.py:166
                const float z0 = d0 * rstd * g0 + be0, z1 = d1 * rstd * g1 + be1;
                if (Z2out) {                                   // x = hi + lo in fp16 for the tensor-core l1
                    __half *zh = Z2out + (size_t)row * 2 * net.K + (size_t)i * QN_D, *zl = zh + net.K;
                    const __half h0 = __float2half_rn(z0), h1 = __float2half_rn(z1);
                    zh[lane] = h0;
                    zh[lane + 32] = h1;
                    zl[lane] = __float2half_rn(z0 - __half2float(h0));
                    zl[lane + 32] = __float2half_rn(z1 - __half2float(h1));
                } else {
                    float *z = Zout + (size_t)row * net.K + (size_t)i * QN_D;
                    z[lane] = z0;
                    z[lane + 32] = z1;
                }
            }
            __syncwarp();
        }
    }
}

// Tiled encoder (the one that runs for the reference's window sizes): one CTA per state.
//   stage 0  tokens, the list of valid (unmasked) keys, and their VF rows -> shared memory
//   per block of ENC_RB query rows:
//     stage A  one warp per row: logits gathered from T, fp32 softmax numerators e[k][row] -> shared
//              (transposed, so stage B reads them as the A operand), row sums kept
//     stage B  [ENC_RB x nk] . [nk x 64] register-tiled product (6 rows x 4 channels per thread,
//              24 FMA per 4 shared loads), then /sum + residual + LayerNorm and the store
// Masked keys contribute exp(-1e9 - max) == 0 in the reference; they are simply not in the key list.
template <int RB> struct EncTile {
    static constexpr int THREADS = RB / 96 * 256;          // 6 rows x 4 channels per thread in stage B
    static constexpr int PS = RB + 2;                      // row stride of the transposed numerators (even)
};

template <int RB>
__global__ void __launch_bounds__(EncTile<RB>::THREADS)
encode_tiled_kernel(const gad_qnet net, const uint8_t *__restrict__ tok_in, const int32_t *__restrict__ active,
                    const int32_t *__restrict__ n_rows_ptr, long long n_rows_fixed, float *__restrict__ Zout,
                    __half *__restrict__ Z2out)
{
    constexpr int PS = EncTile<RB>::PS;
    constexpr int NW = EncTile<RB>::THREADS / 32;
    extern __shared__ __align__(16) float smem[];
    const int L = net.L;
    float *VFs = smem;                                   // [L][64]
    float *Pt = VFs + (size_t)L * QN_D;                  // [L][PS]
    float *rsum = Pt + (size_t)L * PS;                   // [RB]
    int *kj = reinterpret_cast<int *>(rsum + RB);        // [L] valid key positions
    int *koff = kj + L;                                  // [L] tok[j] * L + j of every valid key
    uint8_t *tok = reinterpret_cast<uint8_t *>(koff + L);  // [L]
    __shared__ int s_nk;
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = lane & 15, ty = warp * 2 + (lane >> 4);          // stage-B tile: rows ty*6.., channels tx*4..

    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();
        if (tok_in) {
            for (int i = threadIdx.x; i < L; i += blockDim.x) tok[i] = tok_in[row * L + i];
            __syncthreads();
        } else {
            const int e = active[row];
            build_tokens(tok, net.sub + (size_t)e * net.rows * net.cols, net.heads + (size_t)e * net.rows, net.rows,
                         net.cols, L);
        }
        if (warp == 0) {                                 // ordered list of unmasked keys
            int base = 0;
            for (int j0 = 0; j0 < L; j0 += 32) {
                const int j = j0 + lane;
                const bool v = j < L && tok[j] != 0;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, v);
                if (v) kj[base + __popc(m & ((1u << lane) - 1u))] = j;
                base += __popc(m);
            }
            if (lane == 0) s_nk = base;
        }
        __syncthreads();
        int nk = s_nk;
        const bool all_masked = nk == 0;                 // softmax over all -1e9 is uniform (never an environment state)
        if (all_masked) {
            for (int j = threadIdx.x; j < L; j += blockDim.x) kj[j] = j;
            nk = L;
            __syncthreads();
        }
        for (int i = threadIdx.x; i < nk * (QN_D / 4); i += blockDim.x) {
            const int k = i >> 4, c4 = i & 15;
            const int j = kj[k];
            reinterpret_cast<float4 *>(VFs)[i] =
                __ldg(reinterpret_cast<const float4 *>(net.VF + ((size_t)tok[j] * L + j) * QN_D) + c4);
        }
        for (int k = threadIdx.x; k < nk; k += blockDim.x) {          // offset of key k inside a row of T
            const int j = kj[k];
            koff[k] = tok[j] * L + j;
        }
        for (int rb = 0; rb < L; rb += RB) {
            __syncthreads();
            // ---- stage A: one warp per pair of rows; a lane keeps logits k = lane + 32u in registers (all loads of
            // the pair are issued before the first use), softmax numerators go to shared memory transposed
            constexpr int U = RB / 32;
            for (int r0 = 2 * warp; r0 < RB; r0 += 2 * NW) {
                float sv[2][U];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = rb + r0 + h;
                    const bool live = i < L;
                    const int ii = live ? i : 0;
                    const float *trow = net.T + ((size_t)tok[ii] * L + ii) * (QN_VOCAB * L);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = lane + 32 * u;
                        float v = -INFINITY;
                        if (live && k < nk) v = all_masked ? -1e9f : __ldg(trow + koff[k]);
                        sv[h][u] = v;
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = r0 + h;
                    if (rb + r >= L) {                   // rows beyond the state: stage B multiplies zeros
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            if (lane + 32 * u < nk) Pt[(lane + 32 * u) * PS + r] = 0.f;
                        if (lane == 0) rsum[r] = 1.f;
                        continue;
                    }
                    float mx = sv[h][0];
#pragma unroll
                    for (int u = 1; u < U; ++u) mx = fmaxf(mx, sv[h][u]);
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
                    float sum = 0.f;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = lane + 32 * u;
                        if (k < nk) {
                            const float ev = expf(sv[h][u] - mx);
                            Pt[k * PS + r] = ev;
                            sum += ev;
                        }
                    }
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
                    if (lane == 0) rsum[r] = sum;
                }
            }
            __syncthreads();
            // ---- stage B
            float acc[6][4];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
            const float *pp = Pt + ty * 6;
            const float *vp = VFs + tx * 4;
#pragma unroll 4
            for (int k = 0; k < nk; ++k) {
                const float2 p01 = *reinterpret_cast<const float2 *>(pp + k * PS);
                const float2 p23 = *reinterpret_cast<const float2 *>(pp + k * PS + 2);
                const float2 p45 = *reinterpret_cast<const float2 *>(pp + k * PS + 4);
                const float4 v = *reinterpret_cast<const float4 *>(vp + k * QN_D);
                const float p[6] = {p01.x, p01.y, p23.x, p23.y, p45.x, p45.y};
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    acc[a][0] = fmaf(p[a], v.x, acc[a][0]);
                    acc[a][1] = fmaf(p[a], v.y, acc[a][1]);
                    acc[a][2] = fmaf(p[a], v.z, acc[a][2]);
                    acc[a][3] = fmaf(p[a], v.w, acc[a][3]);
                }
            }
            const float4 g = *reinterpret_cast<const float4 *>(net.ln2_g + tx * 4);
            const float4 be = *reinterpret_cast<const float4 *>(net.ln2_b + tx * 4);
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                const int r = ty * 6 + a, i = rb + r;
                // all 16 lanes of the half-warp share the row: the guard is uniform across the shuffles below
                const bool live = i < L;
                const int ii = live ? i : L - 1;
                const float4 x = __ldg(reinterpret_cast<const float4 *>(net.X + ((size_t)tok[ii] * L + ii) * QN_D) + tx);
                const float inv = 1.f / rsum[r];
                const float y0 = fmaf(acc[a][0], inv, x.x), y1 = fmaf(acc[a][1], inv, x.y);
                const float y2 = fmaf(acc[a][2], inv, x.z), y3 = fmaf(acc[a][3], inv, x.w);
                float sm = (y0 + y1) + (y2 + y3);
                for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xFFFFFFFFu, sm, o);
                const float mu = sm * (1.f / QN_D);
                const float d0 = y0 - mu, d1 = y1 - mu, d2 = y2 - mu, d3 = y3 - mu;
                float vr = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
                for (int o = 8; o > 0; o >>= 1) vr += __shfl_xor_sync(0xFFFFFFFFu, vr, o);
                const float rstd = rsqrtf(vr * (1.f / QN_D) + 1e-6f);          // LayerNorm(eps=1e-6), models.py:166
                const float z[4] = {d0 * rstd * g.x + be.x, d1 * rstd * g.y + be.y, d2 * rstd * g.z + be.z,
                                    d3 * rstd * g.w + be.w};
                if (!live) continue;
                if (Z2out) {                                                   // x = hi + lo in fp16 for the tensor-core l1
                    __half *zh = Z2out + (size_t)row * 2 * net.K + (size_t)i * QN_D + tx * 4, *zl = zh + net.K;
                    __half h[4], l[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        h[b] = __float2half_rn(z[b]);
                        l[b] = __float2half_rn(z[b] - __half2float(h[b]));
                    }
                    *reinterpret_cast<uint2 *>(zh) = *reinterpret_cast<const uint2 *>(h);
                    *reinterpret_cast<uint2 *>(zl) = *reinterpret_cast<const uint2 *>(l);
                } else {
                    *reinterpret_cast<float4 *>(Zout + (size_t)row * net.K + (size_t)i * QN_D + tx * 4) =
                        make_float4(z[0], z[1], z[2], z[3]);
                }
            }
        }
    }
}

// Encoder on the warp-level tensor-core path (mma.sync m16n8k16, fp16 in / fp32 accumulate). One CTA of 8
// warps per state, two CTAs per SM so that one CTA's logit gathers overlap the other's MMAs.
//   per block of 64 query rows:
//     stage A  as above, but the softmax numerators are stored as fp16 hi + lo (E = Eh + El), row-major
//     stage B  Y[64 x 64] = E[64 x nk] . V[nk x 64] with V = Vh + Vl (pre-split, pre-scaled by 2^sv at load
//              time): three MMAs per tile (Eh.Vh + El.Vh + Eh.Vl) - fp32-class accuracy, ~36 MMAs deep
//     stage C  /sum, * 2^-sv, + residual, LayerNorm, store (hi | lo fp16 for the tcgen05 l1, or fp32)
#define EM_RB 64
#define EM_KP 200          // row stride of E in halves (400 B: conflict-free ldmatrix rows)
#define EM_VP 72           // row stride of V in halves (144 B)
#define EM_YP 66           // row stride of the fp32 staging tile

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2)
encode_mma_kernel(const gad_qnet net, const uint8_t *__restrict__ tok_in, const int32_t *__restrict__ active,
                  const int32_t *__restrict__ n_rows_ptr, long long n_rows_fixed, float *__restrict__ Zout,
                  __half *__restrict__ Z2out)
{
    extern __shared__ __align__(16) float smem[];
    const int L = net.L;
    const int LK = (L + 15) & ~15;                                   // keys padded to the MMA k-step
    __half *Vh = reinterpret_cast<__half *>(smem);                   // [LK][EM_VP]
    __half *Vl = Vh + (size_t)LK * EM_VP;
    __half *Eh = Vl + (size_t)LK * EM_VP;                            // [EM_RB][EM_KP]
    __half *El = Eh + (size_t)EM_RB * EM_KP;
    float *rsum = reinterpret_cast<float *>(El + (size_t)EM_RB * EM_KP);   // [EM_RB]
    int *kj = reinterpret_cast<int *>(rsum + EM_RB);                 // [L]
    int *koff = kj + L;                                              // [L]
    uint8_t *tok = reinterpret_cast<uint8_t *>(koff + L);            // [L]
    float *Ybuf = reinterpret_cast<float *>(Eh);                     // [EM_RB][EM_YP], aliases Eh/El after stage B
    __shared__ int s_nk;
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = 8;
    const float vscale_inv = net.vf_inv_scale;

    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();
        if (tok_in) {
            for (int i = threadIdx.x; i < L; i += blockDim.x) tok[i] = tok_in[row * L + i];
            __syncthreads();
        } else {
            const int e = active[row];
            build_tokens(tok, net.sub + (size_t)e * net.rows * net.cols, net.heads + (size_t)e * net.rows, net.rows,
                         net.cols, L);
        }
        if (warp == 0) {                                 // ordered list of unmasked keys
            int base = 0;
            for (int j0 = 0; j0 < L; j0 += 32) {
                const int j = j0 + lane;
                const bool v = j < L && tok[j] != 0;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, v);
                if (v) kj[base + __popc(m & ((1u << lane) - 1u))] = j;
                base += __popc(m);
            }
            if (lane == 0) s_nk = base;
        }
        __syncthreads();
        int nk = s_nk;
        const bool all_masked = nk == 0;                 // softmax over all -1e9 is uniform (never an environment state)
        if (all_masked) {
            for (int j = threadIdx.x; j < L; j += blockDim.x) kj[j] = j;
            nk = L;
            __syncthreads();
        }
        const int nk16 = (nk + 15) & ~15;
        for (int k = threadIdx.x; k < nk; k += blockDim.x) {
            const int j = kj[k];
            koff[k] = tok[j] * L + j;
        }
        // V rows of the valid keys (pre-split fp16 tables, 16 bytes per lane), zero rows up to the k-step
        for (int i = threadIdx.x; i < nk16 * 8; i += blockDim.x) {
            const int k = i >> 3, c8 = i & 7;
            uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = vh;
            if (k < nk) {
                const int j = kj[k];
                const size_t src = ((size_t)tok[j] * L + j) * QN_D + c8 * 8;
                vh = __ldg(reinterpret_cast<const uint4 *>(net.VFh + src));
                vl = __ldg(reinterpret_cast<const uint4 *>(net.VFl + src));
            }
            *reinterpret_cast<uint4 *>(Vh + (size_t)k * EM_VP + c8 * 8) = vh;
            *reinterpret_cast<uint4 *>(Vl + (size_t)k * EM_VP + c8 * 8) = vl;
        }
        for (int rb = 0; rb < L; rb += EM_RB) {
            __syncthreads();
            // ---- stage A: softmax numerators of 64 rows, one warp per pair of rows, logits in registers
            constexpr int U = 6;                         // keys per lane: up to 192
            for (int r0 = 2 * warp; r0 < EM_RB; r0 += 2 * NW) {
                float sv[2][U];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = rb + r0 + h;
                    const bool live = i < L;
                    const int ii = live ? i : 0;
                    const float *trow = net.T + ((size_t)tok[ii] * L + ii) * (QN_VOCAB * L);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = lane + 32 * u;
                        float v = -INFINITY;
                        if (live && k < nk) v = all_masked ? -1e9f : __ldg(trow + koff[k]);
                        sv[h][u] = v;
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = r0 + h;
                    const bool live = rb + r < L;
                    float mx = sv[h][0];
#pragma unroll
                    for (int u = 1; u < U; ++u) mx = fmaxf(mx, sv[h][u]);
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
                    float sum = 0.f;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int k = lane + 32 * u;
                        if (k < nk16) {
                            float ev = 0.f;
                            if (live && k < nk) ev = expf(sv[h][u] - mx);
                            const __half hi = __float2half_rn(ev);
                            Eh[r * EM_KP + k] = hi;
                            El[r * EM_KP + k] = __float2half_rn(ev - __half2float(hi));
                            sum += ev;
                        }
                    }
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
                    if (lane == 0) rsum[r] = live ? sum : 1.f;
                }
            }
            __syncthreads();
            // ---- stage B: warp (mt, nh) owns rows 16*mt.., channels 32*nh.. : 4 n-tiles, 3 MMAs each per k-step
            const int mt = warp & 3, nh = warp >> 2;
            float acc[4][4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
            const uint32_t a_off = (uint32_t)(((mt * 16 + (lane & 15)) * EM_KP + (lane >> 4) * 8) * 2);
            const uint32_t b_off = (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * EM_VP + nh * 32 + (lane >> 4) * 8) * 2);
            const uint32_t eh0 = (uint32_t)__cvta_generic_to_shared(Eh) + a_off;
            const uint32_t el0 = (uint32_t)__cvta_generic_to_shared(El) + a_off;
            const uint32_t vh0 = (uint32_t)__cvta_generic_to_shared(Vh) + b_off;
            const uint32_t vl0 = (uint32_t)__cvta_generic_to_shared(Vl) + b_off;
            for (int k0 = 0; k0 < nk16; k0 += 16) {
                uint32_t ah[4], al[4], bh[2][4], bl[2][4];
                ldsm_x4(ah, eh0 + k0 * 2);
                ldsm_x4(al, el0 + k0 * 2);
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {         // two n-tiles (16 channels) per ldmatrix.x4.trans
                    ldsm_x4_trans(bh[pr], vh0 + (k0 * EM_VP + pr * 16) * 2);
                    ldsm_x4_trans(bl[pr], vl0 + (k0 * EM_VP + pr * 16) * 2);
                }
#pragma unroll
                for (int pr = 0; pr < 2; ++pr)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float (&d)[4] = acc[pr * 2 + q];
                        mma_16816(d, al, bh[pr][2 * q], bh[pr][2 * q + 1]);
                        mma_16816(d, ah, bl[pr][2 * q], bl[pr][2 * q + 1]);
                        mma_16816(d, ah, bh[pr][2 * q], bh[pr][2 * q + 1]);
                    }
            }
            __syncthreads();                             // every warp is done reading E: reuse it as the fp32 staging tile
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int c = nh * 32 + t * 8 + (lane & 3) * 2;
                const int r = mt * 16 + (lane >> 2);
                *reinterpret_cast<float2 *>(Ybuf + r * EM_YP + c) = make_float2(acc[t][0], acc[t][1]);
                *reinterpret_cast<float2 *>(Ybuf + (r + 8) * EM_YP + c) = make_float2(acc[t][2], acc[t][3]);
            }
            __syncthreads();
            // ---- stage C: one warp per row, two channels per lane
            const float2 g = *reinterpret_cast<const float2 *>(net.ln2_g + 2 * lane);
            const float2 be = *reinterpret_cast<const float2 *>(net.ln2_b + 2 * lane);
            for (int r = warp; r < EM_RB; r += NW) {
                const int i = rb + r;
                if (i >= L) break;
                const float2 a = *reinterpret_cast<const float2 *>(Ybuf + r * EM_YP + 2 * lane);
                const float2 x = __ldg(reinterpret_cast<const float2 *>(net.X + ((size_t)tok[i] * L + i) * QN_D) + lane);
                const float inv = vscale_inv / rsum[r];
                const float y0 = fmaf(a.x, inv, x.x), y1 = fmaf(a.y, inv, x.y);
                float sm = y0 + y1;
                for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xFFFFFFFFu, sm, o);
                const float mu = sm * (1.f / QN_D);
                const float d0 = y0 - mu, d1 = y1 - mu;
                float vr = d0 * d0 + d1 * d1;
                for (int o = 16; o > 0; o >>= 1) vr += __shfl_xor_sync(0xFFFFFFFFu, vr, o);
                const float rstd = rsqrtf(vr * (1.f / QN_D) + 1e-6f);          // LayerNorm(eps=1e-6), models.py:166
                const float z0 = d0 * rstd * g.x + be.x, z1 = d1 * rstd * g.y + be.y;
                if (Z2out) {
                    __half *zh = Z2out + (size_t)row * 2 * net.K + (size_t)i * QN_D + 2 * lane, *zl = zh + net.K;
                    const __half h0 = __float2half_rn(z0), h1 = __float2half_rn(z1);
                    *reinterpret_cast<__half2 *>(zh) = __halves2half2(h0, h1);
                    *reinterpret_cast<__half2 *>(zl) =
                        __halves2half2(__float2half_rn(z0 - __half2float(h0)), __float2half_rn(z1 - __half2float(h1)));
                } else {
                    *reinterpret_cast<float2 *>(Zout + (size_t)row * net.K + (size_t)i * QN_D + 2 * lane) = make_float2(z0, z1);
                }
            }
        }
    }
}

int encode_mma_smem_bytes(int L)
{
    const int LK = (L + 15) & ~15;
    return 2 * LK * EM_VP * 2 + 2 * EM_RB * EM_KP * 2 + EM_RB * 4 + 2 * L * 4 + ((L + 15) & ~15);
}

// Encoder as a flash-attention style kernel on mma.sync (the default for the reference's window sizes).
// Nothing is gathered element by element: a state's queries, keys and values are rows of three small
// (token, position) tables (fp16 hi + lo), i.e. whole 128-byte rows.
//   S = Q' K'^T          (16 query rows per warp x 16 keys per trip, 3 split MMAs per tile, log2 domain)
//   online softmax       (running max / sum per row in registers, ex2)
//   O += P V             (P re-used from the S accumulator registers as the A operand, split hi + lo)
//   epilogue             O / sum * 2^-sv + residual, LayerNorm across the lane quad, (hi | lo) store
#define FA_WARPS 6
#define FA_KP 72           // row stride of the K / V tiles in halves (144 B: conflict-free ldmatrix rows)

__device__ __forceinline__ float fast_exp2(float x)       // ex2.approx.ftz: 2 ulp, exp2(-inf) = 0
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b)
{
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

__global__ void __launch_bounds__(FA_WARPS * 32, 2)
encode_fa_kernel(const gad_qnet net, const uint8_t *__restrict__ tok_in, const int32_t *__restrict__ active,
                 const int32_t *__restrict__ n_rows_ptr, long long n_rows_fixed, float *__restrict__ Zout,
                 __half *__restrict__ Z2out)
{
    extern __shared__ __align__(16) float smem[];
    const int L = net.L;
    const int LK = (L + 15) & ~15;
    __half *Ks_h = reinterpret_cast<__half *>(smem);                 // [LK][FA_KP] each
    __half *Ks_l = Ks_h + (size_t)LK * FA_KP;
    __half *Vs_h = Ks_l + (size_t)LK * FA_KP;
    __half *Vs_l = Vs_h + (size_t)LK * FA_KP;
    int *kj = reinterpret_cast<int *>(Vs_l + (size_t)LK * FA_KP);    // [L] valid key positions
    uint8_t *tok = reinterpret_cast<uint8_t *>(kj + L);              // [L]
    __shared__ int s_nk;
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quad = lane & 3, qrow = lane >> 2;
    const float sc = net.s_inv_scale;
    const int n_mtiles = (L + 15) >> 4;

    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();
        if (tok_in) {
            for (int i = threadIdx.x; i < L; i += blockDim.x) tok[i] = tok_in[row * L + i];
            __syncthreads();
        } else {
            const int e = active[row];
            build_tokens(tok, net.sub + (size_t)e * net.rows * net.cols, net.heads + (size_t)e * net.rows, net.rows,
                         net.cols, L);
        }
        if (warp == 0) {                                 // ordered list of unmasked keys
            int base = 0;
            for (int j0 = 0; j0 < L; j0 += 32) {
                const int j = j0 + lane;
                const bool v = j < L && tok[j] != 0;
                const unsigned m = __ballot_sync(0xFFFFFFFFu, v);
                if (v) kj[base + __popc(m & ((1u << lane) - 1u))] = j;
                base += __popc(m);
            }
            if (lane == 0) s_nk = base;
        }
        __syncthreads();
        int nk = s_nk;
        if (nk == 0) {                                   // every key masked: the reference's softmax is uniform over all keys;
            for (int j = threadIdx.x; j < L; j += blockDim.x) kj[j] = j;   // handled by zeroing the logits below
            __syncthreads();
        }
        const bool all_masked = nk == 0;
        if (all_masked) nk = L;
        const int nk16 = (nk + 15) & ~15;
        // K and V rows of the valid keys: 4 tables x 128 bytes per key, 16 bytes per thread; zero rows up to nk16
        for (int i = threadIdx.x; i < nk16 * 8; i += blockDim.x) {
            const int k = i >> 3, c8 = (i & 7) * 8;
            uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a, c = a, d = a;
            if (k < nk) {
                const int j = kj[k];
                const size_t src = ((size_t)tok[j] * L + j) * QN_D + c8;
                a = __ldg(reinterpret_cast<const uint4 *>(net.Kh + src));
                b = __ldg(reinterpret_cast<const uint4 *>(net.Kl + src));
                c = __ldg(reinterpret_cast<const uint4 *>(net.VFh + src));
                d = __ldg(reinterpret_cast<const uint4 *>(net.VFl + src));
            }
            const size_t dst = (size_t)k * FA_KP + c8;
            *reinterpret_cast<uint4 *>(Ks_h + dst) = a;
            *reinterpret_cast<uint4 *>(Ks_l + dst) = b;
            *reinterpret_cast<uint4 *>(Vs_h + dst) = c;
            *reinterpret_cast<uint4 *>(Vs_l + dst) = d;
        }
        __syncthreads();

        const uint32_t kb_off = (uint32_t)((((lane & 7) + ((lane >> 4) & 1) * 8) * FA_KP + ((lane >> 3) & 1) * 8) * 2);
        const uint32_t vb_off = (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * FA_KP + (lane >> 4) * 8) * 2);
        const uint32_t kh0 = (uint32_t)__cvta_generic_to_shared(Ks_h) + kb_off;
        const uint32_t kl0 = (uint32_t)__cvta_generic_to_shared(Ks_l) + kb_off;
        const uint32_t vh0 = (uint32_t)__cvta_generic_to_shared(Vs_h) + vb_off;
        const uint32_t vl0 = (uint32_t)__cvta_generic_to_shared(Vs_l) + vb_off;

        for (int mt = warp; mt < n_mtiles; mt += FA_WARPS) {
            const int iA = mt * 16 + qrow, iB = iA + 8;
            const int rA = min(iA, L - 1), rB = min(iB, L - 1);       // rows beyond the state are computed and dropped
            // Q fragments straight from the tables (A operand layout: a0 rowA k0.., a1 rowB k0.., a2 rowA k8.., a3 rowB k8..)
            uint32_t qh[4][4], ql[4][4];
            {
                const size_t oA = ((size_t)tok[rA] * L + rA) * QN_D + quad * 2, oB = ((size_t)tok[rB] * L + rB) * QN_D + quad * 2;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    qh[ks][0] = __ldg(reinterpret_cast<const uint32_t *>(net.Qh + oA + ks * 16));
                    qh[ks][1] = __ldg(reinterpret_cast<const uint32_t *>(net.Qh + oB + ks * 16));
                    qh[ks][2] = __ldg(reinterpret_cast<const uint32_t *>(net.Qh + oA + ks * 16 + 8));
                    qh[ks][3] = __ldg(reinterpret_cast<const uint32_t *>(net.Qh + oB + ks * 16 + 8));
                    ql[ks][0] = __ldg(reinterpret_cast<const uint32_t *>(net.Ql + oA + ks * 16));
                    ql[ks][1] = __ldg(reinterpret_cast<const uint32_t *>(net.Ql + oB + ks * 16));
                    ql[ks][2] = __ldg(reinterpret_cast<const uint32_t *>(net.Ql + oA + ks * 16 + 8));
                    ql[ks][3] = __ldg(reinterpret_cast<const uint32_t *>(net.Ql + oB + ks * 16 + 8));
                }
            }
            float o[8][4];
#pragma unroll
            for (int t = 0; t < 8; ++t)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[t][c] = 0.f;
            float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;

            // S = Q'K'^T of one 16-key block: six independent accumulation chains (3 split terms x 2 key tiles)
            auto compute_s = [&](int kb, float (&s)[2][4]) {
                float s1[2][4], s2[2][4];
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[t][c] = s1[t][c] = s2[t][c] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t kh[4], kl[4];
                    ldsm_x4(kh, kh0 + (kb * FA_KP + ks * 16) * 2);
                    ldsm_x4(kl, kl0 + (kb * FA_KP + ks * 16) * 2);
                    mma_16816(s1[0], ql[ks], kh[0], kh[1]);
                    mma_16816(s2[0], qh[ks], kl[0], kl[1]);
                    mma_16816(s[0], qh[ks], kh[0], kh[1]);
                    mma_16816(s1[1], ql[ks], kh[2], kh[3]);
                    mma_16816(s2[1], qh[ks], kl[2], kl[3]);
                    mma_16816(s[1], qh[ks], kh[2], kh[3]);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[t][c] += s1[t][c] + s2[t][c];
            };
            // software pipeline: the MMAs of block kb+16 are issued before the softmax of block kb, so the
            // tensor pipe keeps working while this warp does the scalar part
            float s_next[2][4];
            compute_s(0, s_next);
            for (int kb = 0; kb < nk16; kb += 16) {
                float s[2][4];
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[t][c] = s_next[t][c];
                if (kb + 16 < nk16) compute_s(kb + 16, s_next);
                // logits of this thread: key = kb + 8t + 2*quad + {0,1}; c0,c1 -> row A, c2,c3 -> row B
                float bxA = -INFINITY, bxB = -INFINITY;
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int key = kb + 8 * t + 2 * quad + (c & 1);
                        float v = all_masked ? 0.f : s[t][c] * sc;
                        if (key >= nk) v = -INFINITY;
                        s[t][c] = v;
                        if (c < 2) bxA = fmaxf(bxA, v); else bxB = fmaxf(bxB, v);
                    }
                bxA = fmaxf(bxA, __shfl_xor_sync(0xFFFFFFFFu, bxA, 1));
                bxA = fmaxf(bxA, __shfl_xor_sync(0xFFFFFFFFu, bxA, 2));
                bxB = fmaxf(bxB, __shfl_xor_sync(0xFFFFFFFFu, bxB, 1));
                bxB = fmaxf(bxB, __shfl_xor_sync(0xFFFFFFFFu, bxB, 2));
                // Lazy rescaling: the running reference max only moves when a block exceeds it by more than 2^8
                // (numerators then stay <= 256, exact enough in the hi + lo fp16 split), so the accumulator
                // rescale below runs for the first block or two of a row tile and almost never again.
                // Every 16-key block holds at least one valid key, so the maxima are finite.
                const float nA = bxA > mA + 8.f ? bxA : mA, nB = bxB > mB + 8.f ? bxB : mB;
                const float fA = fast_exp2(mA - nA), fB = fast_exp2(mB - nB);      // exp2(-inf) = 0 on the first block
                mA = nA;
                mB = nB;
                float p[2][4];
                float sumA = 0.f, sumB = 0.f;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    p[t][0] = fast_exp2(s[t][0] - nA);
                    p[t][1] = fast_exp2(s[t][1] - nA);
                    p[t][2] = fast_exp2(s[t][2] - nB);
                    p[t][3] = fast_exp2(s[t][3] - nB);
                    sumA += p[t][0] + p[t][1];
                    sumB += p[t][2] + p[t][3];
                }
                lA = fmaf(lA, fA, sumA);
                lB = fmaf(lB, fB, sumB);
                if (__any_sync(0xFFFFFFFFu, fA != 1.f || fB != 1.f)) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        o[t][0] *= fA;
                        o[t][1] *= fA;
                        o[t][2] *= fB;
                        o[t][3] *= fB;
                    }
                }
                // P as the A operand of the next product (16 rows x 16 keys), split hi + lo
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const __half2 hA = __floats2half2_rn(p[t][0], p[t][1]), hB = __floats2half2_rn(p[t][2], p[t][3]);
                    const float2 fAh = __half22float2(hA), fBh = __half22float2(hB);
                    ph[2 * t + 0] = *reinterpret_cast<const uint32_t *>(&hA);
                    ph[2 * t + 1] = *reinterpret_cast<const uint32_t *>(&hB);
                    pl[2 * t + 0] = pack_half2(p[t][0] - fAh.x, p[t][1] - fAh.y);
                    pl[2 * t + 1] = pack_half2(p[t][2] - fBh.x, p[t][3] - fBh.y);
                }
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {         // 16 channels per trip
                    uint32_t vh[4], vl[4];
                    ldsm_x4_trans(vh, vh0 + (kb * FA_KP + pr * 16) * 2);
                    ldsm_x4_trans(vl, vl0 + (kb * FA_KP + pr * 16) * 2);
                    mma_16816(o[2 * pr], pl, vh[0], vh[1]);
                    mma_16816(o[2 * pr], ph, vl[0], vl[1]);
                    mma_16816(o[2 * pr], ph, vh[0], vh[1]);
                    mma_16816(o[2 * pr + 1], pl, vh[2], vh[3]);
                    mma_16816(o[2 * pr + 1], ph, vl[2], vl[3]);
                    mma_16816(o[2 * pr + 1], ph, vh[2], vh[3]);
                }
            }
            // ---- epilogue: the lane quad holds a whole 64-channel row
            lA += __shfl_xor_sync(0xFFFFFFFFu, lA, 1);
            lA += __shfl_xor_sync(0xFFFFFFFFu, lA, 2);
            lB += __shfl_xor_sync(0xFFFFFFFFu, lB, 1);
            lB += __shfl_xor_sync(0xFFFFFFFFu, lB, 2);
            const float invA = net.vf_inv_scale / lA, invB = net.vf_inv_scale / lB;
            const float *xA = net.X + ((size_t)tok[rA] * L + rA) * QN_D + quad * 2;
            const float *xB = net.X + ((size_t)tok[rB] * L + rB) * QN_D + quad * 2;
            float smA = 0.f, smB = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float2 a = __ldg(reinterpret_cast<const float2 *>(xA + 8 * t));
                const float2 b = __ldg(reinterpret_cast<const float2 *>(xB + 8 * t));
                o[t][0] = fmaf(o[t][0], invA, a.x);
                o[t][1] = fmaf(o[t][1], invA, a.y);
                o[t][2] = fmaf(o[t][2], invB, b.x);
                o[t][3] = fmaf(o[t][3], invB, b.y);
                smA += o[t][0] + o[t][1];
                smB += o[t][2] + o[t][3];
            }
            smA += __shfl_xor_sync(0xFFFFFFFFu, smA, 1);
            smA += __shfl_xor_sync(0xFFFFFFFFu, smA, 2);
            smB += __shfl_xor_sync(0xFFFFFFFFu, smB, 1);
            smB += __shfl_xor_sync(0xFFFFFFFFu, smB, 2);
            const float muA = smA * (1.f / QN_D), muB = smB * (1.f / QN_D);
            float vrA = 0.f, vrB = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                o[t][0] -= muA;
                o[t][1] -= muA;
                o[t][2] -= muB;
                o[t][3] -= muB;
                vrA += o[t][0] * o[t][0] + o[t][1] * o[t][1];
                vrB += o[t][2] * o[t][2] + o[t][3] * o[t][3];
            }
            vrA += __shfl_xor_sync(0xFFFFFFFFu, vrA, 1);
            vrA += __shfl_xor_sync(0xFFFFFFFFu, vrA, 2);
            vrB += __shfl_xor_sync(0xFFFFFFFFu, vrB, 1);
            vrB += __shfl_xor_sync(0xFFFFFFFFu, vrB, 2);
            const float rsA = rsqrtf(vrA * (1.f / QN_D) + 1e-6f), rsB = rsqrtf(vrB * (1.f / QN_D) + 1e-6f);   // models.py:166
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int c = 8 * t + quad * 2;
                const float2 g = *reinterpret_cast<const float2 *>(net.ln2_g + c);
                const float2 be = *reinterpret_cast<const float2 *>(net.ln2_b + c);
                const float zA0 = o[t][0] * rsA * g.x + be.x, zA1 = o[t][1] * rsA * g.y + be.y;
                const float zB0 = o[t][2] * rsB * g.x + be.x, zB1 = o[t][3] * rsB * g.y + be.y;
                if (Z2out) {
                    if (iA < L) {
                        __half *zh = Z2out + (size_t)row * 2 * net.K + (size_t)iA * QN_D + c;
                        const __half2 h = __floats2half2_rn(zA0, zA1);
                        const float2 hf = __half22float2(h);
                        *reinterpret_cast<__half2 *>(zh) = h;
                        *reinterpret_cast<__half2 *>(zh + net.K) = __floats2half2_rn(zA0 - hf.x, zA1 - hf.y);
                    }
                    if (iB < L) {
                        __half *zh = Z2out + (size_t)row * 2 * net.K + (size_t)iB * QN_D + c;
                        const __half2 h = __floats2half2_rn(zB0, zB1);
                        const float2 hf = __half22float2(h);
                        *reinterpret_cast<__half2 *>(zh) = h;
                        *reinterpret_cast<__half2 *>(zh + net.K) = __floats2half2_rn(zB0 - hf.x, zB1 - hf.y);
                    }
                } else {
                    if (iA < L) *reinterpret_cast<float2 *>(Zout + (size_t)row * net.K + (size_t)iA * QN_D + c) = make_float2(zA0, zA1);
                    if (iB < L) *reinterpret_cast<float2 *>(Zout + (size_t)row * net.K + (size_t)iB * QN_D + c) = make_float2(zB0, zB1);
                }
            }
        }
    }
}

int encode_fa_smem_bytes(int L)
{
    const int LK = (L + 15) & ~15;
    return 4 * LK * FA_KP * 2 + L * 4 + ((L + 15) & ~15);
}

// ---------------------------------------------------------------------------------------------
// dense layers: C[M][N] = leaky_relu(A[M][K] . W[N][K]^T + b), fp32 SIMT tiles 128x128x16
// ---------------------------------------------------------------------------------------------
#define GM_BM 128
#define GM_BN 128
#define GM_BK 16

__global__ void __launch_bounds__(256)
dense_kernel(const float *__restrict__ A, const float *__restrict__ W, const float *__restrict__ bias,
             float *__restrict__ C, const int32_t *__restrict__ m_ptr, long long m_fixed, int N, int K, int leaky)
{
    __shared__ __align__(16) float As[2][GM_BK][GM_BM + 4];
    __shared__ __align__(16) float Bs[2][GM_BK][GM_BN + 4];
    const long long M = m_ptr ? (long long)*m_ptr : m_fixed;
    const long long m0 = (long long)blockIdx.y * GM_BM;
    const int n0 = blockIdx.x * GM_BN;
    if (m0 >= M) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // loader mapping: each thread moves 2 float4 of A and 2 of W per k-tile
    const int lrow = tid >> 2;            // 0..63 (+64)
    const int lk = (tid & 3) * 4;         // 0,4,8,12
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long r = m0 + lrow + 64 * h;
            const int n = n0 + lrow + 64 * h;
            const int k = k0 + lk;
            ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K) {
                if (r < M) ra[h] = *reinterpret_cast<const float4 *>(A + (size_t)r * K + k);
                if (n < N) rb[h] = __ldg(reinterpret_cast<const float4 *>(W + (size_t)n * K + k));
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + 64 * h;
            As[buf][lk + 0][r] = ra[h].x;
            As[buf][lk + 1][r] = ra[h].y;
            As[buf][lk + 2][r] = ra[h].z;
            As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rb[h].x;
            Bs[buf][lk + 1][r] = rb[h].y;
            Bs[buf][lk + 2][r] = rb[h].z;
            Bs[buf][lk + 3][r] = rb[h].w;
        }
    };
    const int nk = (K + GM_BK - 1) / GM_BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * GM_BK);
#pragma unroll
        for (int k = 0; k < GM_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4 + 64]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4 + 64]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long r = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 4 + (j & 3) + 64 * (j >> 2);
            if (n >= N) continue;
            float v = acc[i][j] + bias[n];
            if (leaky) v = v > 0.f ? v : 0.01f * v;            // nn.LeakyReLU default slope, dqn.py:58,60
            C[(size_t)r * N + n] = v;
        }
    }
}

// Environment.step (env.py:246-259) for episode e with the greedy action; one thread
__device__ __forceinline__ void env_step(const gad_qnet &net, int e, int action, int32_t *next_active, int32_t *next_count)
{
    const int rows = net.rows, cols = net.cols;
    uint8_t *hd = net.heads + (size_t)e * rows;
    for (int r = 0; r < rows; ++r)
        if (!((action >> r) & 1) && hd[r] >= cols) return;    // illegal action: nothing changes, episode ends
    const int t = net.steps[e];
    const int cap = rows * cols;
    const uint8_t *sub = net.sub + (size_t)e * rows * cols;
    uint8_t *al = net.aligned + (size_t)e * rows * cap;
    int left = 0;
    for (int r = 0; r < rows; ++r) {
        uint8_t v = GAD_GAP;
        if (!((action >> r) & 1)) v = sub[r * cols + hd[r]++];
        al[(size_t)r * cap + t] = v;
        left += cols - hd[r];
    }
    net.steps[e] = t + 1;
    if (left > 0) next_active[atomicAdd(next_count, 1)] = e;   // done == 1: the episode continues
}

// ---------------------------------------------------------------------------------------------
// l3 + tanh * max_value + greedy argmax (+ Environment.step when episodes are running)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_step_kernel(const gad_qnet net, const float *__restrict__ H2, const int32_t *__restrict__ n_rows_ptr,
                 long long n_rows_fixed, float max_value, float *__restrict__ Qout, int32_t *__restrict__ act_out,
                 const int32_t *__restrict__ active, int32_t *__restrict__ next_active,
                 int32_t *__restrict__ next_count)
{
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && active && net.d_stats) atomicAdd(net.d_stats, (unsigned long long)n_rows);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int h2 = net.h2, A = net.A;
    const float *h = H2 + (size_t)row * h2;
    float best = 0.f;
    int best_a = -1;
    for (int a = 0; a < A; ++a) {
        const float *w = net.W3 + (size_t)a * h2;
        float s = 0.f;
        for (int k = lane; k < h2; k += 32) s = fmaf(h[k], __ldg(w + k), s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        const float q = tanhf(s + net.b3[a]) * max_value;      // dqn.py:76-77
        if (Qout && lane == 0) Qout[(size_t)row * A + a] = q;
        if (best_a < 0 || q > best) {                          // torch.argmax: first maximum
            best = q;
            best_a = a;
        }
    }
    if (lane != 0) return;
    if (act_out) act_out[row] = best_a;
    if (!active) return;
    env_step(net, active[row], best_a, next_active, next_count);
}

// tanh * max_value + greedy argmax (+ Environment.step) from the l3 pre-activations of the tensor-core chain
__global__ void __launch_bounds__(256)
head_from_logits_kernel(const gad_qnet net, const float *__restrict__ Zpre, int ldz, const int32_t *__restrict__ n_rows_ptr,
                        long long n_rows_fixed, float max_value, float *__restrict__ Qout,
                        int32_t *__restrict__ act_out, const int32_t *__restrict__ active,
                        int32_t *__restrict__ next_active, int32_t *__restrict__ next_count)
{
    const long long n_rows = n_rows_ptr ? (long long)*n_rows_ptr : n_rows_fixed;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && active && net.d_stats) atomicAdd(net.d_stats, (unsigned long long)n_rows);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int A = net.A;
    float best = -INFINITY;
    int best_a = 0x7fffffff;
    for (int a = lane; a < A; a += 32) {
        const float q = tanhf(Zpre[(size_t)row * ldz + a]) * max_value;      // dqn.py:76-77 (bias added by the GEMM)
        if (Qout) Qout[(size_t)row * A + a] = q;
        if (q > best) {                                       // ascending a per lane: keeps the first maximum
            best = q;
            best_a = a;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {                        // torch.argmax: first maximum
        const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const int oa = __shfl_xor_sync(0xFFFFFFFFu, best_a, o);
        if (ob > best || (ob == best && oa < best_a)) {
            best = ob;
            best_a = oa;
        }
    }
    if (lane != 0) return;
    if (best_a == 0x7fffffff) best_a = 0;                     // all NaN: torch returns the first index
    if (act_out) act_out[row] = best_a;
    if (!active) return;
    env_step(net, active[row], best_a, next_active, next_count);
}

// ---------------------------------------------------------------------------------------------
// episodes: set-up and splice
// ---------------------------------------------------------------------------------------------
__global__ void episode_init_kernel(const gad_qnet net, const uint8_t *__restrict__ boards, int R, int stride,
                                    const int32_t *__restrict__ idx, const int32_t *__restrict__ tile, long long M,
                                    int32_t *__restrict__ list, int32_t *__restrict__ counts)
{
    const long long m = blockIdx.x;
    const int rows = net.rows, cols = net.cols;
    const long long p = idx ? idx[m] : m;
    const int fr = tile[2 * m], fc = tile[2 * m + 1];
    if (fr < 0 || fc < 0) {                                // no tile fits this board (gad_worst_tiles): not mutated
        if (threadIdx.x == 0) net.steps[m] = -1;
        return;
    }
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        net.sub[(size_t)m * rows * cols + i] = boards[((size_t)p * R + fr + r) * stride + fc + c];
    }
    if (threadIdx.x < rows) net.heads[(size_t)m * rows + threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        net.steps[m] = 0;
        if (list) list[atomicAdd(&counts[0], 1)] = (int32_t)m;      // order inside the list is irrelevant: episodes are independent
    }
}

// Duplicate-window elimination. The eval-mode network is a pure function of the window, so two episodes
// that start from byte-identical R' x C' windows take the same actions and end with the same aligned rows
// (GA.py:353-366 only ever looks at the window). One warp per episode: a Zobrist hash of the window picks a
// slot of an open-addressing table of episode numbers; the first episode to claim a slot becomes the
// representative that actually runs, later ones compare their bytes with it (a hash collision just probes on)
// and record rep[m]. Which of several identical episodes becomes the representative depends on timing, the
// result does not.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
dedup_kernel(const uint8_t *__restrict__ windows, const int32_t *__restrict__ steps, int rc, long long M,
             int32_t *__restrict__ table, unsigned int mask, int32_t *__restrict__ rep, int32_t *__restrict__ uniq,
             int32_t *__restrict__ n_uniq, int enable)
{
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const int lane = threadIdx.x & 31;
    if (steps && steps[m] < 0) {                          // no tile fits: not mutated, nothing to share
        if (lane == 0) rep[m] = (int32_t)m;
        return;
    }
    const uint8_t *mine = windows + (size_t)m * rc;
    if (!enable) {
        if (lane == 0) {
            rep[m] = (int32_t)m;
            if (uniq) uniq[atomicAdd(n_uniq, 1)] = (int32_t)m;
        }
        return;
    }
    unsigned long long h = 0;
    for (int i = lane; i < rc; i += 32) h ^= mix64(((unsigned long long)i << 8) | mine[i]);
    for (int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(0xFFFFFFFFu, h, o);
    unsigned int slot = (unsigned int)(h ^ (h >> 32)) & mask;
    for (;;) {
        int prev = 0;
        if (lane == 0) prev = atomicCAS(&table[slot], -1, (int32_t)m);
        prev = __shfl_sync(0xFFFFFFFFu, prev, 0);
        if (prev == -1) {                                 // first of its kind: this episode runs
            if (lane == 0) {
                rep[m] = (int32_t)m;
                if (uniq) uniq[atomicAdd(n_uniq, 1)] = (int32_t)m;
            }
            return;
        }
        const uint8_t *other = windows + (size_t)prev * rc;
        bool same = true;
        for (int i = lane; i < rc; i += 32) same &= mine[i] == other[i];
        if (__all_sync(0xFFFFFFFFu, same)) {
            if (lane == 0) rep[m] = prev;
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__global__ void set_i32_kernel(int32_t *dst, int32_t v0, int32_t v1)
{
    dst[0] = v0;
    dst[1] = v1;
}

// rep[m] -> the SMALLEST index among the windows identical to m (dedup_kernel's representative depends on timing;
// ranks of a sharded run must agree on it): min over each group, then relabel.
__global__ void rep_min_kernel(const int32_t *__restrict__ rep, int32_t *__restrict__ lowest, long long M)
{
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) atomicMin(&lowest[rep[m]], (int32_t)m);
}
__global__ void rep_relabel_kernel(int32_t *__restrict__ rep, const int32_t *__restrict__ lowest, long long M)
{
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) rep[m] = lowest[rep[m]];
}

// episode e starts from window list[e] (or e) of an explicit window array
__global__ void episodes_from_windows_kernel(const gad_qnet net, const uint8_t *__restrict__ windows,
                                             const int32_t *__restrict__ list, long long U)
{
    const long long e = blockIdx.x;
    const int rc = net.rows * net.cols;
    const long long src = list ? list[e] : e;
    for (int i = threadIdx.x; i < rc; i += blockDim.x) net.sub[(size_t)e * rc + i] = windows[(size_t)src * rc + i];
    if (threadIdx.x < net.rows) net.heads[(size_t)e * net.rows + threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        net.steps[e] = 0;
        net.uniq[e] = (int32_t)e;
    }
}

// Environment.padding (env.py:344-351) + the splice of GA.py:369-373, one warp per episode.
// The tile's rows become old[:fc] + aligned + leftovers + gaps + old[fc+cols:], i.e. they grow by
// d = steps + longest leftover - cols >= 0; the other rows keep their cells.
// res_of[m] = the episode whose result individual m takes (aligned / heads / steps are indexed by it). m's own
// window (identical content to that episode's) is windows[m] when own_windows is set, else windows[res_of[m]].
__global__ void __launch_bounds__(128)
splice_kernel(int rows, int cols, const int32_t *__restrict__ res_of, int own_windows,
              const uint8_t *__restrict__ aligned, const uint8_t *__restrict__ heads, const int32_t *__restrict__ steps,
              const uint8_t *__restrict__ windows, uint8_t *__restrict__ boards, int32_t *__restrict__ rowlen, int R,
              int stride, const int32_t *__restrict__ idx, const int32_t *__restrict__ tile, long long M,
              int32_t *__restrict__ status)
{
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= M) return;
    const int lane = threadIdx.x & 31;
    const int cap = rows * cols;
    const long long p = idx ? idx[m] : m;
    const int fr = tile[2 * m], fc = tile[2 * m + 1];
    const long long e = res_of[m];                  // the episode that ran for this window (m itself if unique)
    const long long wm = own_windows ? m : e;
    const uint8_t *hd = heads + (size_t)e * rows;
    const int T = steps[e];
    if (T < 0) return;
    int longest = 0;
    for (int r = 0; r < rows; ++r) longest = max(longest, cols - (int)hd[r]);
    const int newlen = T + longest;                 // length of every aligned row after padding
    const int d = newlen - cols;
    int wmax = 0;
    for (int r = 0; r < R; ++r) {
        const int l = rowlen[p * R + r];
        wmax = max(wmax, (r >= fr && r < fr + rows) ? l + d : l);
    }
    if (wmax > stride) {                            // cannot happen when the caller sized the stride (gadpamsa.h)
        if (lane == 0) atomicOr(status, 1);
        return;
    }
    __syncwarp();
    for (int r = 0; r < rows; ++r) {
        uint8_t *row = boards + ((size_t)p * R + fr + r) * stride;
        const int len = rowlen[p * R + fr + r];
        __syncwarp();
        const int tail0 = fc + cols;                // first cell after the window
        // move the tail right by d, highest addresses first so the move is safe in place
        if (d > 0) {
            for (int hi = len; hi > tail0; hi -= 32) {
                const int c = hi - 1 - lane;
                uint8_t v = 0;
                if (c >= tail0) v = row[c];
                __syncwarp();
                if (c >= tail0) row[c + d] = v;
                __syncwarp();
            }
        }
        const uint8_t *al = aligned + ((size_t)e * rows + r) * cap;
        const uint8_t *sub = windows + ((size_t)wm * rows + r) * cols;
        const int h = hd[r];
        for (int c = lane; c < newlen; c += 32) {
            uint8_t v;
            if (c < T) v = al[c];
            else if (c - T < cols - h) v = sub[h + c - T];      // leftovers (env.py:346-347)
            else v = GAD_GAP;                                   // right padding (env.py:349-350)
            row[fc + c] = v;
        }
        if (lane == 0) rowlen[p * R + fr + r] = len + d;
    }
    // layout invariant: every row holds the gap code from its own length up to roundup4(board width)
    __syncwarp();
    const int pad_to = min(stride, (wmax + 3) & ~3);
    for (int r = 0; r < R; ++r) {
        uint8_t *row = boards + ((size_t)p * R + r) * stride;
        const int len = rowlen[p * R + r];
        for (int c = len + lane; c < pad_to; c += 32) row[c] = GAD_GAP;
    }
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
int dev_alloc(T **p, size_t n)
{
    GAD_CUDA(cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)));
    return GAD_OK;
}

int upload(float **dst, const float *src, size_t n)
{
    int rc = dev_alloc(dst, n);
    if (rc) return rc;
    GAD_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return GAD_OK;
}

int ensure_states(gad_qnet *q, long long B)
{
    if (B <= q->cap) return GAD_OK;
    cudaFree(q->tok); cudaFree(q->Z); cudaFree(q->Z2); cudaFree(q->H1); cudaFree(q->H2); cudaFree(q->Q); cudaFree(q->act);
    cudaFree(q->H1s); cudaFree(q->H2s); cudaFree(q->Qpre);
    q->tok = nullptr; q->Z = q->H1 = q->H2 = q->Q = nullptr; q->act = nullptr; q->Z2 = nullptr;
    q->H1s = q->H2s = nullptr; q->Qpre = nullptr;
    q->cap = 0;
    int rc;
    if ((rc = dev_alloc(&q->tok, (size_t)B * q->L))) return rc;
    if (q->tc1) {
        const long long rows = (B + 127) / 128 * 128;      // whole TMA boxes stay inside the allocation
        if ((rc = dev_alloc(&q->Z2, (size_t)rows * 2 * q->K))) return rc;
        GAD_CUDA(cudaMemset(q->Z2, 0, (size_t)rows * 2 * q->K * sizeof(__half)));
        if ((rc = tc_layer_bind_a(q->tc1, q->Z2, rows))) return rc;
        const size_t n1 = (size_t)rows * 2 * q->tc2->Kpad, n2 = (size_t)rows * 2 * q->tc3->Kpad;
        if ((rc = dev_alloc(&q->H1s, n1))) return rc;
        if ((rc = dev_alloc(&q->H2s, n2))) return rc;
        if ((rc = dev_alloc(&q->Qpre, (size_t)rows * TC_BN_L3))) return rc;
        GAD_CUDA(cudaMemset(q->H1s, 0, n1 * sizeof(__half)));          // padding columns stay zero forever
        GAD_CUDA(cudaMemset(q->H2s, 0, n2 * sizeof(__half)));
        if ((rc = tc_layer_bind_a(q->tc2, q->H1s, rows))) return rc;
        if ((rc = tc_layer_bind_a(q->tc3, q->H2s, rows))) return rc;
    } else if ((rc = dev_alloc(&q->Z, (size_t)B * q->K))) return rc;
    if (!q->tc1) {
        if ((rc = dev_alloc(&q->H1, (size_t)B * q->h1))) return rc;
        if ((rc = dev_alloc(&q->H2, (size_t)B * q->h2))) return rc;
    }
    if ((rc = dev_alloc(&q->Q, (size_t)B * q->A))) return rc;
    if ((rc = dev_alloc(&q->act, (size_t)B))) return rc;
    q->cap = B;
    return GAD_OK;
}

int ensure_episodes(gad_qnet *q, long long M)
{
    if (M <= q->ecap) return GAD_OK;
    cudaFree(q->sub); cudaFree(q->heads); cudaFree(q->aligned); cudaFree(q->steps); cudaFree(q->list0); cudaFree(q->list1);
    cudaFree(q->rep); cudaFree(q->uniq); cudaFree(q->table);
    q->sub = q->heads = q->aligned = nullptr; q->steps = q->list0 = q->list1 = nullptr;
    q->rep = q->uniq = q->table = nullptr;
    q->ecap = 0;
    q->tcap = 0;
    const size_t rc_ = (size_t)q->rows * q->cols;
    int rc;
    if ((rc = dev_alloc(&q->sub, (size_t)M * rc_))) return rc;
    if ((rc = dev_alloc(&q->heads, (size_t)M * q->rows))) return rc;
    if ((rc = dev_alloc(&q->aligned, (size_t)M * q->rows * rc_))) return rc;
    if ((rc = dev_alloc(&q->steps, (size_t)M))) return rc;
    if ((rc = dev_alloc(&q->list0, (size_t)M))) return rc;
    if ((rc = dev_alloc(&q->list1, (size_t)M))) return rc;
    if ((rc = dev_alloc(&q->rep, (size_t)M))) return rc;
    if ((rc = dev_alloc(&q->uniq, (size_t)M))) return rc;
    long long slots = 1024;
    while (slots < 2 * M) slots <<= 1;                     // load factor <= 0.5
    if ((rc = dev_alloc(&q->table, (size_t)slots))) return rc;
    q->tcap = slots;
    q->ecap = M;
    return GAD_OK;
}

int encode_smem_bytes(int L, int warps) { return warps * L * ENC_QT * (int)sizeof(float) + ((L + 15) & ~15); }

int encode_tiled_smem_bytes(int L, int RB)
{
    return (L * QN_D + L * (RB + 2) + RB + 2 * L) * (int)sizeof(float) + ((L + 15) & ~15);
}

template <int RB>
int launch_encode_tiled(gad_qnet *q, const uint8_t *tok_in, const int32_t *active, const int32_t *n_ptr,
                        long long n_max, cudaStream_t st)
{
    const int bytes = encode_tiled_smem_bytes(q->L, RB);
    if (bytes > 48 * 1024)
        GAD_CUDA(cudaFuncSetAttribute(encode_tiled_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    long long per_sm = (220 * 1024) / bytes;
    const long long by_threads = 2048 / EncTile<RB>::THREADS;
    if (per_sm > by_threads) per_sm = by_threads;
    const long long grid = n_max < (long long)GAD_NUM_SMS * per_sm ? n_max : (long long)GAD_NUM_SMS * per_sm;
    encode_tiled_kernel<RB><<<(unsigned)grid, EncTile<RB>::THREADS, bytes, st>>>(*q, tok_in, active, n_ptr, n_max,
                                                                                 q->Z, q->Z2);
    return GAD_OK;
}

// one forward over `n` rows (device count in n_ptr, or n_fixed): tokens/episodes -> H2, then the head kernel
int launch_forward(gad_qnet *q, const uint8_t *tok_in, const int32_t *active, const int32_t *n_ptr, long long n_max,
                   float max_value, float *Qout, int32_t *act_out, int32_t *next_active, int32_t *next_count,
                   cudaStream_t st)
{
    // GAD_QNET_ENCODER (A/B testing): unset or "tc" = attention on tcgen05 (attn_tc.cu; windows of 16..192 positions),
    // "fa" = flash-attention style kernel on mma.sync (also what larger or tiny windows get), "mma" = table logits +
    // mma.sync product, "simt" = table logits + fp32 register tiles
    const char *enc = getenv("GAD_QNET_ENCODER");
    if ((!enc || !enc[0] || enc[0] == 't') && attn_tc_supported(q)) {
        int rc = launch_encode_tc(q, tok_in, active, n_ptr, n_max, st);
        if (rc) return rc;
    } else if (encode_fa_smem_bytes(q->L) <= 110 * 1024 && !(enc && (enc[0] == 'm' || enc[0] == 's'))) {
        const int bytes = encode_fa_smem_bytes(q->L);
        if (bytes > 48 * 1024)                             // per device and cheap: set on every launch, like the other encoders
            GAD_CUDA(cudaFuncSetAttribute(encode_fa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        long long per_sm = (224 * 1024) / (bytes + 1024);
        if (per_sm > 4) per_sm = 4;
        const long long grid = n_max < (long long)GAD_NUM_SMS * per_sm ? n_max : (long long)GAD_NUM_SMS * per_sm;
        encode_fa_kernel<<<(unsigned)grid, FA_WARPS * 32, bytes, st>>>(*q, tok_in, active, n_ptr, n_max, q->Z, q->Z2);
    } else if (q->L <= 192 && !(enc && enc[0] == 's')) {
        const int bytes = encode_mma_smem_bytes(q->L);
        if (bytes > 48 * 1024)
            GAD_CUDA(cudaFuncSetAttribute(encode_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        const long long grid = n_max < (long long)GAD_NUM_SMS * 2 ? n_max : (long long)GAD_NUM_SMS * 2;
        encode_mma_kernel<<<(unsigned)grid, 256, bytes, st>>>(*q, tok_in, active, n_ptr, n_max, q->Z, q->Z2);
    } else if (q->L <= 96 && encode_tiled_smem_bytes(q->L, 96) <= 220 * 1024) {
        int rc = launch_encode_tiled<96>(q, tok_in, active, n_ptr, n_max, st);
        if (rc) return rc;
    } else if (q->L <= 192 && encode_tiled_smem_bytes(q->L, 192) <= 220 * 1024) {
        int rc = launch_encode_tiled<192>(q, tok_in, active, n_ptr, n_max, st);
        if (rc) return rc;
    } else {                                             // very large agent windows: warp-per-row fallback
        int warps = 8;
        while (warps > 1 && encode_smem_bytes(q->L, warps) > 200 * 1024) warps >>= 1;
        const int smem = encode_smem_bytes(q->L, warps);
        GAD_REQUIRE(smem <= 227 * 1024, "agent window too large for the encoder kernel (L=%d)", q->L);
        if (smem > 48 * 1024)
            GAD_CUDA(cudaFuncSetAttribute(encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        long long grid = n_max < (long long)GAD_NUM_SMS * 16 ? n_max : (long long)GAD_NUM_SMS * 16;
        encode_kernel<<<(unsigned)grid, warps * 32, smem, st>>>(*q, tok_in, active, n_ptr, n_max, q->Z, q->Z2);
    }
    GAD_LAUNCH_CHECK();
    if (q->tc1) {
        // l1 -> l2 -> l3 on tcgen05; activations travel as (hi | lo) fp16, the last layer emits fp32 pre-activations
        int rc = tc_layer_run(q->tc1, q->b1, q->H1s, 2 * q->tc2->Kpad, q->tc2->Kpad, 1, n_ptr, n_max, 1, st);
        if (!rc) rc = tc_layer_run(q->tc2, q->b2, q->H2s, 2 * q->tc3->Kpad, q->tc3->Kpad, 1, n_ptr, n_max, 1, st);
        if (!rc) rc = tc_layer_run(q->tc3, q->b3, q->Qpre, TC_BN_L3, 0, 0, n_ptr, n_max, 0, st);
        if (rc) return rc;
        head_from_logits_kernel<<<(unsigned)((n_max + 7) / 8), 256, 0, st>>>(*q, q->Qpre, TC_BN_L3, n_ptr, n_max, max_value,
                                                                             Qout, act_out, active, next_active, next_count);
    } else {
        const unsigned mt = (unsigned)((n_max + GM_BM - 1) / GM_BM);
        dense_kernel<<<dim3((q->h1 + GM_BN - 1) / GM_BN, mt), 256, 0, st>>>(q->Z, q->W1, q->b1, q->H1, n_ptr, n_max,
                                                                            q->h1, q->K, 1);
        dense_kernel<<<dim3((q->h2 + GM_BN - 1) / GM_BN, mt), 256, 0, st>>>(q->H1, q->W2, q->b2, q->H2, n_ptr, n_max,
                                                                            q->h2, q->h1, 1);
        head_step_kernel<<<(unsigned)((n_max + 7) / 8), 256, 0, st>>>(*q, q->H2, n_ptr, n_max, max_value, Qout, act_out,
                                                                      active, next_active, next_count);
    }
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

void layer_norm64(const double *x, const float *g, const float *b, double *y)
{
    double mu = 0, var = 0;
    for (int c = 0; c < QN_D; ++c) mu += x[c];
    mu /= QN_D;
    for (int c = 0; c < QN_D; ++c) var += (x[c] - mu) * (x[c] - mu);
    var /= QN_D;
    const double rstd = 1.0 / sqrt(var + 1e-6);
    for (int c = 0; c < QN_D; ++c) y[c] = (x[c] - mu) * rstd * g[c] + b[c];
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
// h_weights: the 16 tensors of the reference state_dict in this order (DPAMSA/dqn.py:187-199):
//  0 encoder.src_word_emb.weight (6,64)       1 encoder.position_enc.pos_table (1,L,64)
//  2 encoder.layer_norm.weight (64)            3 encoder.layer_norm.bias (64)
//  4 w_qs.weight (d_k,64)   5 w_ks.weight (d_k,64)   6 w_vs.weight (d_v,64)   7 fc.weight (64,d_v)
//  8 self_attention.layer_norm.weight (64)     9 self_attention.layer_norm.bias (64)
// 10 l1.weight (h1, L*64)  11 l1.bias  12 l2.weight (h2,h1)  13 l2.bias  14 l3.weight (A,h2)  15 l3.bias
extern "C" int gad_qnet_create(gad_qnet **out, int rows, int cols, int d_k, int d_v, int h1, int h2,
                               const float *const *h_weights)
{
    GAD_REQUIRE(out && h_weights, "gad_qnet_create: null pointer");
    GAD_REQUIRE(rows >= 1 && rows <= 16 && cols >= 1 && cols <= 254, "gad_qnet_create: window %dx%d unsupported", rows, cols);
    GAD_REQUIRE(d_k >= 1 && d_v >= 1 && h1 >= 4 && h1 % 4 == 0 && h2 >= 4 && h2 % 4 == 0,
                "gad_qnet_create: bad layer sizes (h1 and h2 must be multiples of 4)");
    for (int i = 0; i < 16; ++i) GAD_REQUIRE(h_weights[i], "gad_qnet_create: weight tensor %d is null", i);
    gad_qnet *q = new gad_qnet();
    q->rows = rows;
    q->cols = cols;
    q->L = rows * (cols + 1);
    q->A = (1 << rows) - 1;
    q->K = q->L * QN_D;
    q->h1 = h1;
    q->h2 = h2;
    const int L = q->L, V = QN_VOCAB;
    const float *emb = h_weights[0], *pos = h_weights[1], *g1 = h_weights[2], *be1 = h_weights[3];
    const float *wq = h_weights[4], *wk = h_weights[5], *wv = h_weights[6], *wfc = h_weights[7];

    // fold (fp64): X, M = Wq^T Wk / sqrt(dk), N = Wfc Wv
    std::vector<double> Xd((size_t)V * L * QN_D), Mq((size_t)QN_D * QN_D), Nf((size_t)QN_D * QN_D);
    for (int t = 0; t < V; ++t)
        for (int p = 0; p < L; ++p) {
            double e[QN_D];
            for (int c = 0; c < QN_D; ++c) e[c] = (double)(float)(emb[t * QN_D + c] + pos[(size_t)p * QN_D + c]);
            layer_norm64(e, g1, be1, &Xd[((size_t)t * L + p) * QN_D]);
        }
    const double temp = sqrt((double)d_k);
    for (int a = 0; a < QN_D; ++a)
        for (int b = 0; b < QN_D; ++b) {
            double s = 0;
            for (int d = 0; d < d_k; ++d) s += (double)wq[d * QN_D + a] * (double)wk[d * QN_D + b];
            Mq[(size_t)a * QN_D + b] = s / temp;
            double n = 0;
            for (int d = 0; d < d_v; ++d) n += (double)wfc[(size_t)a * d_v + d] * (double)wv[d * QN_D + b];
            Nf[(size_t)a * QN_D + b] = n;
        }
    const size_t n = (size_t)V * L;
    std::vector<double> QM(n * QN_D);
    std::vector<float> Xf(n * QN_D), VFf(n * QN_D), Tf(n * n);
    for (size_t i = 0; i < n; ++i) {
        const double *x = &Xd[i * QN_D];
        for (int b = 0; b < QN_D; ++b) {
            double s = 0, v = 0;
            for (int a = 0; a < QN_D; ++a) s += x[a] * Mq[(size_t)a * QN_D + b];
            for (int a = 0; a < QN_D; ++a) v += Nf[(size_t)b * QN_D + a] * x[a];
            QM[i * QN_D + b] = s;
            VFf[i * QN_D + b] = (float)v;
            Xf[i * QN_D + b] = (float)x[b];
        }
    }
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j) {
            double s = 0;
            const double *a = &QM[i * QN_D], *b = &Xd[j * QN_D];
            for (int c = 0; c < QN_D; ++c) s += a[c] * b[c];
            Tf[i * n + j] = (float)s;
        }
    int rc = GAD_OK;
    auto up = [&](float **dst, const float *src, size_t cnt) { if (!rc) rc = upload(dst, src, cnt); };
    up(&q->X, Xf.data(), Xf.size());
    up(&q->T, Tf.data(), Tf.size());
    up(&q->VF, VFf.data(), VFf.size());
    {
        float mx = 0.f;
        for (float v : VFf) mx = fmaxf(mx, fabsf(v));
        int sv = mx > 0.f ? (int)floorf(log2f(16384.f / mx)) : 0;
        sv = sv > 40 ? 40 : (sv < -40 ? -40 : sv);
        q->vf_inv_scale = ldexpf(1.f, -sv);
        std::vector<__half> vh(VFf.size()), vl(VFf.size());
        for (size_t i = 0; i < VFf.size(); ++i) {
            const float v = ldexpf(VFf[i], sv);
            vh[i] = __float2half_rn(v);
            vl[i] = __float2half_rn(v - __half2float(vh[i]));
        }
        if (!rc) rc = dev_alloc(&q->VFh, vh.size());
        if (!rc) rc = dev_alloc(&q->VFl, vl.size());
        if (!rc && (cudaMemcpy(q->VFh, vh.data(), vh.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(q->VFl, vl.data(), vl.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)) {
            gad_set_error("gad_qnet_create: upload of the split VF tables failed");
            rc = GAD_ECUDA;
        }
    }
    {
        auto split_upload = [&](const std::vector<double> &src, double mul, __half **dh, __half **dl, int *shift) {
            double mx = 0;
            for (double v : src) mx = fmax(mx, fabs(v * mul));
            int sh = mx > 0 ? (int)floor(log2(16384.0 / mx)) : 0;
            sh = sh > 40 ? 40 : (sh < -40 ? -40 : sh);
            *shift = sh;
            std::vector<__half> vh(src.size()), vl(src.size());
            for (size_t i = 0; i < src.size(); ++i) {
                const float v = (float)ldexp(src[i] * mul, sh);
                vh[i] = __float2half_rn(v);
                vl[i] = __float2half_rn(v - __half2float(vh[i]));
            }
            if (!rc) rc = dev_alloc(dh, vh.size());
            if (!rc) rc = dev_alloc(dl, vl.size());
            if (!rc && (cudaMemcpy(*dh, vh.data(), vh.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess ||
                        cudaMemcpy(*dl, vl.data(), vl.size() * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess)) {
                gad_set_error("gad_qnet_create: upload of a split table failed");
                rc = GAD_ECUDA;
            }
        };
        int sq = 0, sk = 0;
        split_upload(QM, 1.4426950408889634, &q->Qh, &q->Ql, &sq);      // logits in the log2 domain: exp -> ex2
        split_upload(Xd, 1.0, &q->Kh, &q->Kl, &sk);
        q->s_inv_scale = ldexpf(1.f, -(sq + sk));
    }
    up(&q->ln2_g, h_weights[8], QN_D);
    up(&q->ln2_b, h_weights[9], QN_D);
    up(&q->W1, h_weights[10], (size_t)h1 * q->K);
    up(&q->b1, h_weights[11], h1);
    up(&q->W2, h_weights[12], (size_t)h2 * h1);
    up(&q->b2, h_weights[13], h2);
    up(&q->W3, h_weights[14], (size_t)q->A * h2);
    up(&q->b3, h_weights[15], q->A);
    const char *simt = getenv("GAD_QNET_SIMT");
    // dense layers on tcgen05 (default); GAD_QNET_SIMT=1, or more than 64 actions, keeps the fp32 SIMT tiles
    if (!rc && !(simt && simt[0] == '1') && q->A <= TC_BN_L3) {
        q->tc1 = new TcLayer();
        q->tc2 = new TcLayer();
        q->tc3 = new TcLayer();
        const char *unf = getenv("GAD_L1_UNFUSED");        // A/B switch: 1 = three separate (hi.hi, lo.hi, hi.lo) k-loops
        rc = tc_layer_init(q->tc1, q->W1, h_weights[10], h1, q->K, TC_BN_L1, (unf && unf[0] == '1') ? 0 : 1);
        if (!rc) rc = tc_layer_init(q->tc2, q->W2, h_weights[12], h2, h1, TC_BN_L2, 0);
        if (!rc) rc = tc_layer_init(q->tc3, q->W3, h_weights[14], q->A, h2, TC_BN_L3, 0);
    }
    if (!rc) rc = dev_alloc(&q->counts, 2);
    if (!rc) rc = dev_alloc(&q->d_stats, 2);
    if (!rc && cudaMallocHost((void **)&q->h_count, 2 * sizeof(int32_t)) != cudaSuccess) {
        gad_set_error("gad_qnet_create: cudaMallocHost failed");
        rc = GAD_ECUDA;
    }
    if (rc) {
        gad_qnet_destroy(q);                               // frees whatever was allocated before the failure
        return rc;
    }
    *out = q;
    return GAD_OK;
}

extern "C" int gad_qnet_destroy(gad_qnet *q)
{
    if (!q) return GAD_OK;
    float *f[] = {q->X, q->T, q->VF, q->ln2_g, q->ln2_b, q->W1, q->b1, q->W2, q->b2, q->W3, q->b3, q->Z, q->H1, q->H2, q->Q};
    for (float *p : f) cudaFree(p);
    cudaFree(q->Z2);
    cudaFree(q->VFh);
    cudaFree(q->VFl);
    cudaFree(q->Qh); cudaFree(q->Ql); cudaFree(q->Kh); cudaFree(q->Kl);
    cudaFree(q->hs_states); cudaFree(q->hs_q); cudaFree(q->hs_act);
    for (TcLayer *t : {q->tc1, q->tc2, q->tc3})
        if (t) {
            tc_layer_free(t);
            delete t;
        }
    cudaFree(q->H1s);
    cudaFree(q->H2s);
    cudaFree(q->Qpre);
    cudaFree(q->tok); cudaFree(q->act); cudaFree(q->sub); cudaFree(q->heads); cudaFree(q->aligned);
    cudaFree(q->steps); cudaFree(q->list0); cudaFree(q->list1); cudaFree(q->counts);
    cudaFree(q->rep); cudaFree(q->uniq); cudaFree(q->table); cudaFree(q->d_stats);
    if (q->h_count) cudaFreeHost(q->h_count);
    delete q;
    return GAD_OK;
}

extern "C" int gad_qnet_dims(const gad_qnet *q, int *L, int *A)
{
    GAD_REQUIRE(q && L && A, "gad_qnet_dims: null pointer");
    *L = q->L;
    *A = q->A;
    return GAD_OK;
}

// Net.forward + argmax for B explicit states (uint8 tokens [B][L], device). d_q [B][A] and d_action [B]
// are optional device outputs. Asynchronous on `stream`.
extern "C" int gad_qnet_forward(gad_qnet *q, const uint8_t *d_states, int64_t B, float max_value, float *d_q,
                                int32_t *d_action, void *stream)
{
    GAD_REQUIRE(q && d_states && B >= 0, "gad_qnet_forward: bad argument");
    cudaStream_t st = gad_stream(stream);
    for (int64_t b0 = 0; b0 < B; b0 += kMaxBatch) {
        const long long nb = B - b0 < kMaxBatch ? B - b0 : kMaxBatch;
        int rc = ensure_states(q, nb);
        if (rc) return rc;
        rc = launch_forward(q, d_states + (size_t)b0 * q->L, nullptr, nullptr, nb, max_value,
                            d_q ? d_q + (size_t)b0 * q->A : nullptr, d_action ? d_action + b0 : nullptr, nullptr,
                            nullptr, st);
        if (rc) return rc;
    }
    return GAD_OK;
}

// Same through host buffers (DQN.predict for one state from python). Synchronises.
extern "C" int gad_qnet_forward_host(gad_qnet *q, const uint8_t *h_states, int64_t B, float max_value, float *h_q,
                                     int32_t *h_action)
{
    GAD_REQUIRE(q && h_states && B >= 0, "gad_qnet_forward_host: bad argument");
    if (B == 0) return GAD_OK;
    int rc = GAD_OK;
    if (B > q->hcap) {
        cudaFree(q->hs_states); cudaFree(q->hs_q); cudaFree(q->hs_act);
        q->hs_states = nullptr; q->hs_q = nullptr; q->hs_act = nullptr;
        q->hcap = 0;
        rc = dev_alloc(&q->hs_states, (size_t)B * q->L);
        if (!rc) rc = dev_alloc(&q->hs_q, (size_t)B * q->A);
        if (!rc) rc = dev_alloc(&q->hs_act, (size_t)B);
        if (rc) return rc;
        q->hcap = B;
    }
    GAD_CUDA(cudaMemcpy(q->hs_states, h_states, (size_t)B * q->L, cudaMemcpyHostToDevice));
    if ((rc = gad_qnet_forward(q, q->hs_states, B, max_value, q->hs_q, q->hs_act, nullptr))) return rc;
    GAD_CUDA(cudaDeviceSynchronize());
    if (h_q) GAD_CUDA(cudaMemcpy(h_q, q->hs_q, (size_t)B * q->A * sizeof(float), cudaMemcpyDeviceToHost));
    if (h_action) GAD_CUDA(cudaMemcpy(h_action, q->hs_act, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return GAD_OK;
}

static int run_lockstep(gad_qnet *q, long long n_unique, float max_value, cudaStream_t st, int64_t *forwards_out);
static int finish_mutate(gad_qnet *q, long long n_unique, int64_t forwards, int64_t *h_forwards, cudaStream_t st);

// GA.mutation for M individuals (GA.py:335-373): d_idx[m] is the individual, d_tile[m] = {from_row,
// from_col} of its worst tile (gad_worst_tiles). Runs the greedy episodes in lock-step chunks, then
// pads and splices the aligned sub-boards back. The stride must leave room for the growth:
// max row length + (rows-1)*cols <= stride, otherwise bit 0 of *d_status is set and that board is
// left unchanged. h_steps_run (optional) receives the number of lock-step forwards. Synchronises.
extern "C" int gad_mutate(gad_qnet *q, uint8_t *d_boards, int32_t *d_rowlen, int R, int stride, const int32_t *d_idx,
                          const int32_t *d_tile, int64_t M, float max_value, int32_t *d_status,
                          int64_t *h_forwards, void *stream)
{
    GAD_REQUIRE(q && d_boards && d_rowlen && d_tile && d_status, "gad_mutate: null pointer");
    GAD_REQUIRE(M >= 0 && R >= q->rows && stride >= 16 && stride % 16 == 0, "gad_mutate: bad shape");
    cudaStream_t st = gad_stream(stream);
    int64_t forwards = 0;
    q->last_stats[0] = M; q->last_stats[1] = q->last_stats[2] = q->last_stats[3] = 0;
    if (M == 0) {
        if (h_forwards) *h_forwards = 0;
        return GAD_OK;
    }
    GAD_REQUIRE(M < (1ll << 30), "gad_mutate: too many episodes in one call (%lld)", (long long)M);
    // GAD_MUTATE_DEDUP=0 (A/B switch): every episode runs, identical windows included
    const char *dd = getenv("GAD_MUTATE_DEDUP");
    const int dedup = !(dd && dd[0] == '0');
    int rc = ensure_episodes(q, M);
    // activation buffers for the largest chunk that can occur (not for this call's distinct windows: their number
    // changes from generation to generation and re-allocating in between costs more than the memory)
    if (!rc) rc = ensure_states(q, M < kMaxBatch ? M : kMaxBatch);
    if (rc) return rc;
    long long slots = 1024;
    while (slots < 2 * M) slots <<= 1;
    GAD_CUDA(cudaMemsetAsync(q->counts, 0, 2 * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(q->d_stats, 0, 2 * sizeof(unsigned long long), st));
    if (dedup) GAD_CUDA(cudaMemsetAsync(q->table, 0xFF, (size_t)slots * sizeof(int32_t), st));
    // 1. every episode's window; 2. one representative per distinct window
    episode_init_kernel<<<(unsigned)M, 64, 0, st>>>(*q, d_boards, R, stride, d_idx, d_tile, M, nullptr, nullptr);
    GAD_LAUNCH_CHECK();
    dedup_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(q->sub, q->steps, q->rows * q->cols, M, q->table,
                                                            (unsigned int)(slots - 1), q->rep, q->uniq, q->counts, dedup);
    GAD_LAUNCH_CHECK();
    GAD_CUDA(cudaMemcpyAsync(q->h_count, q->counts, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GAD_CUDA(cudaStreamSynchronize(st));
    const long long n_unique = q->h_count[0];
    // 3. the representatives' greedy episodes in lock-step chunks
    if ((rc = run_lockstep(q, n_unique, max_value, st, &forwards))) return rc;
    // 4. every individual takes the aligned rows of its window's representative
    splice_kernel<<<(unsigned)((M + 3) / 4), 128, 0, st>>>(q->rows, q->cols, q->rep, 0, q->aligned, q->heads, q->steps,
                                                           q->sub, d_boards, d_rowlen, R, stride, d_idx, d_tile, M, d_status);
    GAD_LAUNCH_CHECK();
    return finish_mutate(q, n_unique, forwards, h_forwards, st);
}

// the greedy episodes of q->uniq[0 .. n_unique) in lock-step chunks (episode state in q->sub / heads / aligned / steps)
static int run_lockstep(gad_qnet *q, long long n_unique, float max_value, cudaStream_t st, int64_t *forwards_out)
{
    int rc;
    int64_t forwards = 0;
    for (long long u0 = 0; u0 < n_unique; u0 += kMaxBatch) {
        const long long nb = n_unique - u0 < kMaxBatch ? n_unique - u0 : kMaxBatch;
        if ((rc = ensure_states(q, nb))) return rc;
        int32_t *lists[2] = {q->uniq + u0, q->list1};       // the chunk's slice of `uniq` doubles as the first active list
        set_i32_kernel<<<1, 1, 0, st>>>(q->counts, (int32_t)nb, 0);
        GAD_LAUNCH_CHECK();
        long long n_max = nb;
        const int max_steps = q->rows * q->cols;            // every step consumes at least one token
        int cur = 0;
        for (int s = 0; s < max_steps && n_max > 0; ++s) {
            rc = launch_forward(q, nullptr, lists[cur], q->counts + cur, n_max, max_value, nullptr, nullptr, lists[cur ^ 1],
                                q->counts + (cur ^ 1), st);
            if (rc) return rc;
            GAD_CUDA(cudaMemsetAsync(q->counts + cur, 0, sizeof(int32_t), st));   // becomes the "next" counter of step s+1
            cur ^= 1;
            ++forwards;
            if ((s & 3) == 3 || s + 1 == max_steps) {       // poll the active count every 4 steps
                GAD_CUDA(cudaMemcpyAsync(q->h_count, q->counts + cur, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                GAD_CUDA(cudaStreamSynchronize(st));
                n_max = q->h_count[0];
            }
        }
    }
    *forwards_out += forwards;
    return GAD_OK;
}

static int finish_mutate(gad_qnet *q, long long n_unique, int64_t forwards, int64_t *h_forwards, cudaStream_t st)
{
    unsigned long long h_stats[2] = {0, 0};
    GAD_CUDA(cudaMemcpyAsync(h_stats, q->d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    GAD_CUDA(cudaStreamSynchronize(st));
    q->last_stats[1] = n_unique;
    q->last_stats[2] = forwards;
    q->last_stats[3] = (long long)h_stats[0];
    if (h_forwards) *h_forwards = forwards;
    return GAD_OK;
}

// Counters of the last gad_mutate call on this handle: out[0] = episodes (individuals mutated, the
// reference-equivalent count), out[1] = episodes actually run (distinct windows), out[2] = lock-step
// forwards, out[3] = states evaluated (sum over the forwards of their active episodes).
extern "C" int gad_mutate_stats(const gad_qnet *q, int64_t *out4)
{
    GAD_REQUIRE(q && out4, "gad_mutate_stats: null pointer");
    for (int i = 0; i < 4; ++i) out4[i] = q->last_stats[i];
    return GAD_OK;
}

// ---------------------------------------------------------------------------------------------
// The same mutation in separate steps, for a population sharded over several GPUs: the windows of ALL ranks
// are gathered, every rank finds the distinct ones (identically), runs its share of them and the results are
// gathered back - each distinct window runs once in the whole job and the ranks do equal work.
//   gad_mutate_windows   windows of the local individuals            (GA.py:338-350)
//   gad_windows_rep      rep[m] = smallest index of a window identical to m (deterministic across ranks)
//   gad_episodes_run     greedy episodes of listed windows           (GA.py:353-366, env.py:231-259)
//   gad_splice_results   padding + splice from gathered results      (env.py:338-351, GA.py:369-373)
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void windows_kernel(int rows, int cols, const uint8_t *__restrict__ boards, int R, int stride,
                               const int32_t *__restrict__ idx, const int32_t *__restrict__ tile, uint8_t *__restrict__ out)
{
    const long long m = blockIdx.x;
    const long long p = idx ? idx[m] : m;
    const int fr = tile[2 * m], fc = tile[2 * m + 1];
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        out[(size_t)m * rows * cols + i] = (fr < 0 || fc < 0) ? (uint8_t)0 : boards[((size_t)p * R + fr + r) * stride + fc + c];
    }
}
}  // namespace

extern "C" int gad_mutate_windows(const gad_qnet *q, const uint8_t *d_boards, int R, int stride, const int32_t *d_idx,
                                  const int32_t *d_tile, int64_t M, uint8_t *d_windows, void *stream)
{
    if (M == 0) return GAD_OK;
    GAD_REQUIRE(q && d_boards && d_tile && d_windows && M > 0 && R >= q->rows, "gad_mutate_windows: bad argument");
    windows_kernel<<<(unsigned)M, 64, 0, gad_stream(stream)>>>(q->rows, q->cols, d_boards, R, stride, d_idx, d_tile, d_windows);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// d_table: int32 scratch of table_slots entries (a power of two >= 2 * M); d_tmp: int32[M] scratch.
extern "C" int gad_windows_rep(const uint8_t *d_windows, int64_t M, int window_bytes, int32_t *d_rep, int32_t *d_table,
                               int64_t table_slots, int32_t *d_tmp, void *stream)
{
    if (M == 0) return GAD_OK;
    GAD_REQUIRE(d_windows && d_rep && d_table && d_tmp && M > 0 && window_bytes >= 1, "gad_windows_rep: bad argument");
    GAD_REQUIRE(table_slots >= 2 * M && (table_slots & (table_slots - 1)) == 0 && M < (1ll << 30),
                "gad_windows_rep: the table needs a power-of-two number of slots >= 2 * M");
    if (M == 0) return GAD_OK;
    cudaStream_t st = gad_stream(stream);
    GAD_CUDA(cudaMemsetAsync(d_table, 0xFF, (size_t)table_slots * sizeof(int32_t), st));
    GAD_CUDA(cudaMemsetAsync(d_tmp, 0x7F, (size_t)M * sizeof(int32_t), st));
    dedup_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(d_windows, nullptr, window_bytes, M, d_table,
                                                            (unsigned int)(table_slots - 1), d_rep, nullptr, nullptr, 1);
    rep_min_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_rep, d_tmp, M);
    rep_relabel_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_rep, d_tmp, M);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// Episodes of the windows d_list[0 .. U) (indices into d_windows; NULL = the first U windows). Results, indexed by
// position in the list: d_aligned uint8 [U][rows][rows*cols] (the first d_steps[e] columns are valid), d_heads
// uint8 [U][rows] (tokens consumed per row), d_steps int32 [U]. Synchronises.
extern "C" int gad_episodes_run(gad_qnet *q, const uint8_t *d_windows, const int32_t *d_list, int64_t U, int64_t reserve,
                                float max_value, uint8_t *d_aligned, uint8_t *d_heads, int32_t *d_steps, int64_t *h_forwards,
                                void *stream)
{
    GAD_REQUIRE(q && U >= 0 && U < (1ll << 30), "gad_episodes_run: bad argument");
    cudaStream_t st = gad_stream(stream);
    q->last_stats[0] = U; q->last_stats[1] = q->last_stats[2] = q->last_stats[3] = 0;
    int64_t forwards = 0;
    if (U == 0) {
        if (h_forwards) *h_forwards = 0;
        return GAD_OK;
    }
    GAD_REQUIRE(d_windows && d_aligned && d_heads && d_steps, "gad_episodes_run: null pointer");
    // buffers for `reserve` episodes (the most this caller can ever ask for): the number of distinct windows changes
    // from generation to generation and re-allocating in between costs more than the memory
    const long long room = reserve > U ? reserve : U;
    int rc = ensure_episodes(q, room);
    if (!rc) rc = ensure_states(q, room < kMaxBatch ? room : kMaxBatch);
    if (rc) return rc;
    GAD_CUDA(cudaMemsetAsync(q->d_stats, 0, 2 * sizeof(unsigned long long), st));
    episodes_from_windows_kernel<<<(unsigned)U, 64, 0, st>>>(*q, d_windows, d_list, U);
    GAD_LAUNCH_CHECK();
    if ((rc = run_lockstep(q, U, max_value, st, &forwards))) return rc;
    const size_t cap = (size_t)q->rows * q->cols;
    GAD_CUDA(cudaMemcpyAsync(d_aligned, q->aligned, (size_t)U * q->rows * cap, cudaMemcpyDeviceToDevice, st));
    GAD_CUDA(cudaMemcpyAsync(d_heads, q->heads, (size_t)U * q->rows, cudaMemcpyDeviceToDevice, st));
    GAD_CUDA(cudaMemcpyAsync(d_steps, q->steps, (size_t)U * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return finish_mutate(q, U, forwards, h_forwards, st);
}

// d_windows: the local individuals' own windows [M][rows*cols]; d_slot[m]: which entry of the result arrays
// (d_aligned / d_heads / d_steps, e.g. all-gathered from every rank) individual m takes. Asynchronous.
extern "C" int gad_splice_results(const gad_qnet *q, uint8_t *d_boards, int32_t *d_rowlen, int R, int stride,
                                  const int32_t *d_idx, const int32_t *d_tile, int64_t M, const uint8_t *d_windows,
                                  const int32_t *d_slot, const uint8_t *d_aligned, const uint8_t *d_heads,
                                  const int32_t *d_steps, int32_t *d_status, void *stream)
{
    if (M == 0) return GAD_OK;
    GAD_REQUIRE(q && d_boards && d_rowlen && d_tile && d_windows && d_slot && d_aligned && d_heads && d_steps && d_status,
                "gad_splice_results: null pointer");
    GAD_REQUIRE(M > 0 && R >= q->rows && stride >= 16 && stride % 16 == 0, "gad_splice_results: bad shape");
    splice_kernel<<<(unsigned)((M + 3) / 4), 128, 0, gad_stream(stream)>>>(q->rows, q->cols, d_slot, 1, d_aligned, d_heads,
                                                                          d_steps, d_windows, d_boards, d_rowlen, R, stride,
                                                                          d_idx, d_tile, M, d_status);
    GAD_LAUNCH_CHECK();
    return GAD_OK;
}

// Timing hook for bench.py's tensor roofline: the l1 tcgen05 GEMM alone over M rows of the current
// activation buffer, `reps` back-to-back launches between CUDA events on `stream`. Synchronises.
extern "C" int gad_qnet_time_l1(gad_qnet *q, int64_t M, int reps, float *ms_per_launch, double *flops_per_launch,
                                void *stream)
{
    GAD_REQUIRE(q && ms_per_launch && flops_per_launch && M >= 1 && M <= kMaxBatch && reps >= 1, "gad_qnet_time_l1: bad argument");
    GAD_REQUIRE(q->tc1, "gad_qnet_time_l1: the tensor-core l1 is disabled (GAD_QNET_SIMT=1)");
    int rc = ensure_states(q, M);
    if (rc) return rc;
    cudaStream_t st = gad_stream(stream);
    GAD_CUDA(cudaMemsetAsync(q->Z2, 0x3c, (size_t)M * 2 * q->K * sizeof(__half), st));      // finite fp16 values
    cudaEvent_t e0, e1;
    GAD_CUDA(cudaEventCreate(&e0));
    GAD_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i)
        if ((rc = tc_layer_run(q->tc1, q->b1, q->H1s, 2 * q->tc2->Kpad, q->tc2->Kpad, 1, nullptr, M, 1, st))) return rc;
    GAD_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i)
        if ((rc = tc_layer_run(q->tc1, q->b1, q->H1s, 2 * q->tc2->Kpad, q->tc2->Kpad, 1, nullptr, M, 1, st))) return rc;
    GAD_CUDA(cudaEventRecord(e1, st));
    GAD_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    GAD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_launch = ms / reps;
    *flops_per_launch = 3.0 * 2.0 * (double)M * (double)q->K * (double)q->h1;      // three fp16 passes (hi.hi, lo.hi, hi.lo)
    return GAD_OK;
}
