// fused_mlp.cuh — row-tile building blocks of the fused minibatch-update kernels (update.cu).
//
// One CTA owns TM = 32 rows of the minibatch and walks a whole MLP chain (forward, or the input-gradient chain of the
// backward) with the activations of its rows in shared memory; only what the weight-gradient GEMMs need later is
// written to global memory.  B = 4096 rows -> 128 CTAs for the 148 SMs.  The dense layers are mma.sync.m16n8k8 TF32
// with fp32 accumulation, error-compensated (3xTF32, see mlp.cuh) so the parity bar of the update (1e-4 against torch
// fp32) holds.  The weights (<= 256 KB per layer) are read from L2 by every CTA, staged through a 3-deep cp.async ring of
// 32-wide K chunks; the layer chain of a whole forward is one launch instead of seven.
//
// Fragment/bank plan: activation tiles are [TM][LDA] with LDA = 260 (== 4 mod 32): the A-fragment pattern (row g, col t)
// touches 32 distinct banks.  Weight stages are [n][36] for W[N][K] (forward: out = in W^T) and [k][N + 8] for the
// transposed use (backward: din = dout W): both conflict-free for the B-fragment pattern.
#pragma once
#include "mlp.cuh"

namespace erlfill {
namespace fused {

using mlp::mma_tf32;
using mlp::split_tf32;

constexpr int TM = 32;            // rows per CTA
constexpr int NTHR = 512;         // threads per CTA (16 warps: the chain is latency bound, not issue bound)
constexpr int KC = 32;            // K chunk staged per cp.async group
constexpr int LDA = 260;          // activation tile row stride
constexpr int LDW_NK = 36;        // weight stage row stride, [n][k] layout
constexpr int WST = 256 * LDW_NK; // floats per weight stage (covers 32 x (256 + 8) of the [k][n] layout too)
constexpr int TILE = TM * LDA;
constexpr int NSTAGE = 3;         // weight ring depth

__device__ __forceinline__ void cp_async16(float *s, const float *g) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g));
}
__device__ __forceinline__ void cp_async8(float *s, const float *g) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g));
}
// 256-float rows [row0 .. row0 + TM) of a global [B][256] matrix -> tile (rows >= B zero-filled); one cp.async group
__device__ __forceinline__ void prefetch_tile256(float *tile, const float *__restrict__ G, int row0, int B);
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void prefetch_tile256(float *tile, const float *__restrict__ G, int row0, int B) {
    for (int v = threadIdx.x; v < TM * 64; v += NTHR) {
        const int r = v >> 6, c4 = v & 63;
        float *dst = tile + r * LDA + c4 * 4;
        if (row0 + r < B) cp_async16(dst, G + (size_t)(row0 + r) * 256 + c4 * 4);
        else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
}

// n-tiles (8 columns) per warp and warps that own columns, for an N-column output tile
template <int N> struct Cols {
    static constexpr int NTW = N >= 128 ? N / 128 : 1;
    static constexpr int WARPS = N / (8 * NTW);
};

// ---- weight chunk loaders -------------------------------------------------------------------------------------
// vec: 4 = rows 16-byte aligned (cp.async 16), 2 = 8-byte aligned (cp.async 8: the actor's flat buffer starts its matrices
// on even, not multiple-of-4, offsets), 0 = guarded scalar loads (ragged rows: Linear(15, .), Linear(5, .)).
// W[n][k] (row stride ldw), rows n < n_real, columns k0 .. k0 + KCc - 1 (k < k_real) -> st[n][k - k0]
template <int N, int KCc>
__device__ __forceinline__ void load_w_nk(float *st, const float *__restrict__ W, int ldw, int n_real, int k_real, int k0, int vec) {
    if (vec) {
        const int VPR = KCc / vec;
        for (int v = threadIdx.x; v < N * VPR; v += NTHR) {
            const int n = v / VPR, kv = v % VPR;
            float *dst = st + n * LDW_NK + kv * vec;
            const float *src = W + (size_t)n * ldw + k0 + kv * vec;
            if (n < n_real) { if (vec == 4) cp_async16(dst, src); else cp_async8(dst, src); }
            else { dst[0] = 0.f; dst[1] = 0.f; if (vec == 4) { dst[2] = 0.f; dst[3] = 0.f; } }
        }
    } else {
        constexpr int IT = (N * KCc + NTHR - 1) / NTHR;       // all loads first (one exposed latency), then the stores
        float tmp[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int e = threadIdx.x + it * NTHR, n = e / KCc, k = e % KCc;
            tmp[it] = (e < N * KCc && n < n_real && k0 + k < k_real) ? W[(size_t)n * ldw + k0 + k] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int e = threadIdx.x + it * NTHR;
            if (e < N * KCc) st[(e / KCc) * LDW_NK + e % KCc] = tmp[it];
        }
    }
}
// W[k][n] (row stride ldw), rows k0 .. k0 + KCc - 1 (k < k_real), columns n < n_real -> st[k - k0][n], stride N + 8
template <int N, int KCc>
__device__ __forceinline__ void load_w_kn(float *st, const float *__restrict__ W, int ldw, int n_real, int k_real, int k0, int vec) {
    constexpr int LDS = N + 8;
    if (vec) {
        const int VPR = N / vec;
        for (int v = threadIdx.x; v < KCc * VPR; v += NTHR) {
            const int k = v / VPR, nv = v % VPR;
            float *dst = st + k * LDS + nv * vec;
            const float *src = W + (size_t)(k0 + k) * ldw + nv * vec;
            if (k0 + k < k_real) { if (vec == 4) cp_async16(dst, src); else cp_async8(dst, src); }
            else { dst[0] = 0.f; dst[1] = 0.f; if (vec == 4) { dst[2] = 0.f; dst[3] = 0.f; } }
        }
    } else {
        constexpr int IT = (N * KCc + NTHR - 1) / NTHR;
        float tmp[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int e = threadIdx.x + it * NTHR, k = e / N, n = e % N;
            tmp[it] = (e < N * KCc && n < n_real && k0 + k < k_real) ? W[(size_t)(k0 + k) * ldw + n] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int e = threadIdx.x + it * NTHR;
            if (e < N * KCc) st[(e / N) * LDS + e % N] = tmp[it];
        }
    }
}

// ---- one dense layer on the CTA's row tile ---------------------------------------------------------------------
// acc[i][j][e] (m16n8 fragments of rows i*16.., columns col0 + j*8..) = in_s[TM][K] * Wop, where
//   KN == false: Wop(k, n) = W[n][k]   (forward,  W is [n_real][k_real], ldw = k_real's row stride)
//   KN == true : Wop(k, n) = W[k][n]   (backward, W is [k_real][n_real])
// vec: see the loaders above.
// All NTHR threads must call (loads and barriers); warps >= Cols<N>::WARPS own no columns.  Ends with a __syncthreads():
// in_s and the stages may be overwritten afterwards.
template <int K, int N, bool KN>
__device__ __forceinline__ void dense(float (&acc)[2][Cols<N>::NTW][4], const float *in_s, const float *__restrict__ W, int ldw,
                                      int n_real, int k_real, float *wst, int vec) {
    constexpr int NTW = Cols<N>::NTW;
    constexpr int KCc = K < KC ? K : KC;
    constexpr int NCH = K / KCc;
    constexpr int LDS = KN ? N + 8 : LDW_NK;
    static_assert(K % KCc == 0 && KCc % 8 == 0, "K must be a multiple of 8 (and of 32 beyond 32)");
    static_assert((KN ? KCc * (N + 8) : N * LDW_NK) <= WST, "weight stage too small");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int col0 = warp * NTW * 8;
    const bool active = warp < Cols<N>::WARPS;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
    auto load = [&](int c) {
        float *st = wst + (c % NSTAGE) * WST;
        if (KN) load_w_kn<N, KCc>(st, W, ldw, n_real, k_real, c * KCc, vec);
        else load_w_nk<N, KCc>(st, W, ldw, n_real, k_real, c * KCc, vec);
        cp_async_commit();
    };
    // 3-stage ring, prefetch distance 2, ONE barrier per chunk: chunk c + 2 goes into the stage chunk c - 1 was read from, and
    // every warp has finished chunk c - 1 once it is past the barrier of iteration c
    load(0);
    if (NCH > 1) load(1);
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        if (c + 2 < NCH) load(c + 2);
        if (active) {
            const float *st = wst + (c % NSTAGE) * WST;
#pragma unroll
            for (int ks = 0; ks < KCc; ks += 8) {
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float *ap = in_s + (i * 16 + g) * LDA + c * KCc + ks + t;
                    split_tf32(ap[0], ah[i][0], al[i][0]);
                    split_tf32(ap[8 * LDA], ah[i][1], al[i][1]);
                    split_tf32(ap[4], ah[i][2], al[i][2]);
                    split_tf32(ap[8 * LDA + 4], ah[i][3], al[i][3]);
                }
                uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const int n = col0 + j * 8 + g;
                    if (KN) {
                        split_tf32(st[(ks + t) * LDS + n], bh[j][0], bl[j][0]);
                        split_tf32(st[(ks + t + 4) * LDS + n], bh[j][1], bl[j][1]);
                    } else {
                        split_tf32(st[n * LDS + ks + t], bh[j][0], bl[j][0]);
                        split_tf32(st[n * LDS + ks + t + 4], bh[j][1], bl[j][1]);
                    }
                }
                // term-major order: MMAs on one accumulator are dependent and a warp issues in order, so all 2 x NTW
                // accumulators are interleaved between the three 3xTF32 terms (small terms first)
#pragma unroll
                for (int j = 0; j < NTW; ++j)
#pragma unroll
                    for (int i = 0; i < 2; ++i) mma_tf32(acc[i][j], al[i], bh[j]);
#pragma unroll
                for (int j = 0; j < NTW; ++j)
#pragma unroll
                    for (int i = 0; i < 2; ++i) mma_tf32(acc[i][j], ah[i], bl[j]);
#pragma unroll
                for (int j = 0; j < NTW; ++j)
#pragma unroll
                    for (int i = 0; i < 2; ++i) mma_tf32(acc[i][j], ah[i], bh[j]);
            }
        }
    }
    __syncthreads();
}

// visit the accumulator fragments: f(row in tile, column, value); columns >= n_real are skipped
template <int N, typename F>
__device__ __forceinline__ void for_each_out(const float (&acc)[2][Cols<N>::NTW][4], int n_real, F f) {
    constexpr int NTW = Cols<N>::NTW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    if (warp >= Cols<N>::WARPS) return;
    const int col0 = warp * NTW * 8;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = i * 16 + g + (e >> 1) * 8, c = col0 + j * 8 + 2 * t + (e & 1);
                if (c < n_real) f(r, c, acc[i][j][e]);
            }
}

}  // namespace fused
}  // namespace erlfill
