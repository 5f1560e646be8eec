// mlp.cuh — building blocks shared by the minibatch-update kernels (update.cu: ER-DDPG; td3_sac.cu: TD3 / SAC):
// the strided tensor-core GEMM, deterministic column sums / sum of squares, and their launch helpers.
// Everything is `static` (one copy per translation unit).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace erlfill {
namespace mlp {

// ---- generic strided GEMM on the tensor cores: C[M,N] = act(A(M,K) * B(K,N) + bias[N]) ------------
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].  gridDim.z > 1: split-K partials to C + z*split_stride.
// mma.sync.m16n8k8 TF32 with fp32 accumulation, error-compensated (3xTF32: a = hi + lo, a*b ~ lo*hi + hi*lo + hi*hi), so
// the result keeps fp32-level accuracy (the parity bar of the update is 1e-4 against torch fp32) while the
// multiply-accumulates run on the tensor pipe.  These GEMMs are M = batch (128..4096) by N, K <= 256: the register-
// fragment MMA fits them; a tcgen05/TMEM tile would be mostly padding.
constexpr int BM = 64, BN = 64, BK = 16, LDS_T = 72;          // 72 % 32 == 8: conflict-free fragment loads
constexpr int kFetch = BM * BK / 256;                         // tile elements per thread
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// v = hi + lo with hi = v truncated to TF32 (one LOP) and lo = v - hi (exact in fp32; the tensor core reads the top 19 bits of
// the register, i.e. truncates lo to TF32 itself: |error| <= 2^-21 |v|).  No cvt.rna: the conversion pipe is quarter rate
// and throttled the MMA loop (ncu: math_pipe_throttle was the top stall with it).
static __device__ __forceinline__ void split_tf32(float v, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
static __device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// one 64 x 64 output tile at (m0, n0) over k in [kbeg, kend), written to Cz
static __device__ __forceinline__ void gemm_tile(int M, int N, const float *__restrict__ A, int sam, int sak,
                                                 const float *__restrict__ B, int sbk, int sbn, float *__restrict__ Cz, int ldc,
                                                 const float *__restrict__ bias, int act, int kbeg, int kend, int m0, int n0) {
    __shared__ __align__(16) float As[BK][LDS_T];
    __shared__ __align__(16) float Bs[BK][LDS_T];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;      // 2 x 4 warps, each a 32 x 16 tile = 2 (m16) x 2 (n8) MMA tiles
    const int g = lane >> 2, t = lane & 3;
    float acc[2][2][4] = {};
    // per-thread tile coordinates (the fastest-varying index follows the unit-stride dimension for coalescing)
    int am[kFetch], ak[kFetch], bn[kFetch], bk[kFetch];
#pragma unroll
    for (int it = 0; it < kFetch; ++it) {
        const int e = tid + it * 256;
        if (sak == 1) { ak[it] = e % BK; am[it] = e / BK; } else { am[it] = e & 63; ak[it] = e >> 6; }
        if (sbn == 1) { bn[it] = e & 63; bk[it] = e >> 6; } else { bk[it] = e % BK; bn[it] = e / BK; }
    }
    float ra[kFetch], rb[kFetch];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int it = 0; it < kFetch; ++it) {
            const int gm = m0 + am[it], gk = k0 + ak[it];
            ra[it] = (gm < M && gk < kend) ? A[(size_t)gm * sam + (size_t)gk * sak] : 0.f;
            const int gn = n0 + bn[it], gkb = k0 + bk[it];
            rb[it] = (gn < N && gkb < kend) ? B[(size_t)gkb * sbk + (size_t)gn * sbn] : 0.f;
        }
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int it = 0; it < kFetch; ++it) { As[ak[it]][am[it]] = ra[it]; Bs[bk[it]][bn[it]] = rb[it]; }
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);           // next tile's global loads fly under this tile's MMAs
#pragma unroll
        for (int ks = 0; ks < BK; ks += 8) {
            uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                split_tf32(As[ks + t][wm + i * 16 + g], ah[i][0], al[i][0]);
                split_tf32(As[ks + t][wm + i * 16 + g + 8], ah[i][1], al[i][1]);
                split_tf32(As[ks + t + 4][wm + i * 16 + g], ah[i][2], al[i][2]);
                split_tf32(As[ks + t + 4][wm + i * 16 + g + 8], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                split_tf32(Bs[ks + t][wn + j * 8 + g], bh[j][0], bl[j][0]);
                split_tf32(Bs[ks + t + 4][wn + j * 8 + g], bh[j][1], bl[j][1]);
            }
            // term-major order: the MMAs on one accumulator are dependent (in-order issue), so the four accumulators are
            // interleaved between the three correction terms (small terms first)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[i][j], al[i], bh[j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[i][j], ah[i], bl[j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(acc[i][j], ah[i], bh[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = m0 + wm + i * 16 + g + (e >> 1) * 8, gn = n0 + wn + j * 8 + 2 * t + (e & 1);
                if (gm >= M || gn >= N) continue;
                float v = acc[i][j][e];
                if (bias) v += bias[gn];
                if (act == ACT_RELU) v = fmaxf(v, 0.f);
                else if (act == ACT_LEAKY) v = v > 0.f ? v : 0.1f * v;
                Cz[(size_t)gm * ldc + gn] = v;
            }
}

static __global__ void __launch_bounds__(256, 3) gemm_kernel(int M, int N, int K, const float *__restrict__ A, int sam, int sak,
                                                   const float *__restrict__ B, int sbk, int sbn, float *__restrict__ C,
                                                   int ldc, const float *__restrict__ bias, int act, int k_per_split,
                                                   size_t split_stride) {
    const int kbeg = blockIdx.z * k_per_split;
    gemm_tile(M, N, A, sam, sak, B, sbk, sbn, C + (size_t)blockIdx.z * split_stride, ldc, bias, act, kbeg, min(K, kbeg + k_per_split),
              blockIdx.y * BM, blockIdx.x * BN);
}

// ---- large-M forward GEMM: C[M,N] = act(A[M,K] W[N,K]^T + bias[N]), A and W row-major with unit K stride -------------------
// The act paths of the batched agents push 16,384 .. 131,072 rows through 256-wide layers; there the 64 x 64 kernel above is
// bound by its fragment traffic (4 instructions per MMA: ncu tensor pipe 34 %, profiles/r01d_td3_gemm.txt).  Here a CTA owns a
// 128 x 128 tile and a warp 64 x 32 of it (4 x 4 MMA tiles): 24 fragment loads + 24 hi/lo splits feed 48 MMAs per k-step
// (2.5 instructions per MMA), tiles arrive through a 3-stage cp.async ring of 32-wide K chunks with one barrier per chunk.
// Per output element the accumulation order (k ascending, lo*hi, hi*lo, hi*hi per k-step) is the one of gemm_kernel: same bits.
constexpr int GB_M = 128, GB_N = 128, GB_K = 32, GB_LD = 36, GB_STAGES = 3;      // 36 % 32 == 4: conflict-free fragment loads
constexpr int kGemmBigSmem = GB_STAGES * (GB_M + GB_N) * GB_LD * 4;               // 110,592 B
static __device__ __forceinline__ void gb_cp_async16(float *s, const float *g) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g));
}
static __global__ void __launch_bounds__(256, 2) gemm_big_kernel(int M, int N, int K, const float *__restrict__ A, int lda,
                                                                 const float *__restrict__ W, int ldw, float *__restrict__ C, int ldc,
                                                                 const float *__restrict__ bias, int act) {
    extern __shared__ __align__(16) float gsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;      // 2 x 4 warps, each 64 x 32 = 4 (m16) x 4 (n8) MMA tiles
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
    float acc[4][4][4] = {};
    const int nk = K / GB_K;
    auto load = [&](int c) {
        float *As = gsm + (c % GB_STAGES) * (GB_M + GB_N) * GB_LD, *Ws = As + GB_M * GB_LD;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int v = tid + it * 256, r = v >> 3, c4 = (v & 7) * 4;
            float *da = As + r * GB_LD + c4, *dw = Ws + r * GB_LD + c4;
            if (m0 + r < M) gb_cp_async16(da, A + (size_t)(m0 + r) * lda + c * GB_K + c4);
            else *reinterpret_cast<float4 *>(da) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < N) gb_cp_async16(dw, W + (size_t)(n0 + r) * ldw + c * GB_K + c4);
            else *reinterpret_cast<float4 *>(dw) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;");
    };
    load(0);
    if (nk > 1) load(1);
#pragma unroll 1
    for (int c = 0; c < nk; ++c) {
        if (c + 1 < nk) asm volatile("cp.async.wait_group 1;"); else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (c + 2 < nk) load(c + 2);           // into the stage chunk c - 1 was read from: every warp is past its MMAs
        const float *As = gsm + (c % GB_STAGES) * (GB_M + GB_N) * GB_LD, *Ws = As + GB_M * GB_LD;
#pragma unroll
        for (int ks = 0; ks < GB_K; ks += 8) {
            uint32_t ah[4][4], al[4][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float *ap = As + (wm + i * 16 + g) * GB_LD + ks + t;
                split_tf32(ap[0], ah[i][0], al[i][0]);
                split_tf32(ap[8 * GB_LD], ah[i][1], al[i][1]);
                split_tf32(ap[4], ah[i][2], al[i][2]);
                split_tf32(ap[8 * GB_LD + 4], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float *bp = Ws + (wn + j * 8 + g) * GB_LD + ks + t;
                split_tf32(bp[0], bh[j][0], bl[j][0]);
                split_tf32(bp[4], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], al[i], bh[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], ah[i], bl[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], ah[i], bh[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = m0 + wm + i * 16 + g + (e >> 1) * 8, gn = n0 + wn + j * 8 + 2 * t + (e & 1);
                if (gm >= M || gn >= N) continue;
                float v = acc[i][j][e];
                if (bias) v += bias[gn];
                if (act == ACT_RELU) v = fmaxf(v, 0.f);
                else if (act == ACT_LEAKY) v = v > 0.f ? v : 0.1f * v;
                C[(size_t)gm * ldc + gn] = v;
            }
}

// ---- the two thin layers of the act MLPs at large M (exact fp32 FMA; the MMA tile kernels are mostly padding there) ------------
// K <= 4 (state -> first hidden layer), N == 256: C[m][n] = act(bias[n] + sum_k A[m][k] W[n][k]).  A warp walks rows; lane l owns
// columns 4 l .. 4 l + 3 and 128 + 4 l .. : its 8 weight rows and biases stay in registers, a row costs two broadcast loads and two
// fully coalesced 512-byte stores: the kernel is the bytes it writes.
constexpr int kThinKN = 256;
static __global__ void __launch_bounds__(256) linear_smallk_kernel(int M, int K, const float *__restrict__ A, int lda,
                                                                   const float *__restrict__ W, int ldw, float *__restrict__ C, int ldc,
                                                                   const float *__restrict__ bias, int act) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float w[8][4], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = (j >> 2) * 128 + 4 * lane + (j & 3);
        b[j] = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) w[j][k] = k < K ? __ldg(W + (size_t)n * ldw + k) : 0.f;
    }
    for (int m = blockIdx.x * wpb + warp; m < M; m += gridDim.x * wpb) {
        float a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = k < K ? __ldg(A + (size_t)m * lda + k) : 0.f;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s = b[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (k < K) s = fmaf(a[k], w[j][k], s);
            if (act == ACT_RELU) s = fmaxf(s, 0.f);
            else if (act == ACT_LEAKY) s = s > 0.f ? s : 0.1f * s;
            v[j] = s;
        }
        float4 *dst = reinterpret_cast<float4 *>(C + (size_t)m * ldc) + lane;
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[32] = make_float4(v[4], v[5], v[6], v[7]);
    }
}
// N <= 16 (last hidden layer -> action heads), K == 256: one warp per row, lane l owns k = l + 32 q; its 8 x 16 weights stay in
// registers, a row costs 8 coalesced loads, 128 FMAs and a 16-shuffle transposing reduction (each step halves the values a lane
// holds and the lanes that share them; fixed order), after which lane 2 n (bit-reversed over lane bits 4..1) holds output n.
constexpr int kThinNMax = 16, kThinNK = 256;
template <int NW>      // columns of W held in registers: N <= NW <= 16 (10 for the action heads: 80 weight registers, two CTAs per SM)
static __global__ void __launch_bounds__(256, NW <= 10 ? 2 : 1) linear_smalln_kernel(int M, int N, const float *__restrict__ A, int lda,
                                                                      const float *__restrict__ W, int ldw, float *__restrict__ C, int ldc,
                                                                      const float *__restrict__ bias, int act) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float w[8][NW];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int n = 0; n < NW; ++n) w[q][n] = n < N ? __ldg(W + (size_t)n * ldw + lane + 32 * q) : 0.f;
    // output this lane ends up with: bit 4 of the lane picks the upper 8, bit 3 the upper 4 of those, ...
    const int n_mine = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const float b_mine = (bias && n_mine < N) ? __ldg(bias + n_mine) : 0.f;
    const int stride = gridDim.x * wpb;
    float an[8];                                        // the next row's values are in flight while this row is reduced
    {
        const int m = blockIdx.x * wpb + warp;
#pragma unroll
        for (int q = 0; q < 8; ++q) an[q] = m < M ? __ldg(A + (size_t)m * lda + lane + 32 * q) : 0.f;
    }
    for (int m = blockIdx.x * wpb + warp; m < M; m += stride) {
        float a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = an[q];
        if (m + stride < M) {
#pragma unroll
            for (int q = 0; q < 8; ++q) an[q] = __ldg(A + (size_t)(m + stride) * lda + lane + 32 * q);
        }
        float s[kThinNMax];
#pragma unroll
        for (int n = 0; n < kThinNMax; ++n) s[n] = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int n = 0; n < NW; ++n) s[n] = fmaf(a[q], w[q][n], s[n]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float send = (lane & 16) ? s[i] : s[i + 8], keep = (lane & 16) ? s[i + 8] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float send = (lane & 8) ? s[i] : s[i + 4], keep = (lane & 8) ? s[i + 4] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float send = (lane & 4) ? s[i] : s[i + 2], keep = (lane & 4) ? s[i + 2] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        {
            const float send = (lane & 2) ? s[0] : s[1], keep = (lane & 2) ? s[1] : s[0];
            s[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
        if ((lane & 1) == 0 && n_mine < N) {
            float v = s[0] + b_mine;
            if (act == ACT_RELU) v = fmaxf(v, 0.f);
            else if (act == ACT_LEAKY) v = v > 0.f ? v : 0.1f * v;
            C[(size_t)m * ldc + n_mine] = v;
        }
    }
}

// ---- grouped weight gradients: up to 4 GEMMs dW_i[N_i][K_i] = dZ_i^T[N_i][Bt] X_i[Bt][K_i] in ONE launch ------------
// Every task is cut into 64 x 64 tiles x `splits` slices of the batch rows; slice s of task i goes to part_i + s*N_i*K_i
// and group_reduce_kernel adds the slices in index order (deterministic).
struct WgTask { const float *dZ, *X; float *part; int ldz, ldx, N, K, tiles_k, tiles, first_block; };
struct WgArgs { WgTask t[4]; int ntask, Bt, splits, kps; };
static __global__ void __launch_bounds__(256, 3) wgrad_group_kernel(const WgArgs a) {
    int ti = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < a.ntask && (int)blockIdx.x >= a.t[i].first_block) ti = i;
    const WgTask &k = a.t[ti];
    const int local = blockIdx.x - k.first_block, tile = local % k.tiles, split = local / k.tiles;
    const int kbeg = split * a.kps;
    gemm_tile(k.N, k.K, k.dZ, 1, k.ldz, k.X, k.ldx, 1, k.part + (size_t)split * k.N * k.K, k.K, nullptr, ACT_NONE, kbeg,
              min(a.Bt, kbeg + a.kps), (tile / k.tiles_k) * BM, (tile % k.tiles_k) * BN);
}
// out_i[j] = sum_s part_i[s*stride_i + j] for up to 12 segments in one launch (256 elements per block)
struct RedSeg { const float *part; float *out; int n, splits, stride, first_block; };
struct RedArgs { RedSeg s[12]; int nseg; };
static __global__ void __launch_bounds__(256) group_reduce_kernel(const RedArgs a) {
    int si = 0;
#pragma unroll
    for (int i = 1; i < 12; ++i)
        if (i < a.nseg && (int)blockIdx.x >= a.s[i].first_block) si = i;
    const RedSeg &g = a.s[si];
    const int j = (blockIdx.x - g.first_block) * 256 + threadIdx.x;
    if (j >= g.n) return;
    float acc = 0.f;
    for (int z = 0; z < g.splits; ++z) acc += g.part[(size_t)z * g.stride + j];
    g.out[j] = acc;
}

// sum split-K partials: out[i] = sum_z part[z*stride + i]
static __global__ void splitk_reduce_kernel(int n, int splits, const float *part, size_t stride, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(size_t)z * stride + i];
    out[i] = s;
}

// column sums, two deterministic stages: part[s][c] = sum over the rows of slice s of X[r*ld + c] * y(r, c), where
// y = 1 (Y == nullptr) or Y[r*ldy + c*ycs] (ycs = 0: one factor per row); then splitk_reduce_kernel adds the slices in
// index order.  grid = (ceil(cols/32), kColSplits): a block is 32 columns x 8 row lanes over rows/kColSplits rows.
constexpr int kColSplits = 32;
static __global__ void __launch_bounds__(256) colsum_part_kernel(int rows, int cols, const float *X, int ld, const float *Y, int ldy, int ycs,
                                                          float *part) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int per = (rows + kColSplits - 1) / kColSplits, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float s = 0.f;
    if (c < cols)
        for (int r = r0 + rl; r < r1; r += 8) {
            const float x = X[(size_t)r * ld + c];
            s = Y ? fmaf(x, Y[(size_t)r * ldy + (size_t)c * ycs], s) : s + x;
        }
    red[rl][threadIdx.x & 31] = s;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
        part[(size_t)blockIdx.y * cols + c] = t;
    }
}

// ---- block-wide deterministic reductions ------------------------------------------------------
static __device__ __forceinline__ float block_sum(float v, float *scratch /*[32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : 0.f;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) scratch[0] = t;
    }
    __syncthreads();
    return scratch[0];
}


static __global__ void relu_bwd_kernel(size_t n, float *g, const float *h) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = h[i] > 0.f ? g[i] : 0.f;
}

// sum of squares of a flat buffer: kSumsqBlocks per-block partials (fixed order inside a block); the consumer adds them in index order
constexpr int kSumsqBlocks = 64;
static __global__ void __launch_bounds__(256) sumsq_kernel(int n, const float *g, float *out /*[kSumsqBlocks]*/) {
    __shared__ float scratch[32];
    const int per = (n + kSumsqBlocks - 1) / kSumsqBlocks, i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    float s = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) s = fmaf(g[i], g[i], s);
    const float t = block_sum(s, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}
// learning rates from the (all-reduced) emotion means: ddpg_agent.py:447-457

static __global__ void fill_kernel(int n, float v, float *p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- launch helpers ---------------------------------------------------------------------------
static inline void gemm(cudaStream_t st, int M, int N, int K, const float *A, int sam, int sak, const float *B, int sbk, int sbn,
                        float *C, int ldc, const float *bias, int act) {
    // large row-major forward GEMMs (the act paths): thin first / last layers in exact fp32, the 128 x 128 tile kernel for the
    // hidden layers; everything else: the 64 x 64 strided kernel
    if (M >= 1024 && sak == 1 && sbk == 1) {      // thin layers: already at minibatch size (the B = 4096 updates), not only on the act path
        if (K <= 4 && N == kThinKN && (ldc & 3) == 0 && ((uintptr_t)C & 15u) == 0) {
            linear_smallk_kernel<<<M >= 65536 ? 148 * 8 : 148 * 2, 256, 0, st>>>(M, K, A, sam, B, sbn, C, ldc, bias, act);
            return;
        }
        if (N <= kThinNMax && K == kThinNK) {
            if (N <= 10) linear_smalln_kernel<10><<<148 * 2, 256, 0, st>>>(M, N, A, sam, B, sbn, C, ldc, bias, act);
            else linear_smalln_kernel<kThinNMax><<<148, 256, 0, st>>>(M, N, A, sam, B, sbn, C, ldc, bias, act);
            return;
        }
    }
    if (M >= 8192 && N >= 64 && sak == 1 && sbk == 1 && K % GB_K == 0 && (sam & 3) == 0 && (sbn & 3) == 0 &&
        (((uintptr_t)A | (uintptr_t)B) & 15u) == 0) {
        static bool attr_set[64] = {};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaFuncSetAttribute(gemm_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmBigSmem);
            attr_set[dev] = true;
        }
        dim3 gridb((N + GB_N - 1) / GB_N, (M + GB_M - 1) / GB_M, 1);
        gemm_big_kernel<<<gridb, 256, kGemmBigSmem, st>>>(M, N, K, A, sam, B, sbn, C, ldc, bias, act);
        return;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
    gemm_kernel<<<grid, 256, 0, st>>>(M, N, K, A, sam, sak, B, sbk, sbn, C, ldc, bias, act, K, 0);
}
// floats of split scratch the helpers below may use
constexpr size_t kPartFloats = (size_t)32 * 256 * 256;
// dW[N,K] = dZ^T[N,Bt] * X[Bt,K], reduction over the batch rows with split-K (`part`: kPartFloats floats of scratch)
static inline void gemm_wgrad(cudaStream_t st, float *part, int Bt, int N, int K, const float *dZ, int ldz, const float *X, int ldx,
                              float *dW) {
    int splits = Bt >= 2048 ? 16 : (Bt >= 512 ? 8 : (Bt >= 128 ? 4 : 1));
    const int kps = ((Bt + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (Bt + kps - 1) / kps;
    dim3 grid((K + BN - 1) / BN, (N + BM - 1) / BM, splits);
    if (splits == 1) {
        gemm_kernel<<<grid, 256, 0, st>>>(N, K, Bt, dZ, 1, ldz, X, ldx, 1, dW, K, nullptr, ACT_NONE, kps, 0);
    } else {
        gemm_kernel<<<grid, 256, 0, st>>>(N, K, Bt, dZ, 1, ldz, X, ldx, 1, part, K, nullptr, ACT_NONE, kps, (size_t)N * K);
        splitk_reduce_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(N * K, splits, part, (size_t)N * K, dW);
    }
}
// out[c] = sum_r X[r*ld + c] * (Y ? Y[r*ldy + c*ycs] : 1); `part` is the slice scratch (stream-ordered reuse)
static inline void colsum(cudaStream_t st, float *part, int rows, int cols, const float *X, int ld, const float *Y, int ldy, int ycs,
                          float *out) {
    dim3 grid((cols + 31) / 32, kColSplits);
    colsum_part_kernel<<<grid, 256, 0, st>>>(rows, cols, X, ld, Y, ldy, ycs, part);
    splitk_reduce_kernel<<<(cols + 255) / 256, 256, 0, st>>>(cols, kColSplits, part, (size_t)cols, out);
}

}  // namespace mlp
}  // namespace erlfill
