// Development probe (not part of the product): MN-major ("transposed") shared-memory operands of tcgen05.mma kind::f16 in the
// no-swizzle canonical layout, as the fused ER-DDPG update (csrc/update_tc.cu) uses them:
//   * the SAME row-major activation image X[row][feature] (K-major for the forward, K = features) is read a second time as the
//     MN-major operand of the weight-gradient GEMM dW = dZ^T X (K = rows), and
//   * the SAME weight image W[out][in] (K-major B operand of the forward, K = in) is read as the MN-major B operand of the
//     input-gradient GEMM dX = dZ W (K = out).
// D[128 x 256] = A B^T with K = 128, small-integer data (exact in fp32), checked against the host for the four major combinations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I erl-fill_b200/csrc -I include tools/umma_mn_probe.cu -o tools/bin/umma_mn_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "umma.cuh"

using namespace erlfill::umma;

constexpr int M = 128, N = 256, K = 128;

// K-major image of a [R][Kd] matrix: element (r, k)
static uint32_t off_kmajor(int r, int k, int Kd) { return (r >> 3) * (Kd >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ uint32_t idesc(int m, int n, int a_mn, int b_mn) {
    return instr_desc_bf16(m, n) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// a_mn / b_mn: operand is the MN-major view of a row-major image whose rows are the K index (image [K][M] resp. [K][N]);
// swap: exchange the roles of LBO / SBO in the MN-major descriptors (to pin the convention)
__global__ void __launch_bounds__(128) probe(const uint8_t *Aimg, const uint8_t *Bimg, int a_mn, int b_mn, int swap, float *Dout) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + M * K * 2;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K * 2 / 16; i += 128) reinterpret_cast<uint4 *>(sA)[i] = reinterpret_cast<const uint4 *>(Aimg)[i];
    for (int i = tid; i < N * K * 2 / 16; i += 128) reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(Bimg)[i];
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tD = tmem_slot;
    if (tid == 0) {
        const uint32_t id = idesc(M, N, a_mn, b_mn);
        for (int s = 0; s < K / 16; ++s) {
            uint64_t da, db;
            if (!a_mn) da = smem_desc(smem_u32(sA) + s * 256, 128, (K / 8) * 128);
            else {
                // image [K rows][M features]: 8 consecutive features = 16 contiguous bytes, next 8 features +128 B (MN group stride),
                // next 8 K rows + (M / 8) * 128 B (K group stride); one MMA consumes 16 K rows
                const uint32_t kgrp = (M / 8) * 128, mngrp = 128;
                da = smem_desc(smem_u32(sA) + s * 2 * kgrp, swap ? mngrp : kgrp, swap ? kgrp : mngrp);
            }
            if (!b_mn) db = smem_desc(smem_u32(sB) + s * 256, 128, (K / 8) * 128);
            else {
                const uint32_t kgrp = (N / 8) * 128, mngrp = 128;
                db = smem_desc(smem_u32(sB) + s * 2 * kgrp, swap ? mngrp : kgrp, swap ? kgrp : mngrp);
            }
            mma_ss(tD, da, db, id, s > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = warp * 32 + lane;
    for (int c = 0; c < N / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tD + lane_base + c * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) Dout[row * N + c * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tD, 256);
}

int main() {
    std::vector<float> A(M * K), B(N * K);
    srand(7);
    for (auto &v : A) v = (float)(rand() % 7 - 3);
    for (auto &v : B) v = (float)(rand() % 5 - 2);
    auto put = [](std::vector<uint8_t> &img, uint32_t off, float v) { __nv_bfloat16 b = __float2bfloat16(v); memcpy(&img[off], &b, 2); };
    // K-major images: rows = M (resp. N), K contiguous.  "Transposed-storage" images: rows = K, features = M (resp. N).
    std::vector<uint8_t> Ak(M * K * 2), Bk(N * K * 2), At(M * K * 2), Bt(N * K * 2);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { put(Ak, off_kmajor(m, k, K), A[m * K + k]); put(At, off_kmajor(k, m, M), A[m * K + k]); }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { put(Bk, off_kmajor(n, k, K), B[n * K + k]); put(Bt, off_kmajor(k, n, N), B[n * K + k]); }
    std::vector<float> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    uint8_t *dAk, *dBk, *dAt, *dBt; float *dD;
    cudaMalloc(&dAk, Ak.size()); cudaMalloc(&dBk, Bk.size()); cudaMalloc(&dAt, At.size()); cudaMalloc(&dBt, Bt.size()); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dAk, Ak.data(), Ak.size(), cudaMemcpyHostToDevice); cudaMemcpy(dBk, Bk.data(), Bk.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dAt, At.data(), At.size(), cudaMemcpyHostToDevice); cudaMemcpy(dBt, Bt.data(), Bt.size(), cudaMemcpyHostToDevice);
    const int smem = (M + N) * K * 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int rc = 0;
    for (int swap = 0; swap < 2; ++swap)
        for (int v = 0; v < 4; ++v) {
            const int a_mn = v & 1, b_mn = v >> 1;
            if (swap && v == 0) continue;
            cudaMemset(dD, 0, M * N * 4);
            probe<<<1, 128, smem>>>(a_mn ? dAt : dAk, b_mn ? dBt : dBk, a_mn, b_mn, swap, dD);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(M * N);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double err = 0; int bad = 0;
            for (int i = 0; i < M * N; ++i) { const double d = fabs(D[i] - ref[i]); err = fmax(err, d); bad += d != 0; }
            printf("A %s, B %s, lbo/sbo %s: %s  max abs err %.1f, %d / %d elements differ\n", a_mn ? "MN-major" : "K-major ", b_mn ? "MN-major" : "K-major ",
                   swap ? "swapped (LBO = MN-group stride)" : "as designed (LBO = K-group stride, SBO = MN-group stride)", cudaGetErrorString(e), err, bad, M * N);
            if (e != cudaSuccess) return 1;
            if (!swap && bad) rc = 2;
        }
    return rc;
}
