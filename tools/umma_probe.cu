// Development probe (not part of the product): validates the tcgen05 primitives in csrc/umma.cuh on a real B200
// and measures the raw rates the emotion-transformer FFN pipeline is designed around.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I erl-fill_b200/csrc -I include tools/umma_probe.cu -o tools/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "umma.cuh"

using namespace erlfill::umma;

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

// ---- test 1: H = X W1c^T (SS), P = bf16(relu(H + b1)), Y = P W2c^T (TS) -----------------------------------
__global__ void __launch_bounds__(128) probe_correct(const uint8_t *Ximg, const uint8_t *W1img, const uint8_t *W2img,
                                                     const float *b1, float *Hout, float *Yout) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sX = smem, *sW1 = smem + 8192, *sW2 = smem + 16384;
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192 / 16; i += 128) {
        reinterpret_cast<uint4 *>(sX)[i] = reinterpret_cast<const uint4 *>(Ximg)[i];
        reinterpret_cast<uint4 *>(sW1)[i] = reinterpret_cast<const uint4 *>(W1img)[i];
        reinterpret_cast<uint4 *>(sW2)[i] = reinterpret_cast<const uint4 *>(W2img)[i];
    }
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t tH = tbase, tY = tbase + 256;
    if (tid == 0) {
        const uint32_t idesc1 = instr_desc_bf16(128, 128);
#pragma unroll
        for (int s = 0; s < 2; ++s)
            mma_ss(tH, smem_desc(smem_u32(sX) + s * 256, 128, 512), smem_desc(smem_u32(sW1) + s * 256, 128, 512), idesc1, s > 0);
        mma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tH + lane_base + c * 32, r);
        tmem_wait_ld();
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 32; ++j) Hout[row * 128 + c * 32 + j] = __uint_as_float(r[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            p[j] = pack_relu_bf16x2(__uint_as_float(r[2 * j]) + b1[c * 32 + 2 * j], __uint_as_float(r[2 * j + 1]) + b1[c * 32 + 2 * j + 1]);
        // all 4 chunks of this thread's row are read before columns [16c, 16c+16) are overwritten only for c == 0;
        // for the probe keep it simple: P goes to a separate region (cols 128..191)
        tmem_st16(tbase + 128 + lane_base + c * 16, p);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc2 = instr_desc_bf16(128, 32);
#pragma unroll
        for (int s = 0; s < 8; ++s)
            mma_ts(tY, tbase + 128 + s * 8, smem_desc(smem_u32(sW2) + s * 256, 128, 2048), idesc2, s > 0);
        mma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after();
    {
        uint32_t r[32];
        tmem_ld32(tY + lane_base, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) Yout[row * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- test 2: raw rates -------------------------------------------------------------------------------------
// mode 0: MMA only (2 SS N=128 + 8 TS N=32 per step);  mode 1: tcgen05.ld only (128 cols/step/thread);
// mode 2: ld + bias + relu/cvt + st (the FFN epilogue);  mode 3: mode 0 and mode 2 concurrently (no dependency)
__global__ void __launch_bounds__(288) probe_rate(int mode, int steps, long long *cycles, float *sink, const float *b1) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 24576 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0 && (mode == 0 || mode == 3)) {
            const uint32_t idesc1 = instr_desc_bf16(128, 128), idesc2 = instr_desc_bf16(128, 32);
            for (int it = 0; it < steps; ++it) {
                const uint32_t tH = tbase + (it & 1) * 128;
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    mma_ss(tH, smem_desc(smem_u32(smem) + s * 256, 128, 512), smem_desc(smem_u32(smem + 8192) + s * 256, 128, 512), idesc1, s > 0);
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    mma_ts(tbase + 384, tbase + 256 + s * 8, smem_desc(smem_u32(smem + 16384) + s * 256, 128, 2048), idesc2, 1);
            }
            mma_commit(&bar);
            mbar_wait(&bar, 0);
        }
    } else if (mode >= 1) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tH = tbase + (warp >> 2) * 128, tP = tbase + 256 + (warp >> 2) * 64;   // races with the MMA are harmless here (rates only)
        float acc = 0.f;
        for (int it = 0; it < steps; it += 2) {      // two warpgroups alternate steps
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(tH + lane_base + c * 32, r);
                tmem_wait_ld();
                if (mode == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
                } else {
                    uint32_t p[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        p[j] = pack_relu_bf16x2(__uint_as_float(r[2 * j]) + b1[c * 32 + 2 * j], __uint_as_float(r[2 * j + 1]) + b1[c * 32 + 2 * j + 1]);
                    tmem_st16(tP + lane_base + c * 16, p);
                }
            }
            if (mode != 1) tmem_wait_st();
        }
        if (acc == 123.f) sink[tid] = acc;
    }
    tc_fence_before();
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[mode] = t1 - t0;
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    // ---------- test 1 ----------
    std::vector<float> X(128 * 32), W1(128 * 32), W2(32 * 128), b1(128);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto &v : X) v = bf16_round(rnd());
    for (auto &v : W1) v = bf16_round(rnd() * 0.3f);
    for (auto &v : W2) v = bf16_round(rnd() * 0.3f);
    for (auto &v : b1) v = rnd() * 0.2f;
    std::vector<uint8_t> Ximg(8192), W1img(8192), W2img(8192);
    auto put = [](std::vector<uint8_t> &img, uint32_t off, float v) {
        __nv_bfloat16 b = __float2bfloat16(v);
        memcpy(&img[off], &b, 2);
    };
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 32; ++k) { put(Ximg, kmajor_off(r, k, 32), X[r * 32 + k]); put(W1img, kmajor_off(r, k, 32), W1[r * 32 + k]); }
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 128; ++k) put(W2img, kmajor_off(n, k, 128), W2[n * 128 + k]);
    uint8_t *dX, *dW1, *dW2; float *db1, *dH, *dY;
    cudaMalloc(&dX, 8192); cudaMalloc(&dW1, 8192); cudaMalloc(&dW2, 8192); cudaMalloc(&db1, 512);
    cudaMalloc(&dH, 128 * 128 * 4); cudaMalloc(&dY, 128 * 32 * 4);
    cudaMemcpy(dX, Ximg.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(dW1, W1img.data(), 8192, cudaMemcpyHostToDevice);
    cudaMemcpy(dW2, W2img.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(db1, b1.data(), 512, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576);
    probe_correct<<<1, 128, 24576>>>(dX, dW1, dW2, db1, dH, dY);
    cudaError_t e = cudaDeviceSynchronize();
    printf("probe_correct: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> H(128 * 128), Y(128 * 32);
    cudaMemcpy(H.data(), dH, H.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double eh = 0, ey = 0;
    for (int r = 0; r < 128; ++r) {
        float P[128];
        for (int n = 0; n < 128; ++n) {
            double s = 0;
            for (int k = 0; k < 32; ++k) s += (double)X[r * 32 + k] * W1[n * 32 + k];
            eh = fmax(eh, fabs(s - H[r * 128 + n]));
            P[n] = bf16_round(fmaxf(H[r * 128 + n] + b1[n], 0.f));
        }
        for (int o = 0; o < 32; ++o) {
            double s = 0;
            for (int n = 0; n < 128; ++n) s += (double)P[n] * W2[o * 128 + n];
            ey = fmax(ey, fabs(s - Y[r * 32 + o]));
        }
    }
    printf("GEMM1 (SS, no-swizzle K-major) max abs err %.3e ; GEMM2 (TS, A in TMEM) max abs err %.3e\n", eh, ey);
    printf("H[0][0..3] = %f %f %f %f ; Y[5][0..3] = %f %f %f %f\n", H[0], H[1], H[2], H[3], Y[160], Y[161], Y[162], Y[163]);

    // ---------- test 2 ----------
    long long *cyc; cudaMallocManaged(&cyc, 64); float *sink; cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576);
    const int steps = 4096;
    const char *names[] = {"MMA only (2xSS N128 + 8xTS N32 per step)", "tcgen05.ld only (8 warps, 128 cols/step)",
                           "epilogue ld+bias+relu/cvt+st (8 warps alternate)", "MMA + epilogue concurrently"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        probe_rate<<<148, 288, 24576>>>(mode, steps, cyc, sink, db1);
        cudaEventRecord(a);
        probe_rate<<<148, 288, 24576>>>(mode, steps, cyc, sink, db1);
        cudaEventRecord(b);
        e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        printf("mode %d %-52s: %s  %.1f cycles/step (CTA 0), %.3f ms for 148 CTAs x %d steps\n", mode, names[mode],
               cudaGetErrorString(e), (double)cyc[mode] / steps, ms, steps);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
