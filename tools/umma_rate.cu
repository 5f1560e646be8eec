// Development probe 2 (not product): per-instruction tcgen05.mma rates by shape/operand source, and epilogue variants.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace erlfill::umma;

// kind 0: SS, kind 1: TS.  n_instr MMAs of shape 128 x N x 16 back to back, alternating between two accumulators.
__global__ void __launch_bounds__(128) mma_rate(int kind, int N, int n_instr, long long *cycles, int n_acc, int n_issuers) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (warp >= 1 && warp <= n_issuers) {
        const uint32_t idesc = instr_desc_bf16(128, N);
        const uint32_t dbase = tbase + (warp - 1) * 128;
        const uint64_t a = smem_desc(smem_u32(smem), 128, 256), b = smem_desc(smem_u32(smem + 8192), 128, 256);
        const long long t0 = clock64();
        for (int i = 0; i < n_instr; ++i) {
            const uint32_t d = dbase + (uint32_t)(i & (n_acc - 1)) * (uint32_t)N;
            if (elect_one()) {
                if (kind == 0) mma_ss(d, a, b, idesc, 1);
                else mma_ts(d, tbase + 384 + (i & 7) * 8, b, idesc, 1);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(&bars[warp]);
        __syncwarp();
        mbar_wait(&bars[warp], 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && tid == 32) cycles[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// epilogue variants over a [128 x 128] fp32 tile per step and warpgroup; 8 warps, two warpgroups
// v0: ld32+wait per 32 cols, scalar add (bias smem), cvt.relu, st16          v1: 4 ld32 in flight, one wait
// v2: v1 + add.f32x2                                                          v3: v1 without bias add
__device__ __forceinline__ void add2(float &a0, float &a1, float2 b) {
    asm("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tadd.rn.f32x2 z, x, y;\n\tmov.b64 {%0, %1}, z;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(b.x), "f"(b.y));
}
template <int V>
__global__ void __launch_bounds__(256) epi_rate(int steps, long long *cycles, float *sink) {
    __shared__ __align__(16) float sb[128];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 128) sb[tid] = 0.01f * tid;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tH = tbase + (warp >> 2) * 128 + lane_base, tP = tbase + 256 + (warp >> 2) * 64 + lane_base;
    const long long t0 = clock64();
    for (int it = 0; it < steps; it += 2) {
        if (V == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32], p[16];
                tmem_ld32(tH + c * 32, r); tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) p[j] = pack_relu_bf16x2(__uint_as_float(r[2 * j]) + sb[c * 32 + 2 * j], __uint_as_float(r[2 * j + 1]) + sb[c * 32 + 2 * j + 1]);
                tmem_st16(tP + c * 16, p);
            }
        } else {
            uint32_t r[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(tH + c * 32, r[c]);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t p[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float a0 = __uint_as_float(r[c][2 * j]), a1 = __uint_as_float(r[c][2 * j + 1]);
                    if (V == 1) { a0 += sb[c * 32 + 2 * j]; a1 += sb[c * 32 + 2 * j + 1]; }
                    if (V == 2) add2(a0, a1, *reinterpret_cast<const float2 *>(&sb[c * 32 + 2 * j]));
                    p[j] = pack_relu_bf16x2(a0, a1);
                }
                tmem_st16(tP + c * 16, p);
            }
        }
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads();
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (warp == 0) tmem_dealloc(tbase, 512);
    if (t1 == 0) sink[0] = 1.f;
}

int main() {
    long long *cyc; cudaMallocManaged(&cyc, 64); float *sink; cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int n_instr = 8192;
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {32, 64, 128})
            for (int n_issuers : {1, 2})
            for (int n_acc : {1, 2, 4}) {
                if (n_acc * N > 128) continue;
                for (int rep = 0; rep < 2; ++rep) { mma_rate<<<148, 128, 65536>>>(kind, N, n_instr, cyc, n_acc, n_issuers); }
                cudaError_t e = cudaDeviceSynchronize();
                printf("mma %s 128x%dx16 acc=%d issuers=%d: %s %.1f cycles/instr/issuer (floor %d)\n", kind ? "TS" : "SS", N, n_acc, n_issuers, cudaGetErrorString(e), (double)cyc[0] / n_instr, N / 2);
                if (e != cudaSuccess) return 1;
            }
    const int steps = 4096;
    for (int v = 3; v < 4; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            if (v == 0) epi_rate<0><<<148, 256>>>(steps, cyc, sink);
            if (v == 1) epi_rate<1><<<148, 256>>>(steps, cyc, sink);
            if (v == 2) epi_rate<2><<<148, 256>>>(steps, cyc, sink);
            if (v == 3) epi_rate<3><<<148, 256>>>(steps, cyc, sink);
        }
        cudaError_t e = cudaDeviceSynchronize();
        printf("epilogue v%d: %s %.1f cycles/step (128x128 tile, two warpgroups alternate)\n", v, cudaGetErrorString(e), (double)cyc[0] / steps);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
