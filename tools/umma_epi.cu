// Development probe 4 (not product): throughput of the FFN hidden-activation epilogue per SM, by variant and
// number of epilogue warps:  TMEM (fp32 H chunk) -> registers -> relu/bf16 pack -> TMEM (bf16 A operand).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I erl-fill_b200/csrc tools/umma_epi.cu -o tools/umma_epi
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"
using namespace erlfill::umma;

// variant 0: ld only; 1: ld + cvt.relu + st; 2: ld + bias(LDS.128 broadcast) + cvt.relu + st; 3: st only;
// 4: ld x32 + cvt + st but wait::ld once per 64 columns (two loads in flight)
template <int V>
__global__ void __launch_bounds__(512) epi_rate(int steps, long long *cycles, float *sink) {
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias[64];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 64) bias[tid] = 0.01f * tid;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tH = tbase + lane_base + (warp >> 2) * 96, tP = tH + 64;   // 64 fp32 columns in, 32 packed columns out
    const long long t0 = clock64();
    float acc = 0.f;
    for (int it = 0; it < steps; ++it) {
        if (V == 4) {
            uint32_t r0[32], r1[32], p[16];
            tmem_ld32(tH, r0);
            tmem_ld32(tH + 32, r1);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = pack_relu_bf16x2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
            tmem_st16(tP, p);
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = pack_relu_bf16x2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
            tmem_st16(tP + 16, p);
            tmem_wait_st();
            continue;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32], p[16];
            if (V != 3) {
                tmem_ld32(tH + c * 32, r);
                tmem_wait_ld();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = it + j;
            }
            if (V == 0) { acc += __uint_as_float(r[0] ^ r[31]); continue; }
            if (V == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = *reinterpret_cast<const float4 *>(&bias[c * 32 + 4 * j]);
                    r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b.x);
                    r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b.y);
                    r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b.z);
                    r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b.w);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = pack_relu_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            tmem_st16(tP + c * 16, p);
        }
        if (V != 0) tmem_wait_st();
    }
    if (acc == 123.f) sink[tid] = acc;
    const long long t1 = clock64();
    tc_fence_before(); __syncthreads();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int V>
void run(long long *cyc, float *sink, const char *name) {
    const int steps = 8192;
    for (int warps : {4, 8, 12, 16}) {
        for (int rep = 0; rep < 2; ++rep) epi_rate<V><<<148, warps * 32, 0>>>(steps, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        const double cps = (double)cyc[0] / steps;
        printf("%-44s warps=%2d: %s %.1f cycles per 32x64 warp-chunk -> %.1f elem/cycle/SM -> %.0f cycles per 128x2048 tile-layer\n", name,
               warps, cudaGetErrorString(e), cps, warps * 2048.0 / cps, 262144.0 / (warps * 2048.0 / cps));
        if (e != cudaSuccess) exit(1);
    }
}
int main() {
    long long *cyc; cudaMallocManaged(&cyc, 64);
    float *sink; cudaMalloc(&sink, 4096);
    run<0>(cyc, sink, "ld only");
    run<3>(cyc, sink, "st only");
    run<1>(cyc, sink, "ld + cvt.relu.bf16x2 + st");
    run<4>(cyc, sink, "ld x2 in flight + cvt.relu.bf16x2 + st");
    run<2>(cyc, sink, "ld + bias(LDS.128) + cvt.relu.bf16x2 + st");
    return 0;
}
