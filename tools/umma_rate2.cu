// Development probe 3 (not product): tcgen05.mma issue overhead by code pattern and number of issuing warps.
#include <cstdio>
#include <cstdlib>
#include "umma.cuh"
using namespace erlfill::umma;

template <int V, int KIND, int N>
__global__ void __launch_bounds__(160) issue_rate(int n_outer, int n_issuers, long long *cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = V == 2 ? 0u : __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp < n_issuers) {
        constexpr uint32_t idesc = instr_desc_bf16(128, N);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 8192);
        const uint32_t dbase = tbase + warp * 96;
        const long long t0 = clock64();
        if (V == 1) {
            if (elect_one()) {
                for (int o = 0; o < n_outer; ++o) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        if (KIND == 0) mma_ss(dbase, smem_desc(sa + s * 256, 128, 2048), smem_desc(sb + s * 256, 128, 2048), idesc, 1);
                        else mma_ts(dbase, tbase + 384 + s * 8, smem_desc(sb + s * 256, 128, 2048), idesc, 1);
                    }
                }
                mma_commit(&bars[warp]);
            }
            __syncwarp();
        } else {
            for (int o = 0; o < n_outer; ++o) {
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    if (elect_one()) {
                        if (KIND == 0) mma_ss(dbase, smem_desc(sa + s * 256, 128, 2048), smem_desc(sb + s * 256, 128, 2048), idesc, 1);
                        else mma_ts(dbase, tbase + 384 + s * 8, smem_desc(sb + s * 256, 128, 2048), idesc, 1);
                    }
                    __syncwarp();
                }
            }
            if (elect_one()) mma_commit(&bars[warp]);
            __syncwarp();
        }
        mbar_wait(&bars[warp], 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && tid == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 4) tmem_dealloc(__shfl_sync(0xffffffffu, tmem_slot, 0), 512);
}

template <int V, int KIND, int N>
void run(long long *cyc) {
    cudaFuncSetAttribute(issue_rate<V, KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int n_outer = 1024;
    for (int issuers : {1, 2, 4}) {
        for (int rep = 0; rep < 2; ++rep) issue_rate<V, KIND, N><<<148, 160, 65536>>>(n_outer, issuers, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        printf("V%d %s N=%d issuers=%d: %s %.1f cycles/instr/issuer -> %.1f cycles/instr/SM (floor %d)\n", V, KIND ? "TS" : "SS", N, issuers,
               cudaGetErrorString(e), (double)cyc[0] / (n_outer * 8), (double)cyc[0] / (n_outer * 8) / issuers, N / 2);
        if (e != cudaSuccess) exit(1);
    }
}
int main() {
    long long *cyc; cudaMallocManaged(&cyc, 64);
    run<0, 1, 32>(cyc); run<1, 1, 32>(cyc); run<2, 1, 32>(cyc);
    run<0, 0, 128>(cyc); run<1, 0, 128>(cyc); run<2, 0, 128>(cyc);
    run<2, 0, 256>(cyc); run<2, 1, 64>(cyc);
    return 0;
}
