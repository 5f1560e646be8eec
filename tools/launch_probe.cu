// Development probe: what a launch of the update's phase kernels costs beyond the work of its CTAs.
// An (almost) empty kernel with the same launch shape — 576 threads, 227 KB of dynamic shared memory, optional cluster of 2, optional
// 512-column TMEM allocation — alternating with a small kernel, timed with CUDA events over a CUDA graph of 20 pairs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/launch_probe.cu -o tools/bin/launch_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void small_kernel(float *p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f; }

template <int kTmem, int kClusterSync>
__global__ void __launch_bounds__(576, 1) big_kernel(float *p, int spin) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    if (kTmem) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __syncthreads();
    }
    if (kClusterSync) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (threadIdx.x == 0) smem[0] = 1;
    __syncthreads();
    if (kClusterSync) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (kTmem && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) p[1] = smem[0];
}

template <typename K>
static float run(K kernel, int grid, int smem, int cluster, int spin, float *d, bool with_small) {
    cudaStream_t st; cudaStreamCreate(&st);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 20; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cluster; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = cluster > 1 ? 1 : 0;
        cudaLaunchKernelEx(&cfg, kernel, d, spin);
        if (with_small) small_kernel<<<256, 256, 0, st>>>(d);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(e0, st);
    for (int w = 0; w < 10; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return ms * 1000.f / 200.f;       // microseconds per (big [+ small]) pair
}

int main() {
    float *d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    const int big = 232448 - 1024;
    for (int spin : {0, 40000}) {
        printf("spin %d cycles per CTA:\n", spin);
        printf("  small smem, no cluster,  64 CTAs + small kernel : %6.2f us\n", run(big_kernel<0, 0>, 64, 1024, 1, spin, d, true));
        printf("  227 KB,     no cluster,  64 CTAs + small kernel : %6.2f us\n", run(big_kernel<0, 0>, 64, big, 1, spin, d, true));
        printf("  227 KB,     no cluster,  64 CTAs, no small      : %6.2f us\n", run(big_kernel<0, 0>, 64, big, 1, spin, d, false));
        printf("  227 KB,     cluster 2,  128 CTAs + small kernel : %6.2f us\n", run(big_kernel<0, 1>, 128, big, 2, spin, d, true));
        printf("  227 KB,     cluster 2,  128 CTAs, no small      : %6.2f us\n", run(big_kernel<0, 1>, 128, big, 2, spin, d, false));
        printf("  227 KB + TMEM, cluster 2, 128 CTAs + small      : %6.2f us\n", run(big_kernel<1, 1>, 128, big, 2, spin, d, true));
        printf("  227 KB + TMEM, no cluster, 64 CTAs + small      : %6.2f us\n", run(big_kernel<1, 0>, 64, big, 1, spin, d, true));
        printf("  small kernel alone (x2)                          : %6.2f us\n", run(big_kernel<0, 0>, 1, 1024, 1, 0, d, true));
    }
    return 0;
}
