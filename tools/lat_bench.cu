// Micro-benchmark (development aid): dependent-issue latency of FP64/INT ops on one warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, long long *cyc, double a, double b, int iters) {
    double x = a; unsigned u = (unsigned)iters; unsigned long long p;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (OP == 0) x = __dadd_rn(x, b);
            if (OP == 1) x = __dmul_rn(x, b);
            if (OP == 2) x = __fma_rn(x, b, a);
            if (OP == 3) { p = (unsigned long long)u * 0xD2511F53u; u = (unsigned)(p >> 32) ^ (unsigned)p; }
            if (OP == 4) { u = (u ^ 0x9E3779B9u) + 7u; }
            if (OP == 5) { x = (x < b) ? x + 1.0 : x - 1.0; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[OP] = t1 - t0; }
    out[threadIdx.x] = x + u;
}
int main() {
    double *o; long long *c; cudaMalloc(&o, 1024 * 8); cudaMallocManaged(&c, 64);
    const int it = 4096;
    k<0><<<1, 32>>>(o, c, 1.0, 1e-9, it); k<1><<<1, 32>>>(o, c, 1.0, 1.0000001, it);
    k<2><<<1, 32>>>(o, c, 1.0, 0.999, it); k<3><<<1, 32>>>(o, c, 1.0, 1.0, it);
    k<4><<<1, 32>>>(o, c, 1.0, 1.0, it); k<5><<<1, 32>>>(o, c, 1.0, 5.0, it);
    cudaDeviceSynchronize();
    const char *names[] = {"DADD", "DMUL", "DFMA", "IMAD.WIDE+LOP", "LOP+IADD", "DSETP+DADD+sel"};
    for (int i = 0; i < 6; ++i) printf("%s: %.2f cycles/op\n", names[i], (double)c[i] / (it * 16.0));
    // throughput: many warps
    return 0;
}
