// Micro-benchmark (development aid, not product): dependent-issue latency of the warp-level ops the simulator uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/lat_bench2.cu -o tools/bin/lat_bench2
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double *out, long long *cyc, double a, double b, int iters) {
    __shared__ double sh[64];
    const int lane = threadIdx.x & 31;
    sh[lane] = b; sh[lane + 32] = b;
    __syncwarp();
    double x = a; unsigned u = (unsigned)iters + lane; int idx = lane;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (OP == 0) x = __dadd_rn(x, b);
            if (OP == 1) u = __reduce_add_sync(0xffffffffu, u) + lane;
            if (OP == 2) u = (unsigned)__reduce_max_sync(0xffffffffu, (int)u) ^ lane;
            if (OP == 3) u = __any_sync(0xffffffffu, u == 77u) + u + 1;
            if (OP == 4) u = __shfl_sync(0xffffffffu, u, (lane + 1) & 31) + 1;
            if (OP == 5) x = __dadd_rn(x, __shfl_sync(0xffffffffu, b, j));                // shfl off the chain
            if (OP == 6) x = __dadd_rn(__shfl_sync(0xffffffffu, x, (lane + 1) & 31), b);  // shfl on the chain
            if (OP == 7) { int v = (int)x; x = (double)(v + 1) + b; }                     // F2I.F64 + I2F.F64 + DADD
            if (OP == 8) { unsigned t; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t)); u = (u ^ t) + 1; }
            if (OP == 9) x = __dadd_rn(x, sh[(j + idx) & 63]);                             // LDS off chain
            if (OP == 10) { idx = (int)sh[idx & 63] + idx; }                               // LDS + F2I on chain
            if (OP == 11) u = __ballot_sync(0xffffffffu, u & 1) + u;
            if (OP == 12) { x = x < b ? __dadd_rn(x, 1.0) : __dadd_rn(x, -1.0); }
            if (OP == 13) { unsigned long long p = (unsigned long long)u * 0xD2511F53u; u = (unsigned)(p >> 32) ^ (unsigned)p; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[OP] = t1 - t0; }
    out[threadIdx.x] = x + u + idx;
}
int main() {
    double *o; long long *c; cudaMalloc(&o, 1024 * 8); cudaMallocManaged(&c, 256);
    const int it = 2048;
#define RUN(n) k<n><<<1, 32>>>(o, c, 1.0, 1e-9, it)
    RUN(0); RUN(1); RUN(2); RUN(3); RUN(4); RUN(5); RUN(6); RUN(7); RUN(8); RUN(9); RUN(10); RUN(11); RUN(12); RUN(13);
    cudaDeviceSynchronize();
    const char *names[] = {"DADD", "REDUX.SUM(+add)", "CREDUX.MAX(+xor)", "VOTE.ANY(+2 add)", "SHFL(+add)", "DADD, shfl off chain",
                           "SHFL64+DADD on chain", "F2I.F64+I2F.F64+DADD", "S2R tid (+xor,add)", "DADD, LDS off chain",
                           "LDS+F2I+IADD chain", "BALLOT(+add)", "DSETP+DADD+sel", "IMAD.WIDE+LOP"};
    for (int i = 0; i < 14; ++i) printf("%-24s %.2f cycles/op\n", names[i], (double)c[i] / (it * 16.0));
    return 0;
}
