// Micro-benchmark (development aid, not product): per-SM throughput of the integer multiplies Philox is made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/tput_bench.cu -o tools/bin/tput_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint4 philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, int rounds) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r < rounds) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
    }
    return make_uint4(c0, c1, c2, c3);
}

template <int OP>
__global__ void k(unsigned *out, long long *cyc, unsigned seed, int iters) {
    unsigned a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x * 8 + j;
    unsigned long long acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) {          // 8 independent IMAD.WIDE chains
#pragma unroll
            for (int j = 0; j < 8; ++j) { unsigned long long p = (unsigned long long)a[j] * 0xD2511F53u; a[j] = (unsigned)(p >> 32) ^ (unsigned)p; }
        } else if (OP == 1) {   // IMAD.HI only
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __umulhi(a[j], 0xD2511F53u) + 1u;
        } else if (OP == 2) {   // 32-bit IMAD
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] * 0xD2511F53u + 1u;
        } else if (OP == 3) {   // Philox-10 x2 per iteration
            const uint4 r0 = philox(seed, 7u, a[0], a[1], (unsigned)i, 0u, 10), r1 = philox(seed, 7u, a[0], a[1], (unsigned)i, 1u, 10);
            acc += (unsigned long long)r0.x + r0.y + r0.z + r0.w + r1.x + r1.y + r1.z + r1.w;
        } else if (OP == 4) {   // Philox-10 x2 + the simulator's block sums (4 wide multiplies per block)
            const uint4 r0 = philox(seed, 7u, a[0], a[1], (unsigned)i, 0u, 10), r1 = philox(seed, 7u, a[0], a[1], (unsigned)i, 1u, 10);
            const unsigned ti = a[2];
            acc += (unsigned long long)ti * r0.x + (unsigned long long)ti * r0.y + (unsigned long long)ti * r0.z + (unsigned long long)ti * r0.w;
            acc += (unsigned long long)ti * r1.x + (unsigned long long)ti * r1.y + (unsigned long long)ti * r1.z + (unsigned long long)ti * r1.w;
        } else if (OP == 5) {   // Philox-10 x2 + block sums as ONE multiply per block: ti * (x0 + x1 + x2 + x3)
            const uint4 r0 = philox(seed, 7u, a[0], a[1], (unsigned)i, 0u, 10), r1 = philox(seed, 7u, a[0], a[1], (unsigned)i, 1u, 10);
            acc += (unsigned long long)r0.x + r0.y + r0.z + r0.w + (unsigned long long)r1.x + r1.y + r1.z + r1.w;
        }
    }
    const long long t1 = clock64();
    unsigned r = (unsigned)acc ^ (unsigned)(acc >> 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) r ^= a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[OP] = t1 - t0;
}

int main() {
    unsigned *o; long long *c; cudaMalloc(&o, 148 * 1024 * 4); cudaMallocManaged(&c, 256);
    const int it = 2000;
    const char *names[] = {"IMAD.WIDE+LOP (8 chains)", "IMAD.HI+IADD (8 chains)", "IMAD (8 chains)", "2 x Philox-10", "2 x Philox-10 + 8 wide mul sums",
                           "2 x Philox-10 + add-only sums"};
    const double per_iter[] = {8, 8, 8, 2, 2, 2};
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<0><<<148, warps * 32>>>(o, c, 1, it); k<1><<<148, warps * 32>>>(o, c, 1, it); k<2><<<148, warps * 32>>>(o, c, 1, it);
        k<3><<<148, warps * 32>>>(o, c, 1, it); k<4><<<148, warps * 32>>>(o, c, 1, it); k<5><<<148, warps * 32>>>(o, c, 1, it);
        cudaDeviceSynchronize();
        for (int i = 0; i < 6; ++i)
            printf("warps/SM %2d  %-34s %.2f cycles per op per warp-per-SMSP (=> %.3f ops/clk/SMSP)\n", warps, names[i],
                   (double)c[i] / (it * per_iter[i]) / (warps / 4.0), (it * per_iter[i]) * (warps / 4.0) / (double)c[i]);
    }
    return 0;
}
