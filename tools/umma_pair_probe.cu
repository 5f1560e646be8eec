// Development probe (not part of the product): the cross-CTA protocol of the paired ER-DDPG update (csrc/update_tc.cu).
// A cluster of two CTAs shares one 128-row tile.  CTA c owns the output columns [c N/2, (c+1) N/2) of every layer: it runs the MMA for
// its half, and its epilogue threads write their bf16 results into the operand image of BOTH CTAs (st.shared::cluster), fence the
// generic-proxy writes for the async proxy, and arrive (release.cluster) on the `a_ready` mbarrier of BOTH CTAs; the MMA issuer of
// each CTA waits (acquire.cluster) for the 2 x 4 warps and reads the complete image through the async proxy (tcgen05.mma).
// Chain of L layers  X_{l+1} = bf16((X_l W^T) mod-ish)  with small integers (exact), checked against the host; also measures the
// cycles per layer so the cost of the cross-CTA handshake is known.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I erl-fill_b200/csrc -I include tools/umma_pair_probe.cu -o tools/bin/umma_pair_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "umma.cuh"

using namespace erlfill::umma;

constexpr int M = 128, F = 128, NH = F / 2, L = 16;      // F features per layer, each CTA computes NH output columns

static uint32_t off_kmajor(int r, int k, int Kd) { return (r >> 3) * (Kd >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// push `bytes` of this CTA's shared memory to the peer's (async proxy on both sides), completion counted on the peer's mbarrier
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, const void *src, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(smem_u32(src)),
                 "r"(bytes), "r"(bar_cluster) : "memory");
}
template <int kPush>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160) probe(const uint8_t *X0, const uint8_t *Wimg, float *Xout, long long *cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem, *sB = smem + M * F * 2;           // A image [128][F]; B = this CTA's NH weight rows [NH][F]
    __shared__ uint64_t bar_d, bar_a, bar_free;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cr = cluster_rank();
    for (int i = tid; i < M * F * 2 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sA)[i] = reinterpret_cast<const uint4 *>(X0)[i];
    // weight rows [cr NH, +NH): contiguous in the K-major image (row groups of 8)
    for (int i = tid; i < NH * F * 2 / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(Wimg + (size_t)cr * NH * F * 2)[i];
    if (tid == 0) { mbar_init(&bar_d, 1); mbar_init(&bar_a, kPush ? 1 : 8); mbar_init(&bar_free, 2); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(&tmem_slot, 64);
    asm volatile("fence.proxy.async;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tD = tmem_slot;
    const long long t0 = clock64();
    if (warp == 4) {
        if (elect_one()) {
            const uint32_t id = instr_desc_bf16(M, NH);
            for (int l = 0; l < L; ++l) {
                if (l > 0) { mbar_wait_cluster(&bar_a, (l - 1) & 1); tc_fence_after(); }
                for (int s = 0; s < F / 16; ++s)
                    mma_ss(tD, smem_desc(smem_u32(sA) + s * 256, 128, (F / 8) * 128), smem_desc(smem_u32(sB) + s * 256, 128, (F / 8) * 128), id, s > 0);
                mma_commit(&bar_d);
                mma_commit_mc(&bar_free, 3);      // the A image of BOTH CTAs may be overwritten once both MMAs have read it
            }
        }
        __syncwarp();
    } else {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const int row = warp * 32 + lane;
        const uint32_t peerA = mapa(smem_u32(sA), cr ^ 1u), peer_bar = mapa(smem_u32(&bar_a), cr ^ 1u);
        for (int l = 0; l < L; ++l) {
            mbar_wait(&bar_d, l & 1);
            tc_fence_after();
            mbar_wait_cluster(&bar_free, l & 1);
            for (int c = 0; c < NH / 16; ++c) {
                uint32_t r[16];
                tmem_ld16(tD + lane_base + c * 16, r);
                tmem_wait_ld();
                float v[16];
                for (int j = 0; j < 16; ++j) {
                    // keep the values small integers: v -> ((int)v mod 7) - 3
                    const int q = (int)__uint_as_float(r[j]);
                    v[j] = (float)(((q % 7) + 7) % 7 - 3);
                }
                if (l == L - 1) for (int j = 0; j < 16; ++j) Xout[row * F + cr * NH + c * 16 + j] = v[j];
                for (int g = 0; g < 2; ++g) {
                    uint4 q;
                    q.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); q.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                    q.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); q.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                    const uint32_t o = (row >> 3) * (F >> 3) * 128 + ((cr * NH + c * 16 + g * 8) >> 3) * 128 + (row & 7) * 16;
                    *reinterpret_cast<uint4 *>(sA + o) = q;
                    if (!kPush) st_cluster_v4(peerA + o, q);
                }
            }
            if (kPush) {
                // local half complete -> one bulk copy per 8-row group pushes it into the peer's image; the peer's a_ready counts the bytes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const uint32_t half_bytes = (NH / 8) * 128;
                if (tid < M / 8) {
                    const uint32_t o = tid * (F >> 3) * 128 + (cr * NH >> 3) * 128;
                    bulk_s2peer(peerA + o, sA + o, half_bytes, peer_bar);
                }
                if (tid == 0) mbar_arrive_expect_tx(&bar_a, (M / 8) * half_bytes);
            } else {
                asm volatile("fence.proxy.async;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive_cluster(smem_u32(&bar_a)); mbar_arrive_cluster(peer_bar); }
            }
        }
    }
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // the peer may still be writing into this CTA's shared memory
    if (warp == 4) tmem_dealloc(tD, 64);
}

int main() {
    std::vector<float> X(M * F), W(F * F);
    srand(11);
    for (auto &v : X) v = (float)(rand() % 7 - 3);
    for (auto &v : W) v = (float)(rand() % 3 - 1);
    auto put = [](std::vector<uint8_t> &img, uint32_t off, float v) { __nv_bfloat16 b = __float2bfloat16(v); memcpy(&img[off], &b, 2); };
    std::vector<uint8_t> Xi(M * F * 2), Wi(F * F * 2);
    for (int m = 0; m < M; ++m) for (int k = 0; k < F; ++k) put(Xi, off_kmajor(m, k, F), X[m * F + k]);
    for (int n = 0; n < F; ++n) for (int k = 0; k < F; ++k) put(Wi, off_kmajor(n, k, F), W[n * F + k]);
    std::vector<float> cur = X, nxt(M * F);
    for (int l = 0; l < L; ++l) {
        for (int m = 0; m < M; ++m) for (int n = 0; n < F; ++n) {
            float s = 0; for (int k = 0; k < F; ++k) s += cur[m * F + k] * W[n * F + k];
            const int q = (int)s; nxt[m * F + n] = (float)(((q % 7) + 7) % 7 - 3);
        }
        cur = nxt;
    }
    uint8_t *dX, *dW; float *dO; long long *dC;
    cudaMalloc(&dX, Xi.size()); cudaMalloc(&dW, Wi.size()); cudaMalloc(&dO, M * F * 4); cudaMalloc(&dC, 16 * 8);
    cudaMemcpy(dX, Xi.data(), Xi.size(), cudaMemcpyHostToDevice); cudaMemcpy(dW, Wi.data(), Wi.size(), cudaMemcpyHostToDevice);
    const int smem = M * F * 2 + NH * F * 2;
    int rc = 0;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemset(dO, 0, M * F * 4);
        if (rep & 1) probe<1><<<2, 160, smem>>>(dX, dW, dO, dC); else probe<0><<<2, 160, smem>>>(dX, dW, dO, dC);
        printf("%s: ", (rep & 1) ? "bulk push " : "remote st ");
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> O(M * F); long long cyc[2];
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < M * F; ++i) bad += O[i] != cur[i];
        printf("pair chain of %d layers: %s, %d / %d elements differ, %lld / %lld cycles (%.0f per layer)\n", L, cudaGetErrorString(e), bad, M * F, cyc[0], cyc[1],
               (double)cyc[0] / L);
        if (e != cudaSuccess) return 1;
        if (bad) rc = 2;
    }
    return rc;
}
