// peer.cu — gradient exchange of the data-parallel update over NVLink peer memory, fused with the clip-norm partials.
//
// Every rank's update workspace lives in one cudaMalloc block that the other ranks of the node map with CUDA IPC
// (erlfill_peer_alloc / erlfill_peer_open).  Between the *_GRAD and *_APPLY phases each rank launches
//   peer_signal_kernel   epoch += 1, then one thread per peer stores the epoch into flag[which][my rank] of that rank's control block
//   peer_reduce_kernel   wait until all `world` flags of the own control block reached the epoch, then ONE pass: read the
//                        bucket of every rank in rank order over NVLink (one-shot all-reduce: 8 x 0.42 MB at world 8), average,
//                        write the reduced bucket locally and the 64 sum-of-squares partials clip_grad_norm_ needs
// so the exchange replaces ncclAllReduce + scale + sumsq (3 launches, 2 of them collective-latency bound) and every rank
// adds the same values in the same order: replicas stay bit-identical.  Buckets are only rewritten by the next update's
// *_GRAD phase, which a rank reaches after the OTHER bucket's exchange — i.e. after every peer has finished reading this one
// (DESIGN.md section 7).  A peer that never arrives trips a bounded spin (ERLFILL_PEER_TIMEOUT_S seconds, default 30): never a
// hang, and never a silent divergence either — the sticky error flag is set and the reduced bucket and its norm partials are
// written as NaN, so the APPLY phase poisons the parameters and the update report of THIS rank; callers poll
// erlfill_peer_error (bench.py and the training loops do after every timed region).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "mlp.cuh"

namespace erlfill {

struct PeerCtrl {                  // one per rank, at the tail of its peer allocation (zero-initialised)
    uint32_t flag[2][ERLFILL_MAX_PEERS];
    uint32_t epoch[2];
    uint32_t error;
};

struct PeerArgs {
    int world, rank, which, n, n_norm;
    long long timeout_cycles;
    const float *bucket[ERLFILL_MAX_PEERS];
    PeerCtrl *ctrl[ERLFILL_MAX_PEERS];
    float *out, *ssq;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one thread per peer: the system-scope release of a flag store waits for this GPU's earlier writes to be visible to that peer, about
// a microsecond over NVLink — eight of them back to back in one thread were most of the exchange's cost at world 8
__global__ void peer_signal_kernel(const PeerArgs a) {
    __shared__ uint32_t s_epoch;
    PeerCtrl *mine = a.ctrl[a.rank];
    if (threadIdx.x == 0) {
        s_epoch = mine->epoch[a.which] + 1;
        mine->epoch[a.which] = s_epoch;
    }
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
        __threadfence_system();                              // this rank's gradients (earlier kernels) before the flag
        st_release_sys(&a.ctrl[threadIdx.x]->flag[a.which][a.rank], s_epoch);
    }
}

constexpr int kPeerThreads = 256;      // the thread / slice geometry of mlp::sumsq_kernel: the norm partials add up in the same order as on the NCCL path
constexpr int kPeerChunks = 7;         // elements per thread whose loads are all issued before the first use (the NVLink round trips overlap)
__global__ void __launch_bounds__(kPeerThreads) peer_reduce_kernel(const PeerArgs a) {
    __shared__ float scratch[32];
    PeerCtrl *mine = a.ctrl[a.rank];
    if (threadIdx.x < a.world) {
        const uint32_t e = *reinterpret_cast<volatile uint32_t *>(&mine->epoch[a.which]);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(&mine->flag[a.which][threadIdx.x]) - e) < 0) {
            if (clock64() - t0 > a.timeout_cycles) { atomicExch(&mine->error, 1u); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    // sticky: once a wait timed out the peers' buckets may be stale or half written; nothing computed from them may look valid
    const bool poisoned = *reinterpret_cast<volatile uint32_t *>(&mine->error) != 0u;
    if (poisoned) {
        const float nan = __int_as_float(0x7fc00000);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) a.out[i] = nan;
        if (threadIdx.x == 0) a.ssq[blockIdx.x] = nan;
        return;
    }
    const int per = (a.n_norm + mlp::kSumsqBlocks - 1) / mlp::kSumsqBlocks;      // the slices of mlp::sumsq_kernel
    const int i0 = blockIdx.x * per, i1 = blockIdx.x == mlp::kSumsqBlocks - 1 ? a.n : min(a.n_norm, i0 + per);
    const float inv = 1.0f / (float)a.world;
    float ss = 0.f;
    for (int base = i0 + threadIdx.x; base < i1; base += kPeerChunks * kPeerThreads) {
        float v[kPeerChunks][ERLFILL_MAX_PEERS];
#pragma unroll
        for (int k = 0; k < kPeerChunks; ++k) {
            const int i = base + k * kPeerThreads;
#pragma unroll
            for (int q = 0; q < ERLFILL_MAX_PEERS; ++q) v[k][q] = (i < i1 && q < a.world) ? __ldcg(a.bucket[q] + i) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kPeerChunks; ++k) {
            const int i = base + k * kPeerThreads;
            if (i >= i1) break;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < ERLFILL_MAX_PEERS; ++q) s += v[k][q];                                        // rank order; absent ranks add +0
            s *= inv;
            a.out[i] = s;
            if (i < a.n_norm) ss = fmaf(s, s, ss);
        }
    }
    const float t = mlp::block_sum(ss, scratch);
    if (threadIdx.x == 0) a.ssq[blockIdx.x] = t;
}

}  // namespace erlfill

using namespace erlfill;

extern "C" {

size_t erlfill_peer_ctrl_bytes(void) { return (sizeof(PeerCtrl) + 255) / 256 * 256; }

int erlfill_peer_alloc(size_t bytes, void **dptr, unsigned char handle[64]) {
    if (!dptr || !handle || bytes == 0) return ERLFILL_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    ERLFILL_CUDA_TRY(cudaMalloc(&p, bytes));
    ERLFILL_CUDA_TRY(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    ERLFILL_CUDA_TRY(cudaIpcGetMemHandle(&h, p));
    memcpy(handle, &h, 64);
    *dptr = p;
    return ERLFILL_OK;
}

int erlfill_peer_open(int peer_device, const unsigned char handle[64], void **mapped) {
    if (!handle || !mapped) return ERLFILL_ERR_ARG;
    int dev = 0;
    ERLFILL_CUDA_TRY(cudaGetDevice(&dev));
    if (peer_device != dev) {
        int can = 0;
        ERLFILL_CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, peer_device));
        if (!can) return ERLFILL_ERR_ARG;
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e);
        (void)cudaGetLastError();
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    ERLFILL_CUDA_TRY(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
    return ERLFILL_OK;
}

int erlfill_peer_close(void *mapped) {
    if (mapped) ERLFILL_CUDA_TRY(cudaIpcCloseMemHandle(mapped));
    return ERLFILL_OK;
}

int erlfill_peer_free(void *dptr) {
    if (dptr) ERLFILL_CUDA_TRY(cudaFree(dptr));
    return ERLFILL_OK;
}

int erlfill_peer_error(const erlfill_peer_group *g, int *error_out) {
    if (!g || !error_out || g->rank < 0 || g->rank >= g->world) return ERLFILL_ERR_ARG;
    PeerCtrl c;
    ERLFILL_CUDA_TRY(cudaMemcpy(&c, g->ctrl[g->rank], sizeof(c), cudaMemcpyDeviceToHost));
    *error_out = (int)c.error;
    return ERLFILL_OK;
}

int erlfill_ddpg_peer_launches(void) { return 4; }       // 2 x (signal, reduce) per update, minus nothing: the APPLY phases skip their sumsq kernels

int erlfill_ddpg_peer_allreduce(const erlfill_peer_group *g, int batch, int which, void *stream) {
    if (!g || g->world < 1 || g->world > ERLFILL_MAX_PEERS || g->rank < 0 || g->rank >= g->world || (which != 0 && which != 1) ||
        batch < 2)
        return ERLFILL_ERR_ARG;
    PeerArgs a;
    a.world = g->world; a.rank = g->rank; a.which = which;
    for (int p = 0; p < g->world; ++p) {
        if (!g->ws[p] || !g->ctrl[p]) return ERLFILL_ERR_ARG;
        float *cg, *ag, *cr, *ar, *ssq;
        int cl, al;
        const int rc = erlfill_ddpg_exchange_buffers(g->ws[p], batch, &cg, &cl, &ag, &al, &cr, &ar, &ssq);
        if (rc != ERLFILL_OK) return rc;
        a.bucket[p] = which == 0 ? cg : ag;
        a.ctrl[p] = (PeerCtrl *)g->ctrl[p];
        if (p == g->rank) {
            a.n = which == 0 ? cl : al;
            a.n_norm = which == 0 ? ERLFILL_CRITIC_PARAMS : ERLFILL_ACTOR_PARAMS;
            a.out = which == 0 ? cr : ar;
            a.ssq = ssq + which * mlp::kSumsqBlocks;
        }
    }
    static long long timeout_cycles = 0;
    if (timeout_cycles == 0) {
        const char *e = getenv("ERLFILL_PEER_TIMEOUT_S");
        double sec = e ? atof(e) : 30.0;
        if (!(sec > 0.0)) sec = 30.0;
        timeout_cycles = (long long)(sec * 1.9e9);          // clock64 ticks at about the SM clock
    }
    a.timeout_cycles = timeout_cycles;
    cudaStream_t st = (cudaStream_t)stream;
    peer_signal_kernel<<<1, 32, 0, st>>>(a);
    peer_reduce_kernel<<<mlp::kSumsqBlocks, kPeerThreads, 0, st>>>(a);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
