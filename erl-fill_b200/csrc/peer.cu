// peer.cu — gradient exchange of the data-parallel update over NVLink peer memory, fused with the clip-norm partials.
//
// Every rank's update workspace lives in one cudaMalloc block that the other ranks of the node map with CUDA IPC
// (erlfill_peer_alloc / erlfill_peer_open).  Between the *_GRAD and *_APPLY phases each rank launches
//   peer_signal_kernel   epoch += 1, then one thread per peer stores the epoch into flag[which][my rank] of that rank's control block
//   peer_reduce_kernel   wait until all `world` flags of the own control block reached the epoch, then ONE pass: read the
//                        bucket of every rank in rank order over NVLink (one-shot all-reduce: 8 x 0.42 MB at world 8), average,
//                        write the reduced bucket locally and the 64 sum-of-squares partials clip_grad_norm_ needs
// (the tcgen05 update raises the flags from the last block of its grad_reduce_kernel instead — peer.cuh — and uses
//   peer_reduce_fine_kernel  the same with one element per thread and one norm partial per 256 elements: erlfill_ddpg_peer_allreduce_fast)
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
#include "peer.cuh"

namespace erlfill {

struct PeerArgs {
    int world, rank, which, n, n_norm;
    long long timeout_cycles;
    const float *bucket[ERLFILL_MAX_PEERS];
    PeerCtrl *ctrl[ERLFILL_MAX_PEERS];
    float *out, *ssq;
};

// one thread per peer: the system-scope release of a flag store waits for this GPU's earlier writes to be visible to that peer, about
// a microsecond over NVLink — eight of them back to back in one thread were most of the exchange's cost at world 8
__global__ void peer_signal_kernel(const PeerSig s) { peer_signal_block(s); }

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

// The fast path's reduce: one element per thread (all ranks' loads in flight at once), one norm partial per block — the tcgen05
// update's apply kernels add any number of partials.
__global__ void __launch_bounds__(kPeerFineThreads) peer_reduce_fine_kernel(const PeerArgs a) {
    __shared__ float scratch[32];
    PeerCtrl *mine = a.ctrl[a.rank];
    if (threadIdx.x < a.world) {
        const uint32_t e = *reinterpret_cast<volatile uint32_t *>(&mine->epoch[a.which]);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(&mine->flag[a.which][threadIdx.x]) - e) < 0) {
            if (clock64() - t0 > a.timeout_cycles) { atomicExch(&mine->error, 1u); break; }
            __nanosleep(40);
        }
    }
    __syncthreads();
    const int i = blockIdx.x * kPeerFineThreads + threadIdx.x;
    if (*reinterpret_cast<volatile uint32_t *>(&mine->error) != 0u) {       // sticky: see peer_reduce_kernel
        const float nan = __int_as_float(0x7fc00000);
        if (i < a.n) a.out[i] = nan;
        if (threadIdx.x == 0) a.ssq[blockIdx.x] = nan;
        return;
    }
    float ss = 0.f;
    if (i < a.n) {
        float v[ERLFILL_MAX_PEERS];
#pragma unroll
        for (int q = 0; q < ERLFILL_MAX_PEERS; ++q) v[q] = q < a.world ? __ldcg(a.bucket[q] + i) : 0.f;
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < ERLFILL_MAX_PEERS; ++q) s += v[q];                  // rank order; absent ranks add +0
        s *= 1.0f / (float)a.world;
        a.out[i] = s;
        if (i < a.n_norm) ss = s * s;
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

int erlfill_ddpg_peer_launches(int fast) { return fast ? 2 : 4; }       // exact: 2 x (signal, reduce); fast: 2 x reduce (the signal rides on grad_reduce_kernel)

static int peer_allreduce_impl(const erlfill_peer_group *g, int batch, int which, bool fine, bool signalled, void *stream) {
    if (!g || g->world < 1 || g->world > ERLFILL_MAX_PEERS || g->rank < 0 || g->rank >= g->world || (which != 0 && which != 1) ||
        batch < 2)
        return ERLFILL_ERR_ARG;
    PeerArgs a;
    PeerSig sig;
    a.world = sig.world = g->world; a.rank = sig.rank = g->rank; a.which = sig.which = which;
    sig.done = nullptr;
    for (int p = 0; p < g->world; ++p) {
        if (!g->ws[p] || !g->ctrl[p]) return ERLFILL_ERR_ARG;
        float *cg, *ag, *cr, *ar, *ssq;
        int cl, al;
        const int rc = erlfill_ddpg_exchange_buffers(g->ws[p], batch, &cg, &cl, &ag, &al, &cr, &ar, &ssq);
        if (rc != ERLFILL_OK) return rc;
        a.bucket[p] = which == 0 ? cg : ag;
        a.ctrl[p] = sig.ctrl[p] = (PeerCtrl *)g->ctrl[p];
        if (p == g->rank) {
            a.n = which == 0 ? cl : al;
            a.n_norm = which == 0 ? ERLFILL_CRITIC_PARAMS : ERLFILL_ACTOR_PARAMS;
            a.out = which == 0 ? cr : ar;
            a.ssq = fine ? ddpg_ssq_fine(g->ws[p], batch, which) : ssq + which * mlp::kSumsqBlocks;
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
    if (!signalled) peer_signal_kernel<<<1, 32, 0, st>>>(sig);
    if (fine) peer_reduce_fine_kernel<<<ddpg_fine_blocks(which), kPeerFineThreads, 0, st>>>(a);
    else peer_reduce_kernel<<<mlp::kSumsqBlocks, kPeerThreads, 0, st>>>(a);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_ddpg_peer_allreduce(const erlfill_peer_group *g, int batch, int which, void *stream) {
    return peer_allreduce_impl(g, batch, which, false, false, stream);
}
int erlfill_ddpg_peer_allreduce_fast(const erlfill_peer_group *g, int batch, int which, void *stream) {
    return peer_allreduce_impl(g, batch, which, true, true, stream);
}

}  // extern "C"
