// peer.cuh — what the data-parallel exchange (peer.cu) and the tcgen05 update (update_tc.cu / update.cu) share.
#pragma once
#include "common.cuh"

namespace erlfill {

struct PeerCtrl {                  // one per rank, at the tail of its peer allocation (zero-initialised)
    uint32_t flag[2][ERLFILL_MAX_PEERS];
    uint32_t epoch[2];
    uint32_t error;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// "this rank's bucket `which` is complete": epoch += 1 (thread 0), then one thread per peer stores the epoch into flag[which][rank] of
// that peer's control block (the system-scope release waits for this GPU's earlier writes to be visible to the peer — about a
// microsecond over NVLink, so the stores go out in parallel).  Called by ALL threads of a block of at least ERLFILL_MAX_PEERS threads.
struct PeerSig {
    int world, rank, which;        // world == 0: no exchange
    PeerCtrl *ctrl[ERLFILL_MAX_PEERS];
    unsigned int *done;            // arrival counter of the signalling kernel's blocks (device memory, zero between launches)
};
__device__ __forceinline__ void peer_signal_block(const PeerSig &s) {
    __shared__ uint32_t s_epoch;
    PeerCtrl *mine = s.ctrl[s.rank];
    if (threadIdx.x == 0) {
        s_epoch = mine->epoch[s.which] + 1;
        mine->epoch[s.which] = s_epoch;
    }
    __syncthreads();
    if ((int)threadIdx.x < s.world) {
        __threadfence_system();
        st_release_sys(&s.ctrl[threadIdx.x]->flag[s.which][s.rank], s_epoch);
    }
}
// the same at the end of a multi-block kernel that produced the bucket: the last block to arrive signals
__device__ __forceinline__ void peer_signal_when_grid_done(const PeerSig &s) {
    if (s.world <= 0) return;
    __shared__ int s_last;
    __threadfence();                                         // this thread's bucket stores before the block's arrival (device scope: the
    __syncthreads();                                         // last block's system-scope fence and release carry them to the peers)
    if (threadIdx.x == 0) s_last = atomicAdd(s.done, 1u) == gridDim.x - 1u ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *s.done = 0u;
    __threadfence();
    peer_signal_block(s);
}

// fast-path (tcgen05 update) exchange geometry: one element per thread, one norm partial per block of 256
constexpr int kPeerFineThreads = 256;
int ddpg_fine_blocks(int which);                             // update.cu
float *ddpg_ssq_fine(void *d_workspace, int batch, int which);
unsigned int *ddpg_peer_done(void *d_workspace, int batch, int which);

}  // namespace erlfill
