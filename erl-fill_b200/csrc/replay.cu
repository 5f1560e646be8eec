// replay.cu — replay buffer kernels (R1/R2).
//
//   erlfill_replay_insert   R1  DDPGAgent.remember + PrioritizedReplayBuffer.add (ddpg_agent.py:359-395, 206-226):
//                               action normalisation clip((a-offset)/scale,-1,1) and the store of
//                               (state, norm_action, reward, next_state, done, emotion)
//   erlfill_replay_sample   R2  index draw.  ER-DDPG's prioritised sampler degenerates to uniform with
//                               replacement, idx = floor(u*n) (all priorities are 1e-4: SURVEY.md section 0.7)
//   erlfill_replay_gather   R2  minibatch assembly (the zip(*batch) + torch.FloatTensor of ddpg_agent.py:434-442)
//
// Layout: one transition = ERLFILL_REC floats (80 B, 16 B aligned):
//   [0:2] state  [2:12] normalised action  [12] reward  [13:15] next_state  [15] done  [16:19] emotion  [19] pad
// The buffer is [capacity][20] row-major, so an insert of N consecutive env transitions is one contiguous
// 80*N-byte write and a gathered sample is five 128-bit loads.  Both kernels are HBM-bandwidth bound:
// insert moves 80 B/transition (+ the 76 B it reads from the step outputs), gather 8 + 80 + 80 B/sample.
#include "common.cuh"

namespace erlfill {

// condition of transition i (0 without a table): column 19 of the record, read back by the update for the actor's scaling
__device__ __forceinline__ int cond_of(const erlfill_cond_scaling &cs, int i) {
    if (cs.n_cond <= 0) return 0;
    return cs.d_cond_id ? cs.d_cond_id[i] : (int)((cs.row_id_base + (uint32_t)i) % (uint32_t)cs.n_cond);
}
// clip((a - offset) / scale, -1, 1): ddpg_agent.py:373-375
__device__ __forceinline__ void normalise_action(float (&a)[10], const float *scale, const float *offset) {
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        float v = __fdiv_rn(__fsub_rn(a[c], __ldg(offset + c)), __ldg(scale + c));
        v = v < -1.f ? -1.f : v;
        a[c] = v > 1.f ? 1.f : v;
    }
}

__global__ void __launch_bounds__(128) replay_insert_kernel(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                                                            const float *d_state, const float *d_action, const float *d_reward,
                                                            const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                                                            const float *d_scale, const float *d_offset, int normalize,
                                                            const erlfill_cond_scaling cs) {
    // One thread per transition: every input byte is fetched exactly once (8-byte loads, rows of a warp are adjacent, so
    // the sectors a load instruction touches are shared by neighbouring lanes and by the next instruction through L1),
    // all 15 loads of a thread are independent and in flight together, and the 80-byte record leaves as five 128-bit
    // stores (L2 merges the five partial-sector writes of a record before anything reaches DRAM).
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 st = __ldg(reinterpret_cast<const float2 *>(d_state) + i);
    const float2 *ap = reinterpret_cast<const float2 *>(d_action) + (size_t)i * 5;
    float a[10];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float2 t = __ldg(ap + j);
        a[2 * j] = t.x; a[2 * j + 1] = t.y;
    }
    const float rw = __ldg(d_reward + i);
    const float2 ns = __ldg(reinterpret_cast<const float2 *>(d_next_state) + i);
    const float dn = __ldg(d_done + i) ? 1.f : 0.f;
    const float e0 = __ldg(d_emotion + (size_t)i * 3), e1 = __ldg(d_emotion + (size_t)i * 3 + 1), e2 = __ldg(d_emotion + (size_t)i * 3 + 2);
    int64_t row;
    if (d_pos) {
        row = d_pos[i];
    } else {
        row = start + i;                       // start < capacity and n <= capacity (checked by the host)
        row = row >= capacity ? row - capacity : row;
    }
    const int cid = cond_of(cs, i);
    if (normalize) normalise_action(a, cs.n_cond > 0 ? cs.d_scale + cid * 10 : d_scale, cs.n_cond > 0 ? cs.d_offset + cid * 10 : d_offset);
    float4 *dst = reinterpret_cast<float4 *>(d_buf + row * ERLFILL_REC);
    dst[0] = make_float4(st.x, st.y, a[0], a[1]);
    dst[1] = make_float4(a[2], a[3], a[4], a[5]);
    dst[2] = make_float4(a[6], a[7], a[8], a[9]);
    dst[3] = make_float4(rw, ns.x, ns.y, dn);
    dst[4] = make_float4(e0, e1, e2, (float)cid);
}

// The ring-order insert (no position table): the records of a CTA are consecutive rows, so they are assembled in shared
// memory (80-byte stride: the five 128-bit stores of a quarter-warp fall on distinct banks) and leave as fully coalesced
// 128-bit stores, 512 contiguous bytes per warp instruction.  A CTA whose rows wrap around the ring end writes directly.
constexpr int kInsertBlock = 128;
__global__ void __launch_bounds__(kInsertBlock) replay_insert_ring_kernel(int n, float *d_buf, int64_t capacity, int64_t start,
                                                                        const float *d_state, const float *d_action,
                                                                        const float *d_reward, const float *d_next_state,
                                                                        const uint8_t *d_done, const float *d_emotion,
                                                                        const float *d_scale, const float *d_offset, int normalize,
                                                                        const erlfill_cond_scaling cs) {
    __shared__ __align__(16) float4 sh[kInsertBlock * 5];
    const int i0 = blockIdx.x * kInsertBlock, i = i0 + threadIdx.x;
    const int cnt = min(kInsertBlock, n - i0);
    int64_t row0 = start + i0;
    row0 = row0 >= capacity ? row0 - capacity : row0;
    const bool contiguous = row0 + cnt <= capacity;
    if (i < n) {
        const float2 st = __ldg(reinterpret_cast<const float2 *>(d_state) + i);
        const float2 *ap = reinterpret_cast<const float2 *>(d_action) + (size_t)i * 5;
        float a[10];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float2 t = __ldg(ap + j);
            a[2 * j] = t.x; a[2 * j + 1] = t.y;
        }
        const float rw = __ldg(d_reward + i);
        const float2 ns = __ldg(reinterpret_cast<const float2 *>(d_next_state) + i);
        const float dn = __ldg(d_done + i) ? 1.f : 0.f;
        const float e0 = __ldg(d_emotion + (size_t)i * 3), e1 = __ldg(d_emotion + (size_t)i * 3 + 1), e2 = __ldg(d_emotion + (size_t)i * 3 + 2);
        const int cid = cond_of(cs, i);
        if (normalize) normalise_action(a, cs.n_cond > 0 ? cs.d_scale + cid * 10 : d_scale, cs.n_cond > 0 ? cs.d_offset + cid * 10 : d_offset);
        float4 *dst = sh + threadIdx.x * 5;
        if (!contiguous) {
            int64_t row = start + i;
            row = row >= capacity ? row - capacity : row;
            dst = reinterpret_cast<float4 *>(d_buf + row * ERLFILL_REC);
        }
        dst[0] = make_float4(st.x, st.y, a[0], a[1]);
        dst[1] = make_float4(a[2], a[3], a[4], a[5]);
        dst[2] = make_float4(a[6], a[7], a[8], a[9]);
        dst[3] = make_float4(rw, ns.x, ns.y, dn);
        dst[4] = make_float4(e0, e1, e2, (float)cid);
    }
    if (!contiguous) return;                      // block-uniform
    __syncthreads();
    float4 *out = reinterpret_cast<float4 *>(d_buf + row0 * ERLFILL_REC);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int q = threadIdx.x + kInsertBlock * k;
        if (q < cnt * 5) out[q] = sh[q];
    }
}

__global__ void replay_sample_kernel(int batch, int64_t n_valid, uint32_t k0, uint32_t k1, uint32_t update_idx,
                                     uint32_t stream_id, int64_t *d_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const uint4 r = philox4x32_10(k0, k1, (uint32_t)(i >> 2), update_idx, stream_id, STREAM_SAMPLE);
    const uint32_t x = (i & 3) == 0 ? r.x : ((i & 3) == 1 ? r.y : ((i & 3) == 2 ? r.z : r.w));
    d_idx[i] = (int64_t)(((unsigned long long)x * (unsigned long long)n_valid) >> 32);   // floor(U * n)
}

__global__ void replay_gather_kernel(int batch, const float *d_buf, const int64_t *d_idx, float *d_out) {
    // thread = (sample, 16-byte part): 5 x LDG.128 per sample, fully coalesced 80*B-byte store
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)batch * (ERLFILL_REC / 4)) return;
    const int64_t s = g / (ERLFILL_REC / 4);
    const int part = (int)(g % (ERLFILL_REC / 4));
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(d_buf) + d_idx[s] * (ERLFILL_REC / 4) + part);
    reinterpret_cast<uint4 *>(d_out)[g] = v;
}

// R2 fused: index draw + minibatch assembly in one launch.  Block = 64 samples x 5 sixteen-byte parts.  Optionally the block also
// leaves the six sums  sum(e_c - 0.5), sum((e_c - 0.5)^2)  of the emotion columns of its 64 records (fixed order inside the block):
// Critic.forward standardises the emotion with BATCH statistics (ddpg_agent.py:176), and the update kernels then only add the blocks'
// sums instead of re-reading the whole minibatch (erlfill_ddpg_hyper::d_stats_partials).
constexpr int kSgSamples = 64;
__global__ void __launch_bounds__(kSgSamples * 5) replay_sample_gather_kernel(int batch, int64_t n_valid, uint32_t k0, uint32_t k1, uint32_t update_idx,
                                                                               uint32_t stream_id, const float *d_buf, int64_t *d_idx, float *d_out,
                                                                               float *d_stats) {
    __shared__ float sh[kSgSamples][6];
    const int sl = threadIdx.x / 5, part = threadIdx.x % 5;
    const int i = blockIdx.x * kSgSamples + sl;
    float4 v = make_float4(0.5f, 0.5f, 0.5f, 0.f);
    if (i < batch) {
        const uint4 r = philox4x32_10(k0, k1, (uint32_t)(i >> 2), update_idx, stream_id, STREAM_SAMPLE);
        const uint32_t x = (i & 3) == 0 ? r.x : ((i & 3) == 1 ? r.y : ((i & 3) == 2 ? r.z : r.w));
        const int64_t row = (int64_t)(((unsigned long long)x * (unsigned long long)n_valid) >> 32);   // floor(U * n)
        if (part == 0 && d_idx) d_idx[i] = row;
        v = __ldg(reinterpret_cast<const float4 *>(d_buf) + row * (ERLFILL_REC / 4) + part);
        reinterpret_cast<float4 *>(d_out)[(size_t)i * (ERLFILL_REC / 4) + part] = v;
    }
    if (!d_stats) return;
    if (part == 4) {
        const float d0 = v.x - 0.5f, d1 = v.y - 0.5f, d2 = v.z - 0.5f;
        sh[sl][0] = d0; sh[sl][1] = d1; sh[sl][2] = d2; sh[sl][3] = d0 * d0; sh[sl][4] = d1 * d1; sh[sl][5] = d2 * d2;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.f;
        for (int q = 0; q < kSgSamples; ++q) t += sh[q][threadIdx.x];
        d_stats[blockIdx.x * 8 + threadIdx.x] = t;
    }
}

}  // namespace erlfill

using namespace erlfill;

extern "C" {

static int replay_insert_impl(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                              const float *d_state, const float *d_action, const float *d_reward,
                              const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                              const float *d_scale, const float *d_offset, const erlfill_cond_scaling *csp, void *stream) {
    if (n <= 0 || !d_buf || capacity <= 0 || !d_state || !d_action || !d_reward || !d_next_state || !d_done || !d_emotion)
        return ERLFILL_ERR_ARG;
    if ((d_scale == nullptr) != (d_offset == nullptr)) return ERLFILL_ERR_ARG;
    erlfill_cond_scaling cs{};
    if (csp) {
        if (csp->n_cond <= 0 || !csp->d_scale || !csp->d_offset) return ERLFILL_ERR_ARG;
        cs = *csp;
    }
    const int normalize = (d_scale || csp) ? 1 : 0;
    if (!d_pos && (start < 0 || n > capacity)) return ERLFILL_ERR_ARG;
    if (((uintptr_t)d_buf & 15u) || ((uintptr_t)d_state & 7u) || ((uintptr_t)d_action & 7u) || ((uintptr_t)d_next_state & 7u))
        return ERLFILL_ERR_ARG;
    const int block = kInsertBlock;
    const int grid = (n + block - 1) / block;
    if (d_pos)
        replay_insert_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(n, d_buf, capacity, start, d_pos, d_state, d_action,
                                                                                 d_reward, d_next_state, d_done, d_emotion, d_scale,
                                                                                 d_offset, normalize, cs);
    else
        replay_insert_ring_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(n, d_buf, capacity, start % capacity, d_state, d_action,
                                                                                      d_reward, d_next_state, d_done, d_emotion,
                                                                                      d_scale, d_offset, normalize, cs);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_replay_insert(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                          const float *d_state, const float *d_action, const float *d_reward,
                          const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                          const float *d_scale, const float *d_offset, void *stream) {
    return replay_insert_impl(n, d_buf, capacity, start, d_pos, d_state, d_action, d_reward, d_next_state, d_done, d_emotion, d_scale,
                              d_offset, nullptr, stream);
}

int erlfill_replay_insert_cond(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                               const float *d_state, const float *d_action, const float *d_reward,
                               const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                               const erlfill_cond_scaling *cs, void *stream) {
    if (!cs) return ERLFILL_ERR_ARG;
    return replay_insert_impl(n, d_buf, capacity, start, d_pos, d_state, d_action, d_reward, d_next_state, d_done, d_emotion, nullptr,
                              nullptr, cs, stream);
}

int erlfill_replay_sample(int batch, int64_t n_valid, uint64_t seed, uint32_t update_idx, uint32_t stream_id,
                          int64_t *d_idx, void *stream) {
    if (batch <= 0 || n_valid <= 0 || n_valid > 0xFFFFFFFFLL || !d_idx) return ERLFILL_ERR_ARG;
    const int block = 256, grid = (batch + block - 1) / block;
    replay_sample_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(batch, n_valid, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                   update_idx, stream_id, d_idx);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_replay_sample_gather_blocks(int batch) { return batch > 0 ? (batch + kSgSamples - 1) / kSgSamples : 0; }

int erlfill_replay_sample_gather(int batch, int64_t n_valid, uint64_t seed, uint32_t update_idx, uint32_t stream_id, const float *d_buf,
                                 int64_t *d_idx, float *d_out, float *d_stats_partials, void *stream) {
    if (batch <= 0 || n_valid <= 0 || n_valid > 0xFFFFFFFFLL || !d_buf || !d_out) return ERLFILL_ERR_ARG;
    if (((uintptr_t)d_buf & 15u) || ((uintptr_t)d_out & 15u)) return ERLFILL_ERR_ARG;
    replay_sample_gather_kernel<<<erlfill_replay_sample_gather_blocks(batch), kSgSamples * 5, 0, (cudaStream_t)stream>>>(
        batch, n_valid, (uint32_t)seed, (uint32_t)(seed >> 32), update_idx, stream_id, d_buf, d_idx, d_out, d_stats_partials);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_replay_gather(int batch, const float *d_buf, const int64_t *d_idx, float *d_out, void *stream) {
    if (batch <= 0 || !d_buf || !d_idx || !d_out) return ERLFILL_ERR_ARG;
    if (((uintptr_t)d_buf & 15u) || ((uintptr_t)d_out & 15u)) return ERLFILL_ERR_ARG;
    const int64_t total = (int64_t)batch * (ERLFILL_REC / 4);
    const int block = 256;
    const int64_t grid = (total + block - 1) / block;
    replay_gather_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(batch, d_buf, d_idx, d_out);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
