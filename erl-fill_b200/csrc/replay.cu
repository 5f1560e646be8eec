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
// insert moves 80 B/transition (+ the 76 B it reads from the step outputs), gather 4 + 80 + 80 B/sample.
#include "common.cuh"

namespace erlfill {

__global__ void replay_insert_kernel(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                                     const float *d_state, const float *d_action, const float *d_reward,
                                     const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                                     const float *d_scale, const float *d_offset, int normalize) {
    // one thread per float of the record: consecutive threads write consecutive addresses
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)n * ERLFILL_REC) return;
    const int i = (int)(g / ERLFILL_REC), f = (int)(g % ERLFILL_REC);
    float v;
    if (f < 2) v = d_state[(size_t)i * 2 + f];
    else if (f < 12) {
        const int c = f - 2;
        v = d_action[(size_t)i * 10 + c];
        if (normalize) {
            v = __fdiv_rn(__fsub_rn(v, d_offset[c]), d_scale[c]);
            v = v < -1.f ? -1.f : v;
            v = v > 1.f ? 1.f : v;
        }
    } else if (f == 12) v = d_reward[i];
    else if (f < 15) v = d_next_state[(size_t)i * 2 + (f - 13)];
    else if (f == 15) v = d_done[i] ? 1.f : 0.f;
    else if (f < 19) v = d_emotion[(size_t)i * 3 + (f - 16)];
    else v = 0.f;
    const int64_t row = d_pos ? d_pos[i] : (start + i) % capacity;
    d_buf[row * ERLFILL_REC + f] = v;
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

}  // namespace erlfill

using namespace erlfill;

extern "C" {

int erlfill_replay_insert(int n, float *d_buf, int64_t capacity, int64_t start, const int64_t *d_pos,
                          const float *d_state, const float *d_action, const float *d_reward,
                          const float *d_next_state, const uint8_t *d_done, const float *d_emotion,
                          const float *d_scale, const float *d_offset, void *stream) {
    if (n <= 0 || !d_buf || capacity <= 0 || !d_state || !d_action || !d_reward || !d_next_state || !d_done || !d_emotion)
        return ERLFILL_ERR_ARG;
    if ((d_scale == nullptr) != (d_offset == nullptr)) return ERLFILL_ERR_ARG;
    if (!d_pos && (start < 0 || n > capacity)) return ERLFILL_ERR_ARG;
    const int64_t total = (int64_t)n * ERLFILL_REC;
    const int block = 256;
    const int64_t grid = (total + block - 1) / block;
    replay_insert_kernel<<<(unsigned)grid, block, 0, (cudaStream_t)stream>>>(n, d_buf, capacity, start, d_pos, d_state, d_action,
                                                                             d_reward, d_next_state, d_done, d_emotion, d_scale,
                                                                             d_offset, d_scale ? 1 : 0);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
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
