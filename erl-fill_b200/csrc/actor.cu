// actor.cu — rollout-side policy kernels.
//
//   erlfill_actor_forward   A1  Actor.forward (DifferentModules/ddpg_agent.py:73-122), optionally fused
//                           with A2: OUNoise.sample + act()'s clip (ddpg_agent.py:348-356, 768-778)
//   erlfill_emotion_update  M1  EmotionModule.update (EmotionModule.py:121-162) + history push (:73-81)
//   erlfill_episode_end     per-env OU reset + sigma anneal on done (ddpg_agent.py:633-634, 762-785)
//
// The actor is one persistent kernel: every CTA stages all 26,216 fp32 parameters in shared memory once
// (transposed to [k][n] so a warp reads 512 contiguous bytes per k), then loops over 64-row tiles with
// register-tiled FFMA.  fp32 throughout: the physical-unit actions feed int() truncation in the simulator,
// so this path stays at fp32 accuracy (51.8 kFLOP/row is <3 % of the rollout step next to the emotion
// transformer); the tensor-core path is reserved for the FFN-dominated transformer.
#include <math.h>

#include "common.cuh"

namespace erlfill {

constexpr int kRows = 96;          // rows per tile: 12 warps share the staged weights (105 KB) — 8 warps per SM issued at 45 %
constexpr int kThreads = 384;      // (a warp owns 8 rows x 128 columns of layers 1-2; 209 KB of shared memory in all)
constexpr int kH = 128;            // hidden width
constexpr int kH3 = 64;

struct ActorArgs {
    erlfill_actor_weights w;
    erlfill_ou ou;
    const float *d_state;
    const float *d_emotion;
    const uint8_t *d_mask1, *d_mask2;
    float *d_scaled, *d_base;
    int n_rows, emotion_broadcast, dropout_mode, has_ou;
    uint32_t call_idx, row_id_base;
    uint64_t seed;
    erlfill_cond_scaling cs;       // n_cond > 0: per-condition scale / offset rows instead of w.scale / w.offset
};

// shared-memory plan (floats)
constexpr int oW1 = 0;                         // [5][128]
constexpr int oW2 = oW1 + 5 * kH;              // [128][128]
constexpr int oW3 = oW2 + kH * kH;             // [128][64]
constexpr int oW4 = oW3 + kH * kH3;            // [64][16] (10 used)
constexpr int oB1 = oW4 + kH3 * 16;
constexpr int oB2 = oB1 + kH;
constexpr int oB3 = oB2 + kH;
constexpr int oB4 = oB3 + kH3;                 // [16]
constexpr int oEW = oB4 + 16;                  // [3][16]
constexpr int oSC = oEW + 48;                  // scale[16], offset[16]
constexpr int oX = oSC + 32;                   // [kRows][8]  state(2) emotion(3)
constexpr int oHa = oX + kRows * 8;            // [kRows][132]
constexpr int kLd = kH + 4;                    // padded row stride (keeps 16 B alignment, breaks conflicts)
constexpr int oHb = oHa + kRows * kLd;         // [kRows][132]
constexpr int kSmemFloats = oHb + kRows * kLd;

__device__ __forceinline__ bool keep_philox(uint32_t k0, uint32_t k1, uint32_t row, uint32_t call, uint32_t col,
                                            uint32_t layer, float p) {
    // one Philox block covers 4 consecutive columns of one row
    const uint4 r = philox4x32_10(k0, k1, row, call, col >> 2, STREAM_DROPOUT + layer);
    const uint32_t x = (col & 3) == 0 ? r.x : ((col & 3) == 1 ? r.y : ((col & 3) == 2 ? r.z : r.w));
    return (float)(x >> 8) * (1.0f / 16777216.0f) >= p;      // keep with probability 1 - p
}

__global__ void __launch_bounds__(kThreads, 1) actor_forward_kernel(const ActorArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    // ---- stage the parameters (transposed) once per CTA ----
    for (int i = tid; i < 5 * kH; i += kThreads) { const int n = i / 5, k = i % 5; sm[oW1 + k * kH + n] = a.w.w_in[i]; }
    for (int i = tid; i < kH * kH; i += kThreads) { const int n = i / kH, k = i % kH; sm[oW2 + k * kH + n] = a.w.w_h2[i]; }
    for (int i = tid; i < kH3 * kH; i += kThreads) { const int n = i / kH, k = i % kH; sm[oW3 + k * kH3 + n] = a.w.w_h5[i]; }
    for (int i = tid; i < kH3 * 16; i += kThreads) { const int k = i / 16, n = i % 16; sm[oW4 + i] = n < 10 ? a.w.w_h7[n * kH3 + k] : 0.f; }
    for (int i = tid; i < kH; i += kThreads) { sm[oB1 + i] = a.w.b_in[i]; sm[oB2 + i] = a.w.b_h2[i]; }
    for (int i = tid; i < kH3; i += kThreads) sm[oB3 + i] = a.w.b_h5[i];
    for (int i = tid; i < 16; i += kThreads) {
        sm[oB4 + i] = i < 10 ? a.w.b_h7[i] : 0.f;
        sm[oSC + i] = i < 10 ? a.w.scale[i] : 0.f;
        sm[oSC + 16 + i] = i < 10 ? a.w.offset[i] : 0.f;
    }
    for (int i = tid; i < 48; i += kThreads) { const int e = i / 16, n = i % 16; sm[oEW + i] = n < 10 ? a.w.emotion_weights[e * 10 + n] : 0.f; }
    __syncthreads();

    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const int n_tiles = (a.n_rows + kRows - 1) / kRows;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * kRows;
        // ---- inputs: x = [state, emotion] ----
        for (int i = tid; i < kRows * 8; i += kThreads) {
            const int r = i >> 3, c = i & 7, row = row0 + r;
            float v = 0.f;
            if (row < a.n_rows) {
                if (c < 2) v = a.d_state[(size_t)row * 2 + c];
                else if (c < 5) v = a.emotion_broadcast ? a.d_emotion[c - 2] : a.d_emotion[(size_t)row * 3 + (c - 2)];
            }
            sm[oX + i] = v;
        }
        __syncthreads();
        // ---- layer 1: 5 -> 128, LeakyReLU(0.1), Dropout ----
        {
            const int cg = tid & 31, rg = tid >> 5;           // 4 cols x 8 rows per thread
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = sm[oB1 + 4 * cg + j];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float4 w = *reinterpret_cast<const float4 *>(&sm[oW1 + k * kH + 4 * cg]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float x = sm[oX + (rg * 8 + i) * 8 + k];
                    acc[i][0] = fmaf(x, w.x, acc[i][0]); acc[i][1] = fmaf(x, w.y, acc[i][1]);
                    acc[i][2] = fmaf(x, w.z, acc[i][2]); acc[i][3] = fmaf(x, w.w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rg * 8 + i, row = row0 + r;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = acc[i][j];
                    v = v > 0.f ? v : 0.1f * v;
                    const int col = 4 * cg + j;
                    if (a.dropout_mode == ERLFILL_DROPOUT_PHILOX) {
                        v = keep_philox(k0, k1, a.row_id_base + row, a.call_idx, col, 0, 0.2f) ? v * 1.25f : 0.f;
                    } else if (a.dropout_mode == ERLFILL_DROPOUT_MASK) {
                        v = (row < a.n_rows && a.d_mask1[(size_t)row * kH + col]) ? v * 1.25f : 0.f;
                    }
                    o[j] = v;
                }
                *reinterpret_cast<float4 *>(&sm[oHa + r * kLd + 4 * cg]) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        // ---- layer 2: 128 -> 128, ReLU, Dropout ----
        {
            const int cg = tid & 31, rg = tid >> 5;
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = sm[oB2 + 4 * cg + j];
            for (int k = 0; k < kH; k += 4) {
                float4 w[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4 *>(&sm[oW2 + (k + kk) * kH + 4 * cg]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 x = *reinterpret_cast<const float4 *>(&sm[oHa + (rg * 8 + i) * kLd + k]);
                    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        acc[i][0] = fmaf(xs[kk], w[kk].x, acc[i][0]); acc[i][1] = fmaf(xs[kk], w[kk].y, acc[i][1]);
                        acc[i][2] = fmaf(xs[kk], w[kk].z, acc[i][2]); acc[i][3] = fmaf(xs[kk], w[kk].w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rg * 8 + i, row = row0 + r;
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = fmaxf(acc[i][j], 0.f);
                    const int col = 4 * cg + j;
                    if (a.dropout_mode == ERLFILL_DROPOUT_PHILOX) {
                        v = keep_philox(k0, k1, a.row_id_base + row, a.call_idx, col, 1, 0.2f) ? v * 1.25f : 0.f;
                    } else if (a.dropout_mode == ERLFILL_DROPOUT_MASK) {
                        v = (row < a.n_rows && a.d_mask2[(size_t)row * kH + col]) ? v * 1.25f : 0.f;
                    }
                    o[j] = v;
                }
                *reinterpret_cast<float4 *>(&sm[oHb + r * kLd + 4 * cg]) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
        // ---- layer 3: 128 -> 64, ReLU (into Ha) ----
        {
            const int cg = tid & 15, rg = tid >> 4;           // 4 cols x 4 rows per thread
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = sm[oB3 + 4 * cg + j];
            for (int k = 0; k < kH; k += 4) {
                float4 w[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4 *>(&sm[oW3 + (k + kk) * kH3 + 4 * cg]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 x = *reinterpret_cast<const float4 *>(&sm[oHb + (rg * 4 + i) * kLd + k]);
                    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        acc[i][0] = fmaf(xs[kk], w[kk].x, acc[i][0]); acc[i][1] = fmaf(xs[kk], w[kk].y, acc[i][1]);
                        acc[i][2] = fmaf(xs[kk], w[kk].z, acc[i][2]); acc[i][3] = fmaf(xs[kk], w[kk].w, acc[i][3]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = rg * 4 + i;
                *reinterpret_cast<float4 *>(&sm[oHa + r * kLd + 4 * cg]) =
                    make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
            }
        }
        __syncthreads();
        // ---- layer 4: 64 -> 10, Softsign, emotion modulation, scaling, (OU noise + clip) ----
        for (int i = tid; i < kRows * 10; i += kThreads) {
            const int r = i / 10, c = i % 10, row = row0 + r;
            if (row >= a.n_rows) continue;
            float z = sm[oB4 + c];
#pragma unroll 8
            for (int k = 0; k < kH3; ++k) z = fmaf(sm[oHa + r * kLd + k], sm[oW4 + k * 16 + c], z);
            const float base = z / (1.0f + fabsf(z));                                  // Softsign
            const float cur = sm[oX + r * 8 + 2], cons = sm[oX + r * 8 + 3], anx = sm[oX + r * 8 + 4];
            float ew = cur * sm[oEW + c];
            ew = fmaf(cons, sm[oEW + 16 + c], ew);
            ew = fmaf(anx, sm[oEW + 32 + c], ew);
            const float raw_effect = tanhf(ew * 1.5f);
            float alpha = 0.1f + 0.9f * anx;
            alpha *= (1.0f - 0.5f * cons);
            alpha *= (1.0f + 0.5f * cur);
            alpha = fminf(fmaxf(alpha, 0.01f), 1.0f);
            const float om = 1.0f - fabsf(base);
            const float sign_flip = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
            const float raw_action = base + alpha * raw_effect * sign_flip * (om * om);
            float scale = sm[oSC + c], offset = sm[oSC + 16 + c];
            if (a.cs.n_cond > 0) {     // the env's own condition scales its action (one agent per condition in the reference)
                const int cid = a.cs.d_cond_id ? a.cs.d_cond_id[row] : (int)((a.cs.row_id_base + (uint32_t)row) % (uint32_t)a.cs.n_cond);
                scale = __ldg(a.cs.d_scale + cid * 10 + c);
                offset = __ldg(a.cs.d_offset + cid * 10 + c);
            }
            const float scaled = raw_action * scale + offset;
            a.d_scaled[(size_t)row * 10 + c] = scaled;
            if (a.d_base) a.d_base[(size_t)row * 10 + c] = base;
            if (a.has_ou) {
                // OUNoise.sample (:774-778) then act()'s clip (:352-356), binary32 like numpy
                float g;
                if (a.ou.d_randn) g = a.ou.d_randn[(size_t)row * 10 + c];
                else {
                    // Box-Muller on Philox lanes: pair p = c/2 uses block p/2, lanes 2*(p%2), 2*(p%2)+1
                    const int p = c >> 1;
                    const uint4 rr = philox4x32_10((uint32_t)a.ou.seed, (uint32_t)(a.ou.seed >> 32),
                                                   a.ou.env_id_base + (uint32_t)row, a.ou.step_idx, (uint32_t)(p >> 1), STREAM_OU);
                    const uint32_t x1 = (p & 1) ? rr.z : rr.x, x2 = (p & 1) ? rr.w : rr.y;
                    const double u1 = ((double)x1 + 1.0) * (1.0 / 4294967296.0);
                    const double u2 = (double)x2 * (1.0 / 4294967296.0);
                    const double rad = sqrt(-2.0 * log(u1)), ang = 6.283185307179586476925286766559 * u2;
                    g = (float)((c & 1) ? rad * sin(ang) : rad * cos(ang));
                }
                float st = a.ou.d_state[(size_t)row * 10 + c];
                // numpy evaluates every product and sum separately in binary32: no FMA contraction here
                float dx = __fmul_rn(0.15f, __fsub_rn(0.0f, st));
                dx = __fadd_rn(dx, __fmul_rn(a.ou.d_sigma[row], g));
                st = __fadd_rn(st, dx);
                a.ou.d_state[(size_t)row * 10 + c] = st;
                float fa = __fadd_rn(scaled, __fmul_rn(st, scale));
                const float lo = __fsub_rn(offset, scale), hi = __fadd_rn(offset, scale);
                fa = fa < lo ? lo : fa;
                fa = fa > hi ? hi : fa;
                a.ou.d_final_action[(size_t)row * 10 + c] = fa;
            }
        }
        __syncthreads();
    }
}

// ---- M1 ------------------------------------------------------------------------------------
__device__ __forceinline__ double clip01(double x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }

__global__ void emotion_update_kernel(int n, const float *d_reward, const float *d_state, const float *d_next_state,
                                      double *d_cur /*[3][n]*/, float *d_hist /*[20][3][n]*/, int32_t *d_len,
                                      int32_t *d_head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double reward = d_reward[i];
    const float2 s = reinterpret_cast<const float2 *>(d_state)[i], s2 = reinterpret_cast<const float2 *>(d_next_state)[i];
    const float mx = s2.x - s.x, my = s2.y - s.y;                  // movement = next_state - state (binary32)
    const double tr = tanh(reward);
    const double mm = sqrt((double)mx * mx + (double)my * my);     // np.linalg.norm
    const double mf = clip01(mm);
    const double exploration = 0.8 * (1 - tr) + 0.2 * mf;
    const double conserv = clip01(tr - 0.1 * mf);
    const double anx = clip01(d_cur[2 * (size_t)n + i] + (-0.2 * tanh(reward * 0.001) + 0.05 * mf));
    const double e0 = clip01(0.8 * d_cur[i] + 0.2 * exploration);
    const double e1 = clip01(0.8 * d_cur[(size_t)n + i] + 0.2 * conserv);
    d_cur[i] = e0; d_cur[(size_t)n + i] = e1; d_cur[2 * (size_t)n + i] = anx;
    // history.append (deque maxlen 20) of the binary32 copy
    int len = d_len[i], head = d_head[i], slot;
    if (len < ERLFILL_HIST) { slot = (head + len) % ERLFILL_HIST; len += 1; }
    else { slot = head; head = (head + 1) % ERLFILL_HIST; }
    d_hist[((size_t)slot * 3 + 0) * n + i] = (float)e0;
    d_hist[((size_t)slot * 3 + 1) * n + i] = (float)e1;
    d_hist[((size_t)slot * 3 + 2) * n + i] = (float)anx;
    d_len[i] = len; d_head[i] = head;
}

// per-episode OU bookkeeping of train_ddpg (ddpg_agent.py:633-634): reset(); anneal()
__global__ void episode_end_kernel(int n, const uint8_t *d_done, float *d_ou_state, float *d_sigma, double *d_sigma64,
                                   long long *d_episode_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool fin = i < n && d_done[i];
    if (d_episode_count) {                              // finished episodes of the batch, one atomic per warp
        const unsigned m = __ballot_sync(0xffffffffu, fin);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(reinterpret_cast<unsigned long long *>(d_episode_count), (unsigned long long)__popc(m));
    }
    if (!fin) return;
#pragma unroll
    for (int c = 0; c < 10; ++c) d_ou_state[(size_t)i * 10 + c] = 0.f;
    const double s = fmax(0.1, d_sigma64[i] * 0.9995);         // sigma is a Python float in the reference
    d_sigma64[i] = s;
    d_sigma[i] = (float)s;
}

static int g_sms = 0;
static int sm_count() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

}  // namespace erlfill

using namespace erlfill;

extern "C" {

int erlfill_actor_forward(const erlfill_actor_weights *w, int n_rows, const float *d_state, const float *d_emotion,
                          int emotion_broadcast, int dropout_mode, uint64_t seed, uint32_t call_idx,
                          uint32_t row_id_base, const uint8_t *d_mask1, const uint8_t *d_mask2, float *d_scaled_action,
                          float *d_base_action, const erlfill_ou *ou, void *stream) {
    return erlfill_actor_forward_cond(w, n_rows, d_state, d_emotion, emotion_broadcast, dropout_mode, seed, call_idx, row_id_base,
                                      d_mask1, d_mask2, d_scaled_action, d_base_action, ou, nullptr, stream);
}

int erlfill_actor_forward_cond(const erlfill_actor_weights *w, int n_rows, const float *d_state, const float *d_emotion,
                               int emotion_broadcast, int dropout_mode, uint64_t seed, uint32_t call_idx,
                               uint32_t row_id_base, const uint8_t *d_mask1, const uint8_t *d_mask2, float *d_scaled_action,
                               float *d_base_action, const erlfill_ou *ou, const erlfill_cond_scaling *cs, void *stream) {
    if (!w || n_rows <= 0 || !d_state || !d_emotion || !d_scaled_action) return ERLFILL_ERR_ARG;
    if (cs && (cs->n_cond <= 0 || !cs->d_scale || !cs->d_offset)) return ERLFILL_ERR_ARG;
    if (!w->w_in || !w->b_in || !w->w_h2 || !w->b_h2 || !w->w_h5 || !w->b_h5 || !w->w_h7 || !w->b_h7 ||
        !w->emotion_weights || !w->scale || !w->offset)
        return ERLFILL_ERR_ARG;
    if (dropout_mode == ERLFILL_DROPOUT_MASK && (!d_mask1 || !d_mask2)) return ERLFILL_ERR_ARG;
    if (ou && (!ou->d_state || !ou->d_sigma || !ou->d_final_action)) return ERLFILL_ERR_ARG;
    ActorArgs a;
    a.w = *w;
    a.has_ou = ou ? 1 : 0;
    if (ou) a.ou = *ou; else a.ou = erlfill_ou{};
    a.d_state = d_state; a.d_emotion = d_emotion; a.d_mask1 = d_mask1; a.d_mask2 = d_mask2;
    a.d_scaled = d_scaled_action; a.d_base = d_base_action;
    a.n_rows = n_rows; a.emotion_broadcast = emotion_broadcast; a.dropout_mode = dropout_mode;
    a.call_idx = call_idx; a.row_id_base = row_id_base; a.seed = seed;
    if (cs) a.cs = *cs; else a.cs = erlfill_cond_scaling{};
    const size_t smem = kSmemFloats * sizeof(float);
    static bool attr_set[64] = {};
    int dev = 0;
    ERLFILL_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        ERLFILL_CUDA_TRY(cudaFuncSetAttribute(actor_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    const int n_tiles = (n_rows + kRows - 1) / kRows;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    actor_forward_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_emotion_update(int n_env, const float *d_reward, const float *d_state, const float *d_next_state,
                           const erlfill_emotion_state *emo, void *stream) {
    if (n_env <= 0 || !d_reward || !d_state || !d_next_state || !emo || !emo->d_current || !emo->d_history ||
        !emo->d_hist_len || !emo->d_hist_head)
        return ERLFILL_ERR_ARG;
    const int block = 128, grid = (n_env + block - 1) / block;
    emotion_update_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(n_env, d_reward, d_state, d_next_state, emo->d_current,
                                                                    emo->d_history, emo->d_hist_len, emo->d_hist_head);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_episode_end(int n_env, const uint8_t *d_done, float *d_ou_state, float *d_sigma, double *d_sigma64,
                        long long *d_episode_count, void *stream) {
    if (n_env <= 0 || !d_done || !d_ou_state || !d_sigma || !d_sigma64) return ERLFILL_ERR_ARG;
    const int block = 128, grid = (n_env + block - 1) / block;
    episode_end_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(n_env, d_done, d_ou_state, d_sigma, d_sigma64, d_episode_count);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
