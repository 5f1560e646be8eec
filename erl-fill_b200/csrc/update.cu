// update.cu — ER-DDPG minibatch update (U1-U5): DDPGAgent.learn, DifferentModules/ddpg_agent.py:422-512.
//
//   U1  Critic.forward (:163-183) incl. the batch standardisation of the emotion columns
//   U2  emotion-driven learning rates (:447-462)
//   U3  target y = 0.01 r + (1-d) 0.99 Q'(s', mu'(s', e'), e'), SmoothL1, clip 5.0, Adam (:465-478)
//   U4  actor loss -mean Q(s, mu(s,e), e) - 0.03 mean_rows(std(a)) + 1e-4 sum ||theta||_2, clip, Adam (:480-499)
//   U5  soft target update tau = 0.01 (:501-507)
//
// The reference quirks are kept: the critic is fed normalised actions for Q(s,a) but PHYSICAL-unit actions
// from the actors for the target and the actor loss; the target uses the agent's
// fresh emotion broadcast over the batch, whose batch standardisation is (e-mean)/(0+1e-6) of identical
// rows: pure summation noise in torch (measured +-0.36), defined here as exactly 0.
//
// Structure: parameters, gradients and Adam moments are flat fp32 buffers in the reference's parameters() order, so
// clip / Adam / soft update / the data-parallel all-reduce are single passes.  Each forward and each input-gradient chain
// is ONE launch (fused_mlp.cuh: a CTA owns 32 minibatch rows, activations in shared memory, weights streamed from L2
// through a cp.async ring, mma.sync TF32 with 3xTF32 compensation); the weight gradients of a network are one grouped
// split-K GEMM launch plus one grouped reduction that also folds the per-CTA bias / LayerNorm / head partial sums.
// 24 launches per update (70 before the fusion); every reduction is ordered and atomics-free, so an update is
// bit-reproducible.  The update stays latency bound at B = 4096 (5.9 GFLOP): see DESIGN.md section 4.5.
#include <math.h>

#include "common.cuh"
#include "ddpg_dropout.cuh"
#include "mlp.cuh"
#include "fused_mlp.cuh"
#include "peer.cuh"

namespace erlfill {
using namespace mlp;
using namespace fused;
using ddrop::Drop;
using ddrop::keep1;
using ddrop::kDropScale;

// ---- flat layouts (parameters() order) -------------------------------------------------------
namespace AO { constexpr int EW = 0, W1 = 30, B1 = 670, W2 = 798, B2 = 17182, W3 = 17310, B3 = 25502, W4 = 25566, B4 = 26206, TOTAL = 26216; }
namespace CO { constexpr int W0 = 0, B0 = 3840, G1 = 4096, BE1 = 4352, W4 = 4608, B4 = 70144, G2 = 70400, BE2 = 70656, W8 = 70912, B8 = 103680, W10 = 103808, B10 = 103936, TOTAL = 103937; }
constexpr int kActorTensors = 9;
__constant__ int c_actor_seg[kActorTensors + 1] = {0, 30, 670, 798, 17182, 17310, 25502, 25566, 26206, 26216};

constexpr int kCriticGradPad = 103968;     // critic grads + 3 emotion means + padding (16 B multiple)
constexpr int kActorGradPad = 26240;

// stats[0:3] mean, stats[3:6] unbiased std of the emotion columns of the batch (cols 16..18); also the
// three means into mean_out (tail of the critic gradient bucket, so a DP all-reduce averages them too).
__global__ void __launch_bounds__(1024) emotion_stats_kernel(int B, const float *batch, float *stats, float *mean_out) {
    __shared__ float scratch[32];
    float s[3] = {0.f, 0.f, 0.f}, mean[3], q[3] = {0.f, 0.f, 0.f};
    for (int r = threadIdx.x; r < B; r += blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] += batch[(size_t)r * ERLFILL_REC + 16 + c];
#pragma unroll
    for (int c = 0; c < 3; ++c) mean[c] = block_sum(s[c], scratch) / (float)B;
    for (int r = threadIdx.x; r < B; r += blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float d = batch[(size_t)r * ERLFILL_REC + 16 + c] - mean[c]; q[c] = fmaf(d, d, q[c]); }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float var = block_sum(q[c], scratch) / (float)(B - 1);
        if (threadIdx.x == 0) { stats[c] = mean[c]; stats[3 + c] = sqrtf(var); mean_out[c] = mean[c]; }
    }
}

// y = 0.01 r + (1-d) gamma Q'; SmoothL1(beta=1) mean loss and dq = dL/dq; loss via deterministic block reduce
__global__ void __launch_bounds__(1024) critic_loss_kernel(int B, const float *batch, const float *q, const float *tq, float gamma,
                                                           float *dq, float *loss_out) {
    __shared__ float scratch[32];
    float ls = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) {
        const float *row = batch + (size_t)r * ERLFILL_REC;
        const float y = row[12] * 0.01f + (1.0f - row[15]) * gamma * tq[r];
        const float d = q[r] - y, ad = fabsf(d);
        ls += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
        float g = ad < 1.0f ? d : (d > 0.f ? 1.f : -1.f);
        if (fabsf(q[r]) >= 1e6f) g = 0.f;                     // clamp blocks the gradient outside (-1e6, 1e6)
        dq[r] = g / (float)B;
    }
    const float t = block_sum(ls, scratch);
    if (threadIdx.x == 0) loss_out[0] = t / (float)B;
}

// actor loss value (reported): -mean(q2) - 0.03*mean_rows(std) + 1e-4 * sum_t ||theta_t||.  Block 0: the two batch means
// into loss_out[0]; block 1 + t: ||theta_t|| into tensor_norms[t] (needed by the L2 gradient); report_kernel adds them up.
__global__ void __launch_bounds__(1024) actor_loss_kernel(int B, const float *q2, const float *out, const float *actor, float *loss_out,
                                                          float *tensor_norms /*[9]*/) {
    __shared__ float scratch[32];
    if (blockIdx.x == 0) {
        float s = 0.f, d = 0.f;
        for (int r = threadIdx.x; r < B; r += blockDim.x) {
            s += q2[r];
            float mean = 0.f;
            for (int c = 0; c < 10; ++c) mean += out[(size_t)r * 10 + c];
            mean *= 0.1f;
            float var = 0.f;
            for (int c = 0; c < 10; ++c) { const float x = out[(size_t)r * 10 + c] - mean; var = fmaf(x, x, var); }
            d += sqrtf(var / 9.0f);
        }
        const float qs = block_sum(s, scratch);
        const float ds = block_sum(d, scratch);
        if (threadIdx.x == 0) loss_out[0] = -qs / (float)B - 0.03f * ds / (float)B;
    } else {
        const int t = blockIdx.x - 1;
        float q = 0.f;
        for (int i = c_actor_seg[t] + threadIdx.x; i < c_actor_seg[t + 1]; i += blockDim.x) q = fmaf(actor[i], actor[i], q);
        const float nrm = sqrtf(block_sum(q, scratch));
        if (threadIdx.x == 0) tensor_norms[t] = nrm;
    }
}
// grad += 1e-4 * theta / ||theta_tensor||
__global__ void actor_l2_grad_kernel(const float *actor, const float *tensor_norms, float *grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= AO::TOTAL) return;
    int t = 0;
    while (i >= c_actor_seg[t + 1]) ++t;
    grad[i] += 1e-4f * actor[i] / tensor_norms[t];
}

// ---- optimiser -------------------------------------------------------------------------------
// per-update scalars in device memory so that a captured CUDA graph of the update can be replayed for every step:
// step[0..1] = (double) lr_decay_rate ** train_step, step[2] = 1 - 0.9^t, step[3] = sqrt(1 - 0.999^t)
__global__ void set_step_kernel(float *step, double decay, float bc1, float bc2_sqrt) {
    *reinterpret_cast<double *>(step) = decay;
    step[2] = bc1; step[3] = bc2_sqrt;
}
__global__ void lr_kernel(const float *means, const float *step, double critic_lr, double actor_lr, float *lrs /*[2]*/) {
    const double decay = *reinterpret_cast<const double *>(step);
    const double cur = means[0], cons = means[1], anx = means[2];
    const double af = 1.0 + 0.8 * cur - 0.4 * cons + 0.2 * anx, cf = 1.0 - 0.3 * anx + 0.6 * cons;
    const double c = fmin(fmax(critic_lr * cf * decay, critic_lr * 0.1), critic_lr * 3);
    const double a = fmin(fmax(actor_lr * af * decay, actor_lr * 0.1), actor_lr * 3);
    lrs[0] = (float)c; lrs[1] = (float)a;
}
// clip_grad_norm_(5.0) + Adam(0.9, 0.999, 1e-8) + soft target update, one pass over the flat buffers
__global__ void adam_soft_kernel(int n, float *p, float *pt, float *m, float *v, const float *g, const float *sumsq,
                                 const float *lr_ptr, float max_norm, const float *step, float tau) {
    const float bc1 = step[2], bc2_sqrt = step[3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ss = 0.f;
#pragma unroll 8
    for (int b = 0; b < kSumsqBlocks; ++b) ss += sumsq[b];
    const float total = sqrtf(ss);
    const float coef = fminf(max_norm / (total + 1e-6f), 1.0f);
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * 0.1f;                             // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * 0.999f + 0.001f * gi * gi;                       // mul_(beta2).addcmul_(g, g, 1-beta2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + 1e-8f;
    const float pi = p[i] - (lr_ptr[0] / bc1) * (mi / denom);
    p[i] = pi;
    if (pt) pt[i] = tau * pi + (1.0f - tau) * pt[i];
}

__global__ void soft_update_kernel(int n, const float *p, float *pt, float tau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pt[i] = tau * p[i] + (1.0f - tau) * pt[i];
}
__global__ void report_kernel(const float *scal, float *out) {
    float l2 = 0.f;
    for (int t = 0; t < kActorTensors; ++t) l2 += scal[6 + t];
    out[0] = scal[4]; out[1] = scal[5] + 1e-4f * l2; out[2] = scal[2]; out[3] = scal[3];
}

// ---- fused row-tile kernels (fused_mlp.cuh): one CTA = 32 minibatch rows through a whole layer chain ------------------
// shared memory: bufA | bufB | bufC (activation tiles) | three weight stages | misc (actor input tile, row statistics)
constexpr int kMisc = 1280;
constexpr int kFusedSmem = (3 * TILE + NSTAGE * WST + kMisc) * 4;
static_assert(kFusedSmem <= 232448, "shared memory plan exceeds the 227 KB opt-in limit");
// per-CTA column partial sums of the critic backward: [db0 | dgamma1 | dbeta1] [db4 | dgamma2 | dbeta2] [db8 | dW10 | db10]
constexpr int kCPartA = 0, kCPartB = 768, kCPartC = 1536, kCPart = 1800;
// per-CTA partials of the actor backward: dEW(30, padded 32) | db4(10, padded 16) | db3(64) | db2(128) | db1(128)
constexpr int kAPartEW = 0, kAPartB4 = 32, kAPartB3 = 48, kAPartB2 = 112, kAPartB1 = 240, kAPart = 368;

// LayerNorm(256) + activation on the tile: z_s (pre-LN rows) becomes h in place; xhat / rstd / h also go to global
// memory (what the backward needs) when the pointers are set.  Warp w owns rows kRowsPerWarp*w ...
constexpr int kRowsPerWarp = TM / (NTHR / 32);
__device__ __forceinline__ void ln_act_tile(float *z_s, const float *gamma, const float *beta, int act, int row0, int B,
                                            float *xhat_g, float *rstd_g, float *h_g, const Drop &dr, int site) {
    const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
    float gm[8], bt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { gm[i] = gamma[l + 32 * i]; bt[i] = beta[l + 32 * i]; }
#pragma unroll 1
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = warp * kRowsPerWarp + rr, gr = row0 + r;
        float v[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = z_s[r * LDA + l + 32 * i]; s += v[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 256.0f);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / 256.0f) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = l + 32 * i;
            const float xh = (v[i] - mean) * rstd;
            float y = xh * gm[i] + bt[i];
            y = act == ACT_LEAKY ? (y > 0.f ? y : 0.1f * y) : fmaxf(y, 0.f);
            if (dr.mode != ERLFILL_DROPOUT_OFF) y = keep1(dr, site, gr < B ? gr : 0, c, 256) ? y * kDropScale : 0.f;   // nn.Dropout(0.2), train()
            z_s[r * LDA + c] = y;
            if (gr < B) {
                if (xhat_g) xhat_g[(size_t)gr * 256 + c] = xh;
                if (h_g) h_g[(size_t)gr * 256 + c] = y;
            }
        }
        if (l == 0 && rstd_g && gr < B) rstd_g[gr] = rstd;
    }
}

// Critic.forward for the CTA's rows (ddpg_agent.py:163-183).  The input row [state(2) | action(10) | emotion standardised(3) | 0]
// is assembled here:  mode 0: (s, stored normalised action, e)   mode 1: (s', actions_ext, 0) [target]   mode 2: (s, actions_ext, e)
struct CriticFwdArgs {
    Drop drop;
    int B, mode, save;
    const float *batch, *actions_ext, *stats, *P;
    float *x0, *xhat1, *h1, *rstd1, *xhat2, *h2, *rstd2, *h3, *q;
};
__global__ void __launch_bounds__(NTHR) critic_fwd_kernel(const CriticFwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *bufA = sm, *bufB = sm + TILE, *wst = sm + 3 * TILE;
    const int tid = threadIdx.x, row0 = blockIdx.x * TM, B = a.B;
    for (int e = tid; e < TM * 16; e += NTHR) {
        const int r = e >> 4, c = e & 15, gr = row0 + r;
        float v = 0.f;
        if (gr < B) {
            const float *row = a.batch + (size_t)gr * ERLFILL_REC;
            if (c < 2) v = a.mode == 1 ? row[13 + c] : row[c];
            else if (c < 12) v = a.mode == 0 ? row[c] : a.actions_ext[(size_t)gr * 10 + (c - 2)];
            else if (c < 15) v = a.mode == 1 ? 0.f : (row[16 + (c - 12)] - a.stats[c - 12]) / (a.stats[3 + (c - 12)] + 1e-6f);
            if (a.save) a.x0[(size_t)gr * 16 + c] = v;
        }
        bufA[r * LDA + c] = v;
    }
    const float *P = a.P;
    {   // Linear(15, 256) -> LayerNorm -> LeakyReLU(0.1)
        float acc[2][Cols<256>::NTW][4];
        dense<16, 256, false>(acc, bufA, P + CO::W0, 15, 256, 15, wst, 0);
        for_each_out<256>(acc, 256, [&](int r, int c, float v) { bufB[r * LDA + c] = v + P[CO::B0 + c]; });
    }
    __syncthreads();
    const int site0 = a.mode == 0 ? 4 : (a.mode == 1 ? 2 : 8);       // dropout sites of this forward (ddpg_dropout.cuh)
    ln_act_tile(bufB, P + CO::G1, P + CO::BE1, ACT_LEAKY, row0, B, a.save ? a.xhat1 : nullptr, a.save ? a.rstd1 : nullptr, a.save ? a.h1 : nullptr,
                a.drop, site0);
    {   // Linear(256, 256) -> LayerNorm -> ReLU
        float acc[2][Cols<256>::NTW][4];
        dense<256, 256, false>(acc, bufB, P + CO::W4, 256, 256, 256, wst, 4);
        for_each_out<256>(acc, 256, [&](int r, int c, float v) { bufA[r * LDA + c] = v + P[CO::B4 + c]; });
    }
    __syncthreads();
    ln_act_tile(bufA, P + CO::G2, P + CO::BE2, ACT_RELU, row0, B, a.save ? a.xhat2 : nullptr, a.save ? a.rstd2 : nullptr, a.save ? a.h2 : nullptr,
                a.drop, site0 + 1);
    {   // Linear(256, 128) -> ReLU
        float acc[2][Cols<128>::NTW][4];
        dense<256, 128, false>(acc, bufA, P + CO::W8, 256, 128, 256, wst, 4);
        for_each_out<128>(acc, 128, [&](int r, int c, float v) {
            v = fmaxf(v + P[CO::B8 + c], 0.f);
            bufB[r * LDA + c] = v;
            if (a.save && row0 + r < B) a.h3[(size_t)(row0 + r) * 128 + c] = v;
        });
    }
    __syncthreads();
    // Linear(128, 1), clamp +-1e6
    const int warp = tid >> 5, l = tid & 31;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = warp * kRowsPerWarp + rr;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s = fmaf(bufB[r * LDA + l + 32 * i], P[CO::W10 + l + 32 * i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (l == 0 && row0 + r < B) a.q[row0 + r] = fminf(fmaxf(s + P[CO::B10], -1e6f), 1e6f);
    }
}

// Actor.forward for the CTA's rows (ddpg_agent.py:73-122) -> physical-unit actions.  Input row [state | emotion | 0]; mode 1:
// (s', next_emotion broadcast) for the target.  keep: store what the backward needs.
struct ActorFwdArgs {
    Drop drop;
    int B, mode, keep, n_cond;
    const float *batch, *next_emotion, *P, *scale, *offset;
    float *out, *xa, *a1, *a2, *a3, *z4, *aux;
};
__global__ void __launch_bounds__(NTHR) actor_fwd_kernel(const ActorFwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *bufA = sm, *bufB = sm + TILE, *wst = sm + 3 * TILE, *xa_s = sm + 3 * TILE + NSTAGE * WST;
    const int tid = threadIdx.x, row0 = blockIdx.x * TM, B = a.B;
    for (int e = tid; e < TM * 8; e += NTHR) {
        const int r = e >> 3, c = e & 7, gr = row0 + r;
        float v = 0.f;
        if (gr < B) {
            const float *row = a.batch + (size_t)gr * ERLFILL_REC;
            if (c < 2) v = a.mode == 1 ? row[13 + c] : row[c];
            else if (c < 5) v = a.mode == 1 ? a.next_emotion[c - 2] : row[16 + (c - 2)];
            else if (c == 5) v = a.n_cond > 1 ? row[19] : 0.f;      // the record's condition index (multiplied by zero weights below)
            if (a.keep) a.xa[(size_t)gr * 8 + c] = v;
        }
        bufA[r * LDA + c] = v;
        xa_s[e] = v;
    }
    const float *P = a.P;
    {   // Linear(5, 128) -> LeakyReLU(0.1)
        float acc[2][Cols<128>::NTW][4];
        dense<8, 128, false>(acc, bufA, P + AO::W1, 5, 128, 5, wst, 0);
        for_each_out<128>(acc, 128, [&](int r, int c, float v) {
            v += P[AO::B1 + c];
            v = v > 0.f ? v : 0.1f * v;
            if (a.drop.mode != ERLFILL_DROPOUT_OFF) v = keep1(a.drop, a.mode == 1 ? 0 : 6, row0 + r < B ? row0 + r : 0, c, 128) ? v * kDropScale : 0.f;
            bufB[r * LDA + c] = v;
            if (a.keep && row0 + r < B) a.a1[(size_t)(row0 + r) * 128 + c] = v;
        });
    }
    {   // Linear(128, 128) -> ReLU
        float acc[2][Cols<128>::NTW][4];
        dense<128, 128, false>(acc, bufB, P + AO::W2, 128, 128, 128, wst, 2);
        for_each_out<128>(acc, 128, [&](int r, int c, float v) {
            v = fmaxf(v + P[AO::B2 + c], 0.f);
            if (a.drop.mode != ERLFILL_DROPOUT_OFF) v = keep1(a.drop, a.mode == 1 ? 1 : 7, row0 + r < B ? row0 + r : 0, c, 128) ? v * kDropScale : 0.f;
            bufA[r * LDA + c] = v;
            if (a.keep && row0 + r < B) a.a2[(size_t)(row0 + r) * 128 + c] = v;
        });
    }
    {   // Linear(128, 64) -> ReLU
        float acc[2][1][4];
        dense<128, 64, false>(acc, bufA, P + AO::W3, 128, 64, 128, wst, 2);
        for_each_out<64>(acc, 64, [&](int r, int c, float v) {
            v = fmaxf(v + P[AO::B3 + c], 0.f);
            bufB[r * LDA + c] = v;
            if (a.keep && row0 + r < B) a.a3[(size_t)(row0 + r) * 64 + c] = v;
        });
    }
    {   // Linear(64, 10)
        float acc[2][1][4];
        dense<64, 16, false>(acc, bufB, P + AO::W4, 64, 10, 64, wst, 2);
        for_each_out<16>(acc, 10, [&](int r, int c, float v) { bufA[r * LDA + c] = v + P[AO::B4 + c]; });
    }
    __syncthreads();
    // Softsign, emotion modulation, scaling (ddpg_agent.py:104-122)
    for (int e = tid; e < TM * 10; e += NTHR) {
        const int r = e / 10, c = e % 10, gr = row0 + r;
        if (gr >= B) continue;
        const float z = bufA[r * LDA + c];
        const float base = z / (1.0f + fabsf(z));
        const float cur = xa_s[r * 8 + 2], cons = xa_s[r * 8 + 3], anx = xa_s[r * 8 + 4];
        float em = cur * P[AO::EW + c];
        em = fmaf(cons, P[AO::EW + 10 + c], em);
        em = fmaf(anx, P[AO::EW + 20 + c], em);
        const float raw = tanhf(em * 1.5f);
        float alpha = 0.1f + 0.9f * anx;
        alpha *= (1.0f - 0.5f * cons);
        alpha *= (1.0f + 0.5f * cur);
        alpha = fminf(fmaxf(alpha, 0.01f), 1.0f);
        const float om = 1.0f - fabsf(base);
        const float sf = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
        const size_t g = (size_t)gr * 10 + c;
        const int so = (int)xa_s[r * 8 + 5] * 10 + c;          // per-condition scaling row (0 for a single condition)
        a.out[g] = (base + alpha * raw * sf * (om * om)) * a.scale[so] + a.offset[so];
        if (a.keep) { a.z4[g] = z; a.aux[g * 3 + 0] = base; a.aux[g * 3 + 1] = raw; a.aux[g * 3 + 2] = alpha; }
    }
}

// backward of (LayerNorm + activation) on the tile.  In: dh in d_s, xhat in x_s.  Out: dz in x_s (and global when dz_g),
// per-CTA column sums [sum dz | sum dy*xhat | sum dy] to part[0:256 | 256:512 | 512:768] when part is set.
__device__ __forceinline__ void ln_act_bwd_tile(float *d_s, float *x_s, float *stat_s, const float *gamma, const float *beta, int act,
                                                const float *rstd_g, int row0, int B, float *dz_g, float *part, const Drop &dr, int site) {
    const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;
    float gmr[8], btr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { gmr[i] = gamma[l + 32 * i]; btr[i] = beta[l + 32 * i]; }
    const float rs_mine = (l < kRowsPerWarp && row0 + warp * kRowsPerWarp + l < B) ? rstd_g[row0 + warp * kRowsPerWarp + l] : 0.f;
#pragma unroll 1
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = warp * kRowsPerWarp + rr;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = l + 32 * i;
            const float xh = x_s[r * LDA + c];
            const float y = xh * gmr[i] + btr[i];
            float g = d_s[r * LDA + c];
            if (dr.mode != ERLFILL_DROPOUT_OFF) g = keep1(dr, site, row0 + r < B ? row0 + r : 0, c, 256) ? g * kDropScale : 0.f;
            g *= act == ACT_LEAKY ? (y > 0.f ? 1.f : 0.1f) : (y > 0.f ? 1.f : 0.f);
            d_s[r * LDA + c] = g;                                  // dy
            const float dxh = g * gmr[i];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (l == 0) { stat_s[r] = s1 * (1.0f / 256.0f); stat_s[TM + r] = s2 * (1.0f / 256.0f); }
        if (l == rr) stat_s[2 * TM + r] = rs_mine;
    }
    __syncthreads();
    {   // two threads per column (16 rows each): dz and the three column sums in one pass; the upper half's sums go through
        // shared memory and are added by the lower half (fixed order)
        const int c = tid & 255, half = tid >> 8;
        const float gm = gamma[c];
        float sdz = 0.f, sg = 0.f, sb = 0.f;
#pragma unroll 4
        for (int r = half * (TM / 2); r < (half + 1) * (TM / 2); ++r) {
            const float dy = d_s[r * LDA + c], xh = x_s[r * LDA + c];
            sb += dy;
            sg = fmaf(dy, xh, sg);
            const float dz = stat_s[2 * TM + r] * (dy * gm - stat_s[r] - xh * stat_s[TM + r]);
            x_s[r * LDA + c] = dz;
            sdz += dz;
            if (dz_g && row0 + r < B) dz_g[(size_t)(row0 + r) * 256 + c] = dz;
        }
        float *hs = stat_s + 3 * TM;                               // [3][256]
        if (half == 1) { hs[c] = sdz; hs[256 + c] = sg; hs[512 + c] = sb; }
        __syncthreads();
        if (half == 0 && part) { part[c] = sdz + hs[c]; part[256 + c] = sg + hs[256 + c]; part[512 + c] = sb + hs[512 + c]; }
    }
    __syncthreads();
}

// backward of critic_fwd_kernel from dq (dq == nullptr: the constant dq_const for every row): the input-gradient chain for
// the CTA's rows.  grads != 0: also dz1/dz2/dz3 to global (operands of the weight-gradient GEMMs) and the per-CTA column
// partial sums (bias, LayerNorm and head gradients) to cpart[cta].  dx0 -> g16 [B][16].
struct CriticBwdArgs {
    Drop drop;
    int B, grads, site0;
    float dq_const;
    const float *P, *dq, *h3, *xhat2, *rstd2, *xhat1, *rstd1;
    float *dz3, *dz2, *dz1, *g16, *cpart;
};
__global__ void __launch_bounds__(NTHR) critic_bwd_kernel(const CriticBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *bufA = sm, *bufB = sm + TILE, *bufC = sm + 2 * TILE, *wst = sm + 3 * TILE, *stat_s = sm + 3 * TILE + NSTAGE * WST;
    const int tid = threadIdx.x, row0 = blockIdx.x * TM, B = a.B;
    const float *P = a.P;
    float *part = a.grads ? a.cpart + (size_t)blockIdx.x * kCPart : nullptr;
    prefetch_tile256(bufC, a.xhat2, row0, B);        // lands while dz3 and dh2 are computed
    // dz3 = dq * w10 * (h3 > 0)
    {
        constexpr int IT = TM * 128 / NTHR;
        float hv[IT], dqv[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int e = tid + it * NTHR, r = e >> 7, c = e & 127, gr = row0 + r;
            hv[it] = gr < B ? a.h3[(size_t)gr * 128 + c] : 0.f;
            dqv[it] = gr < B ? (a.dq ? a.dq[gr] : a.dq_const) : 0.f;
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            const int e = tid + it * NTHR, r = e >> 7, c = e & 127, gr = row0 + r;
            const float v = hv[it] > 0.f ? dqv[it] * P[CO::W10 + c] : 0.f;
            if (a.grads && gr < B) a.dz3[(size_t)gr * 128 + c] = v;
            bufA[r * LDA + c] = v;
        }
    }
    __syncthreads();
    if (a.grads) {
        if (tid < 128) {          // db8[c] = sum_r dz3[r][c]
            float s = 0.f;
            for (int r = 0; r < TM; ++r) s += bufA[r * LDA + tid];
            part[kCPartC + tid] = s;
        } else if (tid < 256) {   // dW10[c] = sum_r dq[r] h3[r][c]; db10 = sum_r dq[r]
            const int c = tid - 128;
            float s = 0.f, sq = 0.f;
            for (int r = 0; r < TM && row0 + r < B; ++r) {
                const float dq = a.dq ? a.dq[row0 + r] : a.dq_const;
                s = fmaf(dq, a.h3[(size_t)(row0 + r) * 128 + c], s);
                sq += dq;
            }
            part[kCPartC + 128 + c] = s;
            if (c == 0) part[kCPartC + 256] = sq;
        }
    }
    {   // dh2 = dz3 W8
        float acc[2][Cols<256>::NTW][4];
        dense<128, 256, true>(acc, bufA, P + CO::W8, 256, 256, 128, wst, 4);
        for_each_out<256>(acc, 256, [&](int r, int c, float v) { bufB[r * LDA + c] = v; });
    }
    prefetch_tile256(bufA, a.xhat1, row0, B);        // bufA (dz3) is free: lands under the LayerNorm backward and dh1
    cp_async_wait<1>();                              // xhat2 tile (the older group)
    __syncthreads();
    ln_act_bwd_tile(bufB, bufC, stat_s, P + CO::G2, P + CO::BE2, ACT_RELU, a.rstd2, row0, B, a.grads ? a.dz2 : nullptr,
                    part ? part + kCPartB : nullptr, a.drop, a.site0 + 1);
    {   // dh1 = dz2 W4
        float acc[2][Cols<256>::NTW][4];
        dense<256, 256, true>(acc, bufC, P + CO::W4, 256, 256, 256, wst, 4);
        for_each_out<256>(acc, 256, [&](int r, int c, float v) { bufB[r * LDA + c] = v; });
    }
    __syncthreads();                                 // (dense drained every cp.async group, the xhat1 tile included)
    ln_act_bwd_tile(bufB, bufA, stat_s, P + CO::G1, P + CO::BE1, ACT_LEAKY, a.rstd1, row0, B, a.grads ? a.dz1 : nullptr,
                    part ? part + kCPartA : nullptr, a.drop, a.site0);
    {   // dx0[:, :15] = dz1 W0
        float acc[2][1][4];
        dense<256, 16, true>(acc, bufA, P + CO::W0, 15, 15, 256, wst, 0);
        for_each_out<16>(acc, 16, [&](int r, int c, float v) {
            if (row0 + r < B) a.g16[(size_t)(row0 + r) * 16 + c] = c < 15 ? v : 0.f;
        });
    }
}

// backward of actor_fwd_kernel for the CTA's rows: d(out) from the critic path (g16 columns 2..11) plus the diversity term
// -0.03 * mean_rows(std_unbiased(out, dim=1)) -> dz4 .. dz1 to global (operands of the weight-gradient GEMMs) and the per-CTA
// partial sums of d(emotion_weights) and the bias gradients to apart[cta].
struct ActorBwdArgs {
    Drop drop;
    int B;
    const float *P, *g16, *out, *aux, *xa, *scale, *z4, *a3, *a2, *a1;
    float *gz4, *ga64, *ga128a, *ga128b, *apart;
};
__global__ void __launch_bounds__(NTHR) actor_bwd_kernel(const ActorBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *bufA = sm, *bufB = sm + TILE, *wst = sm + 3 * TILE, *de_s = sm + 3 * TILE + NSTAGE * WST;   // de_s[TM][10]
    const int tid = threadIdx.x, row0 = blockIdx.x * TM, B = a.B;
    const float *P = a.P;
    float *part = a.apart + (size_t)blockIdx.x * kAPart;
    if (tid < TM) {
        const int r = tid, gr = row0 + r;
        if (gr < B) {
            float o[10], mean = 0.f;
#pragma unroll
            for (int c = 0; c < 10; ++c) { o[c] = a.out[(size_t)gr * 10 + c]; mean += o[c]; }
            mean *= 0.1f;
            float var = 0.f;
#pragma unroll
            for (int c = 0; c < 10; ++c) { const float d = o[c] - mean; var = fmaf(d, d, var); }
            const float sd = sqrtf(var / 9.0f);
#pragma unroll
            for (int c = 0; c < 10; ++c) {
                const size_t g = (size_t)gr * 10 + c;
                float go = a.g16[(size_t)gr * 16 + 2 + c];
                go += (-0.03f / (float)B) * (o[c] - mean) / (9.0f * sd);                      // d std / d o_c
                const float gr_ = go * a.scale[(int)a.xa[(size_t)gr * 8 + 5] * 10 + c];           // d raw_action
                const float base = a.aux[g * 3 + 0], raw = a.aux[g * 3 + 1], alpha = a.aux[g * 3 + 2];
                const float om = 1.0f - fabsf(base);
                const float sf = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
                // d raw_action / d base = 1 + alpha*raw*sf * 2*om*(-sign(base)) = 1 + 2*alpha*raw*om*sf*sf (sign has zero grad)
                const float dbase = gr_ * (1.0f + 2.0f * alpha * raw * om * (sf * sf));
                const float z = a.z4[g], den = 1.0f + fabsf(z);
                const float dz = dbase / (den * den);                                             // softsign'
                bufA[r * LDA + c] = dz;
                a.gz4[g] = dz;
                // d raw_action / d (e @ W_e)[c] = alpha*sf*om^2 * (1 - raw^2) * 1.5
                de_s[r * 10 + c] = gr_ * alpha * sf * (om * om) * (1.0f - raw * raw) * 1.5f;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 10; ++c) { bufA[r * LDA + c] = 0.f; de_s[r * 10 + c] = 0.f; }
        }
#pragma unroll
        for (int c = 10; c < 16; ++c) bufA[r * LDA + c] = 0.f;
    }
    __syncthreads();
    if (tid < 30) {               // dEW[k][c] = sum_r emotion[r][k] * de[r][c]
        const int k = tid / 10, c = tid % 10;
        float s = 0.f;
        for (int r = 0; r < TM && row0 + r < B; ++r) s = fmaf(a.xa[(size_t)(row0 + r) * 8 + 2 + k], de_s[r * 10 + c], s);
        part[kAPartEW + tid] = s;
    } else if (tid >= 32 && tid < 42) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += bufA[r * LDA + tid - 32];
        part[kAPartB4 + tid - 32] = s;
    }
    {   // da3 = dz4 W4, relu'
        float acc[2][1][4];
        dense<16, 64, true>(acc, bufA, P + AO::W4, 64, 64, 10, wst, 2);
        for_each_out<64>(acc, 64, [&](int r, int c, float v) {
            const int gr = row0 + r;
            v = (gr < B && a.a3[(size_t)gr * 64 + c] > 0.f) ? v : 0.f;
            bufB[r * LDA + c] = v;
            if (gr < B) a.ga64[(size_t)gr * 64 + c] = v;
        });
    }
    __syncthreads();
    if (tid < 64) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += bufB[r * LDA + tid];
        part[kAPartB3 + tid] = s;
    }
    {   // da2 = dz3 W3, relu'
        float acc[2][Cols<128>::NTW][4];
        dense<64, 128, true>(acc, bufB, P + AO::W3, 128, 128, 64, wst, 2);
        for_each_out<128>(acc, 128, [&](int r, int c, float v) {
            const int gr = row0 + r;
            // a2 is the dropped activation: > 0 only where the unit was kept and relu passed; kept units carry the 1 / 0.8 scale
            v = (gr < B && a.a2[(size_t)gr * 128 + c] > 0.f) ? (a.drop.mode != ERLFILL_DROPOUT_OFF ? v * kDropScale : v) : 0.f;
            bufA[r * LDA + c] = v;
            if (gr < B) a.ga128a[(size_t)gr * 128 + c] = v;
        });
    }
    __syncthreads();
    if (tid < 128) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += bufA[r * LDA + tid];
        part[kAPartB2 + tid] = s;
    }
    {   // da1 = dz2 W2, leaky'
        float acc[2][Cols<128>::NTW][4];
        dense<128, 128, true>(acc, bufA, P + AO::W2, 128, 128, 128, wst, 2);
        for_each_out<128>(acc, 128, [&](int r, int c, float v) {
            const int gr = row0 + r;
            v = gr < B ? (a.a1[(size_t)gr * 128 + c] > 0.f ? v : 0.1f * v) : 0.f;      // kept units keep the sign of z1
            if (a.drop.mode != ERLFILL_DROPOUT_OFF) v = (gr < B && keep1(a.drop, 6, gr, c, 128)) ? v * kDropScale : 0.f;
            bufB[r * LDA + c] = v;
            if (gr < B) a.ga128b[(size_t)gr * 128 + c] = v;
        });
    }
    __syncthreads();
    if (tid < 128) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += bufB[r * LDA + tid];
        part[kAPartB1 + tid] = s;
    }
}

// ---- workspace -------------------------------------------------------------------------------
struct Ws {
    float *cgrad, *agrad, *cred, *ared;         // flat gradient buckets (cgrad tail holds the 3 emotion means) and their peer-reduced twins
    float *stats, *scal, *ssq, *step;                // stats[6]; scal: [2..3] lrs [4] closs [5] aloss [6..14] norms; ssq: [0:64] critic, [64:128] actor sum-of-squares partials
    float *x0, *xhat1, *h1, *xhat2, *h2, *h3, *rstd1, *rstd2, *q, *tq, *q2, *dq;
    float *xa, *a1, *a2, *a3, *z4, *aout, *aux, *atgt;
    float *dz1, *dz2, *dz3, *g16, *ga128a, *ga128b, *ga64, *gz4, *part, *cpart, *apart;
};
static size_t align16(size_t n) { return (n + 3) & ~(size_t)3; }
static int n_tiles(int B) { return (B + TM - 1) / TM; }
static size_t carve(Ws *w, float *base, int B) {
    size_t o = 0;
    auto take = [&](float **p, size_t n) { if (w) *p = base + o; o += align16(n); };
    float *dummy;
    take(w ? &w->cgrad : &dummy, kCriticGradPad); take(w ? &w->agrad : &dummy, kActorGradPad);
    take(w ? &w->cred : &dummy, kCriticGradPad); take(w ? &w->ared : &dummy, kActorGradPad);
    take(w ? &w->stats : &dummy, 8); take(w ? &w->scal : &dummy, 32); take(w ? &w->ssq : &dummy, 2 * kSumsqBlocks); take(w ? &w->step : &dummy, 4);
    const size_t b = (size_t)B;
    take(w ? &w->x0 : &dummy, b * 16); take(w ? &w->xhat1 : &dummy, b * 256); take(w ? &w->h1 : &dummy, b * 256);
    take(w ? &w->xhat2 : &dummy, b * 256); take(w ? &w->h2 : &dummy, b * 256); take(w ? &w->h3 : &dummy, b * 128);
    take(w ? &w->rstd1 : &dummy, b); take(w ? &w->rstd2 : &dummy, b); take(w ? &w->q : &dummy, b);
    take(w ? &w->tq : &dummy, b); take(w ? &w->q2 : &dummy, b); take(w ? &w->dq : &dummy, b);
    take(w ? &w->xa : &dummy, b * 8); take(w ? &w->a1 : &dummy, b * 128); take(w ? &w->a2 : &dummy, b * 128);
    take(w ? &w->a3 : &dummy, b * 64); take(w ? &w->z4 : &dummy, b * 10); take(w ? &w->aout : &dummy, b * 10);
    take(w ? &w->aux : &dummy, b * 30); take(w ? &w->atgt : &dummy, b * 10);
    take(w ? &w->dz1 : &dummy, b * 256); take(w ? &w->dz2 : &dummy, b * 256); take(w ? &w->dz3 : &dummy, b * 128);
    take(w ? &w->g16 : &dummy, b * 16); take(w ? &w->ga128a : &dummy, b * 128); take(w ? &w->ga128b : &dummy, b * 128);
    take(w ? &w->ga64 : &dummy, b * 64); take(w ? &w->gz4 : &dummy, b * 10);
    take(w ? &w->part : &dummy, mlp::kPartFloats);            // split partials of the grouped weight-gradient GEMMs
    take(w ? &w->cpart : &dummy, (size_t)n_tiles(B) * kCPart); take(w ? &w->apart : &dummy, (size_t)n_tiles(B) * kAPart);
    return o;
}

// ---- launch helpers ---------------------------------------------------------------------------
static cudaError_t fused_smem_optin() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(critic_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(actor_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(critic_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(actor_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem)) != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}
static void critic_forward(cudaStream_t st, const Ws &w, const Drop &dr, int B, int mode, const float *batch, const float *actions_ext, const float *P,
                           float *q_out, bool save) {
    const CriticFwdArgs a{dr, B, mode, save ? 1 : 0, batch, actions_ext, w.stats, P, w.x0, w.xhat1, w.h1, w.rstd1, w.xhat2, w.h2, w.rstd2, w.h3, q_out};
    critic_fwd_kernel<<<n_tiles(B), NTHR, kFusedSmem, st>>>(a);
}
static void actor_forward(cudaStream_t st, const Ws &w, const Drop &dr, int B, int mode, const float *batch, const float *next_emotion, const float *P,
                          const float *scale, const float *offset, float *out, bool keep, int n_cond) {
    const ActorFwdArgs a{dr, B, mode, keep ? 1 : 0, n_cond, batch, next_emotion, P, scale, offset, out, w.xa, w.a1, w.a2, w.a3, w.z4, w.aux};
    actor_fwd_kernel<<<n_tiles(B), NTHR, kFusedSmem, st>>>(a);
}
// the grouped weight-gradient GEMMs and the one reduction launch that finishes every parameter gradient of a network
struct GradPlan {
    mlp::WgArgs wg;
    mlp::RedArgs red;
    int wg_blocks = 0, red_blocks = 0;
    size_t part_used = 0;
    void init(int Bt) {
        wg.ntask = 0; red.nseg = 0; wg.Bt = Bt;
        int splits = Bt >= 2048 ? 16 : (Bt >= 512 ? 8 : (Bt >= 128 ? 4 : 1));
        wg.kps = ((Bt + splits - 1) / splits + BK - 1) / BK * BK;
        wg.splits = (Bt + wg.kps - 1) / wg.kps;
    }
    void seg(const float *part, float *out, int n, int splits, int stride) {
        red.s[red.nseg++] = mlp::RedSeg{part, out, n, splits, stride, red_blocks};
        red_blocks += (n + 255) / 256;
    }
    // dW[N][K] = dZ^T X
    void wgrad(float *part_base, const float *dZ, int ldz, const float *X, int ldx, int N, int K, float *dW) {
        float *part = part_base + part_used;
        const int tiles_k = (K + BN - 1) / BN, tiles = tiles_k * ((N + BM - 1) / BM);
        wg.t[wg.ntask++] = mlp::WgTask{dZ, X, part, ldz, ldx, N, K, tiles_k, tiles, wg_blocks};
        wg_blocks += tiles * wg.splits;
        seg(part, dW, N * K, wg.splits, N * K);
        part_used += (size_t)wg.splits * N * K;
    }
    void launch(cudaStream_t st) const {
        mlp::wgrad_group_kernel<<<wg_blocks, 256, 0, st>>>(wg);
        mlp::group_reduce_kernel<<<red_blocks, 256, 0, st>>>(red);
    }
};

// the tcgen05 path (update_tc.cu) keeps its weight images, scratch and partials behind this workspace
namespace utc {
struct Head { float *cgrad, *agrad, *cred, *ared, *stats, *scal, *ssq, *step; };
size_t tail_bytes(int B);
int pack_images(const erlfill_ddpg_nets *n, uint8_t *tail_base, int B, cudaStream_t st);
int update_fast(const erlfill_ddpg_nets *n, const erlfill_ddpg_hyper *h, const float *d_batch, const float *d_next_emotion, const Head &w,
                uint8_t *tail_base, int phases, float *d_report, cudaStream_t st);
float *tail_ssq_fine(uint8_t *tail_base, int B, int which);
unsigned int *tail_peer_done(uint8_t *tail_base, int B, int which);
int peer_fine_blocks(int which);
}  // namespace utc
static size_t exact_bytes(int batch) { return (carve(nullptr, nullptr, batch) * sizeof(float) + 255) & ~(size_t)255; }
// peer.cuh: what the fast path's exchange needs of the workspace
int ddpg_fine_blocks(int which) { return utc::peer_fine_blocks(which); }
float *ddpg_ssq_fine(void *d_workspace, int batch, int which) { return utc::tail_ssq_fine((uint8_t *)d_workspace + exact_bytes(batch), batch, which); }
unsigned int *ddpg_peer_done(void *d_workspace, int batch, int which) { return utc::tail_peer_done((uint8_t *)d_workspace + exact_bytes(batch), batch, which); }

}  // namespace erlfill

using namespace erlfill;

extern "C" {

size_t erlfill_ddpg_workspace_bytes(int batch) {
    if (batch < 2) return 0;
    return exact_bytes(batch) + utc::tail_bytes(batch);
}

int erlfill_ddpg_pack_images(const erlfill_ddpg_nets *n, void *d_workspace, int batch, void *stream) {
    if (!n || !d_workspace || batch < 2 || !n->actor || !n->actor_target || !n->critic || !n->critic_target) return ERLFILL_ERR_ARG;
    return utc::pack_images(n, (uint8_t *)d_workspace + exact_bytes(batch), batch, (cudaStream_t)stream);
}

int erlfill_ddpg_update_launches(int fast) { return fast ? 6 : 24; }

int erlfill_ddpg_grad_buckets(void *d_workspace, int batch, float **critic_grad, int *critic_len, float **actor_grad, int *actor_len) {
    if (!d_workspace || batch < 2) return ERLFILL_ERR_ARG;
    Ws w;
    carve(&w, (float *)d_workspace, batch);
    if (critic_grad) *critic_grad = w.cgrad;
    if (critic_len) *critic_len = CO::TOTAL + 3;      // 103,937 gradients + the 3 emotion means
    if (actor_grad) *actor_grad = w.agrad;
    if (actor_len) *actor_len = AO::TOTAL;
    return ERLFILL_OK;
}

int erlfill_ddpg_exchange_buffers(void *d_workspace, int batch, float **critic_grad, int *critic_len, float **actor_grad, int *actor_len,
                                  float **critic_reduced, float **actor_reduced, float **sumsq) {
    if (!d_workspace || batch < 2) return ERLFILL_ERR_ARG;
    Ws w;
    carve(&w, (float *)d_workspace, batch);
    if (critic_grad) *critic_grad = w.cgrad;
    if (critic_len) *critic_len = CO::TOTAL + 3;
    if (actor_grad) *actor_grad = w.agrad;
    if (actor_len) *actor_len = AO::TOTAL;
    if (critic_reduced) *critic_reduced = w.cred;
    if (actor_reduced) *actor_reduced = w.ared;
    if (sumsq) *sumsq = w.ssq;
    return ERLFILL_OK;
}

int erlfill_ddpg_set_step(void *d_workspace, int batch, int adam_step, double lr_decay, void *stream) {
    if (!d_workspace || batch < 2 || adam_step < 1) return ERLFILL_ERR_ARG;
    Ws w;
    carve(&w, (float *)d_workspace, batch);
    const double bc1d = 1.0 - pow(0.9, (double)adam_step), bc2d = sqrt(1.0 - pow(0.999, (double)adam_step));
    set_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(w.step, lr_decay, (float)bc1d, (float)bc2d);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_ddpg_update(const erlfill_ddpg_nets *n, const erlfill_ddpg_hyper *h, const float *d_batch,
                        const float *d_next_emotion, void *d_workspace, int phases, float *d_report, void *stream) {
    if (!n || !h || !d_batch || !d_next_emotion || !d_workspace) return ERLFILL_ERR_ARG;
    if (!n->actor || !n->actor_target || !n->critic || !n->critic_target || !n->actor_m || !n->actor_v || !n->critic_m ||
        !n->critic_v || !n->scale || !n->offset)
        return ERLFILL_ERR_ARG;
    const int B = h->batch;
    if (B < 2 || h->adam_step < 0) return ERLFILL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, (float *)d_workspace, B);
    // adam_step >= 1: upload this step's scalars first; adam_step == 0: the caller already did (erlfill_ddpg_set_step), which is
    // how a CUDA graph captured once is replayed for every update
    if (h->adam_step >= 1) {
        const int rc = erlfill_ddpg_set_step(d_workspace, B, h->adam_step, h->lr_decay, stream);
        if (rc != ERLFILL_OK) return rc;
    }

    if (h->precision == ERLFILL_PRECISION_FAST) {
        const utc::Head hd{w.cgrad, w.agrad, w.cred, w.ared, w.stats, w.scal, w.ssq, w.step};
        return utc::update_fast(n, h, d_batch, d_next_emotion, hd, (uint8_t *)d_workspace + exact_bytes(B), phases, d_report, st);
    }
    if (h->precision != ERLFILL_PRECISION_EXACT) return ERLFILL_ERR_ARG;
    Drop dr;
    dr.mode = h->dropout_mode;
    dr.k0 = (uint32_t)h->seed; dr.k1 = (uint32_t)(h->seed >> 32); dr.update_idx = h->update_idx; dr.row_base = h->rank << 20;
    if (dr.mode == ERLFILL_DROPOUT_MASK && !h->masks) return ERLFILL_ERR_ARG;
    for (int s = 0; s < 10; ++s) {
        dr.m[s] = dr.mode == ERLFILL_DROPOUT_MASK ? h->masks->keep[s] : nullptr;
        if (dr.mode == ERLFILL_DROPOUT_MASK && !dr.m[s]) return ERLFILL_ERR_ARG;
    }
    ERLFILL_CUDA_TRY(fused_smem_optin());
    const int nt = n_tiles(B);

    if (phases & ERLFILL_PHASE_CRITIC_GRAD) {
        emotion_stats_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.stats, w.cgrad + CO::TOTAL);
        // target: a' = actor_target(s', e'), Q' = critic_target(s', a', 0)
        actor_forward(st, w, dr, B, 1, d_batch, d_next_emotion, n->actor_target, n->scale, n->offset, w.atgt, false, h->n_cond);
        critic_forward(st, w, dr, B, 1, d_batch, w.atgt, n->critic_target, w.tq, false);
        // Q(s, a_norm, e) and its gradient
        critic_forward(st, w, dr, B, 0, d_batch, nullptr, n->critic, w.q, true);
        critic_loss_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.q, w.tq, h->gamma, w.dq, w.scal + 4);
        const CriticBwdArgs ba{dr, B, 1, 4, 0.f, n->critic, w.dq, w.h3, w.xhat2, w.rstd2, w.xhat1, w.rstd1, w.dz3, w.dz2, w.dz1, w.g16, w.cpart};
        critic_bwd_kernel<<<nt, NTHR, kFusedSmem, st>>>(ba);
        float *G = w.cgrad;
        GradPlan gp;
        gp.init(B);
        gp.wgrad(w.part, w.dz3, 128, w.h2, 256, 128, 256, G + CO::W8);
        gp.wgrad(w.part, w.dz2, 256, w.h1, 256, 256, 256, G + CO::W4);
        gp.wgrad(w.part, w.dz1, 256, w.x0, 16, 256, 15, G + CO::W0);
        gp.seg(w.cpart + kCPartA, G + CO::B0, 768, nt, kCPart);       // db0 | dgamma1 | dbeta1
        gp.seg(w.cpart + kCPartB, G + CO::B4, 768, nt, kCPart);       // db4 | dgamma2 | dbeta2
        gp.seg(w.cpart + kCPartC, G + CO::B8, 257, nt, kCPart);       // db8 | dW10 | db10
        gp.launch(st);
    }
    const bool reduced = (phases & ERLFILL_PHASE_REDUCED) != 0;      // gradient + norm partials come from the peer exchange
    if (phases & ERLFILL_PHASE_CRITIC_APPLY) {
        const float *G = reduced ? w.cred : w.cgrad;
        lr_kernel<<<1, 1, 0, st>>>(G + CO::TOTAL, w.step, h->critic_lr, h->actor_lr, w.scal + 2);
        if (!reduced) sumsq_kernel<<<kSumsqBlocks, 256, 0, st>>>(CO::TOTAL, G, w.ssq);
        adam_soft_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(CO::TOTAL, n->critic, nullptr, n->critic_m, n->critic_v, G,
                                                                  w.ssq, w.scal + 2, h->max_grad_norm, w.step, h->tau);
    }
    if (phases & ERLFILL_PHASE_ACTOR_GRAD) {
        actor_forward(st, w, dr, B, 0, d_batch, d_next_emotion, n->actor, n->scale, n->offset, w.aout, true, h->n_cond);
        critic_forward(st, w, dr, B, 2, d_batch, w.aout, n->critic, w.q2, true);
        actor_loss_kernel<<<1 + kActorTensors, 1024, 0, st>>>(B, w.q2, w.aout, n->actor, w.scal + 5, w.scal + 6);
        // d(-mean q2)/dq2 = -1/B (the +-1e6 clamp is inactive): input gradient of the critic only -> w.g16
        const CriticBwdArgs ba{dr, B, 0, 8, -1.0f / (float)B, n->critic, nullptr, w.h3, w.xhat2, w.rstd2, w.xhat1, w.rstd1, nullptr, nullptr, nullptr,
                               w.g16, nullptr};
        critic_bwd_kernel<<<nt, NTHR, kFusedSmem, st>>>(ba);
        const ActorBwdArgs ab{dr, B, n->actor, w.g16, w.aout, w.aux, w.xa, n->scale, w.z4, w.a3, w.a2, w.a1, w.gz4, w.ga64, w.ga128a, w.ga128b, w.apart};
        actor_bwd_kernel<<<nt, NTHR, kFusedSmem, st>>>(ab);
        float *G = w.agrad;
        GradPlan gp;
        gp.init(B);
        gp.wgrad(w.part, w.gz4, 10, w.a3, 64, 10, 64, G + AO::W4);
        gp.wgrad(w.part, w.ga64, 64, w.a2, 128, 64, 128, G + AO::W3);
        gp.wgrad(w.part, w.ga128a, 128, w.a1, 128, 128, 128, G + AO::W2);
        gp.wgrad(w.part, w.ga128b, 128, w.xa, 8, 128, 5, G + AO::W1);
        gp.seg(w.apart + kAPartEW, G + AO::EW, 30, nt, kAPart);
        gp.seg(w.apart + kAPartB4, G + AO::B4, 10, nt, kAPart);
        gp.seg(w.apart + kAPartB3, G + AO::B3, 64, nt, kAPart);
        gp.seg(w.apart + kAPartB2, G + AO::B2, 128, nt, kAPart);
        gp.seg(w.apart + kAPartB1, G + AO::B1, 128, nt, kAPart);
        gp.launch(st);
        actor_l2_grad_kernel<<<(AO::TOTAL + 255) / 256, 256, 0, st>>>(n->actor, w.scal + 6, G);
    }
    if (phases & ERLFILL_PHASE_ACTOR_APPLY) {
        const float *G = reduced ? w.ared : w.agrad;
        if (!reduced) sumsq_kernel<<<kSumsqBlocks, 256, 0, st>>>(AO::TOTAL, G, w.ssq + kSumsqBlocks);
        adam_soft_kernel<<<(AO::TOTAL + 255) / 256, 256, 0, st>>>(AO::TOTAL, n->actor, n->actor_target, n->actor_m, n->actor_v, G,
                                                                  w.ssq + kSumsqBlocks, w.scal + 3, h->max_grad_norm, w.step, h->tau);
        // both soft updates happen after both optimiser steps (ddpg_agent.py:501-507)
        soft_update_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(CO::TOTAL, n->critic, n->critic_target, h->tau);
        if (d_report) report_kernel<<<1, 1, 0, st>>>(w.scal, d_report);   // [critic_loss, actor_loss, critic_lr, actor_lr]
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
