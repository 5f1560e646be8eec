// td3_sac.cu — T1: the TD3 and fixed-alpha SAC agents of the reference on the GPU.
//
//   erlfill_td3_act      TD3Agent.select_action (DifferentModules/td3_agent.py:221-236) for N envs
//   erlfill_td3_update   TD3Agent.train_step    (td3_agent.py:238-303): clipped double-Q target with target-policy smoothing,
//                        twin-critic MSE, Adam, delayed actor update + soft target updates
//   erlfill_sac_act      SACAgent.act           (DifferentModules/sac_agent.py:186-205)
//   erlfill_sac_update   SACAgent.update        (sac_agent.py:221-283): tanh-Gaussian policy with the reference's
//                        log(1 - a^2 + 1e-7) correction, fixed alpha = 0.3, two Q networks with separate Adam states
//
// All networks are 2-256-256-k / 12-256-256-1 ReLU MLPs: forward and backward are compositions of the shared strided
// tensor-core GEMM (mlp.cuh, mma.sync 3xTF32), deterministic column sums and a few row-wise kernels in this file.  Flat fp32
// parameter buffers in the reference's parameters() order (include/erlfill.h).  Replay records are the 80-byte rows of
// replay.cu: [state(2) | norm_action(10)] is contiguous, so the critic input (s, a) is read in place with a row stride of 20.
// Noise (target smoothing, rsample) is either injected (parity against the reference's torch draws) or Philox Box-Muller.
#include <math.h>

#include "common.cuh"
#include "mlp.cuh"

namespace erlfill {
namespace t1 {
using namespace mlp;

constexpr int H = 256, AD = ERLFILL_ACTION_DIM, SD = ERLFILL_STATE_DIM, QIN = SD + AD;
// one 3-layer MLP inside a flat buffer: W0[H][in] b0[H] W1[H][H] b1[H] W2[out][H] b2[out]
struct Net {
    int in, out;
    size_t W0, B0, W1, B1, W2, B2;
    __host__ __device__ size_t size() const { return B2 + out; }
};
static Net make_net(int in, int out, size_t base) {
    Net n; n.in = in; n.out = out;
    n.W0 = base; n.B0 = n.W0 + (size_t)H * in; n.W1 = n.B0 + H; n.B1 = n.W1 + (size_t)H * H; n.W2 = n.B1 + H; n.B2 = n.W2 + (size_t)out * H;
    return n;
}
static const Net kTd3Actor = make_net(SD, AD, 0);
static const Net kQ1 = make_net(QIN, 1, 0), kQ2 = make_net(QIN, 1, ERLFILL_TD3_CRITIC_PARAMS / 2);
// SAC policy: trunk = first two layers of a Net with out = 10 (mean head in the W2/B2 slot), then log_std.weight / bias
static const Net kSacTrunk = make_net(SD, AD, 0);
constexpr size_t kSacLsW = 2 * 256 + 256 + 256 * 256 + 256 + 10 * 256 + 10, kSacLsB = kSacLsW + 10 * 256;
static_assert(kSacLsB + 10 == ERLFILL_SAC_POLICY_PARAMS, "SAC policy layout");

// y[B][out] = MLP(x); h1/h2 keep the hidden activations for the backward
static void mlp_forward(cudaStream_t st, int B, const float *P, const Net &n, const float *x, int ldx, float *h1, float *h2, float *y, int ldy) {
    gemm(st, B, H, n.in, x, ldx, 1, P + n.W0, 1, n.in, h1, H, P + n.B0, ACT_RELU);
    gemm(st, B, H, H, h1, H, 1, P + n.W1, 1, H, h2, H, P + n.B1, ACT_RELU);
    gemm(st, B, n.out, H, h2, H, 1, P + n.W2, 1, H, y, ldy, P + n.B2, ACT_NONE);
}
// backward from dy[B][out]: parameter gradients into G (nullptr: none), input gradient into dx (nullptr: none)
static void mlp_backward(cudaStream_t st, float *part, int B, const float *P, float *G, const Net &n, const float *x, int ldx, const float *h1,
                         const float *h2, const float *dy, int ldy, float *ga, float *gb, float *dx, int lddx) {
    if (G) { gemm_wgrad(st, part, B, n.out, H, dy, ldy, h2, H, G + n.W2); colsum(st, part, B, n.out, dy, ldy, nullptr, 0, 0, G + n.B2); }
    gemm(st, B, H, n.out, dy, ldy, 1, P + n.W2, H, 1, ga, H, nullptr, ACT_NONE);
    relu_bwd_kernel<<<(unsigned)(((size_t)B * H + 255) / 256), 256, 0, st>>>((size_t)B * H, ga, h2);
    if (G) { gemm_wgrad(st, part, B, H, H, ga, H, h1, H, G + n.W1); colsum(st, part, B, H, ga, H, nullptr, 0, 0, G + n.B1); }
    gemm(st, B, H, H, ga, H, 1, P + n.W1, H, 1, gb, H, nullptr, ACT_NONE);
    relu_bwd_kernel<<<(unsigned)(((size_t)B * H + 255) / 256), 256, 0, st>>>((size_t)B * H, gb, h1);
    if (G) { gemm_wgrad(st, part, B, H, n.in, gb, H, x, ldx, G + n.W0); colsum(st, part, B, H, gb, H, nullptr, 0, 0, G + n.B0); }
    if (dx) gemm(st, B, n.in, H, gb, H, 1, P + n.W0, n.in, 1, dx, lddx, nullptr, ACT_NONE);
}

// ---- noise ------------------------------------------------------------------------------------------------------
// standard normal draw j (0..9) of row `row`: injected, or Philox Box-Muller keyed (row id, idx, j/2 + 8 salt, STREAM_OU).  Salts keep the
// streams apart: 1 = exploration noise at act time (row id = global env id, idx = lock-step index); 4 / 5 = the update's draws (row id =
// rank << 20 | minibatch row, idx = update index), so ranks draw different noise for their different minibatches
__device__ __forceinline__ float normal_draw(const float *inj, size_t row, int j, uint64_t seed, uint32_t row_id, uint32_t idx, uint32_t salt) {
    if (inj) return inj[row * AD + j];
    const uint4 r = philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), row_id, idx, (uint32_t)(j >> 1) + 8u * salt, STREAM_OU);
    const uint32_t a = (j & 1) ? r.z : r.x, b = (j & 1) ? r.w : r.y;
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

// ---- TD3 row kernels ----------------------------------------------------------------------------------------------
// x0n[r] = [s'(2) | clamp(tanh(z') + clamp(0.2 n, +-0.5), -1, 1)(10) | 0 x 4]          (td3_agent.py:268-269)
__global__ void td3_target_action_kernel(int B, const float *batch, const float *z, const float *noise, float policy_noise, float noise_clip,
                                         uint64_t seed, uint32_t update_idx, uint32_t row_id_base, float *x0n) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * 16) return;
    const int r = g >> 4, c = g & 15;
    float v = 0.f;
    if (c < SD) v = batch[(size_t)r * ERLFILL_REC + 13 + c];
    else if (c < QIN) {
        const int j = c - SD;
        const float n = fminf(fmaxf(normal_draw(noise, r, j, seed, row_id_base + (uint32_t)r, update_idx, 4) * policy_noise, -noise_clip), noise_clip);
        v = fminf(fmaxf(tanhf(z[(size_t)r * AD + j]) + n, -1.0f), 1.0f);
    }
    x0n[g] = v;
}
__device__ __forceinline__ float block_sum1024(float v, float *scratch) { return block_sum(v, scratch); }
// y = r + gamma (1-d) min(q1', q2'); loss = mse(q1,y) + mse(q2,y); dq = 2 (q - y) / B; report[0] = loss, [1] = mean(q1)   (:270-275, :284)
__global__ void __launch_bounds__(1024) td3_critic_loss_kernel(int B, const float *batch, const float *q1, const float *q2, const float *q1t,
                                                               const float *q2t, float gamma, float *dq1, float *dq2, float *report) {
    __shared__ float scratch[32];
    float l1 = 0.f, l2 = 0.f, qs = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) {
        const float rew = batch[(size_t)r * ERLFILL_REC + 12], d = batch[(size_t)r * ERLFILL_REC + 15];
        const float y = rew + gamma * (1.0f - d) * fminf(q1t[r], q2t[r]);
        const float e1 = q1[r] - y, e2 = q2[r] - y;
        dq1[r] = 2.0f * e1 / (float)B; dq2[r] = 2.0f * e2 / (float)B;
        l1 = fmaf(e1, e1, l1); l2 = fmaf(e2, e2, l2); qs += q1[r];
    }
    const float s1 = block_sum1024(l1, scratch), s2 = block_sum1024(l2, scratch), sq = block_sum1024(qs, scratch);
    if (threadIdx.x == 0) { report[0] = s1 / (float)B + s2 / (float)B; report[1] = sq / (float)B; }
}
// x0a[r] = [s | tanh(z) | 0]; a kept for the backward
__global__ void td3_policy_action_kernel(int B, const float *batch, const float *z, float *x0a, float *a_out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * 16) return;
    const int r = g >> 4, c = g & 15;
    float v = 0.f;
    if (c < SD) v = batch[(size_t)r * ERLFILL_REC + c];
    else if (c < QIN) { v = tanhf(z[(size_t)r * AD + c - SD]); a_out[(size_t)r * AD + c - SD] = v; }
    x0a[g] = v;
}
// actor loss = -mean(q1); dz = dL/da (1 - a^2) with dL/da = columns 2..11 of the critic's input gradient
__global__ void td3_actor_dz_kernel(int B, const float *g16, const float *a, float *dz) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * AD) return;
    const int r = g / AD, j = g % AD;
    const float av = a[g];
    dz[g] = g16[(size_t)r * 16 + SD + j] * (1.0f - av * av);
}
__global__ void __launch_bounds__(1024) neg_mean_kernel(int B, const float *q, float *out) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) s += q[r];
    const float t = block_sum1024(s, scratch);
    if (threadIdx.x == 0) out[0] = -t / (float)B;
}

// ---- SAC row kernels ----------------------------------------------------------------------------------------------
// GaussianPolicy.sample (sac_agent.py:45-84) from ml[r] = [mean(10) | log_std(10)]: a = tanh(mean + exp(clamp(ls,-20,2)) eps),
// log_prob = sum(-eps^2/2 - ls_c - log sqrt(2 pi) - log(1 - a^2 + 1e-7)).  x0[r] = [state | a | 0]; a, sigma*eps kept for the backward.
__global__ void sac_sample_kernel(int B, const float *batch, int state_col, const float *ml, const float *eps, uint64_t seed, uint32_t update_idx,
                                  uint32_t salt, uint32_t row_id_base, float *x0, float *a_out, float *se_out, float *logp) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    float lp = 0.f;
    float *row = x0 + (size_t)r * 16;
    row[0] = batch[(size_t)r * ERLFILL_REC + state_col]; row[1] = batch[(size_t)r * ERLFILL_REC + state_col + 1];
#pragma unroll
    for (int j = 0; j < AD; ++j) {
        const float mean = ml[(size_t)r * 20 + j], ls = fminf(fmaxf(ml[(size_t)r * 20 + AD + j], -20.0f), 2.0f);
        const float e = normal_draw(eps, r, j, seed, row_id_base + (uint32_t)r, update_idx, salt);
        const float se = expf(ls) * e;
        const float a = tanhf(mean + se);
        lp += -0.5f * e * e - ls - 0.9189385332046727f - logf(1.0f - a * a + 1e-7f);
        row[SD + j] = a;
        if (a_out) { a_out[(size_t)r * AD + j] = a; se_out[(size_t)r * AD + j] = se; }
    }
    row[12] = row[13] = row[14] = row[15] = 0.f;
    logp[r] = lp;
}
// target y = r + (1-d) gamma (min(q1', q2') - alpha logp'); per-network MSE and dq (sac_agent.py:236-252)
__global__ void __launch_bounds__(1024) sac_q_loss_kernel(int B, const float *batch, const float *q1, const float *q2, const float *q1t,
                                                          const float *q2t, const float *logp_next, float gamma, float alpha, float *dq1,
                                                          float *dq2, float *report) {
    __shared__ float scratch[32];
    float l1 = 0.f, l2 = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) {
        const float rew = batch[(size_t)r * ERLFILL_REC + 12], d = batch[(size_t)r * ERLFILL_REC + 15];
        const float y = rew + (1.0f - d) * gamma * (fminf(q1t[r], q2t[r]) - alpha * logp_next[r]);
        const float e1 = q1[r] - y, e2 = q2[r] - y;
        dq1[r] = 2.0f * e1 / (float)B; dq2[r] = 2.0f * e2 / (float)B;
        l1 = fmaf(e1, e1, l1); l2 = fmaf(e2, e2, l2);
    }
    const float s1 = block_sum1024(l1, scratch), s2 = block_sum1024(l2, scratch);
    if (threadIdx.x == 0) { report[0] = s1 / (float)B; report[1] = s2 / (float)B; }
}
// policy loss = mean(alpha logp - q1(s, a)); gradient wrt [mean | log_std]:
//   dL/da = alpha/B * 2a/(1 - a^2 + 1e-7) + g (g = critic input gradient for dq = -1/B), dz = dL/da (1 - a^2),
//   dmean = dz, dlog_std = (dz * sigma*eps - alpha/B) inside the clamp range, 0 outside          (sac_agent.py:265-271)
__global__ void sac_policy_grad_kernel(int B, const float *g16, const float *a, const float *se, const float *ml, float alpha, float *dml) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * AD) return;
    const int r = g / AD, j = g % AD;
    const float av = a[g], om = 1.0f - av * av;
    const float dla = alpha / (float)B * 2.0f * av / (om + 1e-7f) + g16[(size_t)r * 16 + SD + j];
    const float dz = dla * om;
    const float ls = ml[(size_t)r * 20 + AD + j];
    dml[(size_t)r * 20 + j] = dz;
    dml[(size_t)r * 20 + AD + j] = (ls >= -20.0f && ls <= 2.0f) ? dz * se[g] - alpha / (float)B : 0.f;
}
__global__ void __launch_bounds__(1024) sac_policy_loss_kernel(int B, const float *logp, const float *q1, float alpha, float *out) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) s += alpha * logp[r] - q1[r];
    const float t = block_sum1024(s, scratch);
    if (threadIdx.x == 0) out[0] = t / (float)B;
}
// ga = (ga + gb) * (h > 0)
__global__ void add_relu_bwd_kernel(size_t n, float *ga, const float *gb, const float *h) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ga[i] = h[i] > 0.f ? ga[i] + gb[i] : 0.f;
}

// ---- optimiser ----------------------------------------------------------------------------------------------------
// torch.optim.Adam (betas 0.9/0.999, eps 1e-8, no clipping) on a flat buffer; optional soft target update afterwards
__global__ void adam_kernel(int n, float *p, float *m, float *v, const float *g, float lr, float bc1, float bc2_sqrt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * 0.1f;
    const float vi = v[i] * 0.999f + 0.001f * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + 1e-8f));
}
__global__ void soft_kernel(int n, const float *p, float *pt, float tau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pt[i] = tau * p[i] + (1.0f - tau) * pt[i];
}
static void adam(cudaStream_t st, int n, float *p, float *m, float *v, const float *g, double lr, int step) {
    const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = sqrt(1.0 - pow(0.999, (double)step));
    adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, p, m, v, g, (float)lr, (float)bc1, (float)bc2);
}

// ---- action selection (rollout) -----------------------------------------------------------------------------------
// TD3: a = clip(tanh(z) + noise_scale * n, -1, 1) -> 0.5 (a + 1)(hi - lo) + lo                      (td3_agent.py:221-236, 198-219)
// SAC: a = tanh(mean + sigma eps) (or tanh(mean) when deterministic) -> a (hi - lo)/2 + (hi + lo)/2    (sac_agent.py:186-205, 163-184)
__global__ void act_finish_kernel(int n, int sac, int deterministic, const float *zml, const float *noise, float noise_scale, const float *lo,
                                  const float *hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx, int row0, float *action, float *norm_action) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * AD) return;
    const int r = g / AD, j = g % AD;
    const size_t row = (size_t)row0 + r;
    float a;
    if (!sac) {
        a = tanhf(zml[(size_t)r * AD + j]);
        if (noise_scale != 0.f) a += noise_scale * normal_draw(noise, row, j, seed, row_id_base + (uint32_t)row, step_idx, 1);
        a = fminf(fmaxf(a, -1.0f), 1.0f);
    } else {
        const float mean = zml[(size_t)r * 20 + j];
        if (deterministic) a = tanhf(mean);
        else {
            const float ls = fminf(fmaxf(zml[(size_t)r * 20 + AD + j], -20.0f), 2.0f);
            a = tanhf(mean + expf(ls) * normal_draw(noise, row, j, seed, row_id_base + (uint32_t)row, step_idx, 1));
        }
    }
    if (norm_action) norm_action[row * AD + j] = a;
    action[row * AD + j] = !sac ? 0.5f * (a + 1.0f) * (hi[j] - lo[j]) + lo[j] : a * ((hi[j] - lo[j]) * 0.5f) + (hi[j] + lo[j]) * 0.5f;
}

// ---- workspaces -----------------------------------------------------------------------------------------------------
constexpr int kActChunk = 131072;    // rows per pass of the action-selection MLP (2048 tiles of the 128 x 128 GEMM: 6.9 waves of 2 x 148 resident CTAs)
struct Ws {
    float *cgrad, *agrad, *scal;
    float *x0, *h1a, *h2a, *h1b, *h2b, *ah1, *ah2, *z, *ml, *a, *se, *logp, *logp2;
    float *q1, *q2, *q1t, *q2t, *dq1, *dq2, *ga, *gb, *g16, *dz, *dml, *part;
};
static size_t align4(size_t n) { return (n + 3) & ~(size_t)3; }
static size_t carve(Ws *w, float *base, int B) {
    size_t o = 0;
    float *dummy;
    auto take = [&](float **p, size_t n) { if (w) *p = base + o; else (void)p; o += align4(n); };
#define TK(field, n) take(w ? &w->field : &dummy, (n))
    const size_t b = (size_t)B;
    TK(cgrad, ERLFILL_TD3_CRITIC_PARAMS); TK(agrad, ERLFILL_SAC_POLICY_PARAMS); TK(scal, 16);
    TK(x0, b * 16); TK(h1a, b * H); TK(h2a, b * H); TK(h1b, b * H); TK(h2b, b * H); TK(ah1, b * H); TK(ah2, b * H);
    TK(z, b * AD); TK(ml, b * 20); TK(a, b * AD); TK(se, b * AD); TK(logp, b); TK(logp2, b);
    TK(q1, b); TK(q2, b); TK(q1t, b); TK(q2t, b); TK(dq1, b); TK(dq2, b);
    TK(ga, b * H); TK(gb, b * H); TK(g16, b * 16); TK(dz, b * AD); TK(dml, b * 20); TK(part, kPartFloats);
#undef TK
    return o;
}
// Q(x0) for one Q net: h1/h2 as given
static void q_forward(cudaStream_t st, int B, const float *P, const Net &n, const float *x, int ldx, float *h1, float *h2, float *q) {
    mlp_forward(st, B, P, n, x, ldx, h1, h2, q, 1);
}

}  // namespace t1
}  // namespace erlfill

using namespace erlfill;
using namespace erlfill::t1;

extern "C" {

size_t erlfill_t1_workspace_bytes(int batch) {
    if (batch < 1) return 0;
    return carve(nullptr, nullptr, batch) * sizeof(float);
}
size_t erlfill_t1_act_workspace_bytes(void) { return ((size_t)kActChunk * (2 * H + 20)) * sizeof(float); }

int erlfill_t1_grad_buckets(void *d_workspace, int batch, float **critic_grad, float **actor_grad) {
    if (!d_workspace || batch < 1) return ERLFILL_ERR_ARG;
    Ws w;
    carve(&w, (float *)d_workspace, batch);
    if (critic_grad) *critic_grad = w.cgrad;
    if (actor_grad) *actor_grad = w.agrad;
    return ERLFILL_OK;
}

static int act_common(int sac, const float *P, int n, const float *d_state, int ld_state, const float *d_noise, float noise_scale, int deterministic,
                      const float *d_lo, const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx, float *d_action,
                      float *d_norm_action, void *d_workspace, cudaStream_t st) {
    if (!P || n <= 0 || !d_state || !d_lo || !d_hi || !d_action || !d_workspace) return ERLFILL_ERR_ARG;
    float *h1 = (float *)d_workspace, *h2 = h1 + (size_t)kActChunk * H, *out = h2 + (size_t)kActChunk * H;
    for (int r0 = 0; r0 < n; r0 += kActChunk) {
        const int m = n - r0 < kActChunk ? n - r0 : kActChunk;
        const float *x = d_state + (size_t)r0 * ld_state;
        if (!sac) mlp_forward(st, m, P, kTd3Actor, x, ld_state, h1, h2, out, AD);
        else {
            gemm(st, m, H, SD, x, ld_state, 1, P + kSacTrunk.W0, 1, SD, h1, H, P + kSacTrunk.B0, ACT_RELU);
            gemm(st, m, H, H, h1, H, 1, P + kSacTrunk.W1, 1, H, h2, H, P + kSacTrunk.B1, ACT_RELU);
            gemm(st, m, AD, H, h2, H, 1, P + kSacTrunk.W2, 1, H, out, 20, P + kSacTrunk.B2, ACT_NONE);
            if (!deterministic) gemm(st, m, AD, H, h2, H, 1, P + kSacLsW, 1, H, out + AD, 20, P + kSacLsB, ACT_NONE);
        }
        act_finish_kernel<<<(m * AD + 255) / 256, 256, 0, st>>>(m, sac, deterministic, out, d_noise, noise_scale, d_lo, d_hi, seed, row_id_base,
                                                                 step_idx, r0, d_action, d_norm_action);
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_td3_act(const float *d_actor, int n, const float *d_state, const float *d_randn, float noise_scale, const float *d_lo,
                    const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx, float *d_action, float *d_norm_action,
                    void *d_workspace, void *stream) {
    return act_common(0, d_actor, n, d_state, SD, d_randn, noise_scale, 0, d_lo, d_hi, seed, row_id_base, step_idx, d_action, d_norm_action,
                      d_workspace, (cudaStream_t)stream);
}
int erlfill_sac_act(const float *d_policy, int n, const float *d_state, const float *d_eps, int deterministic, const float *d_lo,
                    const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx, float *d_action, float *d_norm_action,
                    void *d_workspace, void *stream) {
    return act_common(1, d_policy, n, d_state, SD, d_eps, 1.0f, deterministic, d_lo, d_hi, seed, row_id_base, step_idx, d_action, d_norm_action,
                      d_workspace, (cudaStream_t)stream);
}

int erlfill_td3_update(const erlfill_td3_nets *n, const erlfill_td3_hyper *h, const float *d_batch, const float *d_noise, void *d_workspace,
                       int phases, float *d_report, void *stream) {
    if (!n || !h || !d_batch || !d_workspace) return ERLFILL_ERR_ARG;
    if (!n->actor || !n->actor_target || !n->critic || !n->critic_target || !n->actor_m || !n->actor_v || !n->critic_m || !n->critic_v)
        return ERLFILL_ERR_ARG;
    const int B = h->batch;
    if (B < 1 || h->critic_step < 1 || (h->update_actor && h->actor_step < 1)) return ERLFILL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, (float *)d_workspace, B);
    const int NC = ERLFILL_TD3_CRITIC_PARAMS, NA = ERLFILL_TD3_ACTOR_PARAMS;
    if (phases & ERLFILL_PHASE_CRITIC_GRAD) {
        // target: a' = clamp(actor_target(s') + clipped noise), y from the twin target critics
        mlp_forward(st, B, n->actor_target, kTd3Actor, d_batch + 13, ERLFILL_REC, w.ah1, w.ah2, w.z, AD);
        td3_target_action_kernel<<<(B * 16 + 255) / 256, 256, 0, st>>>(B, d_batch, w.z, d_noise, h->policy_noise, h->noise_clip, h->seed,
                                                                        h->update_idx, h->rank << 20, w.x0);
        q_forward(st, B, n->critic_target, kQ1, w.x0, 16, w.h1a, w.h2a, w.q1t);
        q_forward(st, B, n->critic_target, kQ2, w.x0, 16, w.h1b, w.h2b, w.q2t);
        // current estimates on (s, a) read in place from the replay rows
        q_forward(st, B, n->critic, kQ1, d_batch, ERLFILL_REC, w.h1a, w.h2a, w.q1);
        q_forward(st, B, n->critic, kQ2, d_batch, ERLFILL_REC, w.h1b, w.h2b, w.q2);
        td3_critic_loss_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.q1, w.q2, w.q1t, w.q2t, h->gamma, w.dq1, w.dq2, w.scal);
        mlp_backward(st, w.part, B, n->critic, w.cgrad, kQ1, d_batch, ERLFILL_REC, w.h1a, w.h2a, w.dq1, 1, w.ga, w.gb, nullptr, 0);
        mlp_backward(st, w.part, B, n->critic, w.cgrad, kQ2, d_batch, ERLFILL_REC, w.h1b, w.h2b, w.dq2, 1, w.ga, w.gb, nullptr, 0);
    }
    if (phases & ERLFILL_PHASE_CRITIC_APPLY) adam(st, NC, n->critic, n->critic_m, n->critic_v, w.cgrad, h->critic_lr, h->critic_step);
    if (h->update_actor) {
        if (phases & ERLFILL_PHASE_ACTOR_GRAD) {
            mlp_forward(st, B, n->actor, kTd3Actor, d_batch, ERLFILL_REC, w.ah1, w.ah2, w.z, AD);
            td3_policy_action_kernel<<<(B * 16 + 255) / 256, 256, 0, st>>>(B, d_batch, w.z, w.x0, w.a);
            q_forward(st, B, n->critic, kQ1, w.x0, 16, w.h1a, w.h2a, w.q1);
            neg_mean_kernel<<<1, 1024, 0, st>>>(B, w.q1, w.scal + 2);
            fill_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, -1.0f / (float)B, w.dq1);
            mlp_backward(st, w.part, B, n->critic, nullptr, kQ1, w.x0, 16, w.h1a, w.h2a, w.dq1, 1, w.ga, w.gb, w.g16, 16);
            td3_actor_dz_kernel<<<(B * AD + 255) / 256, 256, 0, st>>>(B, w.g16, w.a, w.dz);
            mlp_backward(st, w.part, B, n->actor, w.agrad, kTd3Actor, d_batch, ERLFILL_REC, w.ah1, w.ah2, w.dz, AD, w.ga, w.gb, nullptr, 0);
        }
        if (phases & ERLFILL_PHASE_ACTOR_APPLY) {
            adam(st, NA, n->actor, n->actor_m, n->actor_v, w.agrad, h->actor_lr, h->actor_step);
            soft_kernel<<<(NC + 255) / 256, 256, 0, st>>>(NC, n->critic, n->critic_target, h->tau);
            soft_kernel<<<(NA + 255) / 256, 256, 0, st>>>(NA, n->actor, n->actor_target, h->tau);
        }
    }
    if (d_report && (phases & ERLFILL_PHASE_ACTOR_APPLY))      // [critic_loss, mean q1, actor_loss (stale when the actor step was skipped)]
        cudaMemcpyAsync(d_report, w.scal, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_sac_update(const erlfill_sac_nets *n, const erlfill_sac_hyper *h, const float *d_batch, const float *d_eps_next,
                       const float *d_eps_new, void *d_workspace, int phases, float *d_report, void *stream) {
    if (!n || !h || !d_batch || !d_workspace) return ERLFILL_ERR_ARG;
    if (!n->policy || !n->q || !n->q_target || !n->policy_m || !n->policy_v || !n->q_m || !n->q_v) return ERLFILL_ERR_ARG;
    const int B = h->batch;
    if (B < 1 || h->adam_step < 1) return ERLFILL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, (float *)d_workspace, B);
    const int NQ = ERLFILL_SAC_Q_PARAMS, NP = ERLFILL_SAC_POLICY_PARAMS;
    const float *P = n->policy;
    auto policy_forward = [&](const float *x, int ldx) {
        gemm(st, B, H, SD, x, ldx, 1, P + kSacTrunk.W0, 1, SD, w.ah1, H, P + kSacTrunk.B0, ACT_RELU);
        gemm(st, B, H, H, w.ah1, H, 1, P + kSacTrunk.W1, 1, H, w.ah2, H, P + kSacTrunk.B1, ACT_RELU);
        gemm(st, B, AD, H, w.ah2, H, 1, P + kSacTrunk.W2, 1, H, w.ml, 20, P + kSacTrunk.B2, ACT_NONE);
        gemm(st, B, AD, H, w.ah2, H, 1, P + kSacLsW, 1, H, w.ml + AD, 20, P + kSacLsB, ACT_NONE);
    };
    if (phases & ERLFILL_PHASE_CRITIC_GRAD) {
        policy_forward(d_batch + 13, ERLFILL_REC);
        sac_sample_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, d_batch, 13, w.ml, d_eps_next, h->seed, h->update_idx, 4, h->rank << 20, w.x0, nullptr, nullptr, w.logp2);
        q_forward(st, B, n->q_target, kQ1, w.x0, 16, w.h1a, w.h2a, w.q1t);
        q_forward(st, B, n->q_target, kQ2, w.x0, 16, w.h1b, w.h2b, w.q2t);
        q_forward(st, B, n->q, kQ1, d_batch, ERLFILL_REC, w.h1a, w.h2a, w.q1);
        q_forward(st, B, n->q, kQ2, d_batch, ERLFILL_REC, w.h1b, w.h2b, w.q2);
        sac_q_loss_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.q1, w.q2, w.q1t, w.q2t, w.logp2, h->gamma, h->alpha, w.dq1, w.dq2, w.scal);
        mlp_backward(st, w.part, B, n->q, w.cgrad, kQ1, d_batch, ERLFILL_REC, w.h1a, w.h2a, w.dq1, 1, w.ga, w.gb, nullptr, 0);
        mlp_backward(st, w.part, B, n->q, w.cgrad, kQ2, d_batch, ERLFILL_REC, w.h1b, w.h2b, w.dq2, 1, w.ga, w.gb, nullptr, 0);
    }
    // q1 and q2 have separate Adam optimisers with the same hyper-parameters: one element-wise pass over [q1 | q2]
    if (phases & ERLFILL_PHASE_CRITIC_APPLY) adam(st, NQ, n->q, n->q_m, n->q_v, w.cgrad, h->lr, h->adam_step);
    if (phases & ERLFILL_PHASE_ACTOR_GRAD) {
        policy_forward(d_batch, ERLFILL_REC);
        sac_sample_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, d_batch, 0, w.ml, d_eps_new, h->seed, h->update_idx, 5, h->rank << 20, w.x0, w.a, w.se, w.logp);
        q_forward(st, B, n->q, kQ1, w.x0, 16, w.h1a, w.h2a, w.q1);
        sac_policy_loss_kernel<<<1, 1024, 0, st>>>(B, w.logp, w.q1, h->alpha, w.scal + 2);
        fill_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, -1.0f / (float)B, w.dq1);
        mlp_backward(st, w.part, B, n->q, nullptr, kQ1, w.x0, 16, w.h1a, w.h2a, w.dq1, 1, w.ga, w.gb, w.g16, 16);
        sac_policy_grad_kernel<<<(B * AD + 255) / 256, 256, 0, st>>>(B, w.g16, w.a, w.se, w.ml, h->alpha, w.dml);
        float *G = w.agrad;
        // heads: mean (W2/B2 slot of the trunk layout) and log_std
        gemm_wgrad(st, w.part, B, AD, H, w.dml, 20, w.ah2, H, G + kSacTrunk.W2); colsum(st, w.part, B, AD, w.dml, 20, nullptr, 0, 0, G + kSacTrunk.B2);
        gemm_wgrad(st, w.part, B, AD, H, w.dml + AD, 20, w.ah2, H, G + kSacLsW); colsum(st, w.part, B, AD, w.dml + AD, 20, nullptr, 0, 0, G + kSacLsB);
        gemm(st, B, H, AD, w.dml, 20, 1, P + kSacTrunk.W2, H, 1, w.ga, H, nullptr, ACT_NONE);
        gemm(st, B, H, AD, w.dml + AD, 20, 1, P + kSacLsW, H, 1, w.gb, H, nullptr, ACT_NONE);
        add_relu_bwd_kernel<<<(unsigned)(((size_t)B * H + 255) / 256), 256, 0, st>>>((size_t)B * H, w.ga, w.gb, w.ah2);
        gemm_wgrad(st, w.part, B, H, H, w.ga, H, w.ah1, H, G + kSacTrunk.W1); colsum(st, w.part, B, H, w.ga, H, nullptr, 0, 0, G + kSacTrunk.B1);
        gemm(st, B, H, H, w.ga, H, 1, P + kSacTrunk.W1, H, 1, w.gb, H, nullptr, ACT_NONE);
        relu_bwd_kernel<<<(unsigned)(((size_t)B * H + 255) / 256), 256, 0, st>>>((size_t)B * H, w.gb, w.ah1);
        gemm_wgrad(st, w.part, B, H, SD, w.gb, H, d_batch, ERLFILL_REC, G + kSacTrunk.W0); colsum(st, w.part, B, H, w.gb, H, nullptr, 0, 0, G + kSacTrunk.B0);
    }
    if (phases & ERLFILL_PHASE_ACTOR_APPLY) {
        adam(st, NP, n->policy, n->policy_m, n->policy_v, w.agrad, h->lr, h->adam_step);
        soft_kernel<<<(NQ + 255) / 256, 256, 0, st>>>(NQ, n->q, n->q_target, h->tau);
        if (d_report) cudaMemcpyAsync(d_report, w.scal, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st);   // [q1_loss, q2_loss, policy_loss]
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
