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
// Structure: parameters, gradients and Adam moments are flat fp32 buffers in the reference's
// parameters() order, so clip / Adam / soft update / the data-parallel all-reduce are single passes.
// The dense layers run through one strided, split-K capable tensor-core GEMM (mma.sync TF32, 3xTF32 compensated,
// this file); the update is dependent-launch/latency bound at these sizes (5.9 GFLOP at B=4096), see DESIGN.md section 4.  Every reduction is
// tree-ordered without atomics, so an update is bit-reproducible.
#include <math.h>

#include "common.cuh"
#include "mlp.cuh"

namespace erlfill {
using namespace mlp;

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
    for (int c = 0; c < 3; ++c) {
        float s = 0.f;
        for (int r = threadIdx.x; r < B; r += blockDim.x) s += batch[(size_t)r * ERLFILL_REC + 16 + c];
        const float mean = block_sum(s, scratch) / (float)B;
        float q = 0.f;
        for (int r = threadIdx.x; r < B; r += blockDim.x) { const float d = batch[(size_t)r * ERLFILL_REC + 16 + c] - mean; q = fmaf(d, d, q); }
        const float var = block_sum(q, scratch) / (float)(B - 1);
        if (threadIdx.x == 0) { stats[c] = mean; stats[3 + c] = sqrtf(var); mean_out[c] = mean; }
    }
}

// critic input rows [B][16] = [state(2) | action(10) | emotion standardised(3) | 0]
//   mode 0: (s, batch action, e)   mode 1: (s', actions_ext, 0)  [target]   mode 2: (s, actions_ext, e)
__global__ void build_x0_kernel(int B, int mode, const float *batch, const float *actions_ext, const float *stats, float *x0) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * 16) return;
    const int r = g >> 4, c = g & 15;
    const float *row = batch + (size_t)r * ERLFILL_REC;
    float v = 0.f;
    if (c < 2) v = mode == 1 ? row[13 + c] : row[c];
    else if (c < 12) v = mode == 0 ? row[c] : actions_ext[(size_t)r * 10 + (c - 2)];
    else if (c < 15) v = mode == 1 ? 0.f : (row[16 + (c - 12)] - stats[c - 12]) / (stats[3 + (c - 12)] + 1e-6f);
    x0[g] = v;
}
// actor input rows [B][8] = [state | emotion | 0]; mode 1: (s', next_emotion broadcast)
__global__ void build_xa_kernel(int B, int mode, const float *batch, const float *next_emotion, float *xa) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * 8) return;
    const int r = g >> 3, c = g & 7;
    const float *row = batch + (size_t)r * ERLFILL_REC;
    float v = 0.f;
    if (c < 2) v = mode == 1 ? row[13 + c] : row[c];
    else if (c < 5) v = mode == 1 ? next_emotion[c - 2] : row[16 + (c - 2)];
    xa[g] = v;
}

// LayerNorm(256) + activation, one warp per row.  z (in: x W^T + b) is overwritten with xhat.
__global__ void __launch_bounds__(256) ln_act_fwd_kernel(int B, float *z, const float *gamma, const float *beta, int act,
                                                         float *h, float *rstd_out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (row >= B) return;
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = z[(size_t)row * 256 + l + 32 * i]; s += v[i]; }
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
        float y = xh * gamma[c] + beta[c];
        y = act == ACT_LEAKY ? (y > 0.f ? y : 0.1f * y) : fmaxf(y, 0.f);
        z[(size_t)row * 256 + c] = xh;
        h[(size_t)row * 256 + c] = y;
    }
    if (l == 0) rstd_out[row] = rstd;
}

// backward of (LayerNorm + activation): dh -> dy (in place, for the gamma/beta column sums) and dz.
__global__ void __launch_bounds__(256) ln_act_bwd_kernel(int B, float *dh, const float *xhat, const float *rstd,
                                                         const float *gamma, const float *beta, int act, float *dz) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (row >= B) return;
    float dxh[8], xh[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = l + 32 * i;
        xh[i] = xhat[(size_t)row * 256 + c];
        const float y = xh[i] * gamma[c] + beta[c];
        float g = dh[(size_t)row * 256 + c];
        g *= act == ACT_LEAKY ? (y > 0.f ? 1.f : 0.1f) : (y > 0.f ? 1.f : 0.f);
        dh[(size_t)row * 256 + c] = g;                         // dy
        dxh[i] = g * gamma[c];
        s1 += dxh[i];
        s2 = fmaf(dxh[i], xh[i], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float r = rstd[row], m1 = s1 * (1.0f / 256.0f), m2 = s2 * (1.0f / 256.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) dz[(size_t)row * 256 + l + 32 * i] = r * (dxh[i] - m1 - xh[i] * m2);
}

// g *= (h > 0) for ReLU layers without LayerNorm
__global__ void leaky_bwd_kernel(size_t n, float *g, const float *h) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = h[i] > 0.f ? g[i] : 0.1f * g[i];
}

// critic head: q = clamp(h3 . w10 + b10, +-1e6), one warp per row
__global__ void __launch_bounds__(256) critic_head_kernel(int B, const float *h3, const float *w10, const float *b10, float *q) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (row >= B) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s = fmaf(h3[(size_t)row * 128 + l + 32 * i], w10[l + 32 * i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) q[row] = fminf(fmaxf(s + b10[0], -1e6f), 1e6f);
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

// dh3[r][c] = dq[r] * w10[c] * (h3 > 0)    and  dW10 / db10 come from colsum(dq*h3) / sum(dq)
__global__ void critic_head_bwd_kernel(int B, const float *dq, const float *w10, const float *h3, float *dh3) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * 128) return;
    const int r = g >> 7, c = g & 127;
    dh3[g] = h3[g] > 0.f ? dq[r] * w10[c] : 0.f;
}
// ---- actor head: z4 -> physical action, and its backward ---------------------------------------
__global__ void actor_head_fwd_kernel(int B, const float *z4, const float *xa, const float *ew, const float *scale,
                                      const float *offset, float *out, float *aux /*[B][10][3]: base, alpha*sflip*mod2.., see bwd*/) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * 10) return;
    const int r = g / 10, c = g % 10;
    const float z = z4[g];
    const float base = z / (1.0f + fabsf(z));
    const float cur = xa[r * 8 + 2], cons = xa[r * 8 + 3], anx = xa[r * 8 + 4];
    float e = cur * ew[c];
    e = fmaf(cons, ew[10 + c], e);
    e = fmaf(anx, ew[20 + c], e);
    const float raw = tanhf(e * 1.5f);
    float alpha = 0.1f + 0.9f * anx;
    alpha *= (1.0f - 0.5f * cons);
    alpha *= (1.0f + 0.5f * cur);
    alpha = fminf(fmaxf(alpha, 0.01f), 1.0f);
    const float om = 1.0f - fabsf(base);
    const float sf = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
    out[g] = (base + alpha * raw * sf * (om * om)) * scale[c] + offset[c];
    if (aux) { aux[(size_t)g * 3 + 0] = base; aux[(size_t)g * 3 + 1] = raw; aux[(size_t)g * 3 + 2] = alpha; }
}

// d(out) -> dz4, and the per-row contribution to d(emotion_weights): dew_part[r][k][c] folded by colsum later.
// dout already contains the critic path; the diversity term -0.03*mean_rows(std_unbiased(out,dim=1)) is added here.
__global__ void actor_head_bwd_kernel(int B, const float *dout_critic, const float *out, const float *aux, const float *xa,
                                      const float *scale, const float *z4, float *dz4, float *draw_e /*[B][10]*/) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    float o[10], mean = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c) { o[c] = out[(size_t)r * 10 + c]; mean += o[c]; }
    mean *= 0.1f;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 10; ++c) { const float d = o[c] - mean; var = fmaf(d, d, var); }
    const float sd = sqrtf(var / 9.0f);
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        const size_t g = (size_t)r * 10 + c;
        float go = dout_critic[g];
        go += (-0.03f / (float)B) * (o[c] - mean) / (9.0f * sd);                      // d std / d o_c
        const float gr = go * scale[c];                                                   // d raw_action
        const float base = aux[g * 3 + 0], raw = aux[g * 3 + 1], alpha = aux[g * 3 + 2];
        const float om = 1.0f - fabsf(base);
        const float sf = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
        // d raw_action / d base = 1 + alpha*raw*sf * 2*om*(-sign(base)) = 1 + 2*alpha*raw*om*sf*sf (sign has zero grad)
        const float dbase = gr * (1.0f + 2.0f * alpha * raw * om * (sf * sf));
        const float z = z4[g], den = 1.0f + fabsf(z);
        dz4[g] = dbase / (den * den);                                                     // softsign'
        // d raw_action / d (e @ W_e)[c] = alpha*sf*om^2 * (1 - raw^2) * 1.5
        draw_e[g] = gr * alpha * sf * (om * om) * (1.0f - raw * raw) * 1.5f;
    }
}
// dEW[k][c] = sum_r xa[r][2+k] * draw_e[r][c]   (30 outputs): per-slice partials part[slice][30], summed by splitk_reduce_kernel
__global__ void __launch_bounds__(256) ew_grad_part_kernel(int B, const float *xa, const float *draw_e, float *part) {
    __shared__ float red[8][32];
    const int o = threadIdx.x & 31, rl = threadIdx.x >> 5, k = o / 10, c = o % 10;
    const int per = (B + kColSplits - 1) / kColSplits, r0 = blockIdx.x * per, r1 = min(B, r0 + per);
    float s = 0.f;
    if (o < 30)
        for (int r = r0 + rl; r < r1; r += 8) s = fmaf(xa[r * 8 + 2 + k], draw_e[(size_t)r * 10 + c], s);
    red[rl][o] = s;
    __syncthreads();
    if (rl == 0 && o < 30) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][o];
        part[blockIdx.x * 30 + o] = t;
    }
}

// actor loss value (reported): -mean(q2) - 0.03*mean_rows(std) + 1e-4 * sum_t ||theta_t||
__global__ void __launch_bounds__(1024) actor_loss_kernel(int B, const float *q2, const float *out, const float *actor, float *loss_out,
                                                          float *tensor_norms /*[9]*/) {
    __shared__ float scratch[32];
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
    float l2 = 0.f;
    for (int t = 0; t < kActorTensors; ++t) {
        float q = 0.f;
        for (int i = c_actor_seg[t] + threadIdx.x; i < c_actor_seg[t + 1]; i += blockDim.x) q = fmaf(actor[i], actor[i], q);
        const float nrm = sqrtf(block_sum(q, scratch));
        if (threadIdx.x == 0) tensor_norms[t] = nrm;
        l2 += nrm;
    }
    if (threadIdx.x == 0) loss_out[0] = -qs / (float)B - 0.03f * ds / (float)B + 1e-4f * l2;
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

// d(out)[r][c] = dx0[r][2 + c]
__global__ void slice_action_grad_kernel(int B, const float *g16, float *d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 10) d[i] = g16[(i / 10) * 16 + 2 + (i % 10)];
}
__global__ void soft_update_kernel(int n, const float *p, float *pt, float tau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pt[i] = tau * p[i] + (1.0f - tau) * pt[i];
}
__global__ void report_kernel(const float *scal, float *out) {
    out[0] = scal[4]; out[1] = scal[5]; out[2] = scal[2]; out[3] = scal[3];
}

// ---- workspace -------------------------------------------------------------------------------
struct Ws {
    float *cgrad, *agrad;                       // flat gradient buckets (cgrad tail holds the 3 emotion means)
    float *stats, *scal, *ssq, *step;                // stats[6]; scal: [2..3] lrs [4] closs [5] aloss [6..14] norms; ssq: [0:64] critic, [64:128] actor sum-of-squares partials
    float *x0, *xhat1, *h1, *xhat2, *h2, *h3, *rstd1, *rstd2, *q, *tq, *q2, *dq;
    float *xa, *a1, *a2, *a3, *z4, *aout, *aux, *atgt;
    float *g256a, *g256b, *g128, *g16, *ga128a, *ga128b, *ga64, *gz4, *drawe, *part;
    size_t part_floats;
};
static size_t align16(size_t n) { return (n + 3) & ~(size_t)3; }
static size_t carve(Ws *w, float *base, int B) {
    size_t o = 0;
    auto take = [&](float **p, size_t n) { if (w) *p = base + o; o += align16(n); };
    float *dummy;
    take(w ? &w->cgrad : &dummy, kCriticGradPad); take(w ? &w->agrad : &dummy, kActorGradPad);
    take(w ? &w->stats : &dummy, 8); take(w ? &w->scal : &dummy, 32); take(w ? &w->ssq : &dummy, 2 * kSumsqBlocks); take(w ? &w->step : &dummy, 4);
    const size_t b = (size_t)B;
    take(w ? &w->x0 : &dummy, b * 16); take(w ? &w->xhat1 : &dummy, b * 256); take(w ? &w->h1 : &dummy, b * 256);
    take(w ? &w->xhat2 : &dummy, b * 256); take(w ? &w->h2 : &dummy, b * 256); take(w ? &w->h3 : &dummy, b * 128);
    take(w ? &w->rstd1 : &dummy, b); take(w ? &w->rstd2 : &dummy, b); take(w ? &w->q : &dummy, b);
    take(w ? &w->tq : &dummy, b); take(w ? &w->q2 : &dummy, b); take(w ? &w->dq : &dummy, b);
    take(w ? &w->xa : &dummy, b * 8); take(w ? &w->a1 : &dummy, b * 128); take(w ? &w->a2 : &dummy, b * 128);
    take(w ? &w->a3 : &dummy, b * 64); take(w ? &w->z4 : &dummy, b * 10); take(w ? &w->aout : &dummy, b * 10);
    take(w ? &w->aux : &dummy, b * 30); take(w ? &w->atgt : &dummy, b * 10);
    take(w ? &w->g256a : &dummy, b * 256); take(w ? &w->g256b : &dummy, b * 256); take(w ? &w->g128 : &dummy, b * 128);
    take(w ? &w->g16 : &dummy, b * 16); take(w ? &w->ga128a : &dummy, b * 128); take(w ? &w->ga128b : &dummy, b * 128);
    take(w ? &w->ga64 : &dummy, b * 64); take(w ? &w->gz4 : &dummy, b * 10); take(w ? &w->drawe : &dummy, b * 10);
    const size_t part = (size_t)32 * 256 * 256;              // split-K partials (up to 32 splits of a 256x256 dW)
    take(w ? &w->part : &dummy, part);
    if (w) w->part_floats = part;
    return o;
}

// ---- launch helpers ---------------------------------------------------------------------------
// update.cu's launch helpers bind the header's (mlp.cuh) to this workspace's split scratch
static void gemm_wgrad(cudaStream_t st, const Ws &w, int Bt, int N, int K, const float *dZ, int ldz, const float *X, int ldx, float *dW) {
    mlp::gemm_wgrad(st, w.part, Bt, N, K, dZ, ldz, X, ldx, dW);
}
static void colsum(cudaStream_t st, const Ws &w, int rows, int cols, const float *X, int ld, const float *Y, int ldy, int ycs, float *out) {
    mlp::colsum(st, w.part, rows, cols, X, ld, Y, ldy, ycs, out);
}
using mlp::gemm;

// Critic.forward on prepared x0; keeps the activations needed by the backward
static void critic_forward(cudaStream_t st, const Ws &w, int B, const float *P, float *q_out) {
    gemm(st, B, 256, 15, w.x0, 16, 1, P + CO::W0, 1, 15, w.xhat1, 256, P + CO::B0, ACT_NONE);
    ln_act_fwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, w.xhat1, P + CO::G1, P + CO::BE1, ACT_LEAKY, w.h1, w.rstd1);
    gemm(st, B, 256, 256, w.h1, 256, 1, P + CO::W4, 1, 256, w.xhat2, 256, P + CO::B4, ACT_NONE);
    ln_act_fwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, w.xhat2, P + CO::G2, P + CO::BE2, ACT_RELU, w.h2, w.rstd2);
    gemm(st, B, 128, 256, w.h2, 256, 1, P + CO::W8, 1, 256, w.h3, 128, P + CO::B8, ACT_RELU);
    critic_head_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, w.h3, P + CO::W10, P + CO::B10, q_out);
}
// backward of critic_forward from dq; param grads into G (nullptr: input gradient only); dx0 -> w.g16
static void critic_backward(cudaStream_t st, const Ws &w, int B, const float *P, const float *dq, float *G) {
    if (G) {   // dw10[c] = sum_r dq[r] h3[r][c], db10 = sum_r dq[r]
        colsum(st, w, B, 128, w.h3, 128, dq, 1, 0, G + CO::W10);
        colsum(st, w, B, 1, dq, 1, nullptr, 0, 0, G + CO::B10);
    }
    critic_head_bwd_kernel<<<(B * 128 + 255) / 256, 256, 0, st>>>(B, dq, P + CO::W10, w.h3, w.g128);         // dz3
    if (G) { gemm_wgrad(st, w, B, 128, 256, w.g128, 128, w.h2, 256, G + CO::W8); colsum(st, w, B, 128, w.g128, 128, nullptr, 0, 0, G + CO::B8); }
    gemm(st, B, 256, 128, w.g128, 128, 1, P + CO::W8, 256, 1, w.g256a, 256, nullptr, ACT_NONE);               // dh2 = dz3 W8
    ln_act_bwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, w.g256a, w.xhat2, w.rstd2, P + CO::G2, P + CO::BE2, ACT_RELU, w.g256b);
    if (G) {
        colsum(st, w, B, 256, w.g256a, 256, w.xhat2, 256, 1, G + CO::G2); colsum(st, w, B, 256, w.g256a, 256, nullptr, 0, 0, G + CO::BE2);
        gemm_wgrad(st, w, B, 256, 256, w.g256b, 256, w.h1, 256, G + CO::W4); colsum(st, w, B, 256, w.g256b, 256, nullptr, 0, 0, G + CO::B4);
    }
    gemm(st, B, 256, 256, w.g256b, 256, 1, P + CO::W4, 256, 1, w.g256a, 256, nullptr, ACT_NONE);              // dh1 = dz2 W4
    ln_act_bwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, w.g256a, w.xhat1, w.rstd1, P + CO::G1, P + CO::BE1, ACT_LEAKY, w.g256b);
    if (G) {
        colsum(st, w, B, 256, w.g256a, 256, w.xhat1, 256, 1, G + CO::G1); colsum(st, w, B, 256, w.g256a, 256, nullptr, 0, 0, G + CO::BE1);
        // dW0[256][15] = dz1^T x0[:, :15]
        gemm_wgrad(st, w, B, 256, 15, w.g256b, 256, w.x0, 16, G + CO::W0); colsum(st, w, B, 256, w.g256b, 256, nullptr, 0, 0, G + CO::B0);
    }
    gemm(st, B, 15, 256, w.g256b, 256, 1, P + CO::W0, 15, 1, w.g16, 16, nullptr, ACT_NONE);                   // dx0[:, :15]
}
// Actor.forward on prepared xa -> physical actions `out`; keep_act: store activations for the backward
static void actor_forward_train(cudaStream_t st, const Ws &w, int B, const float *P, const float *scale, const float *offset,
                                float *out, bool keep) {
    gemm(st, B, 128, 5, w.xa, 8, 1, P + AO::W1, 1, 5, w.a1, 128, P + AO::B1, ACT_LEAKY);
    gemm(st, B, 128, 128, w.a1, 128, 1, P + AO::W2, 1, 128, w.a2, 128, P + AO::B2, ACT_RELU);
    gemm(st, B, 64, 128, w.a2, 128, 1, P + AO::W3, 1, 128, w.a3, 64, P + AO::B3, ACT_RELU);
    gemm(st, B, 10, 64, w.a3, 64, 1, P + AO::W4, 1, 64, w.z4, 10, P + AO::B4, ACT_NONE);
    actor_head_fwd_kernel<<<(B * 10 + 255) / 256, 256, 0, st>>>(B, w.z4, w.xa, P + AO::EW, scale, offset, out, keep ? w.aux : nullptr);
}

}  // namespace erlfill

using namespace erlfill;

extern "C" {

size_t erlfill_ddpg_workspace_bytes(int batch) {
    if (batch < 2) return 0;
    return carve(nullptr, nullptr, batch) * sizeof(float);
}

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

    if (phases & ERLFILL_PHASE_CRITIC_GRAD) {
        emotion_stats_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.stats, w.cgrad + CO::TOTAL);
        // target: a' = actor_target(s', e'), Q' = critic_target(s', a', 0)
        build_xa_kernel<<<(B * 8 + 255) / 256, 256, 0, st>>>(B, 1, d_batch, d_next_emotion, w.xa);
        actor_forward_train(st, w, B, n->actor_target, n->scale, n->offset, w.atgt, false);
        build_x0_kernel<<<(B * 16 + 255) / 256, 256, 0, st>>>(B, 1, d_batch, w.atgt, w.stats, w.x0);
        critic_forward(st, w, B, n->critic_target, w.tq);
        // Q(s, a_norm, e) and its gradient
        build_x0_kernel<<<(B * 16 + 255) / 256, 256, 0, st>>>(B, 0, d_batch, nullptr, w.stats, w.x0);
        critic_forward(st, w, B, n->critic, w.q);
        critic_loss_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.q, w.tq, h->gamma, w.dq, w.scal + 4);
        critic_backward(st, w, B, n->critic, w.dq, w.cgrad);
    }
    if (phases & ERLFILL_PHASE_CRITIC_APPLY) {
        lr_kernel<<<1, 1, 0, st>>>(w.cgrad + CO::TOTAL, w.step, h->critic_lr, h->actor_lr, w.scal + 2);
        sumsq_kernel<<<kSumsqBlocks, 256, 0, st>>>(CO::TOTAL, w.cgrad, w.ssq);
        adam_soft_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(CO::TOTAL, n->critic, nullptr, n->critic_m, n->critic_v, w.cgrad,
                                                                  w.ssq, w.scal + 2, h->max_grad_norm, w.step, h->tau);
    }
    if (phases & ERLFILL_PHASE_ACTOR_GRAD) {
        build_xa_kernel<<<(B * 8 + 255) / 256, 256, 0, st>>>(B, 0, d_batch, d_next_emotion, w.xa);
        actor_forward_train(st, w, B, n->actor, n->scale, n->offset, w.aout, true);
        build_x0_kernel<<<(B * 16 + 255) / 256, 256, 0, st>>>(B, 2, d_batch, w.aout, w.stats, w.x0);
        critic_forward(st, w, B, n->critic, w.q2);
        actor_loss_kernel<<<1, 1024, 0, st>>>(B, w.q2, w.aout, n->actor, w.scal + 5, w.scal + 6);
        // d(-mean q2)/dq2 = -1/B (the +-1e6 clamp is inactive): input gradient of the critic only -> w.g16
        fill_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, -1.0f / (float)B, w.dq);
        critic_backward(st, w, B, n->critic, w.dq, nullptr);
        slice_action_grad_kernel<<<(B * 10 + 255) / 256, 256, 0, st>>>(B, w.g16, w.drawe);
        actor_head_bwd_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, w.drawe, w.aout, w.aux, w.xa, n->scale, w.z4, w.gz4, w.drawe);
        float *G = w.agrad;
        ew_grad_part_kernel<<<kColSplits, 256, 0, st>>>(B, w.xa, w.drawe, w.part);
        splitk_reduce_kernel<<<1, 256, 0, st>>>(30, kColSplits, w.part, 30, G + AO::EW);
        gemm_wgrad(st, w, B, 10, 64, w.gz4, 10, w.a3, 64, G + AO::W4); colsum(st, w, B, 10, w.gz4, 10, nullptr, 0, 0, G + AO::B4);
        gemm(st, B, 64, 10, w.gz4, 10, 1, n->actor + AO::W4, 64, 1, w.ga64, 64, nullptr, ACT_NONE);
        relu_bwd_kernel<<<(unsigned)(((size_t)B * 64 + 255) / 256), 256, 0, st>>>((size_t)B * 64, w.ga64, w.a3);
        gemm_wgrad(st, w, B, 64, 128, w.ga64, 64, w.a2, 128, G + AO::W3); colsum(st, w, B, 64, w.ga64, 64, nullptr, 0, 0, G + AO::B3);
        gemm(st, B, 128, 64, w.ga64, 64, 1, n->actor + AO::W3, 128, 1, w.ga128a, 128, nullptr, ACT_NONE);
        relu_bwd_kernel<<<(unsigned)(((size_t)B * 128 + 255) / 256), 256, 0, st>>>((size_t)B * 128, w.ga128a, w.a2);
        gemm_wgrad(st, w, B, 128, 128, w.ga128a, 128, w.a1, 128, G + AO::W2); colsum(st, w, B, 128, w.ga128a, 128, nullptr, 0, 0, G + AO::B2);
        gemm(st, B, 128, 128, w.ga128a, 128, 1, n->actor + AO::W2, 128, 1, w.ga128b, 128, nullptr, ACT_NONE);
        leaky_bwd_kernel<<<(unsigned)(((size_t)B * 128 + 255) / 256), 256, 0, st>>>((size_t)B * 128, w.ga128b, w.a1);
        gemm_wgrad(st, w, B, 128, 5, w.ga128b, 128, w.xa, 8, G + AO::W1); colsum(st, w, B, 128, w.ga128b, 128, nullptr, 0, 0, G + AO::B1);
        actor_l2_grad_kernel<<<(AO::TOTAL + 255) / 256, 256, 0, st>>>(n->actor, w.scal + 6, G);
    }
    if (phases & ERLFILL_PHASE_ACTOR_APPLY) {
        sumsq_kernel<<<kSumsqBlocks, 256, 0, st>>>(AO::TOTAL, w.agrad, w.ssq + kSumsqBlocks);
        adam_soft_kernel<<<(AO::TOTAL + 255) / 256, 256, 0, st>>>(AO::TOTAL, n->actor, n->actor_target, n->actor_m, n->actor_v, w.agrad,
                                                                  w.ssq + kSumsqBlocks, w.scal + 3, h->max_grad_norm, w.step, h->tau);
        // both soft updates happen after both optimiser steps (ddpg_agent.py:501-507)
        soft_update_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(CO::TOTAL, n->critic, n->critic_target, h->tau);
        if (d_report) report_kernel<<<1, 1, 0, st>>>(w.scal, d_report);   // [critic_loss, actor_loss, critic_lr, actor_lr]
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
