// esac.cu — EmotionSAC: the emotion-aware SAC agent of DifferentModules/ERL_Fill_agent.py.
//
//   EmotionGaussianPolicy (:63-185)  state encoder 2-256 (Linear, LayerNorm, ReLU) | emotion encoder 3-32 (same) -> 288-256-256 ->
//                                    mean / log_std heads (10 each) + tanh(Linear(3,10)) emotion adapters on both; clamp(log_std, -20, 2)
//   EmotionAwareQNetwork  (:188-239)  state 2-256 | action 10-256 | emotion 3-32 encoders -> 544-256-256-1
//   EmotionSACAgent.act   (:322-348)  a = tanh(mean + std eps) (or tanh(mean)), denormalised
//   EmotionSACAgent.update (:365-448) twin-Q targets with the entropy term, two MSE critics (separate Adam, same hyper-parameters),
//                                    policy loss mean(alpha log_pi - Q1), the emotion regulariser ||e_now - e_prev||_2 (a constant of
//                                    the parameters: reported only), Adam 3e-4, soft target update tau 0.005
//
// Flat fp32 buffers in parameters() order (include/erlfill.h).  Layer-by-layer on the shared tensor-core GEMM of mlp.cuh (3xTF32);
// the encoders write straight into their column block of the concatenated feature matrix, so torch.cat costs nothing.
#include <math.h>

#include "common.cuh"
#include "mlp.cuh"

namespace erlfill {
namespace esac {
using namespace mlp;

constexpr int H = 256, EH = 32, AD = ERLFILL_ACTION_DIM, SD = ERLFILL_STATE_DIM, ED = ERLFILL_EMOTION_DIM;
constexpr int PX = H + EH, QX = 2 * H + EH;            // concatenated feature widths: 288, 544
namespace PO {   // policy offsets
constexpr int SE_W = 0, SE_B = SE_W + H * SD, SE_G = SE_B + H, SE_BE = SE_G + H, EE_W = SE_BE + H, EE_B = EE_W + EH * ED, EE_G = EE_B + EH,
              EE_BE = EE_G + EH, L0_W = EE_BE + EH, L0_B = L0_W + H * PX, L2_W = L0_B + H, L2_B = L2_W + H * H, M_W = L2_B + H,
              M_B = M_W + AD * H, LS_W = M_B + AD, LS_B = LS_W + AD * H, AM_W = LS_B + AD, AM_B = AM_W + AD * ED, AS_W = AM_B + AD,
              AS_B = AS_W + AD * ED, TOTAL = AS_B + AD;
}
namespace QO {   // one Q network
constexpr int SE_W = 0, SE_B = SE_W + H * SD, SE_G = SE_B + H, SE_BE = SE_G + H, AE_W = SE_BE + H, AE_B = AE_W + H * AD, AE_G = AE_B + H,
              AE_BE = AE_G + H, EE_W = AE_BE + H, EE_B = EE_W + EH * ED, EE_G = EE_B + EH, EE_BE = EE_G + EH, L0_W = EE_BE + EH,
              L0_B = L0_W + H * QX, L2_W = L0_B + H, L2_B = L2_W + H * H, L4_W = L2_B + H, L4_B = L4_W + H, TOTAL = L4_B + 1;
}
static_assert(PO::TOTAL == ERLFILL_ESAC_POLICY_PARAMS && QO::TOTAL == ERLFILL_ESAC_Q_PARAMS, "EmotionSAC layouts");

// ---- LayerNorm + ReLU on a column block of a row-major matrix, one warp per row (W = 256 or 32) ---------------------------
template <int W>
__global__ void __launch_bounds__(256) ln_relu_fwd_kernel(int B, float *x, int ld, const float *gamma, const float *beta, float *xhat, float *rstd_out) {
    constexpr int PER = W / 32;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (row >= B) return;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = x[(size_t)row * ld + l + 32 * i]; s += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / W);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / W) + 1e-5f);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = l + 32 * i;
        const float xh = (v[i] - mean) * rstd;
        x[(size_t)row * ld + c] = fmaxf(xh * gamma[c] + beta[c], 0.f);
        if (xhat) xhat[(size_t)row * W + c] = xh;
    }
    if (l == 0 && rstd_out) rstd_out[row] = rstd;
}
// dh (column block of a [B][ldd] matrix) -> dy[B][W] (for the gamma / beta sums) and dz[B][W]
template <int W>
__global__ void __launch_bounds__(256) ln_relu_bwd_kernel(int B, const float *dh, int ldd, const float *h, int ldh, const float *xhat,
                                                          const float *rstd, const float *gamma, float *dy, float *dz) {
    constexpr int PER = W / 32;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (row >= B) return;
    float dxh[PER], xh[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = l + 32 * i;
        xh[i] = xhat[(size_t)row * W + c];
        const float g = h[(size_t)row * ldh + c] > 0.f ? dh[(size_t)row * ldd + c] : 0.f;
        dy[(size_t)row * W + c] = g;
        dxh[i] = g * gamma[c];
        s1 += dxh[i];
        s2 = fmaf(dxh[i], xh[i], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float r = rstd[row], m1 = s1 * (1.0f / W), m2 = s2 * (1.0f / W);
#pragma unroll
    for (int i = 0; i < PER; ++i) dz[(size_t)row * W + l + 32 * i] = r * (dxh[i] - m1 - xh[i] * m2);
}
template <int W>
static void ln_fwd(cudaStream_t st, int B, float *x, int ld, const float *g, const float *b, float *xhat, float *rstd) {
    ln_relu_fwd_kernel<W><<<(B + 7) / 8, 256, 0, st>>>(B, x, ld, g, b, xhat, rstd);
}

// standard normal draw j of row `row`: injected, or Philox Box-Muller keyed (row id, idx, j/2 + 8 salt, STREAM_OU)
__device__ __forceinline__ float normal_draw(const float *inj, size_t row, int j, uint64_t seed, uint32_t row_id, uint32_t idx, uint32_t salt) {
    if (inj) return inj[row * AD + j];
    const uint4 r = philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), row_id, idx, (uint32_t)(j >> 1) + 8u * salt, STREAM_OU);
    const uint32_t a = (j & 1) ? r.z : r.x, b = (j & 1) ? r.w : r.y;
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

// EmotionGaussianPolicy heads -> sample (:135-167).  ml[r] = [base mean(10) | base log_std(10)]; e row stride lde (0 = broadcast).
// a = tanh(mean + std eps) (deterministic: tanh(mean)); log_prob = sum(-eps^2/2 - ls - log sqrt(2 pi) - log(1 - a^2 + 1e-6)).
// a16[r] = [a | 0..] (the action-encoder input); sv[r][c][6] = a, eps, std, adapter_mean, adapter_std, log_std inside the clamp
// (kept for the backward when sv != nullptr).  act mode (d_action != nullptr): physical action + normalised action out.
__global__ void sample_kernel(int B, const float *P, const float *e, int lde, const float *ml, const float *eps, uint64_t seed, uint32_t row_id_base,
                              uint32_t idx, uint32_t salt, int deterministic, float *a16, float *logp, float *sv, const float *lo, const float *hi,
                              float *d_action, float *d_norm_action, int row0) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float e0 = e[(size_t)r * lde], e1 = e[(size_t)r * lde + 1], e2 = e[(size_t)r * lde + 2];
    float lp = 0.f;
#pragma unroll
    for (int j = 0; j < AD; ++j) {
        float am = P[PO::AM_B + j], as = P[PO::AS_B + j];
        am = fmaf(e0, P[PO::AM_W + j * ED], am); am = fmaf(e1, P[PO::AM_W + j * ED + 1], am); am = fmaf(e2, P[PO::AM_W + j * ED + 2], am);
        as = fmaf(e0, P[PO::AS_W + j * ED], as); as = fmaf(e1, P[PO::AS_W + j * ED + 1], as); as = fmaf(e2, P[PO::AS_W + j * ED + 2], as);
        am = tanhf(am); as = tanhf(as);
        const float mean = ml[(size_t)r * 20 + j] + am, lsr = ml[(size_t)r * 20 + AD + j] + as;
        const float ls = fminf(fmaxf(lsr, -20.0f), 2.0f);
        const float ep = deterministic ? 0.f : normal_draw(eps, (size_t)row0 + r, j, seed, row_id_base + (uint32_t)(row0 + r), idx, salt);
        const float sd = expf(ls);
        const float a = tanhf(mean + sd * ep);
        lp += -0.5f * ep * ep - ls - 0.9189385332046727f - logf(1.0f - a * a + 1e-6f);
        if (a16) a16[(size_t)r * 16 + j] = a;
        if (sv) {
            float *s = sv + ((size_t)r * AD + j) * 6;
            s[0] = a; s[1] = ep; s[2] = sd; s[3] = am; s[4] = as; s[5] = (lsr >= -20.0f && lsr <= 2.0f) ? 1.f : 0.f;
        }
        if (d_action) {
            const size_t o = (size_t)(row0 + r) * AD + j;
            d_action[o] = a * ((hi[j] - lo[j]) * 0.5f) + (hi[j] + lo[j]) * 0.5f;      // denormalize_action (:310-320)
            d_norm_action[o] = a;
        }
    }
    if (a16) {
#pragma unroll
        for (int j = AD; j < 16; ++j) a16[(size_t)r * 16 + j] = 0.f;
    }
    if (logp) logp[r] = lp;
}
// target y = r + (1-d) gamma (min(q1', q2') - alpha logp'); per-network MSE and dq (:385-397)
__global__ void __launch_bounds__(1024) q_loss_kernel(int B, const float *batch, const float *q1, const float *q2, const float *q1t, const float *q2t,
                                                      const float *logp_next, float gamma, float alpha, float *dq1, float *dq2, float *report) {
    __shared__ float scratch[32];
    float l1 = 0.f, l2 = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) {
        const float rew = batch[(size_t)r * ERLFILL_REC + 12], d = batch[(size_t)r * ERLFILL_REC + 15];
        const float y = rew + (1.0f - d) * gamma * (fminf(q1t[r], q2t[r]) - alpha * logp_next[r]);
        const float e1 = q1[r] - y, e2 = q2[r] - y;
        dq1[r] = 2.0f * e1 / (float)B; dq2[r] = 2.0f * e2 / (float)B;
        l1 = fmaf(e1, e1, l1); l2 = fmaf(e2, e2, l2);
    }
    const float s1 = block_sum(l1, scratch), s2 = block_sum(l2, scratch);
    if (threadIdx.x == 0) { report[0] = s1 / (float)B; report[1] = s2 / (float)B; }
}
// policy loss mean(alpha logp - q1) and the emotion regulariser ||e_now - e_prev||_2 (:409-430)
__global__ void __launch_bounds__(1024) policy_loss_kernel(int B, const float *logp, const float *q1, float alpha, const float *e_now, const float *e_prev,
                                                           float *report) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) s += alpha * logp[r] - q1[r];
    const float t = block_sum(s, scratch);
    if (threadIdx.x == 0) {
        report[2] = t / (float)B;
        float n2 = 0.f;
        for (int c = 0; c < ED; ++c) { const float d = e_now[c] - e_prev[c]; n2 = fmaf(d, d, n2); }
        report[3] = sqrtf(n2);
    }
}
// gradient of mean(alpha logp - q1(s, a)) wrt the head outputs and the adapter pre-activations:
//   dz = dL/da (1 - a^2) + alpha/B * 2a(1 - a^2)/(1 - a^2 + 1e-6);  dmean = dz;  dlog_std = dz std eps - alpha/B inside the clamp
__global__ void policy_grad_kernel(int B, const float *da16, const float *sv, float alpha, float *dml, float *dam, float *das) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * AD) return;
    const int r = g / AD, j = g % AD;
    const float *s = sv + (size_t)g * 6;
    const float a = s[0], om = 1.0f - a * a;
    const float dz = da16[(size_t)r * 16 + j] * om + alpha / (float)B * 2.0f * a * om / (om + 1e-6f);
    const float dls = s[5] != 0.f ? dz * s[2] * s[1] - alpha / (float)B : 0.f;
    dml[(size_t)r * 20 + j] = dz;
    dml[(size_t)r * 20 + AD + j] = dls;
    dam[g] = dz * (1.0f - s[3] * s[3]);
    das[g] = dls * (1.0f - s[4] * s[4]);
}
__global__ void add_relu_bwd_kernel(size_t n, float *ga, const float *gb, const float *h) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ga[i] = h[i] > 0.f ? ga[i] + gb[i] : 0.f;
}
__global__ void adam_kernel(int n, float *p, float *m, float *v, const float *g, float lr, float bc1, float bc2_sqrt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * 0.1f;
    const float vi = v[i] * 0.999f + 0.001f * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + 1e-8f));
}
__global__ void soft_kernel(int n, const float *p, float *pt, float tau) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pt[i] = tau * p[i] + (1.0f - tau) * pt[i];
}
static void adam(cudaStream_t st, int n, float *p, float *m, float *v, const float *g, double lr, int step) {
    const double bc1 = 1.0 - pow(0.9, (double)step), bc2 = sqrt(1.0 - pow(0.999, (double)step));
    adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, p, m, v, g, (float)lr, (float)bc1, (float)bc2);
}
static void relu_bwd(cudaStream_t st, size_t n, float *g, const float *h) {
    relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, g, h);
}

// ---- workspace ---------------------------------------------------------------------------------------------------
struct QBuf { float *X, *xs, *xa, *xe, *rs, *ra, *re, *h1, *h2; };     // concatenated features + what the backward needs
struct PBuf { float *X, *xs, *xe, *rs, *re, *h1, *h2, *ml; };
struct Ws {
    float *qgrad, *pgrad, *scal;
    QBuf qa, qb;
    PBuf pb;
    float *q1, *q2, *q1t, *q2t, *dq1, *dq2, *logp, *logp2, *a16, *sv, *da16, *dml, *dam, *das;
    float *ga, *gb, *gX, *dy, *dz, *part;
};
constexpr size_t kPart = (size_t)16 * H * QX + 1024;       // 16 batch slices of the largest weight gradient (256 x 544)
static size_t align4(size_t n) { return (n + 3) & ~(size_t)3; }
static size_t carve(Ws *w, float *base, int B) {
    size_t o = 0;
    float *dummy;
    auto take = [&](float **p, size_t n) { if (w) *p = base + o; o += align4(n); };
    const size_t b = (size_t)B;
    take(w ? &w->qgrad : &dummy, 2 * (size_t)QO::TOTAL); take(w ? &w->pgrad : &dummy, PO::TOTAL); take(w ? &w->scal : &dummy, 8);
    for (QBuf *q : {w ? &w->qa : nullptr, w ? &w->qb : nullptr}) {
        take(q ? &q->X : &dummy, b * QX); take(q ? &q->xs : &dummy, b * H); take(q ? &q->xa : &dummy, b * H); take(q ? &q->xe : &dummy, b * EH);
        take(q ? &q->rs : &dummy, b); take(q ? &q->ra : &dummy, b); take(q ? &q->re : &dummy, b); take(q ? &q->h1 : &dummy, b * H);
        take(q ? &q->h2 : &dummy, b * H);
    }
    PBuf *p = w ? &w->pb : nullptr;
    take(p ? &p->X : &dummy, b * PX); take(p ? &p->xs : &dummy, b * H); take(p ? &p->xe : &dummy, b * EH); take(p ? &p->rs : &dummy, b);
    take(p ? &p->re : &dummy, b); take(p ? &p->h1 : &dummy, b * H); take(p ? &p->h2 : &dummy, b * H); take(p ? &p->ml : &dummy, b * 20);
    take(w ? &w->q1 : &dummy, b); take(w ? &w->q2 : &dummy, b); take(w ? &w->q1t : &dummy, b); take(w ? &w->q2t : &dummy, b);
    take(w ? &w->dq1 : &dummy, b); take(w ? &w->dq2 : &dummy, b); take(w ? &w->logp : &dummy, b); take(w ? &w->logp2 : &dummy, b);
    take(w ? &w->a16 : &dummy, b * 16); take(w ? &w->sv : &dummy, b * AD * 6); take(w ? &w->da16 : &dummy, b * 16);
    take(w ? &w->dml : &dummy, b * 20); take(w ? &w->dam : &dummy, b * AD); take(w ? &w->das : &dummy, b * AD);
    take(w ? &w->ga : &dummy, b * H); take(w ? &w->gb : &dummy, b * H); take(w ? &w->gX : &dummy, b * QX); take(w ? &w->dy : &dummy, b * H);
    take(w ? &w->dz : &dummy, b * H); take(w ? &w->part : &dummy, kPart);
    return o;
}

// ---- network passes ----------------------------------------------------------------------------------------------
// encoder: Linear(in, W) into column block c0 of X (row stride ldx), LayerNorm + ReLU in place
template <int W>
static void encoder(cudaStream_t st, int B, const float *in, int ldin, int n_in, const float *Wt, const float *bias, const float *g, const float *be,
                    float *X, int ldx, int c0, float *xhat, float *rstd) {
    gemm(st, B, W, n_in, in, ldin, 1, Wt, 1, n_in, X + c0, ldx, bias, ACT_NONE);
    ln_fwd<W>(st, B, X + c0, ldx, g, be, xhat, rstd);
}
static void q_forward(cudaStream_t st, int B, const float *P, const float *s, int lds, const float *a, int lda, const float *e, int lde,
                      const QBuf &b, float *q, bool keep) {
    encoder<H>(st, B, s, lds, SD, P + QO::SE_W, P + QO::SE_B, P + QO::SE_G, P + QO::SE_BE, b.X, QX, 0, keep ? b.xs : nullptr, keep ? b.rs : nullptr);
    encoder<H>(st, B, a, lda, AD, P + QO::AE_W, P + QO::AE_B, P + QO::AE_G, P + QO::AE_BE, b.X, QX, H, keep ? b.xa : nullptr, keep ? b.ra : nullptr);
    encoder<EH>(st, B, e, lde, ED, P + QO::EE_W, P + QO::EE_B, P + QO::EE_G, P + QO::EE_BE, b.X, QX, 2 * H, keep ? b.xe : nullptr, keep ? b.re : nullptr);
    gemm(st, B, H, QX, b.X, QX, 1, P + QO::L0_W, 1, QX, b.h1, H, P + QO::L0_B, ACT_RELU);
    gemm(st, B, H, H, b.h1, H, 1, P + QO::L2_W, 1, H, b.h2, H, P + QO::L2_B, ACT_RELU);
    gemm(st, B, 1, H, b.h2, H, 1, P + QO::L4_W, 1, H, q, 1, P + QO::L4_B, ACT_NONE);
}
// backward of one encoder: dh = column block c0 of gX -> parameter gradients (G != nullptr) and, for the action encoder, d(input)
template <int W>
static void encoder_backward(cudaStream_t st, const Ws &w, int B, const float *gX, int ldg, const float *X, int ldx, int c0, const float *xhat,
                             const float *rstd, const float *Wt, const float *gamma, const float *in, int ldin, int n_in, float *G, int oW, int oB,
                             int oG, int oBE, float *d_in, int ld_din) {
    ln_relu_bwd_kernel<W><<<(B + 7) / 8, 256, 0, st>>>(B, gX + c0, ldg, X + c0, ldx, xhat, rstd, gamma, w.dy, w.dz);
    if (G) {
        colsum(st, w.part, B, W, w.dy, W, xhat, W, 1, G + oG);
        colsum(st, w.part, B, W, w.dy, W, nullptr, 0, 0, G + oBE);
        gemm_wgrad(st, w.part, B, W, n_in, w.dz, W, in, ldin, G + oW);
        colsum(st, w.part, B, W, w.dz, W, nullptr, 0, 0, G + oB);
    }
    if (d_in) gemm(st, B, n_in, W, w.dz, W, 1, Wt, n_in, 1, d_in, ld_din, nullptr, ACT_NONE);
}
// backward of q_forward from dq[B]: parameter gradients into G (nullptr: none), d(action) into da16 (nullptr: none)
static void q_backward(cudaStream_t st, const Ws &w, int B, const float *P, float *G, const float *s, int lds, const float *a, int lda, const float *e,
                       int lde, const QBuf &b, const float *dq, float *da16) {
    if (G) { gemm_wgrad(st, w.part, B, 1, H, dq, 1, b.h2, H, G + QO::L4_W); colsum(st, w.part, B, 1, dq, 1, nullptr, 0, 0, G + QO::L4_B); }
    gemm(st, B, H, 1, dq, 1, 1, P + QO::L4_W, H, 1, w.ga, H, nullptr, ACT_NONE);
    relu_bwd(st, (size_t)B * H, w.ga, b.h2);
    if (G) { gemm_wgrad(st, w.part, B, H, H, w.ga, H, b.h1, H, G + QO::L2_W); colsum(st, w.part, B, H, w.ga, H, nullptr, 0, 0, G + QO::L2_B); }
    gemm(st, B, H, H, w.ga, H, 1, P + QO::L2_W, H, 1, w.gb, H, nullptr, ACT_NONE);
    relu_bwd(st, (size_t)B * H, w.gb, b.h1);
    if (G) { gemm_wgrad(st, w.part, B, H, QX, w.gb, H, b.X, QX, G + QO::L0_W); colsum(st, w.part, B, H, w.gb, H, nullptr, 0, 0, G + QO::L0_B); }
    gemm(st, B, QX, H, w.gb, H, 1, P + QO::L0_W, QX, 1, w.gX, QX, nullptr, ACT_NONE);
    if (G) {
        encoder_backward<H>(st, w, B, w.gX, QX, b.X, QX, 0, b.xs, b.rs, P + QO::SE_W, P + QO::SE_G, s, lds, SD, G, QO::SE_W, QO::SE_B, QO::SE_G,
                            QO::SE_BE, nullptr, 0);
        encoder_backward<EH>(st, w, B, w.gX, QX, b.X, QX, 2 * H, b.xe, b.re, P + QO::EE_W, P + QO::EE_G, e, lde, ED, G, QO::EE_W, QO::EE_B,
                             QO::EE_G, QO::EE_BE, nullptr, 0);
    }
    if (G || da16)
        encoder_backward<H>(st, w, B, w.gX, QX, b.X, QX, H, b.xa, b.ra, P + QO::AE_W, P + QO::AE_G, a, lda, AD, G, QO::AE_W, QO::AE_B, QO::AE_G,
                            QO::AE_BE, da16, 16);
}
static void policy_forward(cudaStream_t st, int B, const float *P, const float *s, int lds, const float *e, int lde, const PBuf &b, bool keep) {
    encoder<H>(st, B, s, lds, SD, P + PO::SE_W, P + PO::SE_B, P + PO::SE_G, P + PO::SE_BE, b.X, PX, 0, keep ? b.xs : nullptr, keep ? b.rs : nullptr);
    encoder<EH>(st, B, e, lde, ED, P + PO::EE_W, P + PO::EE_B, P + PO::EE_G, P + PO::EE_BE, b.X, PX, H, keep ? b.xe : nullptr, keep ? b.re : nullptr);
    gemm(st, B, H, PX, b.X, PX, 1, P + PO::L0_W, 1, PX, b.h1, H, P + PO::L0_B, ACT_RELU);
    gemm(st, B, H, H, b.h1, H, 1, P + PO::L2_W, 1, H, b.h2, H, P + PO::L2_B, ACT_RELU);
    gemm(st, B, AD, H, b.h2, H, 1, P + PO::M_W, 1, H, b.ml, 20, P + PO::M_B, ACT_NONE);
    gemm(st, B, AD, H, b.h2, H, 1, P + PO::LS_W, 1, H, b.ml + AD, 20, P + PO::LS_B, ACT_NONE);
}

constexpr int kActChunk = 65536;     // rows per pass of the policy MLP (as in td3_sac.cu: whole waves of 128 x 128 GEMM tiles)

}  // namespace esac
}  // namespace erlfill

using namespace erlfill;
using namespace erlfill::esac;

extern "C" {

size_t erlfill_esac_workspace_bytes(int batch) { return batch < 1 ? 0 : carve(nullptr, nullptr, batch) * sizeof(float); }
size_t erlfill_esac_act_workspace_bytes(void) { return (size_t)kActChunk * (PX + 2 * H + 20) * sizeof(float); }

int erlfill_esac_grad_buckets(void *d_workspace, int batch, float **q_grad, float **policy_grad) {
    if (!d_workspace || batch < 1) return ERLFILL_ERR_ARG;
    Ws w;
    carve(&w, (float *)d_workspace, batch);
    if (q_grad) *q_grad = w.qgrad;
    if (policy_grad) *policy_grad = w.pgrad;
    return ERLFILL_OK;
}

int erlfill_esac_act(const float *d_policy, int n, const float *d_state, const float *d_emotion, int emotion_broadcast, const float *d_eps,
                     int deterministic, const float *d_lo, const float *d_hi, uint64_t seed, uint32_t row_id_base, uint32_t step_idx,
                     float *d_action, float *d_norm_action, void *d_workspace, void *stream) {
    if (!d_policy || n < 1 || !d_state || !d_emotion || !d_lo || !d_hi || !d_action || !d_norm_action || !d_workspace) return ERLFILL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    float *base = (float *)d_workspace;
    PBuf b{};
    b.X = base; b.h1 = b.X + (size_t)kActChunk * PX; b.h2 = b.h1 + (size_t)kActChunk * H; b.ml = b.h2 + (size_t)kActChunk * H;
    const int lde = emotion_broadcast ? 0 : ED;
    for (int r0 = 0; r0 < n; r0 += kActChunk) {
        const int m = n - r0 < kActChunk ? n - r0 : kActChunk;
        const float *e = d_emotion + (size_t)r0 * lde;
        policy_forward(st, m, d_policy, d_state + (size_t)r0 * SD, SD, e, lde, b, false);
        sample_kernel<<<(m + 127) / 128, 128, 0, st>>>(m, d_policy, e, lde, b.ml, d_eps, seed, row_id_base, step_idx, 0u, deterministic, nullptr, nullptr,
                                                      nullptr, d_lo, d_hi, d_action, d_norm_action, r0);
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_esac_update(const erlfill_esac_nets *n, const erlfill_esac_hyper *h, const float *d_batch, const float *d_eps_next,
                        const float *d_eps_new, const float *d_emotion_now, const float *d_emotion_prev, void *d_workspace, int phases,
                        float *d_report, void *stream) {
    if (!n || !h || !d_batch || !d_workspace || !d_emotion_now || !d_emotion_prev) return ERLFILL_ERR_ARG;
    if (!n->policy || !n->q || !n->q_target || !n->policy_m || !n->policy_v || !n->q_m || !n->q_v) return ERLFILL_ERR_ARG;
    const int B = h->batch;
    if (B < 1 || h->adam_step < 1) return ERLFILL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, (float *)d_workspace, B);
    const int NQ = QO::TOTAL;
    const float *s = d_batch, *s2 = d_batch + 13, *a = d_batch + 2, *e = d_batch + 16;     // columns of the replay rows
    const int ld = ERLFILL_REC;
    if (phases & ERLFILL_PHASE_CRITIC_GRAD) {
        // target: a' ~ pi(s', e), y from the twin target critics (the reference conditions s' on the STORED emotion e, :385-387)
        policy_forward(st, B, n->policy, s2, ld, e, ld, w.pb, false);
        sample_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, n->policy, e, ld, w.pb.ml, d_eps_next, h->seed, h->rank << 20, h->update_idx, 4u, 0, w.a16, w.logp2, nullptr,
                                                      nullptr, nullptr, nullptr, nullptr, 0);
        q_forward(st, B, n->q_target, s2, ld, w.a16, 16, e, ld, w.qa, w.q1t, false);
        q_forward(st, B, n->q_target + NQ, s2, ld, w.a16, 16, e, ld, w.qb, w.q2t, false);
        q_forward(st, B, n->q, s, ld, a, ld, e, ld, w.qa, w.q1, true);
        q_forward(st, B, n->q + NQ, s, ld, a, ld, e, ld, w.qb, w.q2, true);
        q_loss_kernel<<<1, 1024, 0, st>>>(B, d_batch, w.q1, w.q2, w.q1t, w.q2t, w.logp2, h->gamma, h->alpha, w.dq1, w.dq2, w.scal);
        q_backward(st, w, B, n->q, w.qgrad, s, ld, a, ld, e, ld, w.qa, w.dq1, nullptr);
        q_backward(st, w, B, n->q + NQ, w.qgrad + NQ, s, ld, a, ld, e, ld, w.qb, w.dq2, nullptr);
    }
    // q1 and q2 have separate Adam optimisers with the same hyper-parameters: one element-wise pass over [q1 | q2]
    if (phases & ERLFILL_PHASE_CRITIC_APPLY) adam(st, 2 * NQ, n->q, n->q_m, n->q_v, w.qgrad, h->lr, h->adam_step);
    if (phases & ERLFILL_PHASE_ACTOR_GRAD) {
        const float *P = n->policy;
        policy_forward(st, B, P, s, ld, e, ld, w.pb, true);
        sample_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, P, e, ld, w.pb.ml, d_eps_new, h->seed, h->rank << 20, h->update_idx, 5u, 0, w.a16, w.logp, w.sv, nullptr,
                                                      nullptr, nullptr, nullptr, 0);
        q_forward(st, B, n->q, s, ld, w.a16, 16, e, ld, w.qa, w.q1, true);
        policy_loss_kernel<<<1, 1024, 0, st>>>(B, w.logp, w.q1, h->alpha, d_emotion_now, d_emotion_prev, w.scal);
        fill_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, -1.0f / (float)B, w.dq1);
        q_backward(st, w, B, n->q, nullptr, s, ld, w.a16, 16, e, ld, w.qa, w.dq1, w.da16);
        policy_grad_kernel<<<(B * AD + 255) / 256, 256, 0, st>>>(B, w.da16, w.sv, h->alpha, w.dml, w.dam, w.das);
        float *G = w.pgrad;
        // emotion adapters: tanh(Linear(3, 10)) on the mean and on log_std
        gemm_wgrad(st, w.part, B, AD, ED, w.dam, AD, e, ld, G + PO::AM_W); colsum(st, w.part, B, AD, w.dam, AD, nullptr, 0, 0, G + PO::AM_B);
        gemm_wgrad(st, w.part, B, AD, ED, w.das, AD, e, ld, G + PO::AS_W); colsum(st, w.part, B, AD, w.das, AD, nullptr, 0, 0, G + PO::AS_B);
        // heads
        gemm_wgrad(st, w.part, B, AD, H, w.dml, 20, w.pb.h2, H, G + PO::M_W); colsum(st, w.part, B, AD, w.dml, 20, nullptr, 0, 0, G + PO::M_B);
        gemm_wgrad(st, w.part, B, AD, H, w.dml + AD, 20, w.pb.h2, H, G + PO::LS_W); colsum(st, w.part, B, AD, w.dml + AD, 20, nullptr, 0, 0, G + PO::LS_B);
        gemm(st, B, H, AD, w.dml, 20, 1, P + PO::M_W, H, 1, w.ga, H, nullptr, ACT_NONE);
        gemm(st, B, H, AD, w.dml + AD, 20, 1, P + PO::LS_W, H, 1, w.gb, H, nullptr, ACT_NONE);
        add_relu_bwd_kernel<<<(unsigned)(((size_t)B * H + 255) / 256), 256, 0, st>>>((size_t)B * H, w.ga, w.gb, w.pb.h2);
        gemm_wgrad(st, w.part, B, H, H, w.ga, H, w.pb.h1, H, G + PO::L2_W); colsum(st, w.part, B, H, w.ga, H, nullptr, 0, 0, G + PO::L2_B);
        gemm(st, B, H, H, w.ga, H, 1, P + PO::L2_W, H, 1, w.gb, H, nullptr, ACT_NONE);
        relu_bwd(st, (size_t)B * H, w.gb, w.pb.h1);
        gemm_wgrad(st, w.part, B, H, PX, w.gb, H, w.pb.X, PX, G + PO::L0_W); colsum(st, w.part, B, H, w.gb, H, nullptr, 0, 0, G + PO::L0_B);
        gemm(st, B, PX, H, w.gb, H, 1, P + PO::L0_W, PX, 1, w.gX, PX, nullptr, ACT_NONE);
        encoder_backward<H>(st, w, B, w.gX, PX, w.pb.X, PX, 0, w.pb.xs, w.pb.rs, P + PO::SE_W, P + PO::SE_G, s, ld, SD, G, PO::SE_W, PO::SE_B, PO::SE_G,
                            PO::SE_BE, nullptr, 0);
        encoder_backward<EH>(st, w, B, w.gX, PX, w.pb.X, PX, H, w.pb.xe, w.pb.re, P + PO::EE_W, P + PO::EE_G, e, ld, ED, G, PO::EE_W, PO::EE_B,
                             PO::EE_G, PO::EE_BE, nullptr, 0);
    }
    if (phases & ERLFILL_PHASE_ACTOR_APPLY) {
        adam(st, PO::TOTAL, n->policy, n->policy_m, n->policy_v, w.pgrad, h->lr, h->adam_step);
        soft_kernel<<<(2 * NQ + 255) / 256, 256, 0, st>>>(2 * NQ, n->q, n->q_target, h->tau);
        if (d_report) cudaMemcpyAsync(d_report, w.scal, 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);   // [q1_loss, q2_loss, policy_loss, emo_reg]
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
