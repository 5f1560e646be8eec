// emotion_tf.cu — M2: EmotionTransformer.forward (EmotionModule.py:48-71) batched over envs, exact fp32 path.
//
// embedding Linear(3,32) + learned positions -> 4 x post-norm TransformerEncoderLayer(d=32, 4 heads x 8,
// softmax(QK^T/sqrt(8)), FFN 32->2048->32 ReLU, LayerNorm eps 1e-5) -> last token -> Linear(32,3) -> 0.5*tanh+0.5.
// The sequence is read either from explicit [n][S][3] rows or from the per-env history ring exactly like
// get_state() (EmotionModule.py:83-102): front-padded with the oldest entry to 20, or the single neutral
// token [0.5,0.5,0] when the history is empty.
//
// This is the fp32 SIMT implementation: it is the N=1 drop-in's get_emotion() (1e-5 parity with the
// reference) and the checker-side twin of the bf16 tensor-core kernel (emotion_tf_tc.cu) that the batched
// rollout uses.  One CTA handles kG envs (kG*20 tokens) so every weight read from L2 is used kG*20 times.
// Small batches (the N = 1 drop-in: one launch is a 1.3 M-cycle serial walk of 64 FFN chunks on one SM) run as
// thread-block clusters of kSplit CTAs per env group: every CTA repeats the cheap parts (embedding, q|k|v, attention,
// LayerNorms) and owns FF / kSplit of the hidden units; the partial linear2 sums are exchanged through distributed
// shared memory and added in rank order on every CTA, so the result does not depend on the launch shape of the cluster.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"
#include "tf_dropout.cuh"

namespace erlfill {

constexpr int D = 32, NH = 4, HD = 8, FF = 2048, NL = 4, SMAX = 20;
constexpr int kG = 4;                    // envs per CTA
constexpr int kT = kG * SMAX;            // tokens per CTA
constexpr int kTfThreads = 128;
constexpr int kChunk = 128;              // hidden units per FFN pass

struct TfArgs {
    erlfill_transformer_weights w;
    erlfill_emotion_state emo;
    erlfill_transformer_masks masks;
    const float *d_seq;                  // [n][S][3] or nullptr (then read the history ring)
    float *d_out;                        // [n][3]
    int n, S, dropout_mode, has_masks;
    uint32_t call_idx, row_id_base;
    uint64_t seed;
};

// shared memory plan (floats)
constexpr int ldX = D + 1, ldQ = 3 * D + 1, ldH = kChunk + 4;
constexpr int sX = 0;                                  // [kT][33]  residual stream
constexpr int sQ = sX + kT * ldX;                      // [kT][97]  q|k|v, later attention output in the q slot
constexpr int sA = sQ + kT * ldQ;                      // [kT][33]  sub-layer output before the residual add
constexpr int sHd = sA + kT * ldX;                     // [kT][132] hidden chunk
constexpr int sW2 = sHd + kT * ldH;                    // [kChunk][32] linear2 chunk, transposed
constexpr int sLen = sW2 + kChunk * D;                 // [kG] sequence length per env (as int)
constexpr int kTfSmemFloats = sLen + 8;

__device__ __forceinline__ float keep_scale(const TfArgs &a, uint32_t gid, uint32_t site, uint32_t idx,
                                            const uint8_t *mask_ptr, size_t mask_idx) {
    // returns the multiplier of at::dropout: 0 or 1/(1-p), p = 0.3
    if (a.dropout_mode == ERLFILL_DROPOUT_OFF) return 1.0f;
    bool keep;
    if (a.dropout_mode == ERLFILL_DROPOUT_MASK) keep = mask_ptr[mask_idx] != 0;
    else {
        // the mask definition shared with the tensor-core kernel: tf_dropout.cuh
        const uint4 r = tfdrop::block8((uint32_t)a.seed, (uint32_t)(a.seed >> 32), gid, a.call_idx, site, idx >> 3);
        keep = tfdrop::keep_elem(r, (int)(idx & 7u));
    }
    return keep ? (1.0f / (1.0f - 0.3f)) : 0.0f;
}

__device__ __forceinline__ void layer_norm_rows(float *sm, int n_tok, const float *g, const float *b, int tid) {
    // x <- LN(x + a): one thread per token, two-pass (mean, biased variance), eps inside the sqrt
    for (int t = tid; t < n_tok; t += kTfThreads) {
        float v[D];
        float mean = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { v[d] = sm[sX + t * ldX + d] + sm[sA + t * ldX + d]; mean += v[d]; }
        mean *= (1.0f / D);
        float var = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { const float c = v[d] - mean; var = fmaf(c, c, var); }
        const float rstd = rsqrtf(var * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int d = 0; d < D; ++d) sm[sX + t * ldX + d] = (v[d] - mean) * rstd * __ldg(g + d) + __ldg(b + d);
    }
}

constexpr int kSplitSmall = 8;            // CTAs per cluster in the small-batch variant

template <int kSplit>
__global__ void __launch_bounds__(kTfThreads) emotion_forward_kernel(const TfArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int rank = kSplit > 1 ? (int)(blockIdx.x % kSplit) : 0;      // == cluster rank (1-D clusters of kSplit CTAs)
    const int env0 = (int)(blockIdx.x / kSplit) * kG;
    const int n_env = min(kG, a.n - env0);
    int *s_len = reinterpret_cast<int *>(&sm[sLen]);
    const int nT = n_env * SMAX;          // token rows of the envs this CTA really has (a ragged last CTA, or N = 1: 20 of 80)

    // ---- sequence lengths ----
    if (tid < kG) {
        int S = 0;
        if (tid < n_env) {
            if (a.d_seq) S = a.S;
            else S = a.emo.d_hist_len[env0 + tid] == 0 ? 1 : SMAX;
        }
        s_len[tid] = S;
    }
    __syncthreads();
    // ---- embedding + positional encoding: x[t][d] ----
    for (int i = tid; i < kT * D; i += kTfThreads) {
        const int t = i / D, d = i % D, e = t / SMAX, s = t % SMAX;
        float v = 0.f;
        if (e < n_env && s < s_len[e]) {
            float in[3];
            const int env = env0 + e;
            if (a.d_seq) {
#pragma unroll
                for (int c = 0; c < 3; ++c) in[c] = a.d_seq[((size_t)env * a.S + s) * 3 + c];
            } else {
                const int len = a.emo.d_hist_len[env], head = a.emo.d_hist_head[env];
                if (len == 0) { in[0] = 0.5f; in[1] = 0.5f; in[2] = 0.0f; }
                else {
                    // padded position s holds history[max(0, s - (20 - len))], oldest first
                    const int j = max(0, s - (SMAX - len));
                    const int slot = (head + j) % SMAX;
#pragma unroll
                    for (int c = 0; c < 3; ++c) in[c] = a.emo.d_history[((size_t)slot * 3 + c) * a.n + env];
                }
            }
            v = __ldg(a.w.emb_b + d);
#pragma unroll
            for (int c = 0; c < 3; ++c) v = fmaf(in[c], __ldg(a.w.emb_w + d * 3 + c), v);
            v += __ldg(a.w.pos + s * D + d);
        }
        sm[sX + t * ldX + d] = v;
    }
    __syncthreads();

    for (int l = 0; l < NL; ++l) {
        const float *in_w = a.w.in_proj_w + (size_t)l * 3 * D * D, *in_b = a.w.in_proj_b + l * 3 * D;
        const float *out_w = a.w.out_proj_w + (size_t)l * D * D, *out_b = a.w.out_proj_b + l * D;
        const float *w1 = a.w.lin1_w + (size_t)l * FF * D, *b1 = a.w.lin1_b + (size_t)l * FF;
        const float *w2 = a.w.lin2_w + (size_t)l * D * FF, *b2 = a.w.lin2_b + l * D;
        // ---- q|k|v = x W_in^T + b_in : thread owns output column c (96 of the 128 threads) ----
        if (tid < 3 * D) {
            float wr[D];
#pragma unroll
            for (int d = 0; d < D; ++d) wr[d] = __ldg(in_w + tid * D + d);
            const float bb = __ldg(in_b + tid);
            // four tokens at a time: four independent fma chains (each keeps its d = 0..31 order), nT is a multiple of 20
            for (int t = 0; t < nT; t += 4) {
                float acc[4] = {bb, bb, bb, bb};
#pragma unroll
                for (int d = 0; d < D; ++d)
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = fmaf(sm[sX + (t + u) * ldX + d], wr[d], acc[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) sm[sQ + (t + u) * ldQ + tid] = acc[u];
            }
        }
        __syncthreads();
        // ---- attention: one thread per (env, head, query row) ----
        for (int r = tid; r < kG * NH * SMAX; r += kTfThreads) {
            const int e = r / (NH * SMAX), h = (r / SMAX) % NH, i = r % SMAX;
            const int S = s_len[e];
            if (i >= S) continue;
            const int tq = e * SMAX + i;
            float q[HD], p[SMAX];
#pragma unroll
            for (int d = 0; d < HD; ++d) q[d] = sm[sQ + tq * ldQ + h * HD + d];
            float mx = -INFINITY;
            for (int j = 0; j < S; ++j) {
                float sc = 0.f;
#pragma unroll
                for (int d = 0; d < HD; ++d) sc = fmaf(q[d], sm[sQ + (e * SMAX + j) * ldQ + D + h * HD + d], sc);
                sc *= 0.35355339059327379f;                     // 1/sqrt(8)
                p[j] = sc;
                mx = fmaxf(mx, sc);
            }
            float sum = 0.f;
            for (int j = 0; j < S; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
            const float inv = 1.0f / sum;
            float o[HD];
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = 0.f;
            for (int j = 0; j < S; ++j) {
                float pj = p[j] * inv;
                if (a.dropout_mode != ERLFILL_DROPOUT_OFF)
                    pj *= keep_scale(a, a.row_id_base + env0 + e, 4u * l + 0u, (uint32_t)((h * SMAX + i) * SMAX + j),
                                     a.masks.attn, (((size_t)l * a.n + env0 + e) * NH + h) * a.S * a.S + (size_t)i * a.S + j);
#pragma unroll
                for (int d = 0; d < HD; ++d) o[d] = fmaf(pj, sm[sQ + (e * SMAX + j) * ldQ + 2 * D + h * HD + d], o[d]);
            }
            // each (token, head) slot of q is read only by this thread: reuse it for the attention output
#pragma unroll
            for (int d = 0; d < HD; ++d) sm[sQ + tq * ldQ + h * HD + d] = o[d];
        }
        __syncthreads();
        // ---- out_proj + dropout1 -> sA ; then x = LN1(x + sA) ----
        for (int i = tid; i < nT * D; i += kTfThreads) {
            const int t = i / D, c = i % D, e = t / SMAX, s = t % SMAX;
            float acc = __ldg(out_b + c);
#pragma unroll
            for (int d = 0; d < D; ++d) acc = fmaf(sm[sQ + t * ldQ + d], __ldg(out_w + c * D + d), acc);
            if (a.dropout_mode != ERLFILL_DROPOUT_OFF && e < n_env && s < s_len[e])
                acc *= keep_scale(a, a.row_id_base + env0 + e, 4u * l + 1u, (uint32_t)(s * D + c), a.masks.d1,
                                  (((size_t)l * a.n + env0 + e) * a.S + s) * D + c);
            sm[sA + t * ldX + c] = acc;
        }
        __syncthreads();
        layer_norm_rows(sm, nT, a.w.norm1_w + l * D, a.w.norm1_b + l * D, tid);
        __syncthreads();
        // ---- FFN in chunks of 128 hidden units; accumulators for out[t][o]: thread = (o, token group) ----
        const int o_col = tid & 31, tg = tid >> 5;               // 4 groups x 20 tokens
        float acc[SMAX];
#pragma unroll
        for (int i = 0; i < SMAX; ++i) acc[i] = rank == 0 ? __ldg(b2 + o_col) : 0.f;
        for (int c0 = rank * (FF / kSplit); c0 < (rank + 1) * (FF / kSplit); c0 += kChunk) {
            // hidden: thread j owns hidden unit c0 + j for every token
            {
                float wr[D];
#pragma unroll
                for (int d = 0; d < D; ++d) wr[d] = __ldg(w1 + (size_t)(c0 + tid) * D + d);
                const float bb = __ldg(b1 + c0 + tid);
                for (int t4 = 0; t4 < nT; t4 += 4) {
                    float hh[4] = {bb, bb, bb, bb};
#pragma unroll
                    for (int d = 0; d < D; ++d)
#pragma unroll
                        for (int u = 0; u < 4; ++u) hh[u] = fmaf(sm[sX + (t4 + u) * ldX + d], wr[d], hh[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t4 + u;
                        float h = fmaxf(hh[u], 0.f);
                        if (a.dropout_mode != ERLFILL_DROPOUT_OFF) {
                            const int e = t / SMAX, s = t % SMAX;
                            if (e < n_env && s < s_len[e])
                                h *= keep_scale(a, a.row_id_base + env0 + e, 4u * l + 2u, (uint32_t)(s * FF + c0 + tid),
                                                a.masks.ff, (((size_t)l * a.n + env0 + e) * a.S + s) * FF + c0 + tid);
                        }
                        sm[sHd + t * ldH + tid] = h;
                    }
                }
            }
            // linear2 chunk, transposed into smem: w2s[j][o] = W2[o][c0 + j]
            for (int i = tid; i < kChunk * D; i += kTfThreads) {
                const int o = i / kChunk, j = i % kChunk;
                sm[sW2 + j * D + o] = __ldg(w2 + (size_t)o * FF + c0 + j);
            }
            __syncthreads();
            if (tg < n_env)
            for (int j = 0; j < kChunk; j += 4) {
                const float w0 = sm[sW2 + (j + 0) * D + o_col], w1v = sm[sW2 + (j + 1) * D + o_col];
                const float w2v = sm[sW2 + (j + 2) * D + o_col], w3 = sm[sW2 + (j + 3) * D + o_col];
#pragma unroll
                for (int i = 0; i < SMAX; ++i) {
                    const float4 h = *reinterpret_cast<const float4 *>(&sm[sHd + (tg * SMAX + i) * ldH + j]);
                    acc[i] = fmaf(h.x, w0, acc[i]); acc[i] = fmaf(h.y, w1v, acc[i]);
                    acc[i] = fmaf(h.z, w2v, acc[i]); acc[i] = fmaf(h.w, w3, acc[i]);
                }
            }
            __syncthreads();
        }
        if (kSplit > 1) {
            // partial sums of this CTA's hidden units -> its own smem (the hidden-chunk buffer is free now), then every CTA adds
            // the kSplit partials in rank order
            namespace cg = cooperative_groups;
            cg::cluster_group cluster = cg::this_cluster();
            if (tg < n_env)
#pragma unroll
                for (int i = 0; i < SMAX; ++i) sm[sHd + (tg * SMAX + i) * ldH + o_col] = acc[i];
            cluster.sync();
            if (tg < n_env) {
#pragma unroll
                for (int i = 0; i < SMAX; ++i) acc[i] = 0.f;
                for (int rr = 0; rr < kSplit; ++rr) {
                    const float *remote = cluster.map_shared_rank(sm, rr);
#pragma unroll
                    for (int i = 0; i < SMAX; ++i) acc[i] += remote[sHd + (tg * SMAX + i) * ldH + o_col];
                }
            }
            cluster.sync();                                   // nobody overwrites its partials while a peer still reads them
        }
        if (tg < n_env)
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
            const int t = tg * SMAX + i, e = tg, s = i;
            float v = acc[i];
            if (a.dropout_mode != ERLFILL_DROPOUT_OFF && e < n_env && s < s_len[e])
                v *= keep_scale(a, a.row_id_base + env0 + e, 4u * l + 3u, (uint32_t)(s * D + o_col), a.masks.d2,
                                (((size_t)l * a.n + env0 + e) * a.S + s) * D + o_col);
            sm[sA + t * ldX + o_col] = v;
        }
        __syncthreads();
        layer_norm_rows(sm, nT, a.w.norm2_w + l * D, a.w.norm2_b + l * D, tid);
        __syncthreads();
    }
    // ---- last token -> fc_out -> 0.5*tanh + 0.5 ----
    if (rank == 0 && tid < kG * 3) {
        const int e = tid / 3, c = tid % 3;
        if (e < n_env) {
            const int t = e * SMAX + (s_len[e] - 1);
            float acc = __ldg(a.w.fc_b + c);
#pragma unroll
            for (int d = 0; d < D; ++d) acc = fmaf(sm[sX + t * ldX + d], __ldg(a.w.fc_w + c * D + d), acc);
            a.d_out[(size_t)(env0 + e) * 3 + c] = tanhf(acc) * 0.5f + 0.5f;
        }
    }
}

}  // namespace erlfill

using namespace erlfill;

extern "C" {

int erlfill_emotion_forward(const erlfill_transformer_weights *w, int n_env, const float *d_seq, int seq_len,
                            const erlfill_emotion_state *emo, int dropout_mode, uint64_t seed, uint32_t call_idx,
                            uint32_t row_id_base, const erlfill_transformer_masks *masks, float *d_out, void *stream) {
    if (!w || n_env <= 0 || !d_out) return ERLFILL_ERR_ARG;
    if (!d_seq && (!emo || !emo->d_history || !emo->d_hist_len || !emo->d_hist_head)) return ERLFILL_ERR_ARG;
    if (d_seq && (seq_len < 1 || seq_len > SMAX)) return ERLFILL_ERR_ARG;
    if (dropout_mode == ERLFILL_DROPOUT_MASK && (!masks || !masks->attn || !masks->d1 || !masks->ff || !masks->d2 || !d_seq))
        return ERLFILL_ERR_ARG;
    const float *const *wp = reinterpret_cast<const float *const *>(w);
    for (size_t i = 0; i < sizeof(*w) / sizeof(float *); ++i)
        if (!wp[i]) return ERLFILL_ERR_ARG;
    TfArgs a;
    a.w = *w;
    if (emo) a.emo = *emo; else a.emo = erlfill_emotion_state{};
    if (masks) a.masks = *masks; else a.masks = erlfill_transformer_masks{};
    a.d_seq = d_seq; a.d_out = d_out; a.n = n_env; a.S = d_seq ? seq_len : SMAX;
    a.dropout_mode = dropout_mode; a.has_masks = masks ? 1 : 0;
    a.call_idx = call_idx; a.row_id_base = row_id_base; a.seed = seed;
    const size_t smem = kTfSmemFloats * sizeof(float);
    static bool attr_set[64] = {};
    int dev = 0;
    ERLFILL_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        ERLFILL_CUDA_TRY(cudaFuncSetAttribute(emotion_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ERLFILL_CUDA_TRY(cudaFuncSetAttribute(emotion_forward_kernel<kSplitSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    const int groups = (n_env + kG - 1) / kG;
    if (groups * kSplitSmall <= 160) {
        // small batch (<= 80 envs): kSplitSmall CTAs per env group as one thread-block cluster
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(groups * kSplitSmall), 1, 1);
        cfg.blockDim = dim3(kTfThreads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kSplitSmall; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        ERLFILL_CUDA_TRY(cudaLaunchKernelEx(&cfg, emotion_forward_kernel<kSplitSmall>, a));
    } else {
        emotion_forward_kernel<1><<<groups, kTfThreads, smem, (cudaStream_t)stream>>>(a);
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
