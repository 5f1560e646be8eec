// update_tc.cu — the tcgen05 ("fast") ER-DDPG minibatch update: DDPGAgent.learn, DifferentModules/ddpg_agent.py:422-512,
// rows U1-U5 of SURVEY.md section 8a, with train()-mode dropout (the reference never calls .eval() on this path).
//
// Six launches per update (erlfill_ddpg_update with hyper->precision == ERLFILL_PRECISION_FAST); the two phase kernels run as clusters
// of two CTAs per 128-row tile (the pair splits the output columns of every layer: update_tc.cuh):
//   CRITIC_GRAD   ddpg_critic_grad_tc   2 x B/128 pairs: one half runs the target chain actor_target -> critic_target and hands Q' over,
//                                       then the forward of the CURRENT actor for ACTOR_GRAD (it does not depend on the critic update);
//                                       the other half Q(s, a), SmoothL1 and the whole critic backward incl. the weight-gradient GEMMs
//                                       (per-tile partials).  One pair does all of it when the tiles outnumber the resident pairs.
//                 grad_reduce_kernel    partials -> flat gradient bucket (tile order, deterministic) + clip-norm partials
//                                       (data-parallel: its last block raises this rank's bucket-ready flags on the peers, peer.cuh)
//   CRITIC_APPLY  critic_apply_kernel   emotion-driven lr, clip_grad_norm_(5), Adam, refresh of the critic's bf16 weight images
//   ACTOR_GRAD    ddpg_actor_grad_tc    B/128 pairs: Q(s, actor(s)) with the UPDATED critic, critic input gradient, actor backward
//                 grad_reduce_kernel    (+ the L2-norm regulariser's gradient)
//   ACTOR_APPLY   actor_apply_kernel    clip, Adam, both soft target updates, all target / actor images, report
// Building blocks and the operand-image scheme: update_tc.cuh.  Numerics: bf16 operands, fp32 accumulation, fp32 master weights,
// LayerNorm / losses / optimiser in fp32 (stated tolerance against the fp32 reference: tests/test_update_tc_gpu.py); the exact
// 3xTF32 path of update.cu stays the 1e-4 parity path.
#include <stdlib.h>

#include "update_tc.cuh"

#include "mlp.cuh"
#include "peer.cuh"

namespace erlfill {
namespace utc {

constexpr int ACT_LEAKY = 0, ACT_RELU = 1;

// ---- per-tile global scratch (bytes), shared by the two CTAs of the pair (each writes its own columns) ------------------------
namespace SC {
constexpr size_t X0IMG = 0, H1IMG = X0IMG + 12288, H2IMG = H1IMG + 65536, XAIMG = H2IMG + 65536, A1IMG = XAIMG + 12288, A2IMG = A1IMG + 32768,
                 A3IMG = A2IMG + 32768, ABITS = A3IMG + 16384 /* [128][32] words */, AUX = ABITS + 16384 /* [2 CTAs][64][128] floats */,
                 XCIMG = AUX + 65536 /* critic input image (s, actor(s), e) */, POS3 = XCIMG + 12288 /* [2 CTAs][512] words */, BYTES = POS3 + 4096;
}

struct Args {
    int B, n_cond;
    float gamma;
    const float *batch, *next_emotion;
    const float *actor, *actor_t, *critic, *critic_t;            // fp32 master parameters (small vectors are read from these)
    const uint8_t *img_actor, *img_actor_t, *img_critic, *img_critic_t;
    const float *scale, *offset;
    uint8_t *scratch;                                            // [tiles][SC::BYTES]
    float *cpart, *apart;                                        // [tiles][CP::SIZE], [tiles][AP::SIZE]
    float *mean_out;                                             // 3 batch emotion means (tail of the critic gradient bucket)
    float *scal;                                                 // scal[6 .. 15): L2 norms of the nine actor tensors
    Drop drop;
    int prof;
    // CRITIC_GRAD with the target chain on its own pairs (grid = 2 x 2 x tiles CTAs, all co-resident): pair tiles + t computes Q' of tile
    // t, writes it to tq_g and raises tq_flag[2 t + c] for CTA c of pair t, which consumes it before the loss and lowers its flag again
    int split;
    float *tq_g;
    uint32_t *tq_flag;
    float *stats_g;              // [6] batch emotion mean / std, left by CRITIC_GRAD for ACTOR_GRAD (same minibatch)
    float *stats_part;           // [tiles][8]
    uint32_t *stats_ctr;         // [2]
    const float *stats_ext;      // optional: the sums erlfill_replay_sample_gather left for this minibatch, n_stats_ext rows of 8
    int n_stats_ext;
};

__constant__ int c_actor_seg[10] = {0, 30, 670, 798, 17182, 17310, 25502, 25566, 26206, 26216};

// development profile (erlfill_debug_utc_profile): clock64 of CTA 0 at the numbered points of the two phase kernels
__device__ long long g_utc_prof[2][128];
__device__ long long g_utc_cta[2][256][4];       // per CTA: globaltimer at entry, after setup, at the end of the epilogue role, at exit
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define UCTA(k, i) do { if (a.prof && threadIdx.x == 0 && blockIdx.x < 256) g_utc_cta[k][blockIdx.x][i] = gtimer(); } while (0)
#define UPROF(k, i) do { if (a.prof && blockIdx.x == 0 && e.tid == 0) g_utc_prof[k][i] = clock64(); } while (0)

// ---- common prologue / epilogue of the two phase kernels ---------------------------------------------------------------
__device__ __forceinline__ void setup(uint8_t *smem, uint64_t *bars, uint32_t *tmem_slot, int tid, int warp) {
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bars + BW_FULL + s, 1); mbar_init(bars + BW_EMPTY + s, 1); }
        mbar_init(bars + B_HFULL, 1); mbar_init(bars + B_HFREE, 1); mbar_init(bars + B_DFULL, 1); mbar_init(bars + B_GFULL, 1);
        mbar_init(bars + B_GFREE, NEPI / 32);
        mbar_init(bars + B_AREADY, 1);                           // one local arrival (+ the bytes of the peer's half)
        mbar_init(bars + B_AFREE, 2);                            // the issuers of both CTAs
        mbar_init(bars + B_XRED, NEPI / 32); mbar_init(bars + B_XRED1, NEPI / 32);      // local warps (+ the bytes of the peer's partials)
        mbar_init(bars + B_HPUB, 2 * NEPI / 32);                 // the epilogue warps of both CTAs
        mbar_fence_init();
    }
    if (warp == W_ISSUER) tmem_alloc(tmem_slot, 512);
    // (the MN-major reads of narrow images touch bytes nobody writes: they only feed accumulator lanes nobody reads)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // the peer's barriers exist before anything arrives on them
    tc_fence_after();
}

// The small fp32 vectors the epilogues read (biases, LayerNorm gamma / beta, head weights) into L1 while the first layers run: their
// first touch would otherwise cost every epilogue an L2 round trip.
__device__ __forceinline__ void prefetch_floats(const float *p, int n, int t, int &t0) {
    const int lines = (n * 4 + 127) / 128 + 1, i = t - t0;
    if (i >= 0 && i < lines) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + i * 32));
    t0 += lines;
}
__device__ __forceinline__ void prefetch_critic_vectors(const float *P, int tid, int &t0) {
    prefetch_floats(P + CO::G1, CO::W4 - CO::G1, tid, t0);
    prefetch_floats(P + CO::B4, CO::W8 - CO::B4, tid, t0);
    prefetch_floats(P + CO::B8, CO::TOTAL - CO::B8, tid, t0);
}
__device__ __forceinline__ void prefetch_actor_vectors(const float *P, int tid, int &t0) {
    prefetch_floats(P + AO::EW, 30, tid, t0);
    prefetch_floats(P + AO::B2, 128, tid, t0);
    prefetch_floats(P + AO::B3, 64, tid, t0);
    prefetch_floats(P + AO::B4, 10, tid, t0);
}

// batch mean / unbiased std of the three emotion columns (Critic.forward, ddpg_agent.py:176), by the epilogue threads of every CTA
// (recomputing 3 x B values per CTA is cheaper than a grid-wide dependency); loads unrolled so that their latencies overlap
__device__ __forceinline__ void emotion_stats(Epi &e, const float *batch, float *mean_out) {
    float *st = e.stat();
    float *scr = st + 16;                                        // [16 warps][3]
    float s[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
    for (int r = e.tid; r < e.B; r += NEPI) {
        const float *p = batch + (size_t)r * ERLFILL_REC + 16;
        s[0] += __ldg(p); s[1] += __ldg(p + 1); s[2] += __ldg(p + 2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float t = warp_sum(s[c]); if (e.lane == 0) scr[(e.tid >> 5) * 3 + c] = t; }
    bar_epi();
    if (e.tid < 3) { float t = 0.f; for (int w = 0; w < NEPI / 32; ++w) t += scr[w * 3 + e.tid]; st[e.tid] = t / (float)e.B; }
    bar_epi();
    const float m0 = st[0], m1 = st[1], m2 = st[2];
    float q[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
    for (int r = e.tid; r < e.B; r += NEPI) {
        const float *p = batch + (size_t)r * ERLFILL_REC + 16;
        const float d0 = __ldg(p) - m0, d1 = __ldg(p + 1) - m1, d2 = __ldg(p + 2) - m2;
        q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float t = warp_sum(q[c]); if (e.lane == 0) scr[(e.tid >> 5) * 3 + c] = t; }
    bar_epi();
    if (e.tid < 3) {
        float t = 0.f;
        for (int w = 0; w < NEPI / 32; ++w) t += scr[w * 3 + e.tid];
        st[3 + e.tid] = sqrtf(t / (float)(e.B - 1));
        if (mean_out) mean_out[e.tid] = st[e.tid];
    }
    bar_epi();
}

// The same statistics with the work shared by the tiles of the grid (all co-resident: the split launch of CRITIC_GRAD): CTA 0 of every
// pair sums (e - 0.5) and (e - 0.5)^2 over the tile's 128 rows and publishes the six sums; once all tiles have arrived, BOTH CTAs add
// the tiles' sums in tile order (the same numbers everywhere).  ctr[0] counts arrivals (one per tile), ctr[1] departures (two per
// tile); the last CTA to leave zeroes both for the next launch.
__device__ __forceinline__ void grid_emotion_stats(Epi &e, const float *batch, float *part_g, uint32_t *ctr, int n_tiles, int tile, float *stats_out,
                                                   float *mean_out) {
    float *st = e.stat();
    float *scr = st + 16;                                        // [4 warps][6]
    if (e.cr == 0 && e.wg == 0) {
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (e.valid()) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(batch + (size_t)(e.row0 + e.r) * ERLFILL_REC + 16));
            const float d0 = q.x - 0.5f, d1 = q.y - 0.5f, d2 = q.z - 0.5f;
            v[0] = d0; v[1] = d1; v[2] = d2; v[3] = d0 * d0; v[4] = d1 * d1; v[5] = d2 * d2;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) { const float t = warp_sum(v[c]); if (e.lane == 0) scr[e.wq * 6 + c] = t; }
    }
    bar_epi();
    if (e.tid == 0) {
        if (e.cr == 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) part_g[tile * 8 + c] = (scr[c] + scr[6 + c]) + (scr[12 + c] + scr[18 + c]);
            __threadfence();
            atomicAdd(ctr, 1u);
        }
        const long long t0 = clock64();
        bool ok = true;
        while (*reinterpret_cast<volatile uint32_t *>(ctr) < (uint32_t)n_tiles) {
            if (clock64() - t0 > 4000000000LL) { ok = false; break; }        // ~2 s: a CTA of the grid never became resident
        }
        __threadfence();
        float sum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < n_tiles; ++t)
#pragma unroll
            for (int c = 0; c < 6; ++c) sum[c] += __ldcg(part_g + t * 8 + c);
        const float n = (float)e.B;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float m = sum[c] / n;
            st[c] = ok ? 0.5f + m : __int_as_float(0x7fc00000);
            st[3 + c] = sqrtf(fmaxf(sum[3 + c] - sum[c] * m, 0.f) / (n - 1.0f));
        }
        if (tile == 0 && e.cr == 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) stats_out[c] = st[c];
            mean_out[0] = st[0]; mean_out[1] = st[1]; mean_out[2] = st[2];
        }
        if (atomicAdd(ctr + 1, 1u) == 2u * (uint32_t)n_tiles - 1u) { ctr[0] = 0u; ctr[1] = 0u; __threadfence(); }
    }
    bar_epi();
}

// x[16] -> the K = 48 first-layer operand [x_hi | x_lo | x_hi] of this thread's row (image of 48 features) in THIS CTA's IMG_A (both
// CTAs of the pair build the narrow input images themselves), optionally copied to global
__device__ __forceinline__ void store_split_input(const Epi &e, const float (&x)[16], uint8_t *gimg) {
    float hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        hi[j] = __bfloat162float(__float2bfloat16(x[j]));
        lo[j] = x[j] - hi[j];
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        e.img_store8(e.img_a(), 48, g * 8, hi + g * 8);
        e.img_store8(e.img_a(), 48, 16 + g * 8, lo + g * 8);
        e.img_store8(e.img_a(), 48, 32 + g * 8, hi + g * 8);
        if (gimg) { e.img_store8(gimg, 48, g * 8, hi + g * 8); e.img_store8(gimg, 48, 16 + g * 8, lo + g * 8); e.img_store8(gimg, 48, 32 + g * 8, hi + g * 8); }
    }
}

// Actor head on the 16 accumulator columns of the last layer (ddpg_agent.py:104-122): out[10] in physical units.
// aux (optional): 50 floats of this row at stride 128 ([k][row]: a warp's rows are contiguous): z4[10] base[10] raw[10] alpha[10] out[10]
// (what the backward needs).
__device__ __forceinline__ void actor_head(const float (&z16)[16], const float *P, float cur, float cons, float anx, const float *scale,
                                           const float *offset, float (&out)[10], float *aux) {
    float alpha = 0.1f + 0.9f * anx;
    alpha *= (1.0f - 0.5f * cons);
    alpha *= (1.0f + 0.5f * cur);
    alpha = fminf(fmaxf(alpha, 0.01f), 1.0f);
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        const float z = z16[c] + __ldg(P + AO::B4 + c);
        const float base = z / (1.0f + fabsf(z));
        float em = cur * __ldg(P + AO::EW + c);
        em = fmaf(cons, __ldg(P + AO::EW + 10 + c), em);
        em = fmaf(anx, __ldg(P + AO::EW + 20 + c), em);
        const float raw = tanhf(em * 1.5f);
        const float om = 1.0f - fabsf(base);
        const float sf = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
        out[c] = (base + alpha * raw * sf * (om * om)) * __ldg(scale + c) + __ldg(offset + c);
        if (aux) { aux[c * TR] = z; aux[(10 + c) * TR] = base; aux[(20 + c) * TR] = raw; aux[(30 + c) * TR] = alpha; aux[(40 + c) * TR] = out[c]; }
    }
}

// Hidden layer with N = 128 outputs (this CTA's 64: 16 columns per thread): z = D (+ bias) -> activation -> dropout -> this CTA's half of the image of 128
// features (the peer gets it with a_ready).  pos_out / keep_out: bit j = (z > 0) / kept, of feature cr * 64 + wg * 16 + j.
__device__ __forceinline__ void epi_hidden128(Epi &e, const float *bias, int act, const Drop &dr, int site, uint8_t *gimg, uint32_t *pos_out,
                                              uint32_t *keep_out) {
    float v[16];
    const int cl = e.wg * 16, c0 = e.cr * 64 + cl;
    tmem_ld16w(e.td(cl), v);
    const uint32_t keep = ddrop::keep_bits<16>(dr, site, e.mrow(), c0, 128);
    uint32_t pos = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float z = v[j] + (bias ? __ldg(bias + c0 + j) : 0.f);
        pos |= (z > 0.f ? 1u : 0u) << j;
        z = act == ACT_LEAKY ? (z > 0.f ? z : 0.1f * z) : fmaxf(z, 0.f);
        v[j] = ((keep >> j) & 1u) ? z * (dr.mode == ERLFILL_DROPOUT_OFF ? 1.0f : kDropScale) : 0.f;
    }
    e.wait_afree();
    e.img_store8_a(128, c0, v, gimg);
    e.img_store8_a(128, c0 + 8, v + 8, gimg);
    if (pos_out) *pos_out = pos;
    if (keep_out) *keep_out = keep;
}

// Linear(128, 64) -> ReLU (this CTA's 32 columns: 8 per thread) -> this CTA's half of the image of 64 features; returns the relu' bits
__device__ __forceinline__ uint32_t epi_hidden64(Epi &e, const float *bias, uint8_t *gimg) {
    float v[8];
    const int cl = e.wg * 8, c0 = e.cr * 32 + cl;
    tmem_ld8w(e.td(cl), v);
    uint32_t pos = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = fmaxf(v[j] + __ldg(bias + c0 + j), 0.f);
        pos |= (v[j] > 0.f ? 1u : 0u) << j;
    }
    e.wait_afree();
    e.img_store8_a(64, c0, v, gimg);
    return pos;
}

// LayerNorm(256) + activation + dropout of the critic (this CTA's 128 columns: 32 per thread, held in registers across the exchange
// of the row statistics) -> image of 256 features, then a_ready.
// save != nullptr: what the backward needs stays on chip — xhat (fp32) goes to this thread's TMEM lane (columns TM_X1 / TM_X2 of layer
// 0 / 1), rstd and the keep word to save[0..1] (registers of the caller); gimg: global copy of the image for the weight-gradient
// GEMM.  Both are written AFTER a_ready, off the critical path.
__device__ __forceinline__ void epi_ln256(Epi &e, const float *bias, const float *gamma, const float *beta, int act, const Drop &dr, int site,
                                          int layer, uint32_t *save, uint8_t *gimg) {
    const int cl = e.wg * 32, c0 = e.cr * 128 + cl;
    float v[32];
    tmem_ld32w(e.td(cl), v);
    if (bias) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + c0) + j4);
            v[4 * j4] += b4.x; v[4 * j4 + 1] += b4.y; v[4 * j4 + 2] += b4.z; v[4 * j4 + 3] += b4.w;
        }
    }
    // shifted sums of this thread's 32 columns -> (mean_w, M2_w); the eight statistics of the row combine exactly (Chan et al.)
    const float k = v[0];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { const float d = v[j] - k; s1 += d; s2 = fmaf(d, d, s2); }
    const float mean_w = k + s1 * (1.0f / 32.0f);
    float a1 = mean_w, a2 = (s2 - s1 * s1 * (1.0f / 32.0f)) + 32.0f * mean_w * mean_w;       // M2_w + 32 mean_w^2
    e.row_allreduce2(a1, a2);
    const float mean = a1 * 0.125f;
    const float var = fmaxf(a2 - 256.0f * mean * mean, 0.f) * (1.0f / 256.0f);
    const float rstd = rsqrtf(var + 1e-5f);
    const float ds = dr.mode == ERLFILL_DROPOUT_OFF ? 1.0f : kDropScale;
    const uint32_t kb = keep32(dr, site, e.mrow(), c0, 256);
    uint32_t packed[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(gamma + c0) + j4), b4 = __ldg(reinterpret_cast<const float4 *>(beta + c0) + j4);
        const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
        float hh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = j4 * 4 + i;
            const float xh = (v[j] - mean) * rstd;
            v[j] = xh;
            float y = xh * gm[i] + be[i];
            y = act == ACT_LEAKY ? (y > 0.f ? y : 0.1f * y) : fmaxf(y, 0.f);
            hh[i] = ((kb >> j) & 1u) ? y * ds : 0.f;
        }
        packed[2 * j4] = pack_bf16x2(hh[0], hh[1]);
        packed[2 * j4 + 1] = pack_bf16x2(hh[2], hh[3]);
    }
    e.wait_afree();
#pragma unroll
    for (int g = 0; g < 4; ++g) e.img_store_a_q(256, c0 + g * 8, make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]));
    e.a_ready(256);
    if (gimg) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4 *>(gimg + img_off(e.r, c0 + g * 8, 256)) = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
    }
    if (save) {
        uint32_t xw[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) xw[j] = __float_as_uint(v[j]);
        tmem_st32(e.tx(layer, cl), xw);
        tmem_wait_st();
        save[0] = __float_as_uint(rstd);
        save[1] = kb;
    }
}

// critic layer 8 (N = 128, this CTA's 64: 16 columns per thread): h3 = relu(D + b8), q = clamp(h3 . w10 + b10): every thread of the
// row, in both CTAs, gets q
__device__ __forceinline__ float epi_head(Epi &e, const float *P, float (&h3)[16]) {
    const int cl = e.wg * 16, c0 = e.cr * 64 + cl;
    tmem_ld16w(e.td(cl), h3);
    float part = 0.f, dummy = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        h3[j] = fmaxf(h3[j] + __ldg(P + CO::B8 + c0 + j), 0.f);
        part = fmaf(h3[j], __ldg(P + CO::W10 + c0 + j), part);
    }
    e.row_allreduce2(part, dummy);
    return fminf(fmaxf(part + __ldg(P + CO::B10), -1e6f), 1e6f);
}

// Backward of (LayerNorm(256) + activation + dropout): D holds dh (gradient of the dropped activation) of this CTA's 128 columns;
// writes dz (gradient of the LayerNorm input) into the image of 256 features, then a_ready.  xhat comes back from this thread's TMEM
// lane, rstd / keep word from save[0..1] (epi_ln256).  part != nullptr: per-tile column sums [sum dz | sum dy xhat | sum dy] of this
// CTA's columns (stored after a_ready).
__device__ __forceinline__ void epi_ln256_bwd(Epi &e, const float *gamma, const float *beta, int act, const Drop &dr, int layer, const uint32_t *save,
                                              float *part) {
    const int cl = e.wg * 32, c0 = e.cr * 128 + cl;
    const float ds = dr.mode == ERLFILL_DROPOUT_OFF ? 1.0f : kDropScale;
    const uint32_t kw = save[1];
    const float rstd = __uint_as_float(save[0]);
    const float4 *gp = reinterpret_cast<const float4 *>(gamma + c0), *bp = reinterpret_cast<const float4 *>(beta + c0);
    float s1 = 0.f, s2 = 0.f;
    // pass 1: dy = dh * keep/0.8 * act'(y); row sums of dy*gamma and dy*gamma*xhat; column sums of dy*xhat and dy
    {
        float dh[32], a[32], b[32];
        tmem_ld32w(e.td(cl), dh);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
            float xs[8];
            tmem_ld8w(e.tx(layer, cl + 8 * g8), xs);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 g4 = __ldg(gp + 2 * g8 + h), b4 = __ldg(bp + 2 * g8 + h);
                const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = g8 * 8 + h * 4 + i;
                    const float x = xs[h * 4 + i], y = x * gm[i] + be[i];
                    float g = ((kw >> j) & 1u) ? dh[j] * ds : 0.f;
                    g *= act == ACT_LEAKY ? (y > 0.f ? 1.f : 0.1f) : (y > 0.f ? 1.f : 0.f);
                    const float dxh = g * gm[i];
                    s1 += dxh;
                    s2 = fmaf(dxh, x, s2);
                    a[j] = g * x;
                    b[j] = g;
                }
            }
        }
        if (part) { e.colsum_stage<32>(1, cl, a); e.colsum_stage<32>(2, cl, b); }
    }
    e.row_allreduce2(s1, s2);
    s1 *= (1.0f / 256.0f); s2 *= (1.0f / 256.0f);
    // pass 2: dz = rstd * (dy*gamma - s1 - xhat*s2)
    {
        float dh[32], dz[32];
        tmem_ld32w(e.td(cl), dh);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
            float xs[8];
            tmem_ld8w(e.tx(layer, cl + 8 * g8), xs);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 g4 = __ldg(gp + 2 * g8 + h), b4 = __ldg(bp + 2 * g8 + h);
                const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = g8 * 8 + h * 4 + i;
                    const float x = xs[h * 4 + i], y = x * gm[i] + be[i];
                    float g = ((kw >> j) & 1u) ? dh[j] * ds : 0.f;
                    g *= act == ACT_LEAKY ? (y > 0.f ? 1.f : 0.1f) : (y > 0.f ? 1.f : 0.f);
                    dz[j] = rstd * (g * gm[i] - s1 - x * s2);
                }
            }
        }
        e.wait_afree();
#pragma unroll
        for (int g = 0; g < 4; ++g) e.img_store8_a(256, c0 + g * 8, dz + g * 8);
        e.a_ready(256);
        if (part) e.colsum_stage<32>(0, cl, dz);
    }
    if (part) {
        bar_epi();
        e.colsum_finish(0, 128, part + e.cr * 128);
        e.colsum_finish(1, 128, part + 256 + e.cr * 128);
        e.colsum_finish(2, 128, part + 512 + e.cr * 128);
        bar_epi();
    }
}

// Drain of one weight-gradient block: TMEM lane r = output feature m0 + r, this CTA's `n` accumulator columns = the columns
// [col0, col0 + n) of the block.  The block is stored as [columns / 4][128 rows][4] so that the 32 rows of a warp write 512 contiguous
// bytes (grad_reduce_kernel walks the same order).  dst == nullptr: the accumulators are only released (a block both CTAs computed).
__device__ __forceinline__ void drain_g(Epi &e, float *dst, int n, int col0) {
    e.wait_g();
    float4 *d4 = reinterpret_cast<float4 *>(dst) + (size_t)(col0 >> 2) * TR + e.r;
    if (dst == nullptr) {
    } else if (n >= 128) {
        const int cpw = n / NWG;
        for (int c = 0; c < cpw; c += 32) {
            float v[32];
            tmem_ld32w(e.tg(e.wg * cpw + c), v);
            float4 *o = d4 + (size_t)((e.wg * cpw + c) >> 2) * TR;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[(size_t)j * TR] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    } else if (n == 32) {                   // 8 columns per warpgroup
        float v[8];
        tmem_ld8w(e.tg(e.wg * 8), v);
        float4 *o = d4 + (size_t)(e.wg * 2) * TR;
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[TR] = make_float4(v[4], v[5], v[6], v[7]);
    } else if (n == 64 || e.wg == 0) {      // n = 64 (16 columns per warpgroup) or n = 16 (warpgroup 0 only)
        float v[16];
        const int c = n == 64 ? e.wg * 16 : 0;
        tmem_ld16w(e.tg(c), v);
        float4 *o = d4 + (size_t)(c >> 2) * TR;
#pragma unroll
        for (int j = 0; j < 4; ++j) o[(size_t)j * TR] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    e.g_free();
}

#define UTC_ROLE_PROLOGUE                                                                                      \
    extern __shared__ __align__(128) uint8_t smem[];                                                            \
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);                                               \
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);                                         \
    const int tid = threadIdx.x;                                                                                \
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                                                     \
    const uint32_t cr = cluster_ctarank();                                                                      \
    const int pair = (int)blockIdx.x >> 1;                                                                      \
    setup(smem, bars, tmem_slot, tid, warp);                                                                    \
    const uint32_t tbase = *tmem_slot;

// the last image's a_free has arrived from both issuers (nothing of the peer is still in flight towards this CTA), then both CTAs leave together
#define UTC_EPI_DONE mbar_wait_cluster(bars + B_AFREE, (e.img_n - 1) & 1u);

#define UTC_ROLE_EPILOGUE                                                                                      \
    tc_fence_before();                                                                                          \
    __syncthreads();                                                                                            \
    cluster_sync_all();                                                                                         \
    if (warp == W_ISSUER) tmem_dealloc(tbase, 512);

#define UTC_MAKE_EPI                                                                                                                       \
    Epi e{smem, bars, tbase, tid, tid >> 7, tid & 127, tid & 31, (tid >> 5) & 3, (uint32_t)(((tid >> 5) & 3) * 32) << 16, (int)cr,            \
          mapa_u32(smem_u32(smem), cr ^ 1u)};

// =====================================================================================================================
// CRITIC_GRAD
// =====================================================================================================================
// a = actor(s, e) of ACTOR_GRAD does not depend on the critic update: it runs inside CRITIC_GRAD, on the pairs of the target chain once
// they have handed over Q' (they finish long before the critic chain's backward), or at the end of the unsplit phase
template <int ROLE>
__device__ __forceinline__ void actor_forward_program(Pipe &p, const Args &a) {
    gemm_fwd<ROLE>(p, a.img_actor + AI::W1, 128, 48);
    gemm_fwd<ROLE>(p, a.img_actor + AI::W2, 128, 128);
    gemm_fwd<ROLE>(p, a.img_actor + AI::W3, 64, 128);
    gemm_fwd<ROLE>(p, a.img_actor + AI::W4, 16, 64, false);
}
// part: 0 = the whole phase in one pair, 1 = target chain (+ the actor forward), 2 = everything but the target chain
template <int ROLE>
__device__ __forceinline__ void critic_program(Pipe &p, const Args &a, const uint8_t *scr, int part) {
    const int cr = (int)p.cr;
    if (part != 2) {    // target chain: actor_target, critic_target
        gemm_fwd<ROLE>(p, a.img_actor_t + AI::W1, 128, 48);
        gemm_fwd<ROLE>(p, a.img_actor_t + AI::W2, 128, 128);
        gemm_fwd<ROLE>(p, a.img_actor_t + AI::W3, 64, 128);
        gemm_fwd<ROLE>(p, a.img_actor_t + AI::W4, 16, 64, false);
        gemm_fwd<ROLE>(p, a.img_critic_t + CI::W0, 256, 48);
        gemm_fwd<ROLE>(p, a.img_critic_t + CI::W4, 256, 256);
        gemm_fwd<ROLE>(p, a.img_critic_t + CI::W8, 128, 256);
    }
    if (part == 1) { actor_forward_program<ROLE>(p, a); return; }
    // Q(s, a_norm, e)
    gemm_fwd<ROLE>(p, a.img_critic + CI::W0, 256, 48);
    gemm_fwd<ROLE>(p, a.img_critic + CI::W4, 256, 256);
    gemm_fwd<ROLE>(p, a.img_critic + CI::W8, 128, 256);
    // backward: dh2 = dz3 W8 (this CTA's 128 inputs), dW8 = dz3^T h2 (this CTA's 128 input columns)
    load_h<ROLE>(p, scr + SC::H2IMG, 65536);
    gemm_bwd<ROLE>(p, a.img_critic + CI::W8, 128, 256, 256, true, false);
    wgrad<ROLE>(p, 128, 0, 256, cr * 128, 128, true, false, true, true);
    // dh1 = dz2 W4, dW4 = dz2^T h1 (this CTA's block of 128 output rows)
    load_h<ROLE>(p, scr + SC::H1IMG, 65536);
    gemm_bwd<ROLE>(p, a.img_critic + CI::W4, 256, 256, 256, true, false);
    wgrad<ROLE>(p, 256, cr * 128, 256, 0, 128, true, false, false, false);      // two column halves: 128 accumulator columns at a time
    wgrad<ROLE>(p, 256, cr * 128, 256, 128, 128, false, false, true, true);
    // dW0 = dz1^T x0 (this CTA's block of 128 output rows)
    load_h<ROLE>(p, scr + SC::X0IMG, 12288);
    wgrad<ROLE>(p, 256, cr * 128, 48, 0, 16, true, true, true, true);
    if (part == 0) actor_forward_program<ROLE>(p, a);
}

__global__ void __launch_bounds__(NTHR, 1) ddpg_critic_grad_tc(const Args a) {
    UCTA(0, 0);
    UTC_ROLE_PROLOGUE
    UCTA(0, 1);
    const int n_tiles = (a.B + TR - 1) / TR;
    const int tile = a.split ? pair % n_tiles : pair;
    const int chain = a.split ? (pair >= n_tiles ? 1 : 2) : 0;
    uint8_t *scr = a.scratch + (size_t)tile * SC::BYTES;
    if (warp == W_PRODUCER) {
        if (elect_one()) { Pipe p{smem, bars, tbase, cr}; critic_program<ROLE_PRODUCER>(p, a, scr, chain); }
        __syncwarp();
    } else if (warp == W_ISSUER) {
        if (elect_one()) { Pipe p{smem, bars, tbase, cr}; critic_program<ROLE_ISSUER>(p, a, scr, chain); }
        __syncwarp();
    } else {
        UTC_MAKE_EPI
        e.row0 = tile * TR; e.B = a.B;
        const int grow = e.row0 + e.r;
        const bool valid = e.valid();
        const float *row = a.batch + (size_t)(valid ? grow : 0) * ERLFILL_REC;
        float *part = a.cpart + (size_t)tile * CP::SIZE;
        {
            int t0 = 0;
            if (chain != 2) { prefetch_actor_vectors(a.actor_t, e.tid, t0); prefetch_critic_vectors(a.critic_t, e.tid, t0); prefetch_actor_vectors(a.actor, e.tid, t0); }
            if (chain != 1) prefetch_critic_vectors(a.critic, e.tid, t0);
        }
        UPROF(0, 0);
        if (a.stats_ext) {
            // the minibatch came with its emotion sums: warp c adds column c of the rows (lane l takes rows l, l + 32, ..., then a butterfly:
            // the same order on every CTA), no pass over the records
            if (e.tid < 6 * 32) {
                const int c = e.tid >> 5;
                float t = 0.f;
                for (int q = e.lane; q < a.n_stats_ext; q += 32) t += __ldg(a.stats_ext + q * 8 + c);
                t = warp_sum(t);
                if (e.lane == 0) e.stat()[16 + c] = t;
            }
            bar_epi();
            if (e.tid < 3) {
                float *st_ = e.stat();
                const float n = (float)a.B, m = st_[16 + e.tid] / n;
                st_[e.tid] = 0.5f + m;
                st_[3 + e.tid] = sqrtf(fmaxf(st_[19 + e.tid] - st_[16 + e.tid] * m, 0.f) / (n - 1.0f));
            }
            bar_epi();
            if (chain != 1 && tile == 0 && cr == 0 && e.tid < 6) { a.stats_g[e.tid] = e.stat()[e.tid]; if (e.tid < 3) a.mean_out[e.tid] = e.stat()[e.tid]; }
        } else if (chain == 2) grid_emotion_stats(e, a.batch, a.stats_part, a.stats_ctr, n_tiles, tile, a.stats_g, a.mean_out);
        else if (chain == 0) {
            emotion_stats(e, a.batch, tile == 0 && cr == 0 ? a.mean_out : nullptr);
            if (tile == 0 && cr == 0 && e.tid < 6) a.stats_g[e.tid] = e.stat()[e.tid];
        }
        UPROF(0, 1);
        const float *st = e.stat();
        const int cid = (a.n_cond > 1 && valid) ? (int)__ldg(row + 19) : 0;
        const float *scale = a.scale + cid * 10, *offset = a.offset + cid * 10;
        // ---- a = actor(s, e) for ACTOR_GRAD (actor_forward_program): operand images, activation bits, head values and the critic input
        // (s, a, e) go to the tile's scratch ----
        auto actor_forward = [&]() {
        uint32_t *abits = reinterpret_cast<uint32_t *>(scr + SC::ABITS) + e.r * 32 + (cr * 4 + e.wg);
        float *aux = reinterpret_cast<float *>(scr + SC::AUX) + (size_t)cr * 64 * TR + e.r;
        const float e0 = valid ? __ldg(row + 16) : 0.f, e1 = valid ? __ldg(row + 17) : 0.f, e2 = valid ? __ldg(row + 18) : 0.f;
        e.wait_afree();
        if (e.wg == 0) {
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = 0.f;
            if (valid) { x[0] = __ldg(row); x[1] = __ldg(row + 1); x[2] = e0; x[3] = e1; x[4] = e2; }
            x[15] = 1.0f;
            store_split_input(e, x, cr == 0 ? scr + SC::XAIMG : nullptr);
        }
        e.a_ready_local();
        e.wait_d(); epi_hidden128(e, nullptr, ACT_LEAKY, a.drop, 6, scr + SC::A1IMG, abits, abits + 8); e.a_ready(128);
        e.wait_d(); epi_hidden128(e, a.actor + AO::B2, ACT_RELU, a.drop, 7, scr + SC::A2IMG, abits + 16, abits + 24); e.a_ready(128);
        e.wait_d();
        // relu'(z3) bits of this thread's 8 columns: the backward in ACTOR_GRAD runs on the same thread of the same CTA rank
        reinterpret_cast<uint32_t *>(scr + SC::POS3)[cr * NEPI + e.tid] = epi_hidden64(e, a.actor + AO::B3, scr + SC::A3IMG);
        e.a_ready(64);
        e.wait_d();
        // the head: both CTAs (ACTOR_GRAD's backward reads this CTA rank's copy of aux), the critic input and the std sum by CTA 0
        if (e.wg == 0) {
            float z[16], out[10], x[16];
            tmem_ld16w(e.td(0), z);
            actor_head(z, a.actor, e0, e1, e2, scale, offset, out, aux);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = 0.f;
            float sd = 0.f;
            if (valid) {
                x[0] = __ldg(row); x[1] = __ldg(row + 1);
                float mean = 0.f;
#pragma unroll
                for (int c = 0; c < 10; ++c) { x[2 + c] = out[c]; mean += out[c]; }
                mean *= 0.1f;
                float var = 0.f;
#pragma unroll
                for (int c = 0; c < 10; ++c) { const float d = out[c] - mean; var = fmaf(d, d, var); }
                sd = sqrtf(var / 9.0f);                        // torch.std(action_out, dim=1): unbiased
#pragma unroll
                for (int c = 0; c < 3; ++c) x[12 + c] = (__ldg(row + 16 + c) - st[c]) / (st[3 + c] + 1e-6f);
            }
            x[15] = 1.0f;
            if (cr == 0) {      // the critic input of ACTOR_GRAD, as the K = 48 split image, to global only
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { hi[j] = __bfloat162float(__float2bfloat16(x[j])); lo[j] = x[j] - hi[j]; }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    e.img_store8(scr + SC::XCIMG, 48, g * 8, hi + g * 8);
                    e.img_store8(scr + SC::XCIMG, 48, 16 + g * 8, lo + g * 8);
                    e.img_store8(scr + SC::XCIMG, 48, 32 + g * 8, hi + g * 8);
                }
            }
            const float ssd = warp_sum(sd);
            if (e.lane == 0) e.cs()[(2 * 4 + e.wq) * 256 + 2] = ssd;
        }
        bar_epi();
        if (cr == 0 && e.tid == 0) {
            const float *c = e.cs() + 2 * 4 * 256;
            a.apart[(size_t)tile * AP::SIZE + AP::SSTD] = (c[2] + c[258]) + (c[514] + c[770]);
        }
        };
        // ---- target: a' = actor_target(s', e'), Q' = critic_target(s', a', 0) ----
        const float ne0 = __ldg(a.next_emotion), ne1 = __ldg(a.next_emotion + 1), ne2 = __ldg(a.next_emotion + 2);
        float tq = 0.f;
        if (chain != 2) {
        e.wait_afree();
        if (e.wg == 0) {
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = 0.f;
            if (valid) { x[0] = __ldg(row + 13); x[1] = __ldg(row + 14); x[2] = ne0; x[3] = ne1; x[4] = ne2; }
            x[15] = 1.0f;
            store_split_input(e, x, nullptr);
        }
        e.a_ready_local();
        UPROF(0, 2);
        e.wait_d(); UPROF(0, 3); epi_hidden128(e, nullptr, ACT_LEAKY, a.drop, 0, nullptr, nullptr, nullptr); e.a_ready(128); UPROF(0, 4);
        e.wait_d(); UPROF(0, 5); epi_hidden128(e, a.actor_t + AO::B2, ACT_RELU, a.drop, 1, nullptr, nullptr, nullptr); e.a_ready(128); UPROF(0, 6);
        e.wait_d(); UPROF(0, 7);
        epi_hidden64(e, a.actor_t + AO::B3, nullptr);
        e.a_ready(64);
        UPROF(0, 8);
        e.wait_d();
        UPROF(0, 9);
        e.wait_afree();
        if (e.wg == 0) {
            float z[16], out[10], x[16];
            tmem_ld16w(e.td(0), z);
            actor_head(z, a.actor_t, ne0, ne1, ne2, scale, offset, out, nullptr);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = 0.f;
            if (valid) {
                x[0] = __ldg(row + 13); x[1] = __ldg(row + 14);
#pragma unroll
                for (int c = 0; c < 10; ++c) x[2 + c] = out[c];
            }
            x[15] = 1.0f;                                       // the standardised emotion of identical rows is exactly 0
            store_split_input(e, x, nullptr);
        }
        e.a_ready_local();
        UPROF(0, 10);
        e.wait_d(); UPROF(0, 11); epi_ln256(e, nullptr, a.critic_t + CO::G1, a.critic_t + CO::BE1, ACT_LEAKY, a.drop, 2, 0, nullptr, nullptr); UPROF(0, 12);
        e.wait_d(); UPROF(0, 13); epi_ln256(e, a.critic_t + CO::B4, a.critic_t + CO::G2, a.critic_t + CO::BE2, ACT_RELU, a.drop, 3, 1, nullptr, nullptr); UPROF(0, 14);
        e.wait_d();
        UPROF(0, 15);
        { float h3[16]; tq = epi_head(e, a.critic_t, h3); }
        UPROF(0, 16);
        }
        if (chain == 1) {        // hand Q' of this tile to the pair that runs the rest of the phase
            if (cr == 0 && e.wg == 0 && valid) a.tq_g[grow] = tq;
            __threadfence();
            bar_epi();
            if (cr == 0 && e.tid < 2) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.tq_flag + 2 * tile + e.tid), "r"(1u) : "memory");
            if (!a.stats_ext) emotion_stats(e, a.batch, nullptr);      // without the fused sampler: this pair's own pass, after Q' is on its way
            actor_forward();
        } else {
        // ---- Q(s, a_norm, e) ----
        e.wait_afree();
        if (e.wg == 0) {
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = 0.f;
            if (valid) {
#pragma unroll
                for (int j = 0; j < 12; ++j) x[j] = __ldg(row + j);
#pragma unroll
                for (int c = 0; c < 3; ++c) x[12 + c] = (__ldg(row + 16 + c) - st[c]) / (st[3 + c] + 1e-6f);
            }
            x[15] = 1.0f;
            store_split_input(e, x, cr == 0 ? scr + SC::X0IMG : nullptr);
        }
        e.a_ready_local();
        UPROF(0, 17);
        e.wait_d();
        UPROF(0, 18);
        uint32_t sv1[2], sv2[2];                                // rstd / keep word of this thread's row and columns, LayerNorm 1 / 2
        epi_ln256(e, nullptr, a.critic + CO::G1, a.critic + CO::BE1, ACT_LEAKY, a.drop, 4, 0, sv1, scr + SC::H1IMG);
        UPROF(0, 19);
        e.wait_d();
        UPROF(0, 20);
        epi_ln256(e, a.critic + CO::B4, a.critic + CO::G2, a.critic + CO::BE2, ACT_RELU, a.drop, 5, 1, sv2, scr + SC::H2IMG);
        e.publish_saved();
        UPROF(0, 21);
        if (chain == 2) {
            if (e.tid == 0) {
                uint32_t f;
                const long long t0 = clock64();
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(a.tq_flag + 2 * tile + cr) : "memory");
                    if (clock64() - t0 > 4000000000LL) { e.stat()[8] = __int_as_float(0x7fc00000); break; }     // never a hang: poison instead
                } while (f == 0u);
                if (f != 0u) e.stat()[8] = 0.f;
            }
            bar_epi();
            tq = valid ? __ldcg(a.tq_g + grow) + e.stat()[8] : 0.f;
            bar_epi();
            if (e.tid == 0) a.tq_flag[2 * tile + cr] = 0u;
        }
        e.wait_d();
        UPROF(0, 22);
        {
            float h3[16], dz3[16], gw[16];
            const float q = epi_head(e, a.critic, h3);
            // SmoothL1 (beta = 1, mean) against y = 0.01 r + (1 - d) gamma Q'   (ddpg_agent.py:445, 470-473)
            float g = 0.f, l = 0.f;
            if (valid) {
                const float y = __ldg(row + 12) * 0.01f + (1.0f - __ldg(row + 15)) * a.gamma * tq;
                const float d = q - y, ad = fabsf(d);
                l = ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
                g = ad < 1.0f ? d : (d > 0.f ? 1.f : -1.f);
                if (fabsf(q) >= 1e6f) g = 0.f;                  // the clamp blocks the gradient outside (-1e6, 1e6)
                g /= (float)a.B;
            }
            const int cl = e.wg * 16, c0 = (int)cr * 64 + cl;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                dz3[j] = h3[j] > 0.f ? g * __ldg(a.critic + CO::W10 + c0 + j) : 0.f;
                gw[j] = g * h3[j];
            }
            e.wait_afree();
            e.img_store8_a(128, c0, dz3);
            e.img_store8_a(128, c0 + 8, dz3 + 8);
            e.a_ready(128);
            e.colsum_stage<16>(0, cl, dz3);                     // db8
            e.colsum_stage<16>(1, cl, gw);                      // dW10
            if (e.wg == 0) {                                    // db10 = sum g, loss = sum l over the tile's rows
                const float sg = warp_sum(g), sl = warp_sum(l);
                if (e.lane == 0) { e.cs()[(2 * 4 + e.wq) * 256] = sg; e.cs()[(2 * 4 + e.wq) * 256 + 1] = sl; }
            }
            bar_epi();
            e.colsum_finish(0, 64, part + CP::VC + cr * 64);
            e.colsum_finish(1, 64, part + CP::VC + 128 + cr * 64);
            if (e.tid == 0 && cr == 0) {
                const float *c = e.cs() + 2 * 4 * 256;
                part[CP::VC + 256] = (c[0] + c[256]) + (c[512] + c[768]);
                part[CP::LOSS] = (c[1] + c[257]) + (c[513] + c[769]);
            }
            bar_epi();
        }
        // ---- backward ----  (every weight-gradient block is drained right AFTER the next image has been handed over, while the next
        // input-gradient GEMM runs; the order keeps every drain ahead of the a_free its successor's GEMMs depend on)
        UPROF(0, 23);
        e.wait_d();
        UPROF(0, 25);
        epi_ln256_bwd(e, a.critic + CO::G2, a.critic + CO::BE2, ACT_RELU, a.drop, 1, sv2, part + CP::VB);
        UPROF(0, 26);
        drain_g(e, part + CP::W8, 128, cr * 128);
        UPROF(0, 24);
        drain_g(e, part + CP::W4 + cr * 128 * 256, 128, 0);
        UPROF(0, 27);
        e.wait_d();
        UPROF(0, 29);
        epi_ln256_bwd(e, a.critic + CO::G1, a.critic + CO::BE1, ACT_LEAKY, a.drop, 0, sv1, part + CP::VA);
        UPROF(0, 30);
        drain_g(e, part + CP::W4 + cr * 128 * 256, 128, 128);
        UPROF(0, 28);
        drain_g(e, part + CP::W0 + cr * 128 * 16, 16, 0);
        UPROF(0, 32);
        if (chain == 0) actor_forward();
        }
        UTC_EPI_DONE
        UCTA(0, 2);
    }
    UTC_ROLE_EPILOGUE
    UCTA(0, 3);
}

// =====================================================================================================================
// ACTOR_GRAD
// =====================================================================================================================
template <int ROLE>
__device__ __forceinline__ void actor_program(Pipe &p, const Args &a, const uint8_t *scr) {
    const int cr = (int)p.cr;
    // (a = actor(s, e) ran inside CRITIC_GRAD: actor_forward_program)
    gemm_fwd<ROLE>(p, a.img_critic + CI::W0, 256, 48);
    gemm_fwd<ROLE>(p, a.img_critic + CI::W4, 256, 256);
    gemm_fwd<ROLE>(p, a.img_critic + CI::W8, 128, 256);
    // input gradient of the critic
    gemm_bwd<ROLE>(p, a.img_critic + CI::W8, 128, 256, 256, true, true);
    gemm_bwd<ROLE>(p, a.img_critic + CI::W4, 256, 256, 256, true, true);
    gemm_bwd<ROLE>(p, a.img_critic + CI::W0, 256, 48, 16, false, true);
    // actor backward: every weight gradient is one block of output rows, the pair splits its input columns (dW1: 16 columns, both CTAs)
    load_h<ROLE>(p, scr + SC::A3IMG, 16384);
    gemm_bwd<ROLE>(p, a.img_actor + AI::W4, 16, 64, 64, true, false);
    wgrad<ROLE>(p, 16, 0, 64, cr * 32, 32, true, false, true, true);
    load_h<ROLE>(p, scr + SC::A2IMG, 32768);
    gemm_bwd<ROLE>(p, a.img_actor + AI::W3, 64, 128, 128, true, false);
    wgrad<ROLE>(p, 64, 0, 128, cr * 64, 64, true, false, true, true);
    load_h<ROLE>(p, scr + SC::A1IMG, 32768);
    gemm_bwd<ROLE>(p, a.img_actor + AI::W2, 128, 128, 128, true, false);
    wgrad<ROLE>(p, 128, 0, 128, cr * 64, 64, true, false, true, true);
    load_h<ROLE>(p, scr + SC::XAIMG, 12288);
    wgrad<ROLE>(p, 128, 0, 48, 0, 16, true, true, true, true);
}

__global__ void __launch_bounds__(NTHR, 1) ddpg_actor_grad_tc(const Args a) {
    UCTA(1, 0);
    UTC_ROLE_PROLOGUE
    UCTA(1, 1);
    const int tile = pair;
    uint8_t *scr = a.scratch + (size_t)tile * SC::BYTES;
    if (warp == W_PRODUCER) {
        if (elect_one()) { Pipe p{smem, bars, tbase, cr}; actor_program<ROLE_PRODUCER>(p, a, scr); }
        __syncwarp();
    } else if (warp == W_ISSUER) {
        if (elect_one()) { Pipe p{smem, bars, tbase, cr}; actor_program<ROLE_ISSUER>(p, a, scr); }
        __syncwarp();
    } else {
        UTC_MAKE_EPI
        e.row0 = tile * TR; e.B = a.B;
        const int grow = e.row0 + e.r;
        const bool valid = e.valid();
        const float *row = a.batch + (size_t)(valid ? grow : 0) * ERLFILL_REC;
        float *part = a.apart + (size_t)tile * AP::SIZE;
        // [row][32 words]: pos1 x8 | keep1 x8 | pos2 x8 | keep2 x8, word (4 cr + wg) = this thread's 16 columns
        uint32_t *abits = reinterpret_cast<uint32_t *>(scr + SC::ABITS) + e.r * 32 + (cr * 4 + e.wg);
        float *aux = reinterpret_cast<float *>(scr + SC::AUX) + (size_t)cr * 64 * TR + e.r;      // [64][128 rows], one copy per CTA
        { int t0 = 0; prefetch_actor_vectors(a.actor, e.tid, t0); prefetch_critic_vectors(a.critic, e.tid, t0); }
        // L2 norms of the nine actor tensors (actor_loss += 1e-4 sum ||theta_t||, ddpg_agent.py:490-494), spread over the CTAs
        for (int t = blockIdx.x; t < 9; t += gridDim.x) {
            float qn = 0.f;
            for (int i = c_actor_seg[t] + e.tid; i < c_actor_seg[t + 1]; i += NEPI) { const float w = __ldg(a.actor + i); qn = fmaf(w, w, qn); }
            qn = warp_sum(qn);
            float *scr_s = e.stat() + 16;
            if (e.lane == 0) scr_s[e.tid >> 5] = qn;
            bar_epi();
            if (e.tid == 0) { float tt = 0.f; for (int w = 0; w < NEPI / 32; ++w) tt += scr_s[w]; a.scal[6 + t] = sqrtf(tt); }
            bar_epi();
        }
        UPROF(1, 0);
        const int cid = (a.n_cond > 1 && valid) ? (int)__ldg(row + 19) : 0;
        const float *scale = a.scale + cid * 10;
        const float e0 = valid ? __ldg(row + 16) : 0.f, e1 = valid ? __ldg(row + 17) : 0.f, e2 = valid ? __ldg(row + 18) : 0.f;
        // ---- a = actor(s, e) ran inside CRITIC_GRAD (actor_forward_program): the critic input image (s, a, e) comes from the tile's scratch,
        // the saved operand images / activation bits / head values of the actor are read where the backward needs them ----
        e.publish_saved();                                      // nothing to publish in this kernel: opens the producers' first bulk load
        const uint32_t pos3 = reinterpret_cast<const uint32_t *>(scr + SC::POS3)[cr * NEPI + e.tid];
        e.wait_afree();
        for (int i = e.tid; i < 12288 / 16; i += NEPI)
            reinterpret_cast<uint4 *>(e.img_a())[i] = __ldcg(reinterpret_cast<const uint4 *>(scr + SC::XCIMG) + i);
        e.a_ready_local(); UPROF(1, 9);
        // ---- Q(s, a, e) with the updated critic ----
        e.wait_d(); UPROF(1, 10);
        uint32_t sv1[2], sv2[2];
        epi_ln256(e, nullptr, a.critic + CO::G1, a.critic + CO::BE1, ACT_LEAKY, a.drop, 8, 0, sv1, nullptr);
        UPROF(1, 11);
        e.wait_d(); UPROF(1, 12);
        epi_ln256(e, a.critic + CO::B4, a.critic + CO::G2, a.critic + CO::BE2, ACT_RELU, a.drop, 9, 1, sv2, nullptr);
        UPROF(1, 13);
        e.wait_d(); UPROF(1, 14);
        {
            float h3[16], dz3[16];
            const float q2 = epi_head(e, a.critic, h3);
            const float g = valid ? -1.0f / (float)a.B : 0.f;  // d(-mean Q)/dQ; the +-1e6 clamp is inactive
            const int c0 = (int)cr * 64 + e.wg * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) dz3[j] = h3[j] > 0.f ? g * __ldg(a.critic + CO::W10 + c0 + j) : 0.f;
            e.wait_afree();
            e.img_store8_a(128, c0, dz3);
            e.img_store8_a(128, c0 + 8, dz3 + 8);
            e.a_ready(128); UPROF(1, 15);
            if (e.wg == 0) {
                const float sq = warp_sum(valid ? q2 : 0.f);
                if (e.lane == 0) e.cs()[(2 * 4 + e.wq) * 256 + 3] = sq;
            }
            bar_epi();
            if (e.tid == 0 && cr == 0) {
                const float *c = e.cs() + 2 * 4 * 256;
                part[AP::SQ] = (c[3] + c[259]) + (c[515] + c[771]);      // (AP::SSTD was left by the actor forward in CRITIC_GRAD)
            }
            bar_epi();
        }
        e.wait_d(); UPROF(1, 16);
        epi_ln256_bwd(e, a.critic + CO::G2, a.critic + CO::BE2, ACT_RELU, a.drop, 1, sv2, nullptr);
        UPROF(1, 17);
        e.wait_d(); UPROF(1, 18);
        epi_ln256_bwd(e, a.critic + CO::G1, a.critic + CO::BE1, ACT_LEAKY, a.drop, 0, sv1, nullptr);
        UPROF(1, 19);
        e.wait_d(); UPROF(1, 20);
        // ---- actor backward ----
        e.wait_afree();
        if (e.wg == 0) {
            float g16[16], dz4[16], prod[32];
            tmem_ld16w(e.td(0), g16);                           // dQ/dx0: columns 2..11 are dQ/da
#pragma unroll
            for (int j = 0; j < 32; ++j) prod[j] = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) dz4[j] = 0.f;
            if (valid) {
                float mean = 0.f;
                float o10[10];
#pragma unroll
                for (int c = 0; c < 10; ++c) { o10[c] = aux[(40 + c) * TR]; mean += o10[c]; }
                mean *= 0.1f;
                float var = 0.f;
#pragma unroll
                for (int c = 0; c < 10; ++c) { const float d = o10[c] - mean; var = fmaf(d, d, var); }
                const float sd = sqrtf(var / 9.0f);
#pragma unroll
                for (int c = 0; c < 10; ++c) {
                    float go = g16[2 + c] + (-0.03f / (float)a.B) * (o10[c] - mean) / (9.0f * sd);     // d(-0.03 mean_rows std)/d out
                    const float gr_ = go * __ldg(scale + c);
                    const float z = aux[c * TR], base = aux[(10 + c) * TR], raw = aux[(20 + c) * TR], alpha = aux[(30 + c) * TR];
                    const float om = 1.0f - fabsf(base);
                    const float sf = base > 0.f ? -1.0f : (base < 0.f ? 1.0f : 0.0f);
                    const float dbase = gr_ * (1.0f + 2.0f * alpha * raw * om * (sf * sf));
                    const float den = 1.0f + fabsf(z);
                    dz4[c] = dbase / (den * den);
                    const float de = gr_ * alpha * sf * (om * om) * (1.0f - raw * raw) * 1.5f;
                    prod[c] = e0 * de; prod[10 + c] = e1 * de; prod[20 + c] = e2 * de;
                }
            }
            e.img_store8(e.img_a(), 16, 0, dz4);
            e.img_store8(e.img_a(), 16, 8, dz4 + 8);
            float db4[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) db4[j] = j < 16 ? dz4[j] : 0.f;
            e.colsum_stage<32>(0, 0, prod);                     // d emotion_weights [3][10]
            e.colsum_stage<32>(1, 0, db4);
        }
        e.a_ready_local(); UPROF(1, 21);
        bar_epi();
        if (cr == 0) {
            e.colsum_finish(0, 32, part + AP::EW);
            e.colsum_finish(1, 16, part + AP::B4);
        }
        bar_epi();
        // (each weight-gradient block is drained after the NEXT image has been handed over, while the next GEMM runs)
        e.wait_d(); UPROF(1, 23);
        {   // da3 = dz4 W4 (this CTA's 32 inputs, 8 per thread), relu'
            float v[8];
            const int cl = e.wg * 8, c0 = (int)cr * 32 + cl;
            tmem_ld8w(e.td(cl), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ((pos3 >> j) & 1u) ? v[j] : 0.f;
            e.wait_afree();
            e.img_store8_a(64, c0, v);
            e.a_ready(64); UPROF(1, 24);
            e.colsum_stage<8>(0, cl, v);
            bar_epi();
            e.colsum_finish(0, 32, part + AP::B3 + cr * 32);
            bar_epi();
        }
        drain_g(e, part + AP::W4, 32, cr * 32); UPROF(1, 22);
        e.wait_d(); UPROF(1, 26);
        {   // da2 = dz3 W3 (this CTA's 64 inputs, 16 per thread), dropout (site 7), relu'
            float v[16];
            const int cl = e.wg * 16, c0 = (int)cr * 64 + cl;
            tmem_ld16w(e.td(cl), v);
            const uint32_t pos = abits[16], keep = abits[24];
            const float ds = a.drop.mode == ERLFILL_DROPOUT_OFF ? 1.0f : kDropScale;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (((pos & keep) >> j) & 1u) ? v[j] * ds : 0.f;
            e.wait_afree();
            e.img_store8_a(128, c0, v);
            e.img_store8_a(128, c0 + 8, v + 8);
            e.a_ready(128); UPROF(1, 27);
            e.colsum_stage<16>(0, cl, v);
            bar_epi();
            e.colsum_finish(0, 64, part + AP::B2 + cr * 64);
            bar_epi();
        }
        drain_g(e, part + AP::W3, 64, cr * 64); UPROF(1, 25);
        e.wait_d(); UPROF(1, 29);
        {   // da1 = dz2 W2, dropout (site 6), leaky'
            float v[16];
            const int cl = e.wg * 16, c0 = (int)cr * 64 + cl;
            tmem_ld16w(e.td(cl), v);
            const uint32_t pos = abits[0], keep = abits[8];
            const float ds = a.drop.mode == ERLFILL_DROPOUT_OFF ? 1.0f : kDropScale;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * ds * (((pos >> j) & 1u) ? 1.0f : 0.1f) : 0.f;
            e.wait_afree();
            e.img_store8_a(128, c0, v);
            e.img_store8_a(128, c0 + 8, v + 8);
            e.a_ready(128); UPROF(1, 30);
            e.colsum_stage<16>(0, cl, v);
            bar_epi();
            e.colsum_finish(0, 64, part + AP::B1 + cr * 64);
            bar_epi();
        }
        drain_g(e, part + AP::W2, 64, cr * 64); UPROF(1, 28);
        drain_g(e, cr == 0 ? part + AP::W1 : nullptr, 16, 0); UPROF(1, 31);
        UTC_EPI_DONE
        UCTA(1, 2);
    }
    UTC_ROLE_EPILOGUE
    UCTA(1, 3);
}

}  // namespace utc
}  // namespace erlfill

// =====================================================================================================================
// partial reduction, optimiser, weight images, host side
// =====================================================================================================================
namespace erlfill {
namespace utc {

// position p of a per-CTA partial -> flat gradient index (or -1 for padding).  Matrix blocks are [n / 4][128 rows][4] (drain_g).
__device__ __forceinline__ void block_rc(int q, int &r, int &c) { r = (q & 511) >> 2; c = ((q >> 9) << 2) | (q & 3); }
__device__ __forceinline__ int critic_flat_index(int p) {
    int r, c;
    if (p < CP::VA) { block_rc(p & 2047, r, c); return c < 15 ? CO::W0 + ((p >> 11) * 128 + r) * 15 + c : -1; }
    if (p < CP::W4) return CO::B0 + (p - CP::VA);
    if (p < CP::VB) { const int q = p - CP::W4; block_rc(q & 32767, r, c); return CO::W4 + ((q >> 15) * 128 + r) * 256 + c; }
    if (p < CP::W8) return CO::B4 + (p - CP::VB);
    if (p < CP::VC) { block_rc(p - CP::W8, r, c); return CO::W8 + r * 256 + c; }
    return CO::B8 + (p - CP::VC);
}
__device__ __forceinline__ int actor_flat_index(int p) {
    int r, c;
    if (p < AP::W1) return p < 30 ? AO::EW + p : -1;
    if (p < AP::B1) { block_rc(p - AP::W1, r, c); return c < 5 ? AO::W1 + r * 5 + c : -1; }
    if (p < AP::W2) return AO::B1 + (p - AP::B1);
    if (p < AP::B2) { block_rc(p - AP::W2, r, c); return AO::W2 + r * 128 + c; }
    if (p < AP::W3) return AO::B2 + (p - AP::B2);
    if (p < AP::B3) { block_rc(p - AP::W3, r, c); return r < 64 ? AO::W3 + r * 128 + c : -1; }
    if (p < AP::W4) return AO::B3 + (p - AP::B3);
    if (p < AP::B4) { block_rc(p - AP::W4, r, c); return r < 10 ? AO::W4 + r * 64 + c : -1; }
    return p - AP::B4 < 10 ? AO::B4 + (p - AP::B4) : -1;
}

// G[i] = sum over the CTAs' partials in CTA order (+ the L2-norm regulariser's gradient 1e-4 theta / ||theta_tensor|| for the actor);
// the blocks walk the partial layout (coalesced reads), block b also leaves the sum of squares of its slice (clip_grad_norm_) and
// block 0 the loss value.
// Four threads share one group of four consecutive positions: thread k adds the 16-byte loads of the tiles k, k + 4, ... (in order), the
// four sums combine as (s0 + s1) + (s2 + s3) — a fixed order, no atomics — and thread k finishes component k of the group.
constexpr int kRedThreads = 256, kRedGroups = kRedThreads / 4;
constexpr int red_blocks(bool actor) { return ((actor ? AP::SQ : CP::LOSS) + 3) / 4 / kRedGroups + 1; }
constexpr int kRedBlocksMax = red_blocks(false);
template <bool kActor, bool kSignal>
__global__ void __launch_bounds__(kRedThreads) grad_reduce_kernel(int n_cta, int B, const float *part, const float *theta, float *scal, float *G, float *ssq,
                                                                  const PeerSig sig) {
    __shared__ float scratch[32];
    constexpr int np = kActor ? AP::SQ : CP::LOSS, psize = kActor ? AP::SIZE : CP::SIZE;
    constexpr int nq = (np + 3) / 4;
    const int k = threadIdx.x & 3, q = blockIdx.x * kRedGroups + (threadIdx.x >> 2);
    float ss = 0.f;
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
        const float4 *src = reinterpret_cast<const float4 *>(part) + q;
#pragma unroll 8
        for (int c = k; c < n_cta; c += 4) {
            const float4 v = __ldcg(src + (size_t)c * (psize / 4));
            s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
        }
    }
    float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sv[j] += __shfl_xor_sync(0xffffffffu, sv[j], 1);
        sv[j] += __shfl_xor_sync(0xffffffffu, sv[j], 2);
    }
    const int p = 4 * q + k;
    if (q < nq && p < np) {
        const int i = kActor ? actor_flat_index(p) : critic_flat_index(p);
        if (i >= 0) {
            float sum = k == 0 ? sv[0] : (k == 1 ? sv[1] : (k == 2 ? sv[2] : sv[3]));
            if (kActor) {
                int t = 0;
                while (i >= c_actor_seg[t + 1]) ++t;
                sum += 1e-4f * theta[i] / scal[6 + t];
            }
            G[i] = sum;
            ss = sum * sum;
        }
    }
    const float t = mlp::block_sum(ss, scratch);
    if (threadIdx.x == 0) ssq[blockIdx.x] = t;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (kActor) {
            float sq = 0.f, sd = 0.f;
            for (int c = 0; c < n_cta; ++c) { sq += part[(size_t)c * psize + AP::SQ]; sd += part[(size_t)c * psize + AP::SSTD]; }
            scal[5] = -sq / (float)B - 0.03f * sd / (float)B;
        } else {
            float l = 0.f;
            for (int c = 0; c < n_cta; ++c) l += part[(size_t)c * psize + CP::LOSS];
            scal[4] = l / (float)B;
        }
    }
    if (kSignal) peer_signal_when_grid_done(sig);      // data-parallel: this rank's bucket is complete -> raise its flags on every peer
}

// The clip-norm partials of a gradient bucket that may have changed since grad_reduce_kernel wrote it (an APPLY phase called on its own:
// the caller averaged the bucket over the ranks in between, e.g. ncclAllReduce).  Same thread <-> element mapping and block sum as
// grad_reduce_kernel, so an unchanged bucket gives bit-identical partials.
template <bool kActor>
__global__ void __launch_bounds__(kRedThreads) grad_ssq_kernel(const float *G, float *ssq) {
    __shared__ float scratch[32];
    constexpr int np = kActor ? AP::SQ : CP::LOSS;
    const int p = 4 * (blockIdx.x * kRedGroups + (threadIdx.x >> 2)) + (threadIdx.x & 3);
    float ss = 0.f;
    if (p < np) {
        const int i = kActor ? actor_flat_index(p) : critic_flat_index(p);
        if (i >= 0) { const float g = G[i]; ss = g * g; }
    }
    const float t = mlp::block_sum(ss, scratch);
    if (threadIdx.x == 0) ssq[blockIdx.x] = t;
}

// ---- weight images -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_bf16(uint8_t *img, uint32_t off, float v) { *reinterpret_cast<__nv_bfloat16 *>(img + off) = __float2bfloat16(v); }
// first-layer weight (row n, input k < 16) -> [W_hi | W_hi | W_lo] at columns k, 16 + k, 32 + k
__device__ __forceinline__ void put_split(uint8_t *img, int n, int k, int rows, float v) {
    const float hi = __bfloat162float(__float2bfloat16(v));
    put_bf16(img, wimg_off(n, k, rows, 48), hi);
    put_bf16(img, wimg_off(n, 16 + k, rows, 48), hi);
    put_bf16(img, wimg_off(n, 32 + k, rows, 48), v - hi);
}
__device__ __forceinline__ void critic_image_store(uint8_t *img, int i, float v) {
    if (i < CO::B0) put_split(img + CI::W0, i / 15, i % 15, 256, v);
    else if (i < CO::G1) put_split(img + CI::W0, i - CO::B0, 15, 256, v);                 // bias: column 15 against the column of ones
    else if (i >= CO::W4 && i < CO::B4) put_bf16(img + CI::W4, wimg_off((i - CO::W4) >> 8, (i - CO::W4) & 255, 256, 256), v);
    else if (i >= CO::W8 && i < CO::B8) put_bf16(img + CI::W8, wimg_off((i - CO::W8) >> 8, (i - CO::W8) & 255, 128, 256), v);
}
__device__ __forceinline__ void actor_image_store(uint8_t *img, int i, float v) {
    if (i < AO::W1) return;
    if (i < AO::B1) put_split(img + AI::W1, (i - AO::W1) / 5, (i - AO::W1) % 5, 128, v);
    else if (i < AO::W2) put_split(img + AI::W1, i - AO::B1, 15, 128, v);
    else if (i < AO::B2) put_bf16(img + AI::W2, wimg_off((i - AO::W2) >> 7, (i - AO::W2) & 127, 128, 128), v);
    else if (i >= AO::W3 && i < AO::B3) put_bf16(img + AI::W3, wimg_off((i - AO::W3) >> 7, (i - AO::W3) & 127, 64, 128), v);
    else if (i >= AO::W4 && i < AO::B4) put_bf16(img + AI::W4, wimg_off((i - AO::W4) >> 6, (i - AO::W4) & 63, 16, 64), v);
}
__global__ void pack_images_kernel(const float *actor, const float *actor_t, const float *critic, const float *critic_t, uint8_t *img_actor,
                                   uint8_t *img_actor_t, uint8_t *img_critic, uint8_t *img_critic_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < CO::TOTAL) { critic_image_store(img_critic, i, critic[i]); critic_image_store(img_critic_t, i, critic_t[i]); }
    if (i < AO::TOTAL) { actor_image_store(img_actor, i, actor[i]); actor_image_store(img_actor_t, i, actor_t[i]); }
}

// ---- optimiser ---------------------------------------------------------------------------------------------------------
// step[0..1] = (double) lr_decay_rate ** train_step, step[2] = 1 - 0.9^t, step[3] = sqrt(1 - 0.999^t)   (update.cu: set_step_kernel)
__device__ __forceinline__ float adam_elem(float &p, float &m, float &v, float g, float coef, float lr, float bc1, float bc2_sqrt) {
    const float gi = g * coef;
    const float mi = m + (gi - m) * 0.1f;
    const float vi = v * 0.999f + 0.001f * gi * gi;
    m = mi; v = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + 1e-8f;
    p = p - (lr / bc1) * (mi / denom);
    return p;
}
// n_ssq sum-of-squares partials: red_blocks() from grad_reduce_kernel, mlp::kSumsqBlocks from the peer exchange (REDUCED).  Warp 0 of the
// block adds them (lane l takes partials l, l + 32, ..., then a butterfly: the same order in every block and on every rank) and
// every thread reads the coefficient from shared memory; all threads of the block must call this.
__device__ __forceinline__ float clip_coef(const float *ssq, int n_ssq, float max_norm) {
    __shared__ float s_coef;
    if (threadIdx.x < 32) {
        float ss = 0.f;
        for (int b = threadIdx.x; b < n_ssq; b += 32) ss += ssq[b];
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (threadIdx.x == 0) s_coef = fminf(max_norm / (sqrtf(ss) + 1e-6f), 1.0f);
    }
    __syncthreads();
    return s_coef;
}
// CRITIC_APPLY: emotion-driven learning rates (ddpg_agent.py:447-462), clip_grad_norm_(5), Adam, image refresh
__global__ void critic_apply_kernel(float *p, float *m, float *v, const float *G, const float *ssq, int n_ssq, const float *step, double critic_lr,
                                    double actor_lr, float max_norm, float *scal, uint8_t *img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double decay = *reinterpret_cast<const double *>(step);
    const double cur = G[CO::TOTAL], cons = G[CO::TOTAL + 1], anx = G[CO::TOTAL + 2];
    const double af = 1.0 + 0.8 * cur - 0.4 * cons + 0.2 * anx, cf = 1.0 - 0.3 * anx + 0.6 * cons;
    const float lr_c = (float)fmin(fmax(critic_lr * cf * decay, critic_lr * 0.1), critic_lr * 3);
    if (i == 0) { scal[2] = lr_c; scal[3] = (float)fmin(fmax(actor_lr * af * decay, actor_lr * 0.1), actor_lr * 3); }
    const float coef = clip_coef(ssq, n_ssq, max_norm);
    if (i >= CO::TOTAL) return;
    float pi = p[i], mi = m[i], vi = v[i];
    adam_elem(pi, mi, vi, G[i], coef, lr_c, step[2], step[3]);
    p[i] = pi; m[i] = mi; v[i] = vi;
    critic_image_store(img, i, pi);
}
// ACTOR_APPLY: clip, Adam, both soft target updates (ddpg_agent.py:501-507), all remaining images, report
__global__ void actor_apply_kernel(float *p, float *pt, float *m, float *v, const float *G, const float *ssq, int n_ssq, const float *step, float max_norm,
                                   float tau, const float *critic, float *critic_t, float *scal, uint8_t *img_actor, uint8_t *img_actor_t,
                                   uint8_t *img_critic_t, float *report) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float coef = clip_coef(ssq, n_ssq, max_norm);
    if (i < AO::TOTAL) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_elem(pi, mi, vi, G[i], coef, scal[3], step[2], step[3]);
        p[i] = pi; m[i] = mi; v[i] = vi;
        const float ti = tau * pi + (1.0f - tau) * pt[i];
        pt[i] = ti;
        actor_image_store(img_actor, i, pi);
        actor_image_store(img_actor_t, i, ti);
    }
    if (i < CO::TOTAL) {
        const float ti = tau * critic[i] + (1.0f - tau) * critic_t[i];
        critic_t[i] = ti;
        critic_image_store(img_critic_t, i, ti);
    }
    if (i == 0 && report) {
        float l2 = 0.f;
        for (int t = 0; t < 9; ++t) l2 += scal[6 + t];
        report[0] = scal[4]; report[1] = scal[5] + 1e-4f * l2; report[2] = scal[2]; report[3] = scal[3];
    }
}

// ---- fast-path workspace tail (behind the exact path's workspace, so buckets / exchange buffers keep their addresses) -------------
struct Tail {
    uint8_t *img_actor, *img_actor_t, *img_critic, *img_critic_t, *scratch;
    float *cpart, *apart, *tq_g, *stats_part, *ssq_fine;      // ssq_fine: [2][kRedBlocksMax]
    uint32_t *flags;            // two tq_flags per tile, the two counters of grid_emotion_stats, the two arrival counters of the peer signal
};
static size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }
static int n_cta(int B) { return (B + TR - 1) / TR; }
size_t tail_bytes(int B) {
    const size_t c = (size_t)n_cta(B);
    return up256((2 * c + 4) * 4) + up256(c * 32) + 2 * up256(AI::BYTES) + 2 * up256(CI::BYTES) + up256(c * SC::BYTES) + up256(c * CP::SIZE * 4) + up256(c * AP::SIZE * 4) +
           up256((size_t)B * 4) + up256(2 * kRedBlocksMax * 4);
}
static void carve_tail(Tail &t, uint8_t *base, int B) {
    const size_t c = (size_t)n_cta(B);
    size_t o = 0;
    t.flags = reinterpret_cast<uint32_t *>(base); o += up256((2 * c + 4) * 4);
    t.stats_part = reinterpret_cast<float *>(base + o); o += up256(c * 32);
    t.img_actor = base + o; o += up256(AI::BYTES);
    t.img_actor_t = base + o; o += up256(AI::BYTES);
    t.img_critic = base + o; o += up256(CI::BYTES);
    t.img_critic_t = base + o; o += up256(CI::BYTES);
    t.scratch = base + o; o += up256(c * SC::BYTES);
    t.cpart = reinterpret_cast<float *>(base + o); o += up256(c * CP::SIZE * 4);
    t.apart = reinterpret_cast<float *>(base + o); o += up256(c * AP::SIZE * 4);
    t.tq_g = reinterpret_cast<float *>(base + o); o += up256((size_t)B * 4);
    t.ssq_fine = reinterpret_cast<float *>(base + o);
}

// Both phase kernels run as clusters of two CTAs (one pair per 128-row tile).
static void pair_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute &attr, int n_ctas, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
}
static cudaError_t launch_pairs(void (*kernel)(const Args), int n_ctas, const Args &a, cudaStream_t st) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr;
    pair_config(cfg, attr, n_ctas, st);
    return cudaLaunchKernelEx(&cfg, kernel, a);
}
// shared-memory opt-in; *max_pairs = how many pairs of ddpg_critic_grad_tc the device keeps resident at once (0 if the query fails)
static cudaError_t pair_setup(int *max_pairs) {
    static bool done = false;
    static int pairs = 0;
    if (!done) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(ddpg_critic_grad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(ddpg_actor_grad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)) != cudaSuccess) return e;
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr;
        pair_config(cfg, attr, 2, nullptr);
        if (cudaOccupancyMaxActiveClusters(&pairs, ddpg_critic_grad_tc, &cfg) != cudaSuccess) { pairs = 0; (void)cudaGetLastError(); }
        done = true;
    }
    *max_pairs = pairs;
    return cudaSuccess;
}

int pack_images(const erlfill_ddpg_nets *n, uint8_t *tail_base, int B, cudaStream_t st) {
    Tail t;
    carve_tail(t, tail_base, B);
    ERLFILL_CUDA_TRY(cudaMemsetAsync(t.flags, 0, up256((2 * (size_t)n_cta(B) + 4) * 4), st));
    ERLFILL_CUDA_TRY(cudaMemsetAsync(t.img_actor, 0, 2 * up256(AI::BYTES) + 2 * up256(CI::BYTES), st));
    pack_images_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(n->actor, n->actor_target, n->critic, n->critic_target, t.img_actor, t.img_actor_t,
                                                                t.img_critic, t.img_critic_t);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

// what peer.cu needs of the tail (peer.cuh): the fast exchange's norm partials and the signal's arrival counters
float *tail_ssq_fine(uint8_t *tail_base, int B, int which) {
    Tail t;
    carve_tail(t, tail_base, B);
    return t.ssq_fine + which * kRedBlocksMax;
}
unsigned int *tail_peer_done(uint8_t *tail_base, int B, int which) {
    Tail t;
    carve_tail(t, tail_base, B);
    return t.flags + 2 * n_cta(B) + 2 + which;
}
constexpr int fine_blocks(int which) { return ((which == 0 ? CO::TOTAL + 3 : AO::TOTAL) + kPeerFineThreads - 1) / kPeerFineThreads; }
static_assert(fine_blocks(0) <= kRedBlocksMax && fine_blocks(1) <= kRedBlocksMax, "ssq_fine holds the fast exchange's norm partials");
int peer_fine_blocks(int which) { return fine_blocks(which); }

// the workspace head the two paths share (update.cu)
struct Head { float *cgrad, *agrad, *cred, *ared, *stats, *scal, *ssq, *step; };

int update_fast(const erlfill_ddpg_nets *n, const erlfill_ddpg_hyper *h, const float *d_batch, const float *d_next_emotion, const Head &w,
                uint8_t *tail_base, int phases, float *d_report, cudaStream_t st) {
    const int B = h->batch, nc = n_cta(B);
    int max_pairs = 0;
    ERLFILL_CUDA_TRY(pair_setup(&max_pairs));
    Tail t;
    carve_tail(t, tail_base, B);
    Args a;
    a.B = B; a.n_cond = h->n_cond; a.gamma = h->gamma;
    a.batch = d_batch; a.next_emotion = d_next_emotion;
    a.actor = n->actor; a.actor_t = n->actor_target; a.critic = n->critic; a.critic_t = n->critic_target;
    a.img_actor = t.img_actor; a.img_actor_t = t.img_actor_t; a.img_critic = t.img_critic; a.img_critic_t = t.img_critic_t;
    a.scale = n->scale; a.offset = n->offset;
    a.scratch = t.scratch; a.cpart = t.cpart; a.apart = t.apart;
    a.mean_out = w.cgrad + CO::TOTAL;
    a.scal = w.scal;
    a.drop.mode = h->dropout_mode;
    a.drop.k0 = (uint32_t)h->seed; a.drop.k1 = (uint32_t)(h->seed >> 32);
    a.drop.update_idx = h->update_idx; a.drop.row_base = h->rank << 20;
    {
        static int prof_env = -1;
        if (prof_env < 0) { const char *pe = getenv("ERLFILL_UTC_PROF"); prof_env = (pe && pe[0] == '1') ? 1 : 0; }
        a.prof = prof_env;
    }
    for (int s = 0; s < 10; ++s) a.drop.m[s] = (h->dropout_mode == ERLFILL_DROPOUT_MASK && h->masks) ? h->masks->keep[s] : nullptr;
    if (h->dropout_mode == ERLFILL_DROPOUT_MASK) {
        if (!h->masks) return ERLFILL_ERR_ARG;
        for (int s = 0; s < 10; ++s) if (!h->masks->keep[s]) return ERLFILL_ERR_ARG;
    }
    a.tq_g = t.tq_g; a.tq_flag = t.flags;
    a.stats_g = w.stats; a.stats_part = t.stats_part; a.stats_ctr = t.flags + 2 * nc;
    a.stats_ext = h->n_stats_partials > 0 ? h->d_stats_partials : nullptr; a.n_stats_ext = h->n_stats_partials;
    a.split = (2 * nc <= max_pairs) ? 1 : 0;  // the pairs of both chains of every tile must be resident at once (one CTA per SM)
    const bool reduced = (phases & ERLFILL_PHASE_REDUCED) != 0;
    // REDUCED: the partials of erlfill_ddpg_peer_allreduce_fast (hyper->peer set) or of erlfill_ddpg_peer_allreduce (64 slices)
    const bool fine = reduced && h->peer != nullptr;
    PeerSig sig{};                                           // world == 0: no exchange
    if (h->peer) {
        const erlfill_peer_group *g = (const erlfill_peer_group *)h->peer;
        if (g->world < 1 || g->world > ERLFILL_MAX_PEERS || g->rank < 0 || g->rank >= g->world) return ERLFILL_ERR_ARG;
        sig.world = g->world; sig.rank = g->rank;
        for (int p = 0; p < g->world; ++p) {
            if (!g->ctrl[p]) return ERLFILL_ERR_ARG;
            sig.ctrl[p] = (PeerCtrl *)g->ctrl[p];
        }
    }
    if (phases & ERLFILL_PHASE_CRITIC_GRAD) {
        ERLFILL_CUDA_TRY(launch_pairs(ddpg_critic_grad_tc, 2 * (a.split ? 2 * nc : nc), a, st));
        sig.which = 0; sig.done = t.flags + 2 * nc + 2;
        if (sig.world > 0) grad_reduce_kernel<false, true><<<red_blocks(false), kRedThreads, 0, st>>>(nc, B, t.cpart, nullptr, w.scal, w.cgrad, t.ssq_fine, sig);
        else grad_reduce_kernel<false, false><<<red_blocks(false), kRedThreads, 0, st>>>(nc, B, t.cpart, nullptr, w.scal, w.cgrad, t.ssq_fine, sig);
    }
    if (phases & ERLFILL_PHASE_CRITIC_APPLY) {
        if (!reduced && !(phases & ERLFILL_PHASE_CRITIC_GRAD))
            grad_ssq_kernel<false><<<red_blocks(false), kRedThreads, 0, st>>>(w.cgrad, t.ssq_fine);
        const float *G = reduced ? w.cred : w.cgrad;
        critic_apply_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(n->critic, n->critic_m, n->critic_v, G, reduced && !fine ? w.ssq : t.ssq_fine,
                                                                     fine ? fine_blocks(0) : (reduced ? mlp::kSumsqBlocks : red_blocks(false)), w.step, h->critic_lr, h->actor_lr,
                                                                     h->max_grad_norm, w.scal, t.img_critic);
    }
    if (phases & ERLFILL_PHASE_ACTOR_GRAD) {
        ERLFILL_CUDA_TRY(launch_pairs(ddpg_actor_grad_tc, 2 * nc, a, st));
        sig.which = 1; sig.done = t.flags + 2 * nc + 3;
        if (sig.world > 0) grad_reduce_kernel<true, true><<<red_blocks(true), kRedThreads, 0, st>>>(nc, B, t.apart, n->actor, w.scal, w.agrad, t.ssq_fine + kRedBlocksMax, sig);
        else grad_reduce_kernel<true, false><<<red_blocks(true), kRedThreads, 0, st>>>(nc, B, t.apart, n->actor, w.scal, w.agrad, t.ssq_fine + kRedBlocksMax, sig);
    }
    if (phases & ERLFILL_PHASE_ACTOR_APPLY) {
        if (!reduced && !(phases & ERLFILL_PHASE_ACTOR_GRAD))
            grad_ssq_kernel<true><<<red_blocks(true), kRedThreads, 0, st>>>(w.agrad, t.ssq_fine + kRedBlocksMax);
        const float *G = reduced ? w.ared : w.agrad;
        actor_apply_kernel<<<(CO::TOTAL + 255) / 256, 256, 0, st>>>(n->actor, n->actor_target, n->actor_m, n->actor_v, G,
                                                                    reduced && !fine ? w.ssq + mlp::kSumsqBlocks : t.ssq_fine + kRedBlocksMax,
                                                                    fine ? fine_blocks(1) : (reduced ? mlp::kSumsqBlocks : red_blocks(true)), w.step, h->max_grad_norm, h->tau, n->critic, n->critic_target, w.scal, t.img_actor,
                                                                    t.img_actor_t, t.img_critic_t, d_report);
    }
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // namespace utc
}  // namespace erlfill

// development only (not declared in include/erlfill.h): copies the in-kernel profile of CTA 0 (ERLFILL_UTC_PROF=1)
extern "C" int erlfill_debug_utc_cta(long long *out /*[2][256][4]*/) {
    return cudaMemcpyFromSymbol(out, erlfill::utc::g_utc_cta, sizeof(long long) * 2048) == cudaSuccess ? 0 : -2;
}
extern "C" int erlfill_debug_utc_profile(long long *out /*[2][128]*/) {
    return cudaMemcpyFromSymbol(out, erlfill::utc::g_utc_prof, sizeof(long long) * 256) == cudaSuccess ? 0 : -2;
}
