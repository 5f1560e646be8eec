// update_tc.cuh — building blocks of the tcgen05 ("fast") ER-DDPG update (update_tc.cu).
//
// A CLUSTER OF TWO CTAs owns TR = 128 minibatch rows = the 128 lanes of tensor memory of each of its two SMs.  Every dense layer of
// every forward / input-gradient / weight-gradient pass of DDPGAgent.learn (ddpg_agent.py:422-512) is a tcgen05.mma kind::f16 (bf16
// operands, fp32 accumulation in TMEM) with M = 128; the layer chain of a whole phase runs inside one kernel.  The pair splits the
// OUTPUT COLUMNS of every layer: CTA c (cluster rank) streams the weight rows / columns of its half, runs the MMAs for it and its
// epilogue threads write their results into the operand image of BOTH CTAs over distributed shared memory, so that either SM holds
// the complete activation (or gradient) image the next layer consumes.  The chain is bound by the per-thread epilogue work and its
// latencies (profiles/r02e_critic_grad_tc.txt: issue slots 26 % busy, tensor pipe 7 %), so two SMs per tile halve it.
//
//   roles      16 epilogue warps per CTA (4 warpgroups; warpgroup g of CTA c owns the eighth [(4c + g) N/8, +N/8) of the columns of
//              every row, thread r of a warpgroup owns row r = TMEM lane r), 1 MMA-issuer warp, 1 bulk-copy producer warp.
//   pair sync  a_ready: the epilogue threads write their half of the image into their OWN shared memory, fence it for the async proxy
//              and meet at a named barrier; 16 threads then push the half to the peer with cp.async.bulk shared::cta ->
//              shared::cluster (one 8-row group each), counted in bytes on the peer's a_ready barrier, which the peer's issuer waits on:
//              no remote generic stores, no cluster-scope fences on the epilogue's critical path.  a_free: each issuer multicasts
//              a tcgen05.commit to both CTAs after its last MMA that reads the current image, and the epilogues wait for both before
//              they overwrite it.  Row statistics (LayerNorm, head) are exchanged as 8 partials per row with st.async (8 bytes per
//              thread, counted on the peer's mbarrier).  Saved global images are published once per kernel (h_pub) before the
//              producers bulk-load them (tools/umma_pair_probe.cu, profiles/r02f_umma_pair_probe.txt).
//   operands   activations / gradients are written by the row owners as bf16 "images" in the canonical no-swizzle K-major
//              core-matrix layout (8 rows x 16 bytes) into shared memory (IMG_A).  The SAME image is the K-major A operand of the
//              next layer and the MN-major operand of the weight-gradient GEMM dW = dZ^T X (K = rows); the SAME weight image
//              W[out][in] is the K-major B operand of the forward and the MN-major B operand of dX = dZ W — no transposed copies
//              (MN-major no-swizzle descriptors: tools/umma_mn_probe.cu, profiles/r02a_umma_mn_probe.txt).
//   weights    bf16 images of the fp32 master parameters, chunk-major over K (64 input features per 32 KB chunk), streamed by
//              cp.async.bulk through a 2-stage ring; written by the pack kernel and kept current by the Adam kernel.
//   first layers (K = 15 / 5, physical-unit actions up to 25,000 in the critic input): hi/lo bf16 split, A = [x_hi | x_lo | x_hi],
//              B = [W_hi | W_hi | W_lo] (K = 48), bias in column 15 against a column of ones: ~2^-16 relative instead of 2^-9.
//   weight gradients accumulate in TMEM columns [256, 512) and are drained as per-tile fp32 partials (each CTA of the pair its own
//              output-row block or column half); a second kernel adds the partials in tile order (deterministic, no atomics) and
//              leaves the clip-norm partials.
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "ddpg_dropout.cuh"
#include "umma.cuh"

namespace erlfill {
namespace utc {
using namespace umma;

constexpr int TR = 128;                    // rows per CTA
constexpr int NWG = 4;                     // epilogue warpgroups
constexpr int NEPI = NWG * 128;            // epilogue threads
constexpr int W_ISSUER = NEPI / 32, W_PRODUCER = NEPI / 32 + 1;
constexpr int NTHR = NEPI + 64;
constexpr int KCH = 64;                    // input features per weight chunk
constexpr int STAGE_BYTES = 256 * KCH * 2; // 32 KB: the largest chunk (256 output rows)
constexpr int NSTAGE = 2;

// ---- shared memory map ---------------------------------------------------------------------------------------------
constexpr int SM_IMG_A = 0;                            // 64 KB: current operand image (<= 256 features)
constexpr int SM_IMG_H = SM_IMG_A + TR * 256 * 2;      // 64 KB: saved forward activation reloaded for the weight-gradient GEMM
constexpr int SM_RING = SM_IMG_H + TR * 256 * 2;       // 2 x 32 KB weight stages
constexpr int SM_RED = SM_RING + NSTAGE * STAGE_BYTES; // float red[2 buffers][2 CTAs][NWG][TR][2]
constexpr int SM_CS = SM_RED + 2 * 2 * NWG * TR * 2 * 4; // float cs[3][4 warps][256]: column-sum staging
constexpr int SM_ROW = SM_CS + 3 * 4 * 256 * 4;        // float rowv[4][TR]: per-row scalars handed between layers (tq, dq, ...)
constexpr int SM_STAT = SM_ROW + 4 * TR * 4;           // float stat[8]: batch emotion statistics; scratch[64]
constexpr int SM_BAR = SM_STAT + 128 * 4;
constexpr int NBAR = 2 * NSTAGE + 10;                  // w_full[2] w_empty[2] h_full h_free d_full g_full g_free a_ready a_free xred[2] h_pub
constexpr int SM_TMEM = SM_BAR + NBAR * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16;
static_assert(SMEM_BYTES <= 232448, "shared memory plan exceeds the 227 KB opt-in limit");

enum { BW_FULL = 0, BW_EMPTY = NSTAGE, B_HFULL = 2 * NSTAGE, B_HFREE, B_DFULL, B_GFULL, B_GFREE, B_AREADY, B_AFREE, B_XRED, B_XRED1, B_HPUB };
constexpr uint16_t PAIR_MASK = 3;

// TMEM columns: main accumulators (this CTA's <= 128 output columns) | xhat of LayerNorm 1 | weight-gradient accumulators (<= 128
// columns per block) | xhat of LayerNorm 2   (xhat: what the LayerNorm backward needs, kept on chip from the forward)
constexpr uint32_t TM_D = 0, TM_X1 = 128, TM_G = 256, TM_X2 = 384;

// ---- image geometry ------------------------------------------------------------------------------------------------
// activation image of F features: element (row r, feature f)
__host__ __device__ constexpr uint32_t img_off(int r, int f, int F) {
    return (uint32_t)((r >> 3) * (F >> 3) * 128 + (f >> 3) * 128 + (r & 7) * 16 + (f & 7) * 2);
}
// weight image of a [rows][K] matrix, chunk-major over K in chunks of kc = min(K, 64) features
__host__ __device__ constexpr uint32_t wimg_off(int n, int k, int rows, int K) {
    const int kc = K < KCH ? K : KCH;
    return (uint32_t)((k / kc) * rows * kc * 2 + (n >> 3) * (kc >> 3) * 128 + ((k % kc) >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}
__host__ __device__ constexpr uint32_t wimg_bytes(int rows, int K) { return (uint32_t)(rows * K * 2); }

// ---- weight image blobs ------------------------------------------------------------------------------------------
// critic: W0 [256][48] | W4 [256][256] | W8 [128][256]        actor: W1 [128][48] | W2 [128][128] | W3 [64][128] | W4 [16][64]
namespace CI { constexpr uint32_t W0 = 0, W4 = W0 + 256 * 48 * 2, W8 = W4 + 256 * 256 * 2, BYTES = W8 + 128 * 256 * 2; }
namespace AI { constexpr uint32_t W1 = 0, W2 = W1 + 128 * 48 * 2, W3 = W2 + 128 * 128 * 2, W4 = W3 + 64 * 128 * 2, BYTES = W4 + 16 * 64 * 2; }

// flat fp32 layouts (parameters() order), as in update.cu
namespace AO { constexpr int EW = 0, W1 = 30, B1 = 670, W2 = 798, B2 = 17182, W3 = 17310, B3 = 25502, W4 = 25566, B4 = 26206, TOTAL = 26216; }
namespace CO { constexpr int W0 = 0, B0 = 3840, G1 = 4096, BE1 = 4352, W4 = 4608, B4 = 70144, G2 = 70400, BE2 = 70656, W8 = 70912, B8 = 103680, W10 = 103808, B10 = 103936, TOTAL = 103937; }

// per-tile gradient partials (floats).  A matrix is a sequence of 128-output-row blocks, each stored [n / 4 float4 columns][128 rows][4]
// (the order the weight-gradient accumulators are drained in: update_tc.cu, drain_g).
// critic: W0 2 blocks x 16 columns | db0 dgamma1 dbeta1 | W4 2 blocks x 256 | db4 dgamma2 dbeta2 | W8 1 block x 256 | db8 dW10 db10 | loss
namespace CP { constexpr int W0 = 0, VA = 4096, W4 = VA + 768, VB = W4 + 65536, W8 = VB + 768, VC = W8 + 32768, LOSS = VC + 257,
               SIZE = ((LOSS + 1 + 3) / 4) * 4; }
// actor: EW(32) | W1 block x 16 | B1 | W2 block x 128 | B2 | W3 block x 128 (64 rows used) | B3 | W4 block x 64 (10 rows used) | B4(16) | sum q2 | sum std
namespace AP { constexpr int EW = 0, W1 = 32, B1 = W1 + 128 * 16, W2 = B1 + 128, B2 = W2 + 128 * 128, W3 = B2 + 128, B3 = W3 + 128 * 128,
               W4 = B3 + 64, B4 = W4 + 128 * 64, SQ = B4 + 16, SSTD = SQ + 1, SIZE = ((SSTD + 1 + 3) / 4) * 4; }
static_assert(CP::VA == 2 * 2048 && CP::VB - CP::W4 == 2 * 32768 && CP::VC - CP::W8 == 32768, "critic partial blocks");

// ---- small helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }

__device__ __forceinline__ void tmem_ld32w(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld16w(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    tmem_ld16(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_ld8w(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    tmem_ld8(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// sum over the 32 lanes of a warp of v[c], for W = 32 / 16 / 8 columns at once (W - 1 + log2(32 / W) shuffles): the butterfly halves
// the columns a lane holds while W > 1, then adds the rest; every lane returns the total of column lane >> log2(32 / W)
template <int W>
__device__ __forceinline__ float warp_colsum(float (&v)[W], int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int off = 16 >> s, wb = W >> s;            // columns still held before this step
        if (wb >= 2) {
            const int half = wb / 2;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float send = upper ? v[j] : v[j + half];
                const float keep = upper ? v[j + half] : v[j];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
    return v[0];
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

using ddrop::Drop; using ddrop::keep32; using ddrop::kDropScale;

// ---- role context ----------------------------------------------------------------------------------------------------
struct Pipe {
    uint8_t *smem;
    uint64_t *bars;
    uint32_t tbase;
    uint32_t cr;                               // rank of this CTA in its pair
    // producer / issuer ring position, phases of the single-use barriers
    uint32_t ring_n = 0, h_n = 0, d_n = 0, g_n = 0, gfree_n = 0, a_n = 0;
    __device__ uint64_t *bar(int i) const { return bars + i; }
};

enum { ROLE_PRODUCER = 0, ROLE_ISSUER = 1 };

// One weight chunk through the ring: producer side (wait for the stage, start the bulk copy) / issuer side (wait for the bytes).
__device__ __forceinline__ uint32_t ring_produce(Pipe &p, const uint8_t *src, uint32_t bytes) {
    const uint32_t stage = p.ring_n % NSTAGE, use = p.ring_n / NSTAGE;
    mbar_wait(p.bar(BW_EMPTY + stage), (use & 1u) ^ 1u);
    mbar_arrive_expect_tx(p.bar(BW_FULL + stage), bytes);
    bulk_g2s(p.smem + SM_RING + stage * STAGE_BYTES, src, bytes, p.bar(BW_FULL + stage));
    ++p.ring_n;
    return stage;
}
__device__ __forceinline__ uint32_t ring_consume_begin(Pipe &p) {
    const uint32_t stage = p.ring_n % NSTAGE, use = p.ring_n / NSTAGE;
    mbar_wait(p.bar(BW_FULL + stage), use & 1u);
    return stage;
}
__device__ __forceinline__ void ring_consume_end(Pipe &p, uint32_t stage) {
    mma_commit(p.bar(BW_EMPTY + stage));      // the stage is free once the MMAs that read it have completed
    ++p.ring_n;
}

// The image in IMG_A is complete in BOTH CTAs (issuer side of Epi::a_ready).
__device__ __forceinline__ void wait_a_ready(Pipe &p) {
    mbar_wait(p.bar(B_AREADY), p.a_n & 1u);
    ++p.a_n;
    tc_fence_after();
}
// This CTA's MMAs have read the current image for the last time: tell both CTAs' epilogues (Epi::wait_afree).
__device__ __forceinline__ void image_consumed(Pipe &p) { mma_commit_mc(p.bar(B_AFREE), PAIR_MASK); }

// Forward GEMM  D[128][Nl] = A[128][K] . W[n0 : n0 + Nl][K]^T : A = IMG_A (K-major image of K features), W streamed chunk by chunk.
// split: this CTA computes the output columns [cr N/2, +N/2) (Nl = N/2) into accumulator columns [0, Nl); else all N (the 16-wide head).
template <int ROLE>
__device__ __forceinline__ void gemm_fwd(Pipe &p, const uint8_t *wimg, int N, int K, bool split = true) {
    const int kc = K < KCH ? K : KCH, nch = K / kc;
    const int Nl = split ? N / 2 : N;
    const uint32_t chunk_bytes = (uint32_t)(N * kc * 2), mine = (uint32_t)(Nl * kc * 2), off = split ? p.cr * mine : 0u;   // rows are contiguous in a chunk
    if (ROLE == ROLE_PRODUCER) {
        for (int c = 0; c < nch; ++c) ring_produce(p, wimg + (size_t)c * chunk_bytes + off, mine);
        return;
    }
    wait_a_ready(p);
    const uint32_t idesc = instr_desc_bf16(TR, Nl);
    const uint32_t a0 = smem_u32(p.smem + SM_IMG_A), sbo_a = (uint32_t)(K >> 3) * 128u, sbo_b = (uint32_t)(kc >> 3) * 128u;
    for (int c = 0; c < nch; ++c) {
        const uint32_t stage = ring_consume_begin(p);
        const uint32_t b0 = smem_u32(p.smem + SM_RING + stage * STAGE_BYTES);
        for (int s = 0; s < kc / 16; ++s)
            mma_ss(p.tbase + TM_D, smem_desc(a0 + (uint32_t)(c * (kc >> 3)) * 128u + s * 256u, 128, sbo_a), smem_desc(b0 + s * 256u, 128, sbo_b),
                   idesc, (c | s) != 0);
        ring_consume_end(p, stage);
    }
    mma_commit(p.bar(B_DFULL));
    image_consumed(p);
}

// Input-gradient GEMM  D[128][n_loc] = dZ[128][Kout] . W[Kout][f0 : f0 + n_loc] : dZ = IMG_A (K-major image of Kout features), W read
// MN-major from the same chunk-major image the forward uses (rows = Kout, chunks over the inputs).  split: this CTA computes the input
// features [cr n_use/2, +n_use/2) (whole chunks, or part of the single chunk); else the first n_use features in both CTAs.
// consumed: no weight-gradient GEMM reads this dZ image afterwards.
template <int ROLE>
__device__ __forceinline__ void gemm_bwd(Pipe &p, const uint8_t *wimg, int Kout, int Kin, int n_use, bool split, bool consumed) {
    const int kc = Kin < KCH ? Kin : KCH;
    const int n_loc = split ? n_use / 2 : n_use, f0 = split ? (int)p.cr * n_loc : 0;
    const int nch = n_loc >= kc ? n_loc / kc : 1, c0 = f0 / kc;
    const int ncols = n_loc < kc ? n_loc : kc;                      // output columns per chunk
    const uint32_t chunk_bytes = (uint32_t)(Kout * kc * 2), boff = (uint32_t)((f0 % kc) >> 3) * 128u;
    if (ROLE == ROLE_PRODUCER) {
        for (int c = 0; c < nch; ++c) ring_produce(p, wimg + (size_t)(c0 + c) * chunk_bytes, chunk_bytes);
        return;
    }
    wait_a_ready(p);
    const uint32_t idesc = instr_desc_bf16_major(TR, ncols, 0, 1);
    const uint32_t a0 = smem_u32(p.smem + SM_IMG_A), sbo_a = (uint32_t)(Kout >> 3) * 128u;
    const uint32_t kgrp = (uint32_t)(kc >> 3) * 128u;               // byte stride between 8-row groups of the weight image (K' = out)
    for (int c = 0; c < nch; ++c) {
        const uint32_t stage = ring_consume_begin(p);
        const uint32_t b0 = smem_u32(p.smem + SM_RING + stage * STAGE_BYTES) + boff;
        for (int s = 0; s < Kout / 16; ++s)
            mma_ss(p.tbase + TM_D + (uint32_t)(c * kc), smem_desc(a0 + s * 256u, 128, sbo_a), smem_desc(b0 + s * 2u * kgrp, kgrp, 128), idesc, s != 0);
        ring_consume_end(p, stage);
    }
    mma_commit(p.bar(B_DFULL));
    if (consumed) image_consumed(p);
}

// Reload of a saved activation image into IMG_H (producer) ...
template <int ROLE>
__device__ __forceinline__ void load_h(Pipe &p, const uint8_t *src, uint32_t bytes) {
    if (ROLE != ROLE_PRODUCER) return;
    if (p.h_n == 0) mbar_wait_cluster(p.bar(B_HPUB), 0);            // the epilogues of both CTAs have published the saved images (Epi::publish_saved)
    mbar_wait(p.bar(B_HFREE), (p.h_n & 1u) ^ 1u);
    mbar_arrive_expect_tx(p.bar(B_HFULL), bytes);
    bulk_g2s(p.smem + SM_IMG_H, src, bytes, p.bar(B_HFULL));
    ++p.h_n;
}
// ... and the weight-gradient block  G[128][n] = dZ[:, m0 : m0 + 128]^T . H[:, h0 : h0 + n]  (K = the 128 rows): both operands MN-major.
// first: wait for IMG_H of this layer;  wait_a: no GEMM consumed a_ready for this image yet;  last: IMG_H may be overwritten;
// consumed: last MMA of this CTA that reads the dZ image.
template <int ROLE>
__device__ __forceinline__ void wgrad(Pipe &p, int Fdz, int m0, int Fh, int h0, int n, bool first, bool wait_a, bool last, bool consumed) {
    if (ROLE != ROLE_ISSUER) return;
    if (wait_a) wait_a_ready(p);
    if (first) { mbar_wait(p.bar(B_HFULL), p.h_n & 1u); ++p.h_n; }
    mbar_wait(p.bar(B_GFREE), (p.gfree_n & 1u) ^ 1u);               // the previous block has been drained
    ++p.gfree_n;
    tc_fence_after();
    const uint32_t idesc = instr_desc_bf16_major(TR, n, 1, 1);
    const uint32_t a0 = smem_u32(p.smem + SM_IMG_A) + (uint32_t)(m0 >> 3) * 128u, ka = (uint32_t)(Fdz >> 3) * 128u;
    const uint32_t b0 = smem_u32(p.smem + SM_IMG_H) + (uint32_t)(h0 >> 3) * 128u, kb = (uint32_t)(Fh >> 3) * 128u;
    for (int s = 0; s < TR / 16; ++s)
        mma_ss(p.tbase + TM_G, smem_desc(a0 + s * 2u * ka, ka, 128), smem_desc(b0 + s * 2u * kb, kb, 128), idesc, s != 0);
    mma_commit(p.bar(B_GFULL));
    if (last) mma_commit(p.bar(B_HFREE));
    if (consumed) image_consumed(p);
}

// ---- epilogue-side context -----------------------------------------------------------------------------------------
struct Epi {
    uint8_t *smem;
    uint64_t *bars;
    uint32_t tbase;
    int tid, wg, r, lane, wq;          // wq: warp inside the warpgroup (= TMEM lane quarter)
    uint32_t lane_base;                // TMEM lane field of this warp
    int cr;                            // rank of this CTA in its pair
    uint32_t peer;                     // shared::cluster address of the peer CTA's dynamic shared memory
    uint32_t d_n = 0, g_n = 0, red_n = 0, img_n = 0;
    int row0, B;                       // first minibatch row of the tile, minibatch size

    __device__ float *red() const { return reinterpret_cast<float *>(smem + SM_RED); }
    __device__ float *cs() const { return reinterpret_cast<float *>(smem + SM_CS); }
    __device__ float *rowv() const { return reinterpret_cast<float *>(smem + SM_ROW); }
    __device__ float *stat() const { return reinterpret_cast<float *>(smem + SM_STAT); }
    __device__ uint8_t *img_a() const { return smem + SM_IMG_A; }
    __device__ bool valid() const { return row0 + r < B; }
    __device__ int mrow() const { return row0 + r < B ? row0 + r : 0; }      // row index that is always in range (mask / counter lookups)
    __device__ uint32_t td(int col) const { return tbase + TM_D + lane_base + (uint32_t)col; }
    __device__ uint32_t tg(int col) const { return tbase + TM_G + lane_base + (uint32_t)col; }
    __device__ uint32_t tx(int layer, int col) const { return tbase + (layer ? TM_X2 : TM_X1) + lane_base + (uint32_t)col; }
    __device__ uint32_t peer_bar(int i) const { return peer + (uint32_t)(SM_BAR + i * 8); }

    // Before the first store of a new image into IMG_A: the MMAs of BOTH CTAs have read the previous one (Pipe: image_consumed).
    // Every epilogue thread calls it once per image, whether it writes or not.
    __device__ void wait_afree() {
        if (img_n > 0) mbar_wait_cluster(bars + B_AFREE, (img_n - 1) & 1u);
        ++img_n;
    }
    // The image of F features in IMG_A is complete on this CTA's side: generic-proxy writes fenced for the async proxy, TMEM reads
    // retired, all epilogue threads met.  Then this CTA's half [cr F/2, +F/2) goes to the peer (one bulk copy per 8-row group, counted
    // on the peer's barrier) and this CTA's barrier expects the peer's half.
    __device__ void a_ready(int F) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        bar_epi();
        const uint32_t half = (uint32_t)F * 8u;                   // bytes of F / 2 features of 8 rows
        if (tid < TR / 8) {
            const uint32_t o = (uint32_t)SM_IMG_A + (uint32_t)tid * 2u * half + (uint32_t)cr * half;
            bulk_s2peer(peer + o, smem + o, half, peer_bar(B_AREADY));
        }
        if (tid == 0) mbar_arrive_expect_tx(bars + B_AREADY, (uint32_t)(TR / 8) * half);
    }
    // the same for an image this CTA built alone (the narrow input images, dz4): nothing is exchanged
    __device__ void a_ready_local() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        bar_epi();
        if (tid == 0) mbar_arrive(bars + B_AREADY);
    }
    // Once per kernel, after the last global store of a saved image / input image: every warp of both CTAs releases its stores at
    // cluster scope (and fences them for the async proxy); the producers wait for all 32 warps before the first bulk load (load_h).
    __device__ void publish_saved() {
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive_cluster(peer_bar(B_HPUB)); mbar_arrive_cluster(smem_u32(bars + B_HPUB)); }
    }
    __device__ void wait_d() { mbar_wait(bars + B_DFULL, d_n & 1u); ++d_n; tc_fence_after(); }
    __device__ void wait_g() { mbar_wait(bars + B_GFULL, g_n & 1u); ++g_n; tc_fence_after(); }
    __device__ void g_free() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_GFREE);
    }
    // sum of (a, b) over the 2 x 4 warpgroups' threads of this row; every thread of the row — in both CTAs — gets the same totals
    // (added in (CTA, warpgroup) order).  The partial goes into this CTA's buffer with a plain store and into the peer's with
    // st.async, counted on the peer's barrier (16 warp arrivals + 512 x 8 bytes).  Two buffers and two barriers alternate, so a warp
    // that runs ahead cannot disturb the exchange the others are still reading.
    __device__ void row_allreduce2(float &a, float &b) {
        const uint32_t buf = red_n & 1u, par = (red_n >> 1) & 1u;
        ++red_n;
        const uint32_t o = (buf * (2 * NWG * TR) + (uint32_t)((cr * NWG + wg) * TR + r)) * 2u;
        float *rd = red();
        *reinterpret_cast<float2 *>(rd + o) = make_float2(a, b);
        st_async_f32x2(peer + (uint32_t)SM_RED + o * 4u, a, b, peer_bar(B_XRED + (int)buf));
        __syncwarp();
        if (tid == 0) mbar_arrive_expect_tx(bars + B_XRED + buf, (uint32_t)NEPI * 8u);
        else if (lane == 0) mbar_arrive(bars + B_XRED + buf);
        mbar_wait_cluster(bars + B_XRED + buf, par);
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int g = 0; g < 2 * NWG; ++g) {
            const float2 q = *reinterpret_cast<const float2 *>(rd + (buf * (2 * NWG * TR) + (uint32_t)(g * TR + r)) * 2u);
            sa += q.x; sb += q.y;
        }
        a = sa; b = sb;
    }
    // 8 consecutive features [f0, f0 + 8) of this thread's row into an image of F features: any image (a global copy) ...
    __device__ void img_store8(uint8_t *img, int F, int f0, const float *v) const {
        uint4 q;
        q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4 *>(img + img_off(r, f0, F)) = q;
    }
    // ... or this CTA's IMG_A (the peer receives this CTA's half with a_ready), optionally with a global copy
    __device__ void img_store8_a(int F, int f0, const float *v, uint8_t *gimg = nullptr) const {
        uint4 q;
        q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]); q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
        const uint32_t o = img_off(r, f0, F);
        *reinterpret_cast<uint4 *>(smem + SM_IMG_A + o) = q;
        if (gimg) *reinterpret_cast<uint4 *>(gimg + o) = q;
    }
    __device__ void img_store_a_q(int F, int f0, uint4 q) const { *reinterpret_cast<uint4 *>(smem + SM_IMG_A + img_off(r, f0, F)) = q; }
    // column sums over the tile's 128 rows, stage 1: v[W] are this thread's values for the CTA-local columns [c0, c0 + W)
    // -> cs[slot][wq][c0 + column]
    template <int W>
    __device__ void colsum_stage(int slot, int c0, float (&v)[W]) {
        const float t = warp_colsum<W>(v, lane);
        constexpr int per = 32 / W;
        if ((lane & (per - 1)) == 0) cs()[(slot * 4 + wq) * 256 + c0 + lane / per] = t;
    }
    // stage 2 (after bar_epi): out[c] = sum over the four row quarters, for the CTA-local columns c < n
    __device__ void colsum_finish(int slot, int n, float *out) const {
        if (tid < n) {
            const float *c = cs() + slot * 4 * 256 + tid;
            out[tid] = (c[0] + c[256]) + (c[512] + c[768]);
        }
    }
};

}  // namespace utc
}  // namespace erlfill
