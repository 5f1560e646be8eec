// emotion_tf_tc.cu — M2: EmotionTransformer.forward (EmotionModule.py:48-71) batched over envs on the sm_100a
// tensor cores (tcgen05.mma, accumulators in tensor memory); dropout off (module.eval()), or train() mode as the reference runs it
// (it never calls .eval() on this path): keep-masks from Philox or injected by the caller (tf_dropout.cuh), template kDrop.
//
// Same network as emotion_tf.cu (the exact fp32 twin): embedding Linear(3,32) + positions -> 4 x post-norm
// TransformerEncoderLayer(d=32, 4 heads x 8, FFN 32->2048->32 ReLU, LayerNorm eps 1e-5) -> last token ->
// Linear(32,3) -> 0.5*tanh+0.5.  Every dense contraction (in_proj, out_proj, linear1, linear2 = 99 % of the
// 21.8 MFLOP per env) is a bf16 x bf16 -> fp32 tcgen05.mma; the per-(env, head) 20x20 attention runs on register
// fragments with mma.sync.m16n8k8 (bf16 q/k/v/p, fp32 scores and softmax); the LayerNorms, the residual stream
// and the embedding stay fp32 SIMT.
//
// Work decomposition (persistent kernel, one CTA per SM):
//   token tile   = 128 accumulator rows (TMEM lanes) = 6 envs x 20 tokens (+8 idle rows); thread r of the
//                  tile's warpgroup owns row r: its residual stream x[32] lives in that thread's registers, so
//                  LayerNorm, softmax and the residual adds need no cross-thread traffic.
//   CTA          = kTiles token tiles (one warpgroup each) + one MMA-issue warp per tile + one weight-producer warp.
//                  The tiles share the weight ring, so every weight byte fetched from L2 feeds all of them.
//   activations  = the bf16 copy of x that the projections read lives in tensor memory too ([x|1|1|0..], 24 packed
//                  columns written with tcgen05.st by the row owners), so in_proj and linear1 are TS-form MMAs whose
//                  only shared-memory operand is the weight chunk (an SS-form N=64 MMA is bound by re-reading the
//                  128x16 A tile from shared memory, ~48 instead of 32 cycles).
//   FFN pipeline = linear1/linear2 in 32 chunks of 64 hidden units.  Per chunk and tile:
//                  MMA1  H[128x64] = [x|1|1] (TMEM, K=48) . [W1|b1_hi|b1_lo]^T (smem)         -> TMEM (3-deep ring)
//                  epilogue: tcgen05.ld H -> relu -> bf16x2 -> tcgen05.st P in place (A operand of MMA2)
//                  MMA2  Y[128x32] += P (TMEM, K=64) . W2chunk^T (smem)                        -> TMEM
//                  The 2048-wide hidden activation never leaves tensor memory.  The linear1 bias rides in two
//                  extra K columns (hi + lo bf16 halves against two columns of ones), which removes the bias
//                  add from the epilogue: measured on B200 the epilogue then costs ~2.9k cycles per tile-layer
//                  against ~5.1k cycles of MMA, with the bias add it costs 5.8-7.1k (tools/umma_epi.cu).
//   weights      = pre-packed once (erlfill_emotion_tc_pack) into bf16 "images" in the canonical no-swizzle
//                  K-major core-matrix layout, streamed two chunks at a time with cp.async.bulk into a 3-stage ring.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "tf_dropout.cuh"
#include "umma.cuh"

namespace erlfill {
namespace tc {
using namespace umma;

constexpr int D = 32, FF = 2048, NL = 4, SMAX = 20, HD = 8;
#ifndef ERLFILL_TC_TILES
#define ERLFILL_TC_TILES 2
#endif
constexpr int kTiles = ERLFILL_TC_TILES;   // token tiles per CTA: 2 (one CTA per SM, the tiles share the weight stream) or 1 (two CTAs per
                                           // SM, each with its own stream: measured 3 % slower on B200 with both CTAs starting together, i.e.
                                           // in lock-step — a forced phase offset between them is untested; the variant no longer fits the
                                           // shared memory since the parked-row buffers were added: DESIGN.md section 4.2)
constexpr int kCtasPerSm = kTiles == 1 ? 2 : 1;
constexpr int EPT = 6;                     // envs per tile
constexpr int ROWS = EPT * SMAX;           // 120 live rows of 128
constexpr int KX = 48;                     // K extent of the activation operand: 32 features | 1 | 1 | 14 x 0
constexpr int CH = 64;                     // hidden units per FFN chunk
constexpr int NCH = FF / CH;               // 32
constexpr int NS = 3;                      // weight ring stages
constexpr int CPS = 2;                     // FFN chunks per ring stage
constexpr int W1_BYTES = CH * KX * 2;      // 6144
constexpr int W2_BYTES = D * CH * 2;       // 4096
constexpr int CHUNK_BYTES = W1_BYTES + W2_BYTES;
constexpr int STAGE_BYTES = CPS * CHUNK_BYTES;
constexpr int INP_BYTES = 96 * KX * 2;     // 9216
constexpr int OUTP_BYTES = D * KX * 2;     // 3072
constexpr int ATT_BYTES = INP_BYTES + OUTP_BYTES;
// fp32 block: per layer [norm1_w norm1_b norm2_w norm2_b lin2_b] then emb_w[32][3] emb_b[32] pos[20][33] fc_w[3][32] fc_b[4]
constexpr int MISC_LAYER = 5 * D;
constexpr int MISC_EMBW = NL * MISC_LAYER, MISC_EMBB = MISC_EMBW + 96, MISC_POS = MISC_EMBB + 32;
constexpr int POS_LD = 33;
constexpr int MISC_FCW = MISC_POS + SMAX * POS_LD, MISC_FCB = MISC_FCW + 96;
constexpr int MISC_FLOATS = ((MISC_FCB + 4 + 31) / 32) * 32;
// weight stream: per layer one attention stage [in_proj | out_proj] (ATT_BYTES used of a STAGE_BYTES slot) followed by
// NCH / CPS FFN stages; the producer walks it front to back once per token-tile group
constexpr int SPL = 1 + NCH / CPS;         // stages per layer
constexpr int IMG_STREAM = 0;
constexpr int IMG_MISC = IMG_STREAM + NL * SPL * STAGE_BYTES;
constexpr int IMAGE_BYTES = IMG_MISC + MISC_FLOATS * 4;
static_assert(ATT_BYTES <= STAGE_BYTES, "attention weights must fit one ring stage");

constexpr int XS_BYTES = 128 * KX * 2;     // 12288 per tile
// attention operands of one tile, bf16: q[136][40] k[136][40] (token rows, 32 features + 8 pad) and v transposed
// vT[32][136] (feature rows, 128 tokens + 8 pad); the pads make every mma fragment load bank-conflict free
constexpr int QK_LD = 40, QK_ROWS = 136, VT_LD = 136;
constexpr int SQ_OFF = 0, SK_OFF = QK_ROWS * QK_LD * 2, SVT_OFF = 2 * QK_ROWS * QK_LD * 2;
constexpr int KV_BYTES = SVT_OFF + D * VT_LD * 2 + 32;      // + per-env sequence lengths (6 ints)
constexpr int SLEN_OFF = SVT_OFF + D * VT_LD * 2;
// dynamic shared memory map
constexpr int SM_RING = 0;
constexpr int SM_MISC = SM_RING + NS * STAGE_BYTES;
constexpr int SM_XS = SM_MISC + MISC_FLOATS * 4;
constexpr int SM_KV = SM_XS + kTiles * XS_BYTES;
// deferred last layer: only the LAST token of an env feeds fc_out, so in layer 4 everything after the attention (out_proj,
// LayerNorm1, FFN, LayerNorm2) is needed for one row in 20.  Those rows are parked here — the attention output row in a second
// operand image, the residual stream and the env index beside it — and processed kTiles*128 at a time ("tail pass").
constexpr int TAIL_ROWS = kTiles * 128;
constexpr int XT_LD = 36;                  // floats per parked residual row (32 + 4 pad: conflict-free 16-byte row access)
constexpr int SM_XSTAIL = SM_KV + kTiles * KV_BYTES;
constexpr int SM_XTAIL = SM_XSTAIL + kTiles * XS_BYTES;
constexpr int SM_ENVTAIL = SM_XTAIL + TAIL_ROWS * XT_LD * 4;
constexpr int SM_BAR = SM_ENVTAIL + TAIL_ROWS * 4;
constexpr int NHB = 3;                     // H buffers (TMEM) per tile: MMA1 runs two chunks ahead of the epilogue
constexpr int kBarPerTile = 6 + 2 * NHB;   // x_full qkv_full o_full out_full x2_full y_full h_full[NHB] p_full[NHB]
constexpr int kNumBars = kTiles * kBarPerTile + 2 * NS + 1;
constexpr int SM_TMEM_SLOT = SM_BAR + kNumBars * 8;
constexpr int SMEM_BYTES = SM_TMEM_SLOT + 16;
static_assert(kCtasPerSm * (SMEM_BYTES + 1024) <= 233472, "shared memory plan exceeds the SM");
static_assert(SM_MISC % 128 == 0 && SM_XS % 128 == 0 && SM_KV % 16 == 0 && SM_XSTAIL % 16 == 0 && SM_XTAIL % 16 == 0 && SM_BAR % 8 == 0, "alignment");

constexpr int kThreads = kTiles * 128 + kTiles * 32 + 32;   // tile warpgroups | one MMA issuer warp per tile | weight producer
// TMEM columns per tile: H ring [0,192) | Y [192,224) | X = bf16 [x|1|1|0..] [224,248); QKV aliases [0,96), OUT aliases Y
constexpr int TM_TILE = 256;
constexpr int TM_Y = NHB * CH;
constexpr int TM_X = TM_Y + D;
constexpr int TM_COLS = kTiles * TM_TILE;    // 256 or 512: a power of two; two resident CTAs take 256 each
static_assert(TM_X + KX / 2 <= TM_TILE && kCtasPerSm * TM_COLS <= 512, "tensor memory plan");
constexpr float kQScale = 0.51006973f;     // log2(e) / sqrt(8): folded into the q rows of in_proj, softmax then uses exp2

enum { B_XFULL = 0, B_QKV = 1, B_OFULL = 2, B_OUT = 3, B_X2 = 4, B_Y = 5, B_H = 6, B_P = 6 + NHB };

struct TcArgs {
    const uint8_t *img;
    erlfill_emotion_state emo;
    const float *d_seq;
    float *d_out;
    int n, S, n_groups;
    // train() mode (kDrop != 0)
    uint32_t k0, k1, call_idx, row_id_base;
    erlfill_transformer_masks masks;
};
constexpr int DROP_OFF = 0, DROP_PHILOX = 1, DROP_MASK = 2;

// development profile: clock64 ticks of block 0 / thread 0 per phase (see erlfill_debug_tc_profile)
__device__ unsigned long long g_tc_prof[16];
#define TC_PROF_DECL long long tp = clock64(); unsigned int pacc[16] = {}
#define TC_PROF(i)                                   \
    do {                                             \
        if (prof) {                                  \
            const long long now_ = clock64();        \
            pacc[i] += (unsigned int)(now_ - tp);    \
            tp = now_;                               \
        }                                            \
    } while (0)
#define TC_PROF_FLUSH                                \
    do {                                             \
        if (prof) {                                  \
            _Pragma("unroll") for (int i_ = 0; i_ < 16; ++i_) if (pacc[i_]) g_tc_prof[i_] += pacc[i_]; \
        }                                            \
    } while (0)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// D(16x8, fp32) += A(16x8, bf16, row) . B(8x8, bf16, col) on register fragments (the legacy warp-level tensor path: the
// per-(env, head) attention blocks are 20x20x8, far below a tcgen05 tile)
__device__ __forceinline__ void mma_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// this thread's 32 features, rounded to bf16, into its TMEM lane of the tile's X operand (16 packed columns)
__device__ __forceinline__ void store_row_tmem(uint32_t tX, const float (&x)[D]) {
    uint32_t p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
    tmem_st16(tX, p);
    tmem_wait_st();
}

// x <- LayerNorm(x) over the 32 features held by this thread; gamma/beta in shared memory (broadcast reads)
__device__ __forceinline__ void layer_norm32(float (&x)[D], const float *g, const float *b) {
    float mean = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) mean += x[d];
    mean *= (1.0f / D);
    float var = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { const float c = x[d] - mean; var = fmaf(c, c, var); }
    const float rstd = rsqrtf(var * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int d4 = 0; d4 < D; d4 += 4) {
        const float4 gg = *reinterpret_cast<const float4 *>(g + d4), bb = *reinterpret_cast<const float4 *>(b + d4);
        x[d4 + 0] = (x[d4 + 0] - mean) * rstd * gg.x + bb.x;
        x[d4 + 1] = (x[d4 + 1] - mean) * rstd * gg.y + bb.y;
        x[d4 + 2] = (x[d4 + 2] - mean) * rstd * gg.z + bb.z;
        x[d4 + 3] = (x[d4 + 3] - mean) * rstd * gg.w + bb.w;
    }
}

// One schedule, walked identically by every warp role of the CTA (the roles only meet through barriers):
//   for each group of kTiles*6 envs:  layers 0..2 in full, layer 3 up to the attention (its last-token rows are parked);
//   a tail pass (second half of layer 3 on the parked rows) whenever the next group would not fit, and after the last group.
struct Sched {
    int pending = 0;            // parked rows
    uint32_t nA = 0, nB = 0;    // first halves / second halves done so far: phases of their barriers
    uint32_t gs = 0;            // weight stages consumed so far
    __device__ bool flush_after_group(int it, int n_it) const { return pending + kTiles * EPT > TAIL_ROWS || it == n_it - 1; }
};

template <int kDrop>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) emotion_forward_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *w_full = bars + kTiles * kBarPerTile, *w_empty = w_full + NS, *c_full = w_empty + NS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM_SLOT);
    const float *misc = reinterpret_cast<const float *>(smem + SM_MISC);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n_it = (a.n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // ---- one-time setup ----------------------------------------------------------------------------------
    if (tid == 0) {
        for (int t = 0; t < kTiles; ++t) {
            uint64_t *b = bars + t * kBarPerTile;
            // the tile's 4 warps arrive once each (lane 0 after __syncwarp); MMA completions arrive through tcgen05.commit
            mbar_init(b + B_XFULL, 4); mbar_init(b + B_QKV, 1); mbar_init(b + B_OFULL, 4); mbar_init(b + B_OUT, 1);
            mbar_init(b + B_X2, 4); mbar_init(b + B_Y, 1);
            for (int i = 0; i < NHB; ++i) { mbar_init(b + B_H + i, 1); mbar_init(b + B_P + i, 4); }
        }
        for (int s = 0; s < NS; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, kTiles); }
        mbar_init(c_full, 1);
        mbar_fence_init();
    }
    if (warp == kTiles * 5) tmem_alloc(tmem_slot, TM_COLS);
    // attention-output images (per-layer and tail): zero (rows the attention never writes are still read by the out_proj), then
    // the constant columns 32..47 of every row: [1, 1, 0 x 14] (bias hi/lo ride on the two ones)
    for (int i = tid; i < kTiles * XS_BYTES / 16; i += kThreads) {
        reinterpret_cast<uint4 *>(smem + SM_XS)[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(smem + SM_XSTAIL)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < TAIL_ROWS * XT_LD / 4; i += kThreads) reinterpret_cast<uint4 *>(smem + SM_XTAIL)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < TAIL_ROWS; i += kThreads) reinterpret_cast<int *>(smem + SM_ENVTAIL)[i] = -1;
    __syncthreads();
    for (int i = tid; i < kTiles * 128; i += kThreads) {
        const int off = (i >> 7) * XS_BYTES + ((i & 127) >> 3) * (KX / 8) * 128 + (i & 7) * 16 + 4 * 128;
        *reinterpret_cast<uint4 *>(smem + SM_XS + off) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(smem + SM_XSTAIL + off) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    }
    // attention operand buffers start finite (rows/columns past the 120 live tokens are read, never written)
    for (int i = tid; i < kTiles * KV_BYTES / 16; i += kThreads) reinterpret_cast<uint4 *>(smem + SM_KV)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp == kTiles * 5) {
        // =============================== weight producer ===============================================
        if (elect_one()) {
            mbar_arrive_expect_tx(c_full, MISC_FLOATS * 4);
            bulk_g2s(smem + SM_MISC, a.img + IMG_MISC, MISC_FLOATS * 4, c_full);
            uint32_t g = 0;
            const bool prof = blockIdx.x == 0;
            TC_PROF_DECL;
            // stages [first, first + count) of layer l's block of the stream
            auto load = [&](int l, int first, int count) {
                for (int sidx = first; sidx < first + count; ++sidx, ++g) {
                    const uint32_t stage = g % NS, use = g / NS;
                    const uint32_t bytes = sidx == 0 ? ATT_BYTES : STAGE_BYTES;
                    TC_PROF(13);
                    mbar_wait(w_empty + stage, (use & 1u) ^ 1u);
                    TC_PROF(14);
                    mbar_arrive_expect_tx(w_full + stage, bytes);
                    bulk_g2s(smem + SM_RING + stage * STAGE_BYTES, a.img + IMG_STREAM + ((size_t)l * SPL + sidx) * STAGE_BYTES, bytes, w_full + stage);
                }
            };
            Sched sc;
            for (int it = 0; it < n_it; ++it) {
                for (int l = 0; l < NL; ++l) load(l, 0, l < NL - 1 ? SPL : 1);
                sc.pending += kTiles * EPT;
                if (sc.flush_after_group(it, n_it)) { load(NL - 1, 0, SPL); sc.pending = 0; }
            }
            TC_PROF_FLUSH;
        }
        __syncwarp();
    } else if (warp >= kTiles * 4) {
        // =============================== MMA issuer of tile t ==========================================
        const int t = warp - kTiles * 4;
        constexpr uint32_t idesc_qkv = instr_desc_bf16(128, 96), idesc_32 = instr_desc_bf16(128, 32),
                           idesc_h = instr_desc_bf16(128, CH);
        constexpr uint32_t SBO_X = (KX / 8) * 128, SBO_W2 = (CH / 8) * 128;
        const uint32_t xs = smem_u32(smem + SM_XS) + t * XS_BYTES, xs_tail = smem_u32(smem + SM_XSTAIL) + t * XS_BYTES;
        const uint32_t ring0 = smem_u32(smem + SM_RING);
        const uint32_t tT = tbase + t * TM_TILE;
        uint64_t *b = bars + t * kBarPerTile;
        // descriptor of ring byte offset `off` = base descriptor + (off >> 4): the start-address field holds address/16 and never
        // carries out of its 14 bits inside the ring, so the issue loop needs one add per MMA instead of rebuilding the descriptor
        const uint64_t dX = smem_desc(ring0, 128, SBO_X), dW2 = smem_desc(ring0, 128, SBO_W2);
        mbar_wait(c_full, 0);
        const bool prof = blockIdx.x == 0 && t == 0 && (tid & 31) == 0;
        TC_PROF_DECL;
        Sched sc;
        // first half of a layer: q|k|v = [x|1|1] (TMEM) . [W_in|b_hi|b_lo]^T from the layer's attention stage (stage sc.gs)
        auto first_half = [&](bool release_stage) {
            const uint32_t att_stage = sc.gs % NS, att = att_stage * (STAGE_BYTES >> 4);
            mbar_wait(w_full + att_stage, (sc.gs / NS) & 1u);
            mbar_wait(b + B_XFULL, sc.nA & 1u);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < KX / 16; ++s) mma_ts(tT, tT + TM_X + s * 8, dX + (att + s * 16), idesc_qkv, s > 0);
                mma_commit(b + B_QKV);
                if (release_stage) mma_commit(w_empty + att_stage);
            }
            __syncwarp();
            ++sc.nA;
        };
        // second half: out_proj from the image at `img` and the attention stage sc.gs, then the FFN from stages sc.gs+1 ..
        auto second_half = [&](uint32_t img) {
            const uint32_t par = sc.nB & 1u;
            const uint32_t att_stage = sc.gs % NS, att = att_stage * (STAGE_BYTES >> 4);
            mbar_wait(w_full + att_stage, (sc.gs / NS) & 1u);
            mbar_wait(b + B_OFULL, par);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < KX / 16; ++s)
                    mma_ss(tT + TM_Y, smem_desc(img + s * 256, 128, SBO_X), dX + (att + (INP_BYTES >> 4) + s * 16), idesc_32, s > 0);
                mma_commit(b + B_OUT);
                mma_commit(w_empty + att_stage);
            }
            __syncwarp();
            // ---- FFN: MMA1 runs NHB-1 chunks ahead of MMA2, so the epilogue always has a finished H chunk waiting ----
            mbar_wait(b + B_X2, par);
            tc_fence_after();
            TC_PROF(8);
            const uint32_t g0 = sc.nB * NCH, gs0 = sc.gs + 1;   // chunk counter (H ring / P phases) and first FFN stage
            for (int c = 0; c < NCH + NHB - 1; ++c) {
                if (c < NCH) {
                    const uint32_t g = g0 + c, sg = gs0 + c / CPS, stage = sg % NS, hb = g % NHB;
                    if ((c & (CPS - 1)) == 0) {
                        mbar_wait(w_full + stage, (sg / NS) & 1u);
                        tc_fence_after();
                    }
                    TC_PROF(9);
                    if (elect_one()) {
                        const uint64_t w1 = dX + (stage * (STAGE_BYTES >> 4) + (c & (CPS - 1)) * (CHUNK_BYTES >> 4));
#pragma unroll
                        for (int s = 0; s < KX / 16; ++s) mma_ts(tT + hb * CH, tT + TM_X + s * 8, w1 + s * 16, idesc_h, s > 0);
                        mma_commit(b + B_H + hb);
                    }
                    __syncwarp();
                    TC_PROF(10);
                }
                if (c >= NHB - 1) {
                    const int cc = c - (NHB - 1);
                    const uint32_t g = g0 + cc, sg = gs0 + cc / CPS, stage = sg % NS, hb = g % NHB;
                    mbar_wait(b + B_P + hb, (g / NHB) & 1u);
                    tc_fence_after();
                    TC_PROF(11);
                    if (elect_one()) {
                        const uint64_t w2 = dW2 + (stage * (STAGE_BYTES >> 4) + (cc & (CPS - 1)) * (CHUNK_BYTES >> 4) + (W1_BYTES >> 4));
#pragma unroll
                        for (int s = 0; s < CH / 16; ++s)
                            mma_ts(tT + TM_Y, tT + hb * CH + s * 8, w2 + s * 16, idesc_32, (cc > 0 || s > 0) ? 1u : 0u);
                        if ((cc & (CPS - 1)) == CPS - 1) mma_commit(w_empty + stage);
                        if (cc == NCH - 1) mma_commit(b + B_Y);
                    }
                    __syncwarp();
                    TC_PROF(12);
                }
            }
            ++sc.nB;
            sc.gs += SPL;
        };
        for (int it = 0; it < n_it; ++it) {
            for (int l = 0; l < NL; ++l) {
                first_half(l == NL - 1);
                if (l < NL - 1) second_half(xs);
                else sc.gs += 1;
            }
            sc.pending += kTiles * EPT;
            if (sc.flush_after_group(it, n_it)) { second_half(xs_tail); sc.pending = 0; }
        }
        TC_PROF_FLUSH;
    } else {
        // =============================== token-tile warpgroups ==========================================
        const int t = warp >> 2, r = tid & 127, lane = tid & 31, wq = warp & 3;
        const int e_loc = r / SMAX, s_pos = r % SMAX;
        uint64_t *b = bars + t * kBarPerTile;
        uint8_t *xs = smem + SM_XS + t * XS_BYTES;
        uint8_t *att = smem + SM_KV + t * KV_BYTES;
        __nv_bfloat16 *sq = reinterpret_cast<__nv_bfloat16 *>(att + SQ_OFF), *sk = reinterpret_cast<__nv_bfloat16 *>(att + SK_OFF),
                      *svt = reinterpret_cast<__nv_bfloat16 *>(att + SVT_OFF);
        int *slen = reinterpret_cast<int *>(att + SLEN_OFF);
        float *xtail = reinterpret_cast<float *>(smem + SM_XTAIL);
        int *envtail = reinterpret_cast<int *>(smem + SM_ENVTAIL);
        const uint32_t tT = tbase + t * TM_TILE + ((uint32_t)(wq * 32) << 16);
        // constant K columns 32..47 of the X operand: [1, 1, 0 x 14]
        {
            const uint32_t ones[8] = {0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(tT + TM_X + 16, ones);
            tmem_wait_st();
        }
        mbar_wait(c_full, 0);
        const bool prof = blockIdx.x == 0 && tid == 0;
        TC_PROF_DECL;
        Sched sc;
        float x[D];

        // ---- first half of a layer: x -> q|k|v (tensor core) -> attention on register fragments -> bf16 rows of o.
        //      park == false: o goes to the tile's operand image for the out_proj that follows;
        //      park == true (last layer): only each env's last-token row is kept, in the tail image at row sc.pending + t*6 + env ----
        // global ids of the tile's six envs (Philox counter word / mask row) in the current group
        int grp_env0 = 0;
        auto first_half = [&](bool park, int l) {
            store_row_tmem(tT + TM_X, x);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b + B_XFULL);
            TC_PROF(0);
            mbar_wait(b + B_QKV, sc.nA & 1u);
            tc_fence_after();
            TC_PROF(1);
            {
                // q|k|v rows -> bf16 attention operands: q[r][:], k[r][:] row-major, v transposed vT[:][r]
                uint32_t rq[32];
                tmem_ld32(tT, rq);
                tmem_wait_ld();
                uint4 *qrow = reinterpret_cast<uint4 *>(sq + r * QK_LD);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    qrow[j] = make_uint4(pack_bf16x2(__uint_as_float(rq[8 * j]), __uint_as_float(rq[8 * j + 1])),
                                         pack_bf16x2(__uint_as_float(rq[8 * j + 2]), __uint_as_float(rq[8 * j + 3])),
                                         pack_bf16x2(__uint_as_float(rq[8 * j + 4]), __uint_as_float(rq[8 * j + 5])),
                                         pack_bf16x2(__uint_as_float(rq[8 * j + 6]), __uint_as_float(rq[8 * j + 7])));
                tmem_ld32(tT + 32, rq);
                tmem_wait_ld();
                uint4 *krow = reinterpret_cast<uint4 *>(sk + r * QK_LD);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    krow[j] = make_uint4(pack_bf16x2(__uint_as_float(rq[8 * j]), __uint_as_float(rq[8 * j + 1])),
                                         pack_bf16x2(__uint_as_float(rq[8 * j + 2]), __uint_as_float(rq[8 * j + 3])),
                                         pack_bf16x2(__uint_as_float(rq[8 * j + 4]), __uint_as_float(rq[8 * j + 5])),
                                         pack_bf16x2(__uint_as_float(rq[8 * j + 6]), __uint_as_float(rq[8 * j + 7])));
                tmem_ld32(tT + 64, rq);
                tmem_wait_ld();
#pragma unroll
                for (int d = 0; d < D; ++d) svt[d * VT_LD + r] = __float2bfloat16_rn(__uint_as_float(rq[d]));
            }
            tc_fence_before();                                // q|k|v loads ordered before the MMAs that reuse those columns
            named_bar_sync(1 + t, 128);
            TC_PROF(2);
            // ---- attention on register fragments: this warp owns 6 of the tile's 24 (env, head) blocks ----
            {
                const int g8 = lane >> 2, tq = lane & 3;
#pragma unroll 3
                for (int u = wq * 6; u < wq * 6 + 6; ++u) {
                    const int e = u >> 2, h = u & 3, row0 = e * SMAX;
                    const int Se = max(slen[e], 1);             // envs past the batch end: keep every value finite
                    // scores S = Q_h K_h^T (already in the log2 domain): 2 query m-tiles x 3 key n-tiles
                    uint32_t qa[2][2], kb[3];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        qa[mt][0] = *reinterpret_cast<const uint32_t *>(sq + (row0 + mt * 16 + g8) * QK_LD + h * HD + 2 * tq);
                        qa[mt][1] = *reinterpret_cast<const uint32_t *>(sq + (row0 + mt * 16 + g8 + 8) * QK_LD + h * HD + 2 * tq);
                    }
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt)
                        kb[nt] = *reinterpret_cast<const uint32_t *>(sk + (row0 + nt * 8 + g8) * QK_LD + h * HD + 2 * tq);
                    float scr[2][3][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 3; ++nt) {
                            scr[mt][nt][0] = scr[mt][nt][1] = scr[mt][nt][2] = scr[mt][nt][3] = 0.f;
                            mma_16x8x8(scr[mt][nt], qa[mt][0], qa[mt][1], kb[nt]);
                        }
                    // V fragments: vT[feature g8 of head h][key 8*ks + 2*tq, +1]
                    uint32_t vb[3];
#pragma unroll
                    for (int ks = 0; ks < 3; ++ks)
                        vb[ks] = *reinterpret_cast<const uint32_t *>(svt + (h * HD + g8) * VT_LD + row0 + ks * 8 + 2 * tq);
                    // masked keys belong to another env (or to no token): their V is zeroed, not just weighted by p = 0, so a
                    // non-finite value in one env can never reach its neighbours in the tile
                    if (Se == SMAX) { if (tq >= 2) vb[2] = 0u; }
                    else {
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) {
                            const int k0 = ks * 8 + 2 * tq;
                            if (k0 >= Se) vb[ks] = 0u; else if (k0 + 1 >= Se) vb[ks] &= 0x0000FFFFu;
                        }
                    }
                    // destination of a query row: the tile image, or (parked) the tail image row of this env if it is the last token
                    const int park_row = sc.pending + t * EPT + e;
                    uint8_t *park_img = smem + SM_XSTAIL + (park_row >> 7) * XS_BYTES;
                    auto store_o = [&](int qrow, float o0, float o1) {
                        if (!park) {
                            const int rr = row0 + qrow;
                            *reinterpret_cast<uint32_t *>(xs + (rr >> 3) * (KX / 8) * 128 + h * 128 + (rr & 7) * 16 + tq * 4) = pack_bf16x2(o0, o1);
                        } else if (qrow == Se - 1) {
                            const int rr = park_row & 127;
                            *reinterpret_cast<uint32_t *>(park_img + (rr >> 3) * (KX / 8) * 128 + h * 128 + (rr & 7) * 16 + tq * 4) = pack_bf16x2(o0, o1);
                        }
                    };
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        // rows g8 (elements 0,1) and g8+8 (elements 2,3) of this m-tile; key of element i: 8*nt + 2*tq + (i&1).
                        // Rows 24..31 (second half of m-tile 1) never belong to this env: nothing is computed for them.
                        const bool lo_rows_only = mt == 1;
                        if (Se == SMAX) {                 // full history (the common case): only keys 20..23 are masked
                            if (tq >= 2) scr[mt][2][0] = scr[mt][2][1] = scr[mt][2][2] = scr[mt][2][3] = -INFINITY;
                        } else {
#pragma unroll
                            for (int nt = 0; nt < 3; ++nt)
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    if (nt * 8 + 2 * tq + (i & 1) >= Se) scr[mt][nt][i] = -INFINITY;
                        }
                        float mx0 = fmaxf(fmaxf(fmaxf(scr[mt][0][0], scr[mt][0][1]), fmaxf(scr[mt][1][0], scr[mt][1][1])),
                                          fmaxf(scr[mt][2][0], scr[mt][2][1]));
                        float mx1 = lo_rows_only ? 0.f
                                                 : fmaxf(fmaxf(fmaxf(scr[mt][0][2], scr[mt][0][3]), fmaxf(scr[mt][1][2], scr[mt][1][3])),
                                                         fmaxf(scr[mt][2][2], scr[mt][2][3]));
                        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                        if (!lo_rows_only) {
                            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                        }
                        float sum0 = 0.f, sum1 = 0.f;
                        uint32_t pa[3][2];
                        // dropout on the attention probabilities (the sums below stay those of the undropped softmax): the pair (query i,
                        // keys j, j + 1) sits in one half-word pair of one Philox block since its element index is even.  The six blocks of
                        // this thread's fragment are drawn unconditionally and back to back — six independent chains the scheduler can
                        // interleave; one call at a time inside the range checks, the ten dependent rounds of every call were exposed
                        // (31.7 k of the 92 k cycles of a train()-mode layer step).
                        uint32_t am[3][2];
                        if (kDrop == DROP_PHILOX) {
                            // The four threads of a quad hold the 24 consecutive elements (query row i, keys 0..23) of the site: element
                            // index R .. R + 23 with R = (h * 20 + i) * 20, R % 8 in {0, 4} — three or four Philox blocks.  Lane tq of the
                            // quad draws block R / 8 + tq (one call per row instead of three), the words travel by shuffle: thread tq
                            // needs word (R % 8) / 2 + 4 nt + tq of the quad's 16.
                            const uint32_t env_g = (uint32_t)(grp_env0 + e);
                            const int qbase = lane & ~3;
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                const uint32_t R = (uint32_t)((h * SMAX + (mt * 16 + g8 + 8 * hf)) * SMAX);
                                const uint4 rr = tfdrop::block8(a.k0, a.k1, a.row_id_base + env_g, a.call_idx, 4u * l, (R >> 3) + (uint32_t)tq);
                                const uint32_t off = (R & 7u) >> 1;
#pragma unroll
                                for (int nt = 0; nt < 3; ++nt) {
                                    const uint32_t W = off + 4u * nt + (uint32_t)tq;          // word of the quad's 16
                                    const int src = qbase + (int)(W >> 2);
                                    const uint32_t w0 = __shfl_sync(0xffffffffu, rr.x, src), w1 = __shfl_sync(0xffffffffu, rr.y, src);
                                    const uint32_t w2 = __shfl_sync(0xffffffffu, rr.z, src), w3 = __shfl_sync(0xffffffffu, rr.w, src);
                                    const uint32_t ws = W & 3u;
                                    am[nt][hf] = tfdrop::pair_mask(ws == 0 ? w0 : (ws == 1 ? w1 : (ws == 2 ? w2 : w3)));
                                }
                            }
                        }
#pragma unroll
                        for (int nt = 0; nt < 3; ++nt) {
                            const float p0 = fast_exp2(scr[mt][nt][0] - mx0), p1 = fast_exp2(scr[mt][nt][1] - mx0);
                            sum0 += p0 + p1;
                            pa[nt][0] = pack_bf16x2(p0, p1);
                            if (!lo_rows_only) {
                                const float p2 = fast_exp2(scr[mt][nt][2] - mx1), p3 = fast_exp2(scr[mt][nt][3] - mx1);
                                sum1 += p2 + p3;
                                pa[nt][1] = pack_bf16x2(p2, p3);
                            } else pa[nt][1] = 0u;
                            if (kDrop != DROP_OFF) {
                                const int env_g = grp_env0 + e;
                                const int j = nt * 8 + 2 * tq;
#pragma unroll
                                for (int hf = 0; hf < 2; ++hf) {
                                    if (hf == 1 && lo_rows_only) continue;
                                    const int i = mt * 16 + g8 + 8 * hf;
                                    uint32_t m = 0xffffffffu;
                                    if (i < Se && j < Se && env_g < a.n) {
                                        if (kDrop == DROP_PHILOX) m = am[nt][hf];
                                        else {
                                            const uint8_t *mp = a.masks.attn + ((((size_t)l * a.n + env_g) * 4 + h) * a.S + i) * a.S + j;
                                            m = (mp[0] ? 0xffffu : 0u) | ((j + 1 < Se && mp[1]) ? 0xffff0000u : 0u);
                                        }
                                    }
                                    pa[nt][hf] &= m;
                                }
                            }
                        }
                        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
                        if (!lo_rows_only) {
                            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
                        }
                        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int ks = 0; ks < 3; ++ks) mma_16x8x8(o, pa[ks][0], pa[ks][1], vb[ks]);
                        // rows past the env's 20 tokens belong to the next env: not stored
                        const int q0 = mt * 16 + g8;
                        const float dsc = kDrop != DROP_OFF ? tfdrop::kScale : 1.0f;      // kept probabilities carry 1 / 0.7
                        if (q0 < SMAX) {
                            const float i0 = fast_rcp(sum0) * dsc;
                            store_o(q0, o[0] * i0, o[1] * i0);
                        }
                        if (!lo_rows_only) {
                            const float i1 = fast_rcp(sum1) * dsc;
                            store_o(q0 + 8, o[2] * i1, o[3] * i1);
                        }
                    }
                }
            }
            fence_proxy_async();                              // o rows (generic-proxy stores) -> visible to the out_proj MMA
            TC_PROF(3);
            ++sc.nA;
        };

        // ---- second half of a layer on this thread's row: out_proj + residual + LayerNorm1, FFN, + residual + LayerNorm2 ----
        // drow_env / drow_s: the env (index in this call) and token of the row this thread processes in second_half (-1: padding row)
        int drow_env = -1, drow_s = 0;
        // 32 keep bits (features 0..31) of this row at site 4l + which (1: dropout1, 3: dropout2)
        auto keep32_row = [&](int l, int which) -> uint32_t {
            if (kDrop == DROP_OFF || drow_env < 0) return 0xffffffffu;
            uint32_t bits = 0;
            if (kDrop == DROP_PHILOX) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    bits |= tfdrop::keep_bits8(tfdrop::block8(a.k0, a.k1, a.row_id_base + (uint32_t)drow_env, a.call_idx, 4u * l + which,
                                                              (uint32_t)(drow_s * 4 + q))) << (8 * q);
            } else {
                const uint8_t *mp = (which == 1 ? a.masks.d1 : a.masks.d2) + (((size_t)l * a.n + drow_env) * a.S + drow_s) * D;
#pragma unroll
                for (int d = 0; d < D; ++d) bits |= (mp[d] ? 1u : 0u) << d;
            }
            return bits;
        };
        auto second_half = [&](int l) {
            const uint32_t par = sc.nB & 1u;
            const float *lw = misc + l * MISC_LAYER;
            __syncwarp();
            if (lane == 0) mbar_arrive(b + B_OFULL);
            mbar_wait(b + B_OUT, par);
            tc_fence_after();
            TC_PROF(4);
            {
                uint32_t ro[32];
                tmem_ld32(tT + TM_Y, ro);
                tmem_wait_ld();
                if (kDrop == DROP_OFF) {
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] += __uint_as_float(ro[d]);
                } else {
                    const uint32_t kb = keep32_row(l, 1);
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] += ((kb >> d) & 1u) ? __uint_as_float(ro[d]) * tfdrop::kScale : 0.f;
                }
            }
            layer_norm32(x, lw, lw + D);
            store_row_tmem(tT + TM_X, x);
            tc_fence_before();                                // also orders the tcgen05.ld of OUT before MMA2 overwrites that region
            __syncwarp();
            if (lane == 0) mbar_arrive(b + B_X2);
            TC_PROF(5);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                const uint32_t g = sc.nB * NCH + c, hb = g % NHB;
                // dropout on the hidden activation: packed word w of the chunk holds units 64c + 2w, 2w + 1 = the two halves of word w % 4
                // of the row's Philox block s * 256 + 8c + w / 4.  The masks do not depend on H: they are drawn BEFORE waiting for the
                // chunk's MMA, in the time the row owner would otherwise idle.  (The 1 / 0.7 scale is applied to Y, which is linear in P.)
                uint32_t dm[32];
                if (kDrop == DROP_PHILOX && drow_env >= 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint4 rr = tfdrop::block8(a.k0, a.k1, a.row_id_base + (uint32_t)drow_env, a.call_idx, 4u * l + 2u,
                                                        (uint32_t)(drow_s * (FF / 8) + c * (CH / 8) + q));
                        dm[4 * q] = tfdrop::pair_mask(rr.x); dm[4 * q + 1] = tfdrop::pair_mask(rr.y);
                        dm[4 * q + 2] = tfdrop::pair_mask(rr.z); dm[4 * q + 3] = tfdrop::pair_mask(rr.w);
                    }
                }
                mbar_wait(b + B_H + hb, (g / NHB) & 1u);
                tc_fence_after();
                uint32_t h0[32], h1[32], p[32];
                tmem_ld32(tT + hb * CH, h0);
                tmem_ld32(tT + hb * CH + 32, h1);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    p[j] = pack_relu_bf16x2(__uint_as_float(h0[2 * j]), __uint_as_float(h0[2 * j + 1]));
                    p[16 + j] = pack_relu_bf16x2(__uint_as_float(h1[2 * j]), __uint_as_float(h1[2 * j + 1]));
                }
                if (kDrop == DROP_PHILOX && drow_env >= 0) {
#pragma unroll
                    for (int w = 0; w < 32; ++w) p[w] &= dm[w];
                } else if (kDrop == DROP_MASK && drow_env >= 0) {
                    const uint8_t *mp = a.masks.ff + (((size_t)l * a.n + drow_env) * a.S + drow_s) * FF + c * CH;
#pragma unroll
                    for (int w = 0; w < 32; ++w) p[w] &= (mp[2 * w] ? 0xffffu : 0u) | (mp[2 * w + 1] ? 0xffff0000u : 0u);
                }
                tmem_st32(tT + hb * CH, p);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b + B_P + hb);
            }
            TC_PROF(6);
            mbar_wait(b + B_Y, par);
            tc_fence_after();
            {
                uint32_t ry[32];
                tmem_ld32(tT + TM_Y, ry);
                tmem_wait_ld();
                if (kDrop == DROP_OFF) {
#pragma unroll
                    for (int d4 = 0; d4 < D; d4 += 4) {
                        const float4 b2 = *reinterpret_cast<const float4 *>(lw + 4 * D + d4);
                        x[d4 + 0] += __uint_as_float(ry[d4 + 0]) + b2.x;
                        x[d4 + 1] += __uint_as_float(ry[d4 + 1]) + b2.y;
                        x[d4 + 2] += __uint_as_float(ry[d4 + 2]) + b2.z;
                        x[d4 + 3] += __uint_as_float(ry[d4 + 3]) + b2.w;
                    }
                } else {
                    // Y = W2 (keep o relu(h)): the FFN dropout's 1 / 0.7 first, then the bias, then dropout2 (with its own 1 / 0.7)
                    const uint32_t kb = keep32_row(l, 3);
#pragma unroll
                    for (int d = 0; d < D; ++d)
                        x[d] += ((kb >> d) & 1u) ? (__uint_as_float(ry[d]) * tfdrop::kScale + lw[4 * D + d]) * tfdrop::kScale : 0.f;
                }
            }
            layer_norm32(x, lw + 2 * D, lw + 3 * D);
            tc_fence_before();                                // tcgen05.ld of Y / H ordered before the next MMAs into those columns
            TC_PROF(7);
            ++sc.nB;
        };

        for (int it = 0; it < n_it; ++it) {
            const int group = blockIdx.x + it * gridDim.x;
            const int env = (group * kTiles + t) * EPT + e_loc;
            const bool env_ok = r < ROWS && env < a.n;
            // ---- sequence of this env (EmotionModule.get_state, :83-102) and this row's token ----
            int S = 0;
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = 0.f;
            if (env_ok) {
                float in0, in1, in2;
                if (a.d_seq) {
                    S = a.S;
                    if (s_pos < S) {
                        const float *p = a.d_seq + ((size_t)env * a.S + s_pos) * 3;
                        in0 = p[0]; in1 = p[1]; in2 = p[2];
                    }
                } else {
                    const int len = a.emo.d_hist_len[env];
                    if (len == 0) { S = 1; in0 = 0.5f; in1 = 0.5f; in2 = 0.0f; }
                    else {
                        S = SMAX;
                        const int slot = (a.emo.d_hist_head[env] + max(0, s_pos - (SMAX - len))) % SMAX;
                        in0 = a.emo.d_history[((size_t)slot * 3 + 0) * a.n + env];
                        in1 = a.emo.d_history[((size_t)slot * 3 + 1) * a.n + env];
                        in2 = a.emo.d_history[((size_t)slot * 3 + 2) * a.n + env];
                    }
                }
                if (s_pos < S) {
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        float v = misc[MISC_EMBB + d];
                        v = fmaf(in0, misc[MISC_EMBW + d * 3 + 0], v);
                        v = fmaf(in1, misc[MISC_EMBW + d * 3 + 1], v);
                        v = fmaf(in2, misc[MISC_EMBW + d * 3 + 2], v);
                        x[d] = v + misc[MISC_POS + s_pos * POS_LD + d];
                    }
                }
            }
            // the previous group's attention (last reader of slen) finished before this point: every warp of the tile has since
            // passed a barrier that all four warps arrive on after their attention
            if (r < ROWS && s_pos == 0) slen[e_loc] = S;
            grp_env0 = (group * kTiles + t) * EPT;
            drow_env = (env_ok && s_pos < S) ? env : -1;
            drow_s = s_pos;

            for (int l = 0; l < NL - 1; ++l) {
                first_half(false, l);
                second_half(l);
            }
            // last layer: park the residual stream of each env's last token (and which env it is); the attention parks its o row
            if (r < ROWS && s_pos == max(S, 1) - 1) {
                const int pr = sc.pending + t * EPT + e_loc;
                float4 *dst = reinterpret_cast<float4 *>(xtail + pr * XT_LD);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                envtail[pr] = (env_ok && S > 0) ? (env | ((S - 1) << 24)) : -1;      // env (< 2^24) and the token index of the parked row
            }
            first_half(true, NL - 1);
            sc.pending += kTiles * EPT;
            if (sc.flush_after_group(it, n_it)) {
                // ---- tail pass: second half of the last layer on the parked rows; thread r of tile t takes parked row t*128 + r ----
                asm volatile("bar.sync %0, %1;" ::"r"(kTiles + 1), "r"(kTiles * 128) : "memory");     // all parked rows written
                const int pr = t * 128 + r;
                const int ptag = pr < sc.pending ? envtail[pr] : -1;
                const int penv = ptag < 0 ? -1 : (ptag & 0xffffff);
                drow_env = penv; drow_s = ptag < 0 ? 0 : (ptag >> 24);
                {
                    const float4 *src = reinterpret_cast<const float4 *>(xtail + pr * XT_LD);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float4 v = src[j]; x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w; }
                }
                second_half(NL - 1);
                // ---- fc_out -> 0.5*tanh + 0.5 ----
                if (penv >= 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float acc = misc[MISC_FCB + c];
#pragma unroll
                        for (int d = 0; d < D; ++d) acc = fmaf(x[d], misc[MISC_FCW + c * D + d], acc);
                        a.d_out[(size_t)penv * 3 + c] = tanhf(acc) * 0.5f + 0.5f;
                    }
                }
                // the parked rows may be overwritten only after every thread has read its own
                asm volatile("bar.sync %0, %1;" ::"r"(kTiles + 1), "r"(kTiles * 128) : "memory");
                sc.pending = 0;
            }
        }
        TC_PROF_FLUSH;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTiles * 5) tmem_dealloc(tbase, TM_COLS);
}

// ---- weight packing: fp32 state_dict tensors -> bf16 operand images + fp32 block -----------------------------
__device__ __forceinline__ void put_bf16(uint8_t *img, uint32_t off, float v) {
    *reinterpret_cast<__nv_bfloat16 *>(img + off) = __float2bfloat16_rn(v);
}
// value of column k (0..47) of an augmented weight row: 32 weights, bias hi, bias lo, zeros
__device__ __forceinline__ float aug_col(const float *wrow, float bias, int k, float scale) {
    if (k < D) return wrow[k] * scale;
    const float bs = bias * scale;
    const float hi = __bfloat162float(__float2bfloat16_rn(bs));
    if (k == D) return hi;
    if (k == D + 1) return bs - hi;
    return 0.f;
}

__global__ void emotion_tc_pack_kernel(const erlfill_transformer_weights w, uint8_t *img) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    // FFN chunks
    for (int i = i0; i < NL * NCH * (CH * KX + D * CH); i += stride) {
        const int per = CH * KX + D * CH;
        const int lc = i / per, j = i % per, l = lc / NCH, c = lc % NCH;
        uint8_t *st = img + IMG_STREAM + ((size_t)l * SPL + 1 + c / CPS) * STAGE_BYTES + (size_t)(c % CPS) * CHUNK_BYTES;
        if (j < CH * KX) {
            const int nn = j / KX, k = j % KX, hid = c * CH + nn;
            put_bf16(st, kmajor_off(nn, k, KX), aug_col(w.lin1_w + ((size_t)l * FF + hid) * D, w.lin1_b[l * FF + hid], k, 1.0f));
        } else {
            const int jj = j - CH * KX, o = jj / CH, k = jj % CH;
            put_bf16(st + W1_BYTES, kmajor_off(o, k, CH), w.lin2_w[((size_t)l * D + o) * FF + c * CH + k]);
        }
    }
    // attention projections (q rows pre-scaled by log2(e)/sqrt(8))
    for (int i = i0; i < NL * (96 + D) * KX; i += stride) {
        const int per = (96 + D) * KX;
        const int l = i / per, j = i % per;
        uint8_t *at = img + IMG_STREAM + (size_t)l * SPL * STAGE_BYTES;
        if (j < 96 * KX) {
            const int nn = j / KX, k = j % KX;
            put_bf16(at, kmajor_off(nn, k, KX), aug_col(w.in_proj_w + ((size_t)l * 96 + nn) * D, w.in_proj_b[l * 96 + nn], k, nn < D ? kQScale : 1.0f));
        } else {
            const int jj = j - 96 * KX, nn = jj / KX, k = jj % KX;
            put_bf16(at + INP_BYTES, kmajor_off(nn, k, KX), aug_col(w.out_proj_w + ((size_t)l * D + nn) * D, w.out_proj_b[l * D + nn], k, 1.0f));
        }
    }
    float *misc = reinterpret_cast<float *>(img + IMG_MISC);
    for (int i = i0; i < MISC_FLOATS; i += stride) {
        float v = 0.f;
        if (i < MISC_EMBW) {
            const int l = i / MISC_LAYER, j = i % MISC_LAYER, f = j / D, d = j % D;
            const float *src = f == 0 ? w.norm1_w : (f == 1 ? w.norm1_b : (f == 2 ? w.norm2_w : (f == 3 ? w.norm2_b : w.lin2_b)));
            v = src[l * D + d];
        } else if (i < MISC_EMBB) v = w.emb_w[i - MISC_EMBW];
        else if (i < MISC_POS) v = w.emb_b[i - MISC_EMBB];
        else if (i < MISC_FCW) {
            const int j = i - MISC_POS, s = j / POS_LD, d = j % POS_LD;
            v = d < D ? w.pos[s * D + d] : 0.f;
        } else if (i < MISC_FCB) v = w.fc_w[i - MISC_FCW];
        else if (i < MISC_FCB + 3) v = w.fc_b[i - MISC_FCB];
        misc[i] = v;
    }
}

}  // namespace tc
}  // namespace erlfill

using namespace erlfill;

extern "C" {

/* development aid (not in erlfill.h): clock64 ticks block 0 spent per phase since the last reset:
 * 0 image store, 1 wait q|k|v, 2 q|k|v epilogue, 3 attention, 4 wait out_proj, 5 LN1 + image store, 6 FFN chunk loop,
 * 7 wait Y + LN2; issuer warp of tile 0: 8 non-FFN part of the layer, 9 wait weights, 10 issue MMA1, 11 wait P, 12 issue MMA2 */
int erlfill_debug_tc_profile(unsigned long long *out16, int reset) {
    if (out16) ERLFILL_CUDA_TRY(cudaMemcpyFromSymbol(out16, tc::g_tc_prof, sizeof(tc::g_tc_prof)));
    if (reset) {
        unsigned long long z[16] = {};
        ERLFILL_CUDA_TRY(cudaMemcpyToSymbol(tc::g_tc_prof, z, sizeof(z)));
    }
    return ERLFILL_OK;
}

/* development aid: compile-time shape of the kernel (token tiles per CTA, resident CTAs per SM) */
int erlfill_debug_tc_config(int *tiles, int *ctas_per_sm) {
    if (tiles) *tiles = tc::kTiles;
    if (ctas_per_sm) *ctas_per_sm = tc::kCtasPerSm;
    return ERLFILL_OK;
}

size_t erlfill_emotion_tc_image_bytes(void) { return (size_t)tc::IMAGE_BYTES; }

int erlfill_emotion_tc_pack(const erlfill_transformer_weights *w, void *d_image, void *stream) {
    if (!w || !d_image) return ERLFILL_ERR_ARG;
    if (((uintptr_t)d_image & 15) != 0) return ERLFILL_ERR_ARG;
    const float *const *wp = reinterpret_cast<const float *const *>(w);
    for (size_t i = 0; i < sizeof(*w) / sizeof(float *); ++i)
        if (!wp[i]) return ERLFILL_ERR_ARG;
    tc::emotion_tc_pack_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(*w, static_cast<uint8_t *>(d_image));
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

int erlfill_emotion_forward_tc(const void *d_image, int n_env, const float *d_seq, int seq_len,
                               const erlfill_emotion_state *emo, float *d_out, void *stream) {
    return erlfill_emotion_forward_tc_train(d_image, n_env, d_seq, seq_len, emo, ERLFILL_DROPOUT_OFF, 0, 0, 0, nullptr, d_out, stream);
}

int erlfill_emotion_forward_tc_train(const void *d_image, int n_env, const float *d_seq, int seq_len, const erlfill_emotion_state *emo,
                                     int dropout_mode, uint64_t seed, uint32_t call_idx, uint32_t row_id_base,
                                     const erlfill_transformer_masks *masks, float *d_out, void *stream) {
    if (!d_image || n_env <= 0 || n_env >= (1 << 24) || !d_out) return ERLFILL_ERR_ARG;
    if (dropout_mode < ERLFILL_DROPOUT_OFF || dropout_mode > ERLFILL_DROPOUT_MASK) return ERLFILL_ERR_ARG;
    if (dropout_mode == ERLFILL_DROPOUT_MASK && (!masks || !masks->attn || !masks->d1 || !masks->ff || !masks->d2 || !d_seq)) return ERLFILL_ERR_ARG;
    if (!d_seq && (!emo || !emo->d_history || !emo->d_hist_len || !emo->d_hist_head)) return ERLFILL_ERR_ARG;
    if (d_seq && (seq_len < 1 || seq_len > tc::SMAX)) return ERLFILL_ERR_ARG;
    tc::TcArgs a;
    a.img = static_cast<const uint8_t *>(d_image);
    if (emo) a.emo = *emo; else a.emo = erlfill_emotion_state{};
    a.d_seq = d_seq; a.d_out = d_out; a.n = n_env; a.S = d_seq ? seq_len : tc::SMAX;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32); a.call_idx = call_idx; a.row_id_base = row_id_base;
    if (masks) a.masks = *masks; else a.masks = erlfill_transformer_masks{};
    a.n_groups = (n_env + tc::kTiles * tc::EPT - 1) / (tc::kTiles * tc::EPT);
    static int n_sm[64] = {};
    int dev = 0;
    ERLFILL_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return ERLFILL_ERR_UNSUPPORTED;
    if (!n_sm[dev]) {
        int v = 0;
        ERLFILL_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        ERLFILL_CUDA_TRY(cudaFuncSetAttribute(tc::emotion_forward_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        ERLFILL_CUDA_TRY(cudaFuncSetAttribute(tc::emotion_forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        ERLFILL_CUDA_TRY(cudaFuncSetAttribute(tc::emotion_forward_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        n_sm[dev] = v;
    }
    const int max_ctas = n_sm[dev] * tc::kCtasPerSm;
    const int grid = a.n_groups < max_ctas ? a.n_groups : max_ctas;
    if (dropout_mode == ERLFILL_DROPOUT_OFF) tc::emotion_forward_tc_kernel<0><<<grid, tc::kThreads, tc::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    else if (dropout_mode == ERLFILL_DROPOUT_PHILOX) tc::emotion_forward_tc_kernel<1><<<grid, tc::kThreads, tc::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    else tc::emotion_forward_tc_kernel<2><<<grid, tc::kThreads, tc::SMEM_BYTES, (cudaStream_t)stream>>>(a);
    ERLFILL_CUDA_TRY(cudaGetLastError());
    return ERLFILL_OK;
}

}  // extern "C"
