// tf_dropout.cuh — keep-masks of the EmotionTransformer's dropout layers (p = 0.3: EmotionModule.py:9,35), shared by the exact fp32
// kernel (emotion_tf.cu) and the tcgen05 kernel (emotion_tf_tc.cu) so that ERLFILL_DROPOUT_PHILOX means the same masks in both.
//
// Sites per encoder layer l, in at::dropout's draw order: 4l + 0 attention probabilities, 4l + 1 dropout1 (after out_proj),
// 4l + 2 FFN hidden (after relu), 4l + 3 dropout2 (after linear2).  Element index inside a site:
//   attention (head h, query i, key j): (h * 20 + i) * 20 + j        dropout1 / dropout2 (token s, feature c): s * 32 + c
//   FFN hidden (token s, unit u): s * 2048 + u
// One Philox-4x32-7 call (common.cuh: the fewest rounds that pass BigCrush; these masks are never rebuilt outside the two kernels, and
// drawing them bounds the train()-mode kernel) covers 8 consecutive elements: counter (global env id, call index, element / 8, STREAM_DROPOUT + 8 + site);
// its eight 16-bit halves h decide  keep = ((h & 0x7fff) >= 9830), i.e. P(drop) = 9830 / 32768 = 0.29999.  The low 15 bits are used so
// that two halves can be compared at once with one 32-bit add (no carry between the halves) — the packed form the tensor-core kernel
// needs, where one 32-bit word holds two bf16 activations.  Kept values are scaled by 1 / 0.7 like at::dropout.
#pragma once
#include "common.cuh"

namespace erlfill {
namespace tfdrop {

constexpr uint32_t kT15 = 9830u;
constexpr float kScale = 1.0f / (1.0f - 0.3f);

__device__ __forceinline__ uint4 block8(uint32_t k0, uint32_t k1, uint32_t gid, uint32_t call_idx, uint32_t site, uint32_t block) {
    return philox4x32_r<7>(k0, k1, gid, call_idx, block, STREAM_DROPOUT + 8u + site);
}
// keep flag of element `lane8` (0..7) of a block
__device__ __forceinline__ bool keep_elem(const uint4 &r, int lane8) {
    const uint32_t w = (lane8 >> 1) == 0 ? r.x : ((lane8 >> 1) == 1 ? r.y : ((lane8 >> 1) == 2 ? r.z : r.w));
    return ((w >> (16 * (lane8 & 1))) & 0x7fffu) >= kT15;
}
// AND-mask for a packed pair of 16-bit values from the word that holds the pair's two halves: 0xffff per kept half
__device__ __forceinline__ uint32_t pair_mask(uint32_t w) {
    const uint32_t t = (w & 0x7fff7fffu) + (0x8000u - kT15) * 0x00010001u;      // bit 15 / 31 set iff the half keeps
    uint32_t m;      // sign-replicating byte select: bytes 0,1 <- sign(byte 1), bytes 2,3 <- sign(byte 3)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(t), "r"(0u), "r"(0xBB99u));
    return m;
}
// 8 keep bits of a block (bit e = element e)
__device__ __forceinline__ uint32_t keep_bits8(const uint4 &r) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t bits = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint32_t s = (w[t] & 0x7fff7fffu) + (0x8000u - kT15) * 0x00010001u;
        bits |= ((s >> 15) & 1u) << (2 * t);
        bits |= ((s >> 31) & 1u) << (2 * t + 1);
    }
    return bits;
}

}  // namespace tfdrop
}  // namespace erlfill
