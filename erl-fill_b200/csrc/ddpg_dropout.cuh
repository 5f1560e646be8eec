// ddpg_dropout.cuh — keep-masks of the ten Dropout(0.2) layers one DDPGAgent.learn() evaluates (shared by update.cu and update_tc.cu).
#pragma once
#include "common.cuh"

namespace erlfill {
namespace ddrop {

// train() mode of the reference: Dropout(0.2) in actor and critic, ddpg_agent.py:40,43,150,155
// site s of the ten dropout layers one learn() evaluates, in at::dropout's draw order:
//   0,1 actor_target   2,3 critic_target   4,5 critic(s, a)   6,7 actor   8,9 critic(s, actor(s))
// PHILOX: one Philox-4x32-10 call covers 8 consecutive columns of a row: counter (global row, update index, column / 8,
// STREAM_DROPOUT + 16 + site), the eight 16-bit halves h give keep = (h >= 13107), i.e. P(drop) = 13107 / 65536 = 0.2 - 3e-6.
// MASK: caller-supplied uint8 keep arrays [B][width].  Kept activations are scaled by 1 / 0.8 like at::dropout.
struct Drop {
    int mode;
    uint32_t k0, k1, update_idx, row_base;
    const uint8_t *m[10];
};
constexpr float kDropScale = 1.25f;
constexpr uint32_t kDropThresh = 13107u;
// keep bits of the W = 8 / 16 / 32 columns [c0, c0 + W) (c0 % 8 == 0) of global row `grow` at dropout site `site` of width `width`
// (bit j = column c0 + j)
template <int W>
__device__ __forceinline__ uint32_t keep_bits(const Drop &d, int site, int grow_local, int c0, int width) {
    if (d.mode == ERLFILL_DROPOUT_OFF) return W == 32 ? 0xffffffffu : ((1u << (W & 31)) - 1u);
    uint32_t bits = 0;
    if (d.mode == ERLFILL_DROPOUT_MASK) {
        const uint8_t *p = d.m[site] + (size_t)grow_local * width + c0;
#pragma unroll
        for (int q = 0; q < W / 8; ++q) {
            const uint2 w = *reinterpret_cast<const uint2 *>(p + 8 * q);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                bits |= (((w.x >> (8 * b)) & 0xffu) ? 1u : 0u) << (8 * q + b);
                bits |= (((w.y >> (8 * b)) & 0xffu) ? 1u : 0u) << (8 * q + 4 + b);
            }
        }
        return bits;
    }
#pragma unroll
    for (int q = 0; q < W / 8; ++q) {
        const uint4 r = philox4x32_10(d.k0, d.k1, d.row_base + (uint32_t)grow_local, d.update_idx, (uint32_t)((c0 >> 3) + q),
                                      STREAM_DROPOUT + 16u + (uint32_t)site);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t h = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
            bits |= (h >= kDropThresh ? 1u : 0u) << (8 * q + j);
        }
    }
    return bits;
}
__device__ __forceinline__ uint32_t keep32(const Drop &d, int site, int grow_local, int c0, int width) { return keep_bits<32>(d, site, grow_local, c0, width); }

// one element (the exact path's scattered fragment layout)
__device__ __forceinline__ bool keep1(const Drop &d, int site, int grow_local, int c, int width) {
    if (d.mode == ERLFILL_DROPOUT_OFF) return true;
    if (d.mode == ERLFILL_DROPOUT_MASK) return d.m[site][(size_t)grow_local * width + c] != 0;
    const uint4 r = philox4x32_10(d.k0, d.k1, d.row_base + (uint32_t)grow_local, d.update_idx, (uint32_t)(c >> 3), STREAM_DROPOUT + 16u + (uint32_t)site);
    const int j = c & 7;
    const uint32_t w = (j >> 1) == 0 ? r.x : ((j >> 1) == 1 ? r.y : ((j >> 1) == 2 ? r.z : r.w));
    return ((w >> (16 * (j & 1))) & 0xffffu) >= kDropThresh;
}

}  // namespace ddrop
}  // namespace erlfill
