// common.cuh — shared device/host helpers for liberlfill.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "erlfill.h"

namespace erlfill {

// ---- error plumbing (C ABI never throws) -------------------------------------------------
extern thread_local int g_last_cuda_error;
inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return ERLFILL_ERR_CUDA;
}
#define ERLFILL_CUDA_TRY(expr)                                     \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::erlfill::cuda_fail(_e);    \
    } while (0)

// ---- Philox-4x32-10, counter = (env_global_id, step_idx, block, stream) -------------------
// Same function as oracle/erlfill_oracle.c:orc_philox4x32 (Salmon et al., SC'11).
enum : uint32_t { STREAM_TICK = 0u, STREAM_MISC = 1u, STREAM_OU = 2u, STREAM_SAMPLE = 3u, STREAM_DROPOUT = 4u };

// Philox-4x32 with R rounds (Salmon et al., SC'11).  R = 10 is the standard generator used everywhere a stream is also rebuilt on the
// host (simulator noise, OU noise, replay sampling, the update's dropout masks); R = 7 — the fewest rounds that still pass BigCrush in
// the paper — only draws the EmotionTransformer's train()-mode keep-masks (tf_dropout.cuh), where mask generation bounds the kernel.
template <int R>
__host__ __device__ __forceinline__ uint4 philox4x32_r(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                        uint32_t c2, uint32_t c3) {
    return philox4x32_r<10>(k0, k1, c0, c1, c2, c3);
}

// U in [0,1) on the 2^-32 grid, exact in binary64
__device__ __forceinline__ double u32_to_unit(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }

}  // namespace erlfill
