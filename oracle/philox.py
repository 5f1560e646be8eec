"""Philox-4x32-10 in numpy (vectorised) — TEST INFRASTRUCTURE ONLY: the CPU restatement of the counter-based streams the CUDA kernels
draw (csrc/common.cuh: philox4x32_10; the C oracle has the same function as orc_philox4x32), used to rebuild on the host the
dropout keep-masks of the ER-DDPG update's PHILOX mode (csrc/ddpg_dropout.cuh)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_DROPOUT = 4


def philox4x32_10(k0, k1, c0, c1, c2, c3):
    c0, c1, c2, c3 = (np.asarray(x, np.uint32).astype(np.uint64) for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in (c0, c1, c2, c3)]


DDPG_SITE_WIDTHS = [128, 128, 256, 256, 256, 256, 128, 128, 256, 256]


def ddpg_keep_masks(seed, update_idx, rank, B):
    """The ten uint8 keep arrays [B][width] of one update in ERLFILL_DROPOUT_PHILOX mode: one Philox call per 8 columns, counter
    (rank << 20 | row, update_idx, column // 8, STREAM_DROPOUT + 16 + site); the eight 16-bit halves h keep when h >= 13107."""
    out = []
    rows = ((int(rank) << 20) + np.arange(B, dtype=np.uint64)).astype(np.uint32)[:, None]
    for site, w in enumerate(DDPG_SITE_WIDTHS):
        blocks = np.arange(w // 8, dtype=np.uint32)[None, :]
        r = philox4x32_10(seed & 0xFFFFFFFF, seed >> 32, rows, np.uint32(update_idx), blocks, np.uint32(STREAM_DROPOUT + 16 + site))
        halves = np.stack([(r[j >> 1] >> np.uint32(16 * (j & 1))) & np.uint32(0xFFFF) for j in range(8)], axis=-1)     # [B, w/8, 8]
        out.append((halves.reshape(B, w) >= 13107).astype(np.uint8))
    return out
