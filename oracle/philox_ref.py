"""Dropout keep-masks of the path, CPU restatement (test infrastructure only; never imported by ``sgcdet_b200``).

The reference draws its dropout masks with ``nn.Dropout`` (FFN of ``mmdet3d_plugin/models/im2voxel/transformer_utils/
encoder.py:262-340`` through mmcv's ``FFN``; the attention residual of ``deformable_cross_attention.py:835-837``), i.e. ATen's
Philox stream -- a contract no other implementation can reproduce bit for bit.  What the product fixes instead is stated here:
keep-mask[i] = (Philox4x32-10(counter, key)[i % 4] <= keep * 2^32 - 1) with

    key     = the 64-bit seed (low word, high word)
    counter = (w, (w >> 32) ^ (job << 24), step, step >> 32),  w = 4 * (i // 16) + (i % 16) // 4

(Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11; known-answer vectors of the
Random123 distribution are pinned in ``tests/test_oracle_cpu.py``.)  ``job`` numbers the masks of one launch, ``step`` the
launches of one call site.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """counter [...,4] uint32, key = (k0, k1) -> [...,4] uint32."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def keep_mask(n: int, keep: float, seed: int, job: int, step: int) -> np.ndarray:
    """uint8 [n] keep-mask of mask number ``job`` of launch number ``step``."""
    seed &= 2 ** 64 - 1
    thr = min(int(np.float64(np.float32(keep)) * 4294967296.0 - 1.0), 0xFFFFFFFF)
    words = np.arange((n + 15) // 16 * 4, dtype=np.uint64)
    ctr = np.stack([words & np.uint64(0xFFFFFFFF), (words >> np.uint64(32)) ^ np.uint64(job << 24),
                    np.full_like(words, step & 0xFFFFFFFF), np.full_like(words, step >> 32)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32))
    return (r.reshape(-1) <= np.uint32(thr)).astype(np.uint8)[:n]
