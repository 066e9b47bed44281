"""Philox4x32-10 counter-based RNG (numpy, vectorised).

Host-side twin of the device generator in csrc/sampler.cuh.  It is used for things that are *inputs* to the hot path
(synthetic weights) and, through oracle/sampler.py, as the CPU statement of the explicit sampling contract
(DESIGN.md "Sampler contract").  Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11) constants.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(counter, key):
    """counter: uint32 array [...,4]; key: uint32 array [...,2] (broadcastable). Returns uint32 [...,4]."""
    c = np.asarray(counter, dtype=np.uint64)
    k = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (np.array(c[..., i]) for i in range(4))
    k0 = np.array(np.broadcast_to(k[..., 0], c0.shape))
    k1 = np.array(np.broadcast_to(k[..., 1], c0.shape))
    for r in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform01(stream: int, n: int, seed: int = 0) -> np.ndarray:
    """n float64 uniforms in [0,1) for a named stream: counter=(i, 0, stream, 0x5eed), key=(seed_lo, seed_hi)."""
    i = np.arange(n, dtype=np.uint64)
    ctr = np.stack([i & _MASK, i >> _S32, np.full(n, stream, np.uint64), np.full(n, 0x5EED, np.uint64)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    out = philox4x32(ctr, key).astype(np.uint64)
    bits = (out[:, 0] | (out[:, 1] << _S32)) >> np.uint64(11)  # 53 bits
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)
