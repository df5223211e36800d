"""ORACLE (test infrastructure): Philox4x32-10 counter-based RNG restated in numpy.

Published algorithm: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
(SC'11), Random123 ``philox4x32_R(10, ctr, key)``.  Pinned by the Random123 known-answer
vectors in ``tests/test_oracle_philox.py``.

The reference draws with ``torch.bernoulli`` / ``torch.multinomial`` (bandit_sampler.py:98,423;
ladies_sampler.py:68,181), whose stream depends on device and tensor length.  The B200 path
replaces that by one Philox draw per (seed, step, layer, global node id); tests inject the same
numbers into the oracle as the uniform draws (SURVEY.md §8b "RNG injection").
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 arrays (or scalars) of uint32; key: 2 uint32 → 4 arrays of uint32."""
    c = [np.asarray(x, dtype=np.uint32).astype(np.uint64) for x in ctr]
    c = list(np.broadcast_arrays(*c))
    k0 = np.uint32(key[0])
    k1 = np.uint32(key[1])
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c[0]
            p1 = M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return [x.astype(np.uint32) for x in c]


def uniform_for_nodes(seed: int, step: int, layer: int, nids) -> np.ndarray:
    """The draw the device makes for a node: word 0 of
    ``philox(ctr=(nid, layer, step_lo, step_hi), key=(seed_lo, seed_hi))`` → ``(x >> 8)·2^-24``
    in [0, 1) as float32 (csrc/philox.cuh)."""
    nids = np.asarray(nids, dtype=np.uint32)
    out = philox4x32_10(
        (nids, np.uint32(layer), np.uint32(step & 0xFFFFFFFF), np.uint32((step >> 32) & 0xFFFFFFFF)),
        (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return ((out[0] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def edge_keys(seed: int, step: int, layer: int, csc_pos) -> np.ndarray:
    """The 32-bit key the device gives an in-edge for uniform neighbour sampling (csrc/sampler.cu ``edge_key``):
    word 0 of ``philox(ctr=(pos_lo, layer | 0x8000 | pos_hi << 16, step_lo, step_hi), key=seed)``."""
    pos = np.asarray(csc_pos, dtype=np.uint64)
    c1 = (np.uint64(layer) | np.uint64(0x8000) | ((pos >> np.uint64(32)) << np.uint64(16))).astype(np.uint32)
    out = philox4x32_10(((pos & MASK).astype(np.uint32), c1, np.uint32(step & 0xFFFFFFFF),
                         np.uint32((step >> 32) & 0xFFFFFFFF)), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return out[0]
