"""Philox4x32-10 counter-based RNG and the epsilon / base-sample draw spec.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it.

Philox4x32-10 is restated from its published description (Salmon, Moraes, Dror,
Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11) and pinned by the
Random123 known-answer vectors in ``tests/test_oracle_philox.py``.

The draw spec below is what the CUDA kernels implement for
``eps_kind = gaussian | rademacher`` (in-kernel noise, BASELINE.json
north_star): the value of row ``r`` of sample ``b`` depends on
``(seed, stream, b, r)`` only, never on the RK stage or step, so the noise is
constant over one solve exactly as the reference's host draw is
(/root/reference/src/core/base_icnf.jl:258-259 draws once per solve, the
closure at :62-78 captures it).
"""

from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)

STREAM_EPS = 0x45505331  # "EPS1": Hutchinson probe
STREAM_BASE = 0x42415345  # "BASE": base-distribution sample for generate()
STREAM_STEER = 0x53544552  # "STER": STEER end-time perturbation

_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for rnd in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            if rnd != 9:
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def _words(seed: int, stream: int, rows: int, n: int, offset: int) -> np.ndarray:
    """uint32 words, shape (rows_padded_to_4, n); word for (row r, sample b) is
    philox(ctr=(b_lo, b_hi, r//4, stream), key=(seed_lo, seed_hi))[r%4] with b
    the GLOBAL sample index (offset + local column), so a batch sharded across
    GPUs draws the same numbers as the unsharded batch."""
    nblk = (rows + 3) // 4
    b = np.arange(offset, offset + n, dtype=np.uint64)
    ctr = np.zeros((nblk, n, 4), dtype=np.uint32)
    ctr[..., 0] = (b & _MASK32).astype(np.uint32)[None, :]
    ctr[..., 1] = (b >> np.uint64(32)).astype(np.uint32)[None, :]
    ctr[..., 2] = np.arange(nblk, dtype=np.uint32)[:, None]
    ctr[..., 3] = np.uint32(stream)
    key = np.zeros((nblk, n, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w = philox4x32_10(ctr, key)  # (nblk, n, 4)
    return np.transpose(w, (0, 2, 1)).reshape(nblk * 4, n)


def rademacher(seed: int, rows: int, n: int, offset: int = 0,
               stream: int = STREAM_EPS) -> np.ndarray:
    """(rows, n) float32 of +-1: +1 where the word's top bit is set."""
    w = _words(seed, stream, rows, n, offset)[:rows]
    return np.where((w >> np.uint32(31)) != 0, np.float32(1.0), np.float32(-1.0)).astype(np.float32)


def gaussian(seed: int, rows: int, n: int, offset: int = 0,
             stream: int = STREAM_EPS) -> np.ndarray:
    """(rows, n) float32 standard normals by Box-Muller on word pairs
    (w[2p], w[2p+1]) of each Philox block: u1 = ((w0>>8)+1)*2^-24 in (0,1],
    u2 = (w1>>8)*2^-24 in [0,1); even row -> r*cos(2 pi u2), odd -> r*sin."""
    w = _words(seed, stream, rows, n, offset)
    nblk4 = w.shape[0]
    w = w.reshape(nblk4 // 2, 2, n)
    u1 = ((w[:, 0] >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (w[:, 1] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(2.0 * np.pi) * u2).astype(np.float32)
    out = np.empty((nblk4, n), dtype=np.float32)
    out[0::2] = r * np.cos(ang)
    out[1::2] = r * np.sin(ang)
    return out[:rows]


def uniform_pm(seed: int, rate: float, stream: int = STREAM_STEER) -> float:
    """One U(-rate, rate) float32 draw (STEER, base_icnf.jl:23-39)."""
    w = _words(seed, stream, 1, 1, 0)[0, 0]
    u = np.float32((int(w) >> 8) * 2.0 ** -24)
    return float(np.float32(rate) * (np.float32(2.0) * u - np.float32(1.0)))
