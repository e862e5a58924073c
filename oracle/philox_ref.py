"""numpy statement of the counter-based normal generator the CUDA rollout uses in production mode.

TEST INFRASTRUCTURE ONLY (see oracle/rollout_oracle.py header).  It plays two roles:
(1) tests/golden fixtures do not store Brownian increments: both the golden generator and the
tests regenerate ``noise[K, B, d]`` from a seed with this function and stream it into the
reference / the oracle / the CUDA validation mode; (2) ``tests/test_rng_gpu.py`` checks the
in-kernel generator against it.

Spec (one Philox4x32-10 call yields four standard normals):
    counter = (particle index, step index, dim block j // 4, stream id),  key = (seed lo, seed hi)
    r0..r3 = philox4x32_10(counter, key)
    u1 = ((r >> 8) + 1) * 2^-24 in (0, 1],  u2 = (r' >> 8) * 2^-24 in [0, 1)
    rad = sqrt(-2 ln u1), theta = 2 pi u2 - pi
    (z[4b], z[4b+1]) = rad(r0) * (cos, sin)(theta(r1)),  (z[4b+2], z[4b+3]) = rad(r2) * (cos, sin)(theta(r3))
The reference itself draws from torch's global generator (losses/oc.py:277, eq/sdes.py:537) - any
N(0, I) stream is a valid Brownian path; parity is always checked on identical increments.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_NOISE = 0
STREAM_PRIOR = 1
STREAM_ACCEPT = 2  # the uniforms of the MALA accept test (lrds_mala): u = ((r0 >> 8) + 1) * 2^-24 of counter (chain, step, 0, 2)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10 (Salmon et al. 2011).  c* are uint32 arrays, k* python ints."""
    c0 = c0.astype(np.uint64)
    c1 = c1.astype(np.uint64)
    c2 = c2.astype(np.uint64)
    c3 = c3.astype(np.uint64)
    for r in range(10):
        ka = np.uint64((k0 + r * W0) & 0xFFFFFFFF)
        kb = np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ ka), lo1, (hi0 ^ c3 ^ kb), lo0
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _box_muller(ra, rb):
    u1 = ((ra >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u2 = (rb >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    rad = np.sqrt(-2.0 * np.log(u1))
    theta = 2.0 * np.pi * u2 - np.pi
    return rad * np.cos(theta), rad * np.sin(theta)


def normals(seed: int, B: int, K: int, d: int, particle_offset: int = 0, stream: int = STREAM_NOISE,
            dtype=np.float32) -> np.ndarray:
    """Standard normals z[K, B, d] per the spec above (computed in float64, rounded to ``dtype``)."""
    nblk = (d + 3) // 4
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    kk, bb, jj = np.meshgrid(np.arange(K, dtype=np.uint32),
                             np.arange(particle_offset, particle_offset + B, dtype=np.uint32),
                             np.arange(nblk, dtype=np.uint32), indexing="ij")
    r0, r1, r2, r3 = philox4x32_10(bb, kk, jj, np.full_like(bb, stream), k0, k1)
    z0, z1 = _box_muller(r0, r1)
    z2, z3 = _box_muller(r2, r3)
    z = np.stack([z0, z1, z2, z3], axis=-1).reshape(K, B, nblk * 4)[:, :, :d]
    return np.ascontiguousarray(z.astype(dtype))


def uniforms(seed: int, B: int, K: int, stream: int = STREAM_ACCEPT, dtype=np.float32) -> np.ndarray:
    """u[K, B] in (0, 1]: the first Philox word of counter (b, k, 0, stream), as the MALA kernel draws them."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    kk, bb = np.meshgrid(np.arange(K, dtype=np.uint32), np.arange(B, dtype=np.uint32), indexing="ij")
    r0, _, _, _ = philox4x32_10(bb, kk, np.zeros_like(bb), np.full_like(bb, stream), k0, k1)
    return ((r0 >> np.uint32(8)).astype(np.float64) + 1.0).astype(dtype) * dtype(2.0 ** -24)
