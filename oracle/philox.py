"""Philox4x32-10 + Box-Muller, numpy restatement of the kernel's random stream (csrc/common.cuh).

Stream definition: key = (seed_lo, seed_hi), counter = (gid, sample, attempt, call_id); the four 32-bit outputs
give u1..u4 in (0,1) via ((x >> 9) + 0.5) * 2^-23 (exact in fp32), and
    eps = (sqrt(-2 ln u1) cos(2 pi u2), sqrt(-2 ln u1) sin(2 pi u2), sqrt(-2 ln u3) cos(2 pi u4)).
This plays the role of torch's `_standard_normal` draw inside MultivariateNormal.rsample
(torch/distributions/multivariate_normal.py:251-254; call site gauss_to_pc.py:149): eps[s, i, :] is the draw for
sample s of the i-th still-unfinished Gaussian.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays; returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n2 = hi0 ^ c3 ^ k1
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _unit(x):
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def draw_eps(gids, k, attempt, seed, call_id=0):
    """eps of shape (k, n', 3) float32 for samples 0..k-1 of the Gaussians `gids` (global ids) in `attempt`."""
    gids = np.asarray(gids, dtype=np.int64)
    g = (gids & 0xFFFFFFFF).astype(np.uint32)[None, :]
    s = np.arange(k, dtype=np.uint32)[:, None]
    x, y, z, w = philox4x32_10(g, s, np.uint32(attempt), np.uint32(call_id),
                               np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    u1, u2, u3, u4 = _unit(x), _unit(y), _unit(z), _unit(w)
    two_pi = np.float32(6.283185307179586)
    ra = np.sqrt(np.maximum(np.float32(0.0), np.float32(-2.0) * np.log(u1)))
    rb = np.sqrt(np.maximum(np.float32(0.0), np.float32(-2.0) * np.log(u3)))
    e = np.empty((k, gids.shape[0], 3), dtype=np.float32)
    e[..., 0] = ra * np.cos(two_pi * u2)
    e[..., 1] = ra * np.sin(two_pi * u2)
    e[..., 2] = rb * np.cos(two_pi * u4)
    return e
