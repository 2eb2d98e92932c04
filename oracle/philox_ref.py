"""TEST INFRASTRUCTURE (oracle): Philox4x32-10 counter-based generator (Salmon et al., "Parallel random numbers: as
easy as 1, 2, 3", SC'11; Random123) and the disturbance model the engine's device-side generator implements.

The reference draws ``sigMat @ random.normal(0, 1, 4)`` once per ``noise_length`` control steps and uses the two
position entries (``src/trajectorySimulate.py:268, 351-356``); the engine's ``mpcb_noise_fill`` reproduces that MODEL (one
4-word Philox block per (lane, refresh): two Box-Muller normals used, two words left unused), not numpy's Mersenne
stream -- the batched API takes the draws as an input tensor, so any generator can feed it.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = np.array(ctr, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(np.asarray(key, dtype=np.uint32), c.shape[:-1] + (2,)), copy=True)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c[..., 0].astype(np.uint64)
            p1 = M1 * c[..., 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = np.stack([hi1 ^ c[..., 1] ^ k[..., 0], lo1, hi0 ^ c[..., 3] ^ k[..., 1], lo0], axis=-1)
            k = np.stack([k[..., 0] + W0, k[..., 1] + W1], axis=-1)
    return c


def noise_fill(B, n_refresh, sigma_x, sigma_y, seed, lane_offset=0):
    """noise[n_refresh, 2, B] exactly as csrc/sim.cuh noise_fill_kernel builds it."""
    lane = np.arange(B, dtype=np.uint64) + np.uint64(lane_offset)
    r = np.arange(n_refresh, dtype=np.uint64)
    ctr = np.zeros((n_refresh, B, 4), dtype=np.uint32)
    ctr[..., 0] = (lane & MASK).astype(np.uint32)[None, :]
    ctr[..., 1] = (lane >> np.uint64(32)).astype(np.uint32)[None, :]
    ctr[..., 2] = r.astype(np.uint32)[:, None]
    key = np.array([np.uint64(seed) & MASK, np.uint64(seed) >> np.uint64(32)], dtype=np.uint64).astype(np.uint32)
    x = philox4x32_10(ctr, key)
    u0 = (x[..., 0].astype(np.float64) + 0.5) / 4294967296.0
    u1 = (x[..., 1].astype(np.float64) + 0.5) / 4294967296.0
    rad = np.sqrt(-2.0 * np.log(u0))
    out = np.empty((n_refresh, 2, B))
    out[:, 0, :] = sigma_x * rad * np.cos(2.0 * np.pi * u1)
    out[:, 1, :] = sigma_y * rad * np.sin(2.0 * np.pi * u1)
    return out
