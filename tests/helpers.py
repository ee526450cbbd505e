"""Shared test helpers (numpy Philox4x32-10, ray sets)."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
RNG_DOMAIN = 0x52544232


def philox4x32_10(ctr, key):
    """ctr: [n,4] uint32, key: [n,2] uint32 -> [n,4] uint32 (Salmon et al., Random123)."""
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k = np.array(key, dtype=np.uint64).reshape(-1, 2).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[:, 0]
        p1 = np.uint64(M1) * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], axis=1)
        k[:, 0] = (k[:, 0] + np.uint64(W0)) & mask
        k[:, 1] = (k[:, 1] + np.uint64(W1)) & mask
    return c.astype(np.uint32)


def u01(x):
    return (np.asarray(x, np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def camera_block0(seed, pixel, sample):
    """The Philox block the kernels draw the pixel jitter and shutter time from:
    counter (pixel, sample, bounce 0 | block 0, domain), key = seed lo/hi."""
    n = len(pixel)
    ctr = np.stack([np.asarray(pixel, np.uint32), np.asarray(sample, np.uint32), np.zeros(n, np.uint32),
                    np.full(n, RNG_DOMAIN, np.uint32)], axis=1)
    key = np.tile(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32), (n, 1))
    return philox4x32_10(ctr, key)


def camera_rays(cam, nx, ny, n, rng):
    """n jittered camera rays of the thin-lens camera with aperture 0 (float32, as the GPU sees them)."""
    cam = np.asarray(cam, np.float64)
    origin, lleft, horiz, vert = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    i = rng.integers(0, nx, n)
    j = rng.integers(0, ny, n)
    s = (i + rng.random(n)) / nx
    t = (j + rng.random(n)) / ny
    d = lleft[None] + s[:, None] * horiz[None] + t[:, None] * vert[None] - origin[None]
    o = np.tile(origin, (n, 1))
    tm = rng.random(n)
    return o.astype(np.float32), d.astype(np.float32), tm.astype(np.float32)
