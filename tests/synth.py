"""Counter-hash synthetic pixels: identical on every host and
GPU without transfers.  v = mix32(seed ^ (c + C*(x + W*y)))."""
import numpy as np


def mix32(v):
    v = v.astype(np.uint64)
    v ^= v >> np.uint64(16)
    v = (v * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(13)
    v = (v * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    return v.astype(np.uint32)


def image(w, h, c, dtype, seed=20261019, hdr=False):
    idx = np.arange(w * h * c, dtype=np.uint64)
    v = mix32(idx ^ np.uint64(seed))
    dtype = np.dtype(dtype)
    if dtype == np.uint8:
        a = (v >> np.uint32(24)).astype(np.uint8)
    elif dtype == np.uint16:
        a = (v >> np.uint32(16)).astype(np.uint16)
    else:
        u = v.astype(np.float64) * (1.0 / 4294967296.0)
        if hdr:
            u = np.exp(6.0 * u - 3.0)
        a = u.astype(dtype)
    return a.reshape(h, w, c)
