"""Portable deterministic synthetic data for fixtures and tests -- TEST INFRASTRUCTURE.

`hashrand` is a counter-based generator (splitmix64 finaliser on uint64 indices) so that the
same (shape, seed) gives bit-identical fp32 arrays on every machine / numpy / torch version;
golden fixtures therefore only need to store outputs plus the seeds of their big inputs.
"""
import numpy as np
import torch


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def hashrand_np(shape, seed, lo=-0.5, hi=0.5):
    n = int(np.prod(shape))
    with np.errstate(over='ignore'):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
        bits = _splitmix64(_splitmix64(idx)) >> np.uint64(40)            # top 24 bits
    u = bits.astype(np.float32) * np.float32(1.0 / (1 << 24))            # exact in fp32, [0,1)
    return (u * np.float32(hi - lo) + np.float32(lo)).reshape(shape)


def hashrand(shape, seed, lo=-0.5, hi=0.5):
    return torch.from_numpy(hashrand_np(tuple(shape), seed, lo, hi))


def hash_state_dict(template, seed, scale=None):
    """Fill a dict {name: shape} with hashrand tensors; per-key seeds derive from key order."""
    out = {}
    for i, (k, shape) in enumerate(template.items()):
        s = (scale or {}).get(k, 0.05)
        out[k] = hashrand(shape, seed * 1000 + i, -s, s)
    return out
