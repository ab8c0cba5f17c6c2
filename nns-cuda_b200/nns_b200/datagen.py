"""Synthetic workloads of BASELINE.json: counter-based, reproducible in any
language -- x[i] = float(splitmix64(seed ^ stream-salt, i) >> 40) * 2^-24 in [0, 1) -- so the
data does not depend on glibc rand().  `reference_rand_sample` reproduces the reference's own
generator (main.cu:10-13, 24-35: srand(1000); rand()/double(RAND_MAX); queries first, then
references) through libc for config C1."""
from __future__ import annotations

import ctypes

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, stream: int, start: int, count: int) -> np.ndarray:
    """splitmix64 of counters start..start+count-1 in stream `stream` (uint64 array)."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed ^ (stream * 0xD1B54A32D192ED03)) & 0xFFFFFFFFFFFFFFFF)
        i = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = base + i * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed: int, stream: int, count: int, start: int = 0, chunk: int = 1 << 24) -> np.ndarray:
    out = np.empty(count, dtype=np.float32)
    for c0 in range(0, count, chunk):
        c = min(chunk, count - c0)
        out[c0:c0 + c] = (splitmix64(seed, stream, start + c0, c) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
    return out


def uniform_points(count: int, k: int, seed: int = 1000, stream: int = 0) -> np.ndarray:
    """AoS float32 [count][k], i.i.d. U[0,1) (main.cu:10-13 in distribution)."""
    return uniform01(seed, stream, count * k).reshape(count, k)


def _normal(seed: int, stream: int, count: int) -> np.ndarray:
    u1 = uniform01(seed, stream, count).astype(np.float64)
    u2 = uniform01(seed, stream + 1, count).astype(np.float64)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def clustered_points(count: int, k: int = 3, seed: int = 1000, stream: int = 3, centres: int = 1024,
                     sigma: float = 0.01, grid_bits: int = 10) -> np.ndarray:
    """Clustered Gaussian points snapped to a 2^-grid_bits grid (exact FP32 ties), SURVEY 8d."""
    cen = uniform_points(centres, k, seed, 2).astype(np.float64)
    which = (splitmix64(seed, stream + 100, 0, count) % np.uint64(centres)).astype(np.int64)
    pts = cen[which]
    for t in range(k):
        pts[:, t] += sigma * _normal(seed, stream * 16 + 2 * t + 200, count)
    scale = float(1 << grid_bits)
    return (np.round(pts * scale) / scale).astype(np.float32)


def clustered_workload(m: int, n: int, k: int = 3, seed: int = 1000):
    """Config C5: clustered refs (stream 3) with duplicated points r[j] = r[j-37] for j % 64 == 0,
    j >= 64; clustered queries (stream 4) where every 2nd query is an exact copy of reference
    (i * 2654435761) mod n.  Returns (queries, refs)."""
    refs = clustered_points(n, k, seed, 3)
    j = np.arange(64, n, 64, dtype=np.int64)
    refs[j] = refs[j - 37]
    queries = clustered_points(m, k, seed, 4)
    i = np.arange(0, m, 2, dtype=np.int64)
    queries[i] = refs[(i * 2654435761) % n]
    return queries, refs


def reference_rand_sample(k: int, m: int, n: int, seed: int = 1000):
    """The reference's literal generator for one sample right after srand(seed)
    (main.cu:24-35, 64): libc rand()/double(RAND_MAX), queries first then references."""
    libc = ctypes.CDLL(None)
    libc.srand(ctypes.c_uint(seed))
    libc.rand.restype = ctypes.c_int
    rand_max = 2147483647.0
    s = np.array([libc.rand() / rand_max for _ in range(k * m)], dtype=np.float64).astype(np.float32)
    r = np.array([libc.rand() / rand_max for _ in range(k * n)], dtype=np.float64).astype(np.float32)
    return s.reshape(m, k), r.reshape(n, k)
