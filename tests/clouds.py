"""Deterministic synthetic inputs shared by tests and bench:
counter-based generators, identical bytes on every machine."""
import numpy as np

F = np.float32


def _hash_u32(seed, n, d):
    """splitmix64-style counter hash -> uint32; element (i, d) depends only on (seed, i, d)."""
    i = np.arange(n, dtype=np.uint64)
    off = np.uint64((int(seed) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)
    x = (i * np.uint64(3) + np.uint64(d)) * np.uint64(0x9E3779B97F4A7C15) + off
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(32)).astype(np.uint32)


def uniform01(seed, n, dims=3):
    """[n, dims] float32 in [0, 1): u = (hash >> 8) * 2^-24."""
    cols = [(_hash_u32(seed, n, d) >> np.uint32(8)).astype(F) * F(2.0 ** -24) for d in range(dims)]
    return np.stack(cols, 1)


def filled_box(seed, n):
    """bvh_driver filled-box cloud: uniform in [-a, a]^3, a = cbrt(n)
    (benchmarks/bvh_driver/benchmark_registration.hpp:129-149)."""
    a = F(np.cbrt(float(n)))
    return (a * (F(2) * uniform01(seed, n) - F(1))).astype(F)


def bvh_driver_radius(k=10):
    """r = cbrt(k * 6 / pi): about k results per query (benchmark_registration.hpp:179-183)."""
    return F(np.cbrt(k * 6.0 / np.pi))


def clustered(seed, n, n_clusters=10, domain=1.0e6, spread=100.0, noise_frac=1e-4, length=50.0):
    """Clustered cloud in the spirit of the GanTao seed spreader used by the reference's DBSCAN
    benchmark (benchmarks/cluster/data_timpl.hpp:253-335): n_clusters elongated dense tubes (a seed
    walking `length * spread` through the domain, scattering points within ~spread of itself) plus a
    fraction of uniform noise.  Counter-based, so CPU and GPU sides get identical bytes."""
    g = uniform01(seed + 1, n, 3) + uniform01(seed + 2, n, 3) + uniform01(seed + 3, n, 3) - F(1.5)
    centers = (uniform01(seed + 4, n_clusters, 3) * F(domain * 0.6) + F(domain * 0.2)).astype(F)
    dirs = uniform01(seed + 7, n_clusters, 3) - F(0.5)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(F)
    cid = (_hash_u32(seed + 5, n, 0) % np.uint32(n_clusters)).astype(np.int64)
    t = uniform01(seed + 6, n, 1) - F(0.5)
    pts = centers[cid] + dirs[cid] * (t * F(length * spread)) + g * F(spread)
    n_noise = int(n * noise_frac)
    if n_noise:
        pts[:n_noise] = uniform01(seed, n_noise, 3) * F(domain)
    return np.clip(pts, 0, domain).astype(F)
