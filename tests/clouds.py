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


CLOUD_KINDS = {"filled_box": 0, "hollow_box": 1, "filled_sphere": 2, "hollow_sphere": 3}


def point_cloud(kind, seed, n):
    """The four bvh_driver cloud kinds (benchmarks/utils/ArborXBenchmark_PointClouds.hpp:25-178), scaled by
    a = cbrt(n) like constructPoints (benchmark_registration.hpp:129-149).  Unit shapes: filled_box uniform in
    [-1, 1]^3; hollow_box on the faces (point i on axis (i / 2) % 3, side i % 2); filled_sphere a normal
    direction scaled by u^(1/3); hollow_sphere a normalised normal direction."""
    a = F(np.cbrt(float(n)))
    if kind == "filled_box":
        return filled_box(seed, n)
    if kind == "hollow_box":
        p = F(2) * uniform01(seed, n) - F(1)
        i = np.arange(n)
        p[i, (i // 2) % 3] = np.where(i % 2 == 0, F(-1), F(1))
        return (a * p).astype(F)
    g = _normals(seed + 31, n)
    norm = np.linalg.norm(g, axis=1, keepdims=True)
    norm[norm == 0] = 1.0
    if kind == "filled_sphere":
        scale = np.cbrt(uniform01(seed + 37, n, 1).astype(np.float64)) / norm
        return (a * (g * scale)).astype(F)
    if kind == "hollow_sphere":
        return (a * (g / norm)).astype(F)
    raise ValueError("unknown point cloud kind " + str(kind))


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


def _normals(seed, n, dims=3):
    """[n, dims] standard normals (Box-Muller on the counter hash)."""
    u1 = np.maximum(uniform01(seed, n, dims).astype(np.float64), 1e-12)
    u2 = uniform01(seed + 7919, n, dims).astype(np.float64)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def gan_tao(seed, n, variable_density=False, num_clusters=10, c_reset=100, rho_noise=1e-4, L=1.0e6):
    """Vectorised restatement of the GanTao seed spreader of the reference's DBSCAN benchmark
    (benchmarks/cluster/data_timpl.hpp:253-335): a seed random-walks through [0, L]^3 (step r_shift
    every c_reset points, restart at a random location with probability (num_clusters-1)/n per point)
    and scatters points uniformly in the ball of radius r_vicinity around itself; a fraction rho_noise
    of uniform noise is appended.  3-D constants: r_vicinity = 100 (x (restart % 10 + 1) with
    variable_density), r_shift = 1.5 r_vicinity.  The reference's std::default_random_engine streams are
    implementation-defined, so the counter-based hash of this module is used instead; parity only needs
    identical bytes on the CPU and GPU sides."""
    n_wo = n - int(n * rho_noise)
    rho_restart = (num_clusters - 1) / max(n_wo, 1)
    restart_after = uniform01(seed, n_wo, 1)[:, 0].astype(np.float64) < rho_restart
    start = np.zeros(n_wo, bool)
    start[0] = True
    start[1:] = restart_after[:-1]
    cid = np.cumsum(start) - 1
    first = np.nonzero(start)[0]
    j = np.arange(n_wo) - first[cid]
    new_seg = start | (j % c_reset == 0)
    seg = np.cumsum(new_seg) - 1
    seg_first = np.nonzero(new_seg)[0]
    n_seg = len(seg_first)
    seg_cluster = cid[seg_first]
    seg_is_cluster_first = start[seg_first]
    dens = (seg_cluster % 10 + 1) if variable_density else np.ones(n_seg, np.int64)
    r_vic_seg = 100.0 * dens
    r_shift_seg = 1.5 * r_vic_seg
    d = _normals(seed + 11, n_seg)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    shift = d * r_shift_seg[:, None]
    shift[seg_is_cluster_first] = 0.0
    pos = np.cumsum(shift, 0)
    n_clu = int(cid[-1]) + 1
    origin_c = uniform01(seed + 13, n_clu, 3).astype(np.float64) * L
    clu_first_seg = np.nonzero(seg_is_cluster_first)[0]
    offset_c = origin_c - pos[clu_first_seg]
    origin_seg = pos + offset_c[seg_cluster]
    b = _normals(seed + 17, n_wo)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    rad = np.cbrt(uniform01(seed + 19, n_wo, 1).astype(np.float64)) * r_vic_seg[seg][:, None]
    pts = origin_seg[seg] + b * rad
    noise = uniform01(seed + 23, n - n_wo, 3).astype(np.float64) * L
    return np.concatenate([pts, noise]).astype(F)
