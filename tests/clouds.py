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


# ---- config 5: triangulated sphere (benchmarks/triangulated_surface_distance/generator.hpp) -------------------
def _icosahedron():
    """generator.hpp:29-71: vertices and triangles of the icosahedron, then the edge form of :112-140
    (edges numbered in order of first appearance, a triangle = its three edge indices)."""
    a, b = F((1 + 5 ** 0.5) / 2), F(1)
    v = np.array([[0, b, -a], [b, a, 0], [-b, a, 0], [0, b, a], [0, -b, a], [-a, 0, b], [0, -b, -a], [a, 0, -b],
                  [a, 0, b], [-a, 0, -b], [b, -a, 0], [-b, -a, 0]], F)
    t = [(2, 1, 0), (1, 2, 3), (5, 4, 3), (4, 8, 3), (7, 6, 0), (6, 9, 0), (11, 10, 4), (10, 11, 6), (9, 5, 2),
         (5, 9, 11), (8, 7, 1), (7, 8, 10), (2, 5, 3), (8, 1, 3), (9, 2, 0), (1, 7, 0), (11, 9, 6), (7, 10, 6),
         (5, 11, 4), (10, 8, 4)]
    edges, index, tri_e = [], {}, []
    for tri in t:
        e = []
        for j in range(3):
            key = (min(tri[j], tri[(j + 1) % 3]), max(tri[j], tri[(j + 1) % 3]))
            if key not in index:
                index[key] = len(edges)
                edges.append(key)
            e.append(index[key])
        tri_e.append(e)
    return v, np.array(edges, np.int64), np.array(tri_e, np.int64)


def icosphere(refinements, radius=1.0):
    """Vectorised restatement of buildTriangles("ball") (generator.hpp:176-321): `refinements` rounds of the
    edge-form 1 -> 4 subdivision (:176-241, same vertex / edge / triangle numbering), projection of the vertices
    onto the sphere (:243-257), conversion to vertex form (:142-163).  -> (vertices [V, 3] float32, triangles
    [T, 3] int32); refinements = 10 gives the 20 971 520 triangles / 10 485 762 vertices of BASELINE config 5."""
    v, edges, tris = _icosahedron()
    for _ in range(refinements):
        nv, ne, nt = len(v), len(edges), len(tris)
        mid = ((v[edges[:, 0]] + v[edges[:, 1]]) / F(2)).astype(F)
        v = np.concatenate([v, mid])
        new_edges = np.empty((2 * ne + 3 * nt, 2), np.int64)
        new_edges[0:2 * ne:2, 0] = edges[:, 0]
        new_edges[1:2 * ne:2, 0] = edges[:, 1]
        new_edges[0:2 * ne:2, 1] = nv + np.arange(ne)
        new_edges[1:2 * ne:2, 1] = nv + np.arange(ne)
        new_tris = np.empty((4 * nt, 3), np.int64)
        off = 2 * ne + 3 * np.arange(nt)
        for j in range(3):
            e0, e1 = 2 * tris[:, j], 2 * tris[:, (j + 1) % 3]
            c1 = new_edges[e0, 0] == new_edges[e1 + 1, 0]
            c2 = ~c1 & (new_edges[e0 + 1, 0] == new_edges[e1, 0])
            c3 = ~c1 & ~c2 & (new_edges[e0 + 1, 0] == new_edges[e1 + 1, 0])
            e1 = e1 + (c1 | c3)
            e0 = e0 + (c2 | c3)
            assert bool((new_edges[e0, 0] == new_edges[e1, 0]).all())
            new_edges[off + j, 0] = new_edges[e0, 1]
            new_edges[off + j, 1] = new_edges[e1, 1]
            new_tris[4 * np.arange(nt) + j] = np.stack([e0, e1, off + j], 1)
        new_tris[4 * np.arange(nt) + 3] = np.stack([off, off + 1, off + 2], 1)
        edges, tris = new_edges, new_tris
    norm = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]).astype(F)).astype(F)
    v = (v * (F(radius) / norm)[:, None]).astype(F)
    e0, e1 = edges[tris[:, 0]], edges[tris[:, 1]]
    third = np.where((e0[:, 0] == e1[:, 0]) | (e0[:, 1] == e1[:, 0]), e1[:, 1], e1[:, 0])
    return v, np.stack([e0[:, 0], e0[:, 1], third], 1).astype(np.int32)


def triangle_soup(vertices, triangles):
    """[T, 9] float32: the three corners of every triangle (the flat form the C ABI takes)."""
    return np.ascontiguousarray(vertices[triangles].reshape(-1, 9), F)


def shell_points(seed, n, r_lo=0.9, r_hi=1.1):
    """Query points with norm in [r_lo, r_hi] (second query set of config 5: close to the unit sphere)."""
    g = _normals(seed, n)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    rad = r_lo + (r_hi - r_lo) * uniform01(seed + 3, n, 1).astype(np.float64)
    return (g * rad).astype(F)


def ball_rays(seed, n):
    """Rays for config 5 (modelled on examples/raytracing/example_raytracing.cpp:255-281): origins uniform in the
    unit ball (scaled by 0.9: strictly inside the triangulated sphere), directions uniform on S^2."""
    g = _normals(seed, n)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    o = g * (0.9 * np.cbrt(uniform01(seed + 5, n, 1).astype(np.float64)))
    d = _normals(seed + 9, n)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], 1).astype(F)


# ---- the same filled-box cloud generated with torch (any device): identical bytes, used for the 100M-point runs --
def filled_box_torch(seed, n, device, a=None, first=0):
    """filled_box(seed, n) computed with torch int64 arithmetic (wrap-around multiply, logical shifts emulated):
    bit-identical to the numpy generator; elements [first, first + n) of the stream when `first` is given."""
    import torch
    m64 = (1 << 64) - 1

    def s64(x):  # python int -> the int64 with the same bits
        x &= m64
        return x - (1 << 64) if x >= (1 << 63) else x

    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)

    i = torch.arange(first, first + n, dtype=torch.int64, device=device)
    off = s64(int(seed) * 0xD1B54A32D192ED03)
    cols = []
    for d in range(3):
        x = (i * 3 + d) * s64(0x9E3779B97F4A7C15) + off
        x = x ^ lsr(x, 30)
        x = x * s64(0xBF58476D1CE4E5B9)
        x = x ^ lsr(x, 27)
        x = x * s64(0x94D049BB133111EB)
        x = x ^ lsr(x, 31)
        u = (lsr(x, 32) >> 8).to(torch.float32) * (2.0 ** -24)
        cols.append(u)
    u = torch.stack(cols, 1)
    aa = float(F(np.cbrt(float(n)))) if a is None else float(a)
    return (aa * (2.0 * u - 1.0)).contiguous()
