// abx_traverse.cuh -- predicates and traversal cores shared by the query and
// DBSCAN kernels.  See abx_query.cu for the behavioural contract.
#pragma once
#include "abx_common.cuh"

namespace abx
{

constexpr int kThreads = 128;
constexpr int kStackSize = 96; // >= 63 code bits + 31 index bits of tie-breaking

// ---- predicates ------------------------------------------------------------------
template <int PRED>
struct Pred;

template <>
struct Pred<ABX_PRED_SPHERE3F>
{
  float cx, cy, cz, r, t; // t: largest d2 with sqrt(d2) <= r
  __device__ __forceinline__ void load(float const *__restrict__ p, int64_t i)
  {
    float4 v = *reinterpret_cast<float4 const *>(p + 4 * i);
    cx = v.x;
    cy = v.y;
    cz = v.z;
    r = v.w;
    t = sqrtThreshold(r);
  }
  // intersects(Sphere, Box): distance(centre, box) <= radius (Intersects.hpp:84-91)
  __device__ __forceinline__ bool box(float4 lo, float4 hi) const
  {
    return pointBoxDist2(cx, cy, cz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z) <= t;
  }
  // intersects(Sphere, Point): distance(centre, point) <= radius (Intersects.hpp:107-114);
  // tmp = point - centre, the same value the box path computes for a degenerate box
  __device__ __forceinline__ bool point(float4 p) const
  {
    float tx = __fsub_rn(p.x, cx), ty = __fsub_rn(p.y, cy), tz = __fsub_rn(p.z, cz);
    float d2 = __fmul_rn(tx, tx);
    d2 = __fadd_rn(d2, __fmul_rn(ty, ty));
    d2 = __fadd_rn(d2, __fmul_rn(tz, tz));
    return d2 <= t;
  }
};

template <>
struct Pred<ABX_PRED_BOX3F>
{
  float lx, ly, lz, hx, hy, hz;
  __device__ __forceinline__ void load(float const *__restrict__ p, int64_t i)
  {
    lx = p[6 * i];
    ly = p[6 * i + 1];
    lz = p[6 * i + 2];
    hx = p[6 * i + 3];
    hy = p[6 * i + 4];
    hz = p[6 * i + 5];
  }
  // intersects(Box, Box): Intersects.hpp:53-65
  __device__ __forceinline__ bool box(float4 lo, float4 hi) const
  {
    return !(lx > hi.x || hx < lo.x || ly > hi.y || hy < lo.y || lz > hi.z || hz < lo.z);
  }
  __device__ __forceinline__ bool point(float4 p) const { return box(p, p); }
};

template <>
struct Pred<ABX_PRED_POINT3F>
{
  float x, y, z;
  __device__ __forceinline__ void load(float const *__restrict__ p, int64_t i)
  {
    x = p[3 * i];
    y = p[3 * i + 1];
    z = p[3 * i + 2];
  }
  // intersects(Point, Box): Intersects.hpp:69-80
  __device__ __forceinline__ bool box(float4 lo, float4 hi) const
  {
    return !(x > hi.x || x < lo.x || y > hi.y || y < lo.y || z > hi.z || z < lo.z);
  }
  __device__ __forceinline__ bool point(float4 p) const { return box(p, p); }
};

// Experimental::Ray (geometry/ArborX_Ray.hpp): direction normalised in double (:47-55), slab
// test with explicit +-inf for zero components (:107-157), intersects = hit && tmax >= 0 (:159-167)
template <>
struct Pred<ABX_PRED_RAY3F>
{
  float ox, oy, oz, dx, dy, dz;
  __device__ __forceinline__ void load(float const *__restrict__ p, int64_t i)
  {
    ox = p[6 * i];
    oy = p[6 * i + 1];
    oz = p[6 * i + 2];
    double const gx = p[6 * i + 3], gy = p[6 * i + 4], gz = p[6 * i + 5];
    double const m = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), __dmul_rn(gz, gz)));
    dx = (float)__ddiv_rn(gx, m);
    dy = (float)__ddiv_rn(gy, m);
    dz = (float)__ddiv_rn(gz, m);
  }
  __device__ __forceinline__ void slab(float o, float d, float lo, float hi, float &tmin, float &tmax) const
  {
    float const inf = __int_as_float(0x7f800000);
    float tdmin, tdmax;
    if (d == 0.f)
    {
      float const min_orig = __fsub_rn(lo, o);
      if (min_orig == 0.f)
        return;
      float const max_orig = __fsub_rn(hi, o);
      tdmin = signbit(__fmul_rn(d, max_orig)) ? inf : -inf;
      tdmax = signbit(__fmul_rn(d, min_orig)) ? inf : -inf;
    }
    else if (d > 0.f)
    {
      tdmin = __fdiv_rn(__fsub_rn(lo, o), d);
      tdmax = __fdiv_rn(__fsub_rn(hi, o), d);
    }
    else
    {
      tdmin = __fdiv_rn(__fsub_rn(hi, o), d);
      tdmax = __fdiv_rn(__fsub_rn(lo, o), d);
    }
    if (tmin < tdmin)
      tmin = tdmin;
    if (tmax > tdmax)
      tmax = tdmax;
  }
  __device__ __forceinline__ bool box(float4 lo, float4 hi) const
  {
    float tmin = -__int_as_float(0x7f800000), tmax = __int_as_float(0x7f800000);
    slab(ox, dx, lo.x, hi.x, tmin, tmax);
    slab(oy, dy, lo.y, hi.y, tmin, tmax);
    slab(oz, dz, lo.z, hi.z, tmin, tmax);
    return tmin <= tmax && tmax >= 0.f;
  }
  // distance(Ray, Box) (geometry/ArborX_Ray.hpp:433-444): where the ray enters the box, +inf when it misses
  __device__ __forceinline__ float distance(float4 lo, float4 hi) const
  {
    float const inf = __int_as_float(0x7f800000);
    float tmin = -inf, tmax = inf;
    slab(ox, dx, lo.x, hi.x, tmin, tmax);
    slab(oy, dy, lo.y, hi.y, tmin, tmax);
    slab(oz, dz, lo.z, hi.z, tmin, tmax);
    return (tmin <= tmax && tmax >= 0.f) ? fmaxf(tmin, 0.f) : inf;
  }
  __device__ __forceinline__ bool point(float4 p) const { return box(p, p); }
};

// ---- point-triangle distance (ClosestPoint.hpp:69-153, Distance.hpp:112-123) ----
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
  // misc/ArborX_Vector.hpp dot(): accumulate from 0 in index order, unfused
  float r = __fmul_rn(ax, bx);
  r = __fadd_rn(r, __fmul_rn(ay, by));
  r = __fadd_rn(r, __fmul_rn(az, bz));
  return r;
}

__device__ inline float pointTriangleDist2(float px, float py, float pz, float4 A, float4 B, float4 C)
{
  float abx_ = __fsub_rn(B.x, A.x), aby = __fsub_rn(B.y, A.y), abz = __fsub_rn(B.z, A.z);
  float acx = __fsub_rn(C.x, A.x), acy = __fsub_rn(C.y, A.y), acz = __fsub_rn(C.z, A.z);
  float apx = __fsub_rn(px, A.x), apy = __fsub_rn(py, A.y), apz = __fsub_rn(pz, A.z);
  float qx, qy, qz; // closest point
  float const d1 = dot3(abx_, aby, abz, apx, apy, apz);
  float const d2 = dot3(acx, acy, acz, apx, apy, apz);
  bool done = false;
  float u = 0, v = 0, w = 0;
  if (d1 <= 0 && d2 <= 0)
  {
    qx = A.x, qy = A.y, qz = A.z;
    done = true;
  }
  float d3 = 0, d4 = 0, d5 = 0, d6 = 0;
  if (!done)
  {
    float bpx = __fsub_rn(px, B.x), bpy = __fsub_rn(py, B.y), bpz = __fsub_rn(pz, B.z);
    d3 = dot3(abx_, aby, abz, bpx, bpy, bpz);
    d4 = dot3(acx, acy, acz, bpx, bpy, bpz);
    if (d3 >= 0 && d4 <= d3)
    {
      qx = B.x, qy = B.y, qz = B.z;
      done = true;
    }
  }
  if (!done)
  {
    float cpx = __fsub_rn(px, C.x), cpy = __fsub_rn(py, C.y), cpz = __fsub_rn(pz, C.z);
    d5 = dot3(abx_, aby, abz, cpx, cpy, cpz);
    d6 = dot3(acx, acy, acz, cpx, cpy, cpz);
    if (d6 >= 0 && d5 <= d6)
    {
      qx = C.x, qy = C.y, qz = C.z;
      done = true;
    }
  }
  if (!done)
  {
    bool comb = false;
    float const vc = __fsub_rn(__fmul_rn(d1, d4), __fmul_rn(d3, d2));
    if (vc <= 0 && d1 >= 0 && d3 <= 0)
    {
      float const t = __fdiv_rn(d1, __fsub_rn(d1, d3));
      u = __fsub_rn(1.f, t), v = t, w = 0.f;
      comb = true;
    }
    float vb = 0;
    if (!comb)
    {
      vb = __fsub_rn(__fmul_rn(d5, d2), __fmul_rn(d1, d6));
      if (vb <= 0 && d2 >= 0 && d6 <= 0)
      {
        float const t = __fdiv_rn(d2, __fsub_rn(d2, d6));
        u = __fsub_rn(1.f, t), v = 0.f, w = t;
        comb = true;
      }
    }
    if (!comb)
    {
      float const va = __fsub_rn(__fmul_rn(d3, d6), __fmul_rn(d5, d4));
      float const e43 = __fsub_rn(d4, d3), e56 = __fsub_rn(d5, d6);
      if (va <= 0 && e43 >= 0 && e56 >= 0)
      {
        float const t = __fdiv_rn(e43, __fadd_rn(e43, e56));
        u = 0.f, v = __fsub_rn(1.f, t), w = t;
      }
      else
      {
        float const denom = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(va, vb), vc));
        float const vv = __fmul_rn(vb, denom);
        float const ww = __fmul_rn(vc, denom);
        u = __fsub_rn(__fsub_rn(1.f, vv), ww), v = vv, w = ww;
      }
    }
    // combine(): r[d] = u*a[d] + v*b[d] + w*c[d]
    qx = __fadd_rn(__fadd_rn(__fmul_rn(u, A.x), __fmul_rn(v, B.x)), __fmul_rn(w, C.x));
    qy = __fadd_rn(__fadd_rn(__fmul_rn(u, A.y), __fmul_rn(v, B.y)), __fmul_rn(w, C.y));
    qz = __fadd_rn(__fadd_rn(__fmul_rn(u, A.z), __fmul_rn(v, B.z)), __fmul_rn(w, C.z));
  }
  float tx = __fsub_rn(qx, px), ty = __fsub_rn(qy, py), tz = __fsub_rn(qz, pz);
  float r2 = __fmul_rn(tx, tx);
  r2 = __fadd_rn(r2, __fmul_rn(ty, ty));
  r2 = __fadd_rn(r2, __fmul_rn(tz, tz));
  return r2;
}

// ---- spatial traversal core -----------------------------------------------------
// Subtrees with at most kBucket leaves are not descended: their leaves are
// contiguous in the sorted leaf array (a node's range [lo, hi] of sorted positions
// is known from its parent's record), so they are scanned linearly -- independent
// 16/32-byte loads from one or two cache lines instead of three more levels of
// dependent 64-byte node loads.  Same result set, shorter dependent chain.
constexpr int kBucket = 8;

// LEAF_F4: float4s per sorted leaf (1: points (xyz, orig); 2: boxes (lo, orig)(hi, -))
// emit(orig, pos) returns true to stop the whole traversal (early exit).
template <int LEAF_F4, class P, class Emit>
__device__ __forceinline__ bool scanLeaves(float4 const *__restrict__ leaf_box, int lo, int hi, P const &pred,
                                           Emit &&emit)
{
  for (int j = lo; j <= hi; ++j)
  {
    if (LEAF_F4 == 1)
    {
      float4 const p = __ldg(leaf_box + j);
      if (pred.point(p) && emit(__float_as_uint(p.w), j))
        return true;
    }
    else
    {
      float4 const l = __ldg(leaf_box + 2 * (size_t)j), h = __ldg(leaf_box + 2 * (size_t)j + 1);
      if (pred.box(l, h) && emit(__float_as_uint(l.w), j))
        return true;
    }
  }
  return false;
}

template <int LEAF_F4, int BUCKET = kBucket, class P, class Emit>
__device__ __forceinline__ void traverseSpatial(Node64 const *__restrict__ nodes,
                                                float4 const *__restrict__ leaf_box, P const &pred, Emit &&emit)
{
  int stack[kStackSize];
  int sp = 0;
  int node = 0;
  while (true)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
    float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
    int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
    int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
    // BUCKET == 0: the reference's leaf rule (TreeTraversal.hpp:97-119): a leaf has no box of its own, its value is
    // handed to the predicate whenever its parent is visited.  Needed where the exact leaf test can accept what the
    // leaf's bounding box rejects: ray - triangle, whose watertight test carries tolerances (ArborX_Ray.hpp:340-355).
    bool hit_l = (BUCKET == 0 && refIsLeaf(lref)) || pred.box(a0, a1);
    bool hit_r = (BUCKET == 0 && refIsLeaf(rref)) || pred.box(a2, a3);
    // left child covers [rl, l_hi], right child [r_lo, rr] (an internal left child's
    // Karras index is its last leaf, an internal right child's its first)
    int const l_hi = refIsLeaf(lref) ? rl : lref;
    int const r_lo = refIsLeaf(rref) ? rr : rref;
    if (hit_l)
    {
      if (refIsLeaf(lref))
      {
        if (emit(refOrig(lref), rl))
          return;
        hit_l = false;
      }
      else if (BUCKET > 1 && l_hi - rl < BUCKET)
      {
        if (scanLeaves<LEAF_F4>(leaf_box, rl, l_hi, pred, emit))
          return;
        hit_l = false;
      }
    }
    if (hit_r)
    {
      if (refIsLeaf(rref))
      {
        if (emit(refOrig(rref), rr))
          return;
        hit_r = false;
      }
      else if (BUCKET > 1 && rr - r_lo < BUCKET)
      {
        if (scanLeaves<LEAF_F4>(leaf_box, r_lo, rr, pred, emit))
          return;
        hit_r = false;
      }
    }
    if (hit_l)
    {
      if (hit_r)
        stack[sp++] = rref;
      node = lref;
    }
    else if (hit_r)
      node = rref;
    else
    {
      if (sp == 0)
        return;
      node = stack[--sp];
    }
  }
}

// ---- spatial traversal with deferred leaf tests ------------------------------------
// ncu (profiles/r01_ncu_v2_summary.md) shows the immediate form above spending two thirds of
// its issue slots on leaf tests and emits that run with 2-3 of 32 lanes: a lane reaches a leaf
// at its own time and the rest of the warp waits.  Here a lane only RECORDS the sorted leaf
// positions it has to test (a hit leaf child, or every leaf of a small subtree) in a per-thread
// queue in shared memory and keeps walking internal nodes; when its walk ends (or the queue is
// nearly full) the warp reconverges and all lanes test their queued leaves together, one queue
// slot per iteration.  Same tests on the same leaves, hence the same result set; only the
// order in which a query's results are emitted changes.
// queue: QCAP rows of blockDim.x entries (row-major, so a warp's accesses are conflict-free).
// Every lane of the warp must call this (active = false for lanes without a query).
// after >= 0: half traversal (HalfTraversal.hpp:52-74), only the leaves at sorted positions > after.
template <int LEAF_F4, int BUCKET, int QCAP, class P, class Emit>
__device__ __forceinline__ void traverseSpatialDeferred(Node64 const *__restrict__ nodes,
                                                        float4 const *__restrict__ leaf_box, P const &pred,
                                                        bool active, unsigned *queue, Emit &&emit, int after = -1)
{
  // a queue entry is a run of sorted leaf positions: (first << 2) | (length - 1); n < 2^30
  static_assert(BUCKET >= 1 && BUCKET <= 4, "run length must fit in two bits");
  static_assert(QCAP >= 3, "one node visit (two runs) must fit");
  int const stride = blockDim.x;
  unsigned *const myq = queue + threadIdx.x;
  int stack[kStackSize];
  int sp = 0;
  int node = 0;
  int cnt = 0; // queued runs
  int tot = 0; // queued leaves
  while (true)
  {
    while (active && cnt + 2 <= QCAP)
    {
      float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
      float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
      int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
      int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
      int const l_hi = refIsLeaf(lref) ? rl : lref;
      // a subtree holds leaves > after iff its last position is > after
      bool hit_l = l_hi > after && pred.box(a0, a1);
      bool hit_r = rr > after && pred.box(a2, a3);
      int const l_len1 = l_hi - rl; // leaves of the left child - 1
      int const r_len1 = rr - (refIsLeaf(rref) ? rr : rref);
      if (hit_l && l_len1 < BUCKET)
      {
        int const first = max(rl, after + 1);
        myq[(cnt++) * stride] = ((unsigned)first << 2) | (unsigned)(l_hi - first);
        tot += l_hi - first + 1;
        hit_l = false;
      }
      if (hit_r && r_len1 < BUCKET)
      {
        int const first = max(rr - r_len1, after + 1);
        myq[(cnt++) * stride] = ((unsigned)first << 2) | (unsigned)(rr - first);
        tot += rr - first + 1;
        hit_r = false;
      }
      if (hit_l)
      {
        if (hit_r)
          stack[sp++] = rref;
        node = lref;
      }
      else if (hit_r)
        node = rref;
      else if (sp == 0)
        active = false;
      else
        node = stack[--sp];
    }
    // the warp is converged here: every lane tests its queued leaves, one per iteration
    int const maxt = __reduce_max_sync(0xffffffffu, tot);
    int e = 0, o = 0;
    unsigned cur = myq[0];
    for (int k = 0; k < maxt; ++k)
    {
      if (k < tot)
      {
        int const j = (int)(cur >> 2) + o;
        if (o == (int)(cur & 3u))
        {
          ++e;
          o = 0;
          cur = myq[min(e, QCAP - 1) * stride];
        }
        else
          ++o;
        bool hit;
        unsigned orig;
        if (LEAF_F4 == 1)
        {
          float4 const p = __ldg(leaf_box + j);
          hit = pred.point(p);
          orig = __float_as_uint(p.w);
        }
        else
        {
          float4 const l = __ldg(leaf_box + 2 * (size_t)j), h = __ldg(leaf_box + 2 * (size_t)j + 1);
          hit = pred.box(l, h);
          orig = __float_as_uint(l.w);
        }
        if (hit && emit(orig, j))
        {
          active = false; // early exit (CountUpToN): drop the rest of the queue
          tot = 0;
        }
      }
    }
    cnt = 0;
    tot = 0;
    if (!__any_sync(0xffffffffu, active))
      return;
  }
}

// ---- experimental: the same deferred traversal over 4-wide nodes (Wide64, abx_common.cuh) -----------------------
// Half the dependent node loads of the Node64 walk (scripts/wide_node_study.py).  The quantised child boxes only
// cull; every leaf of a reported run is tested exactly from leaf_box in the converged leaf phase, so the result
// set is the one of traverseSpatialDeferred.
template <int LEAF_F4, int QCAP, class P, class Emit>
__device__ __forceinline__ void traverseWideDeferred(Wide64 const *__restrict__ wide, float4 const *__restrict__ leaf_box,
                                                     P const &pred, bool active, unsigned *queue, Emit &&emit)
{
  static_assert(QCAP >= 5, "one node visit (four runs) must fit");
  int const stride = blockDim.x;
  unsigned *const myq = queue + threadIdx.x;
  // up to three pushes per wide level, (63 code bits + 31 index bits) / 2 wide levels
  int stack[3 * (kStackSize / 2) + 8];
  int sp = 0;
  int node = 0;
  int cnt = 0; // queued runs
  int tot = 0; // queued leaves
  while (true)
  {
    while (active && cnt + 4 <= QCAP)
    {
      uint4 const *w = wide[node].w;
      uint4 const w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
      float const ox = __uint_as_float(w0.x), oy = __uint_as_float(w0.y), oz = __uint_as_float(w0.z);
      float const sx = __uint_as_float(w0.w), sy = __uint_as_float(w1.x), sz = __uint_as_float(w1.y);
      unsigned const q[6] = {w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
      int const ref[4] = {(int)w3.x, (int)w3.y, (int)w3.z, (int)w3.w};
      int next = -1;
#pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        if (ref[k] == kWideEmpty)
          continue;
        // bytes 6k .. 6k + 5 of q: min x y z, max x y z
        constexpr int kB[4] = {0, 6, 12, 18};
        int const b = kB[k];
        float4 lo, hi;
        lo.x = wideLo(wideByte(q[(b + 0) >> 2], (b + 0) & 3), sx, ox);
        lo.y = wideLo(wideByte(q[(b + 1) >> 2], (b + 1) & 3), sy, oy);
        lo.z = wideLo(wideByte(q[(b + 2) >> 2], (b + 2) & 3), sz, oz);
        hi.x = wideHi(wideByte(q[(b + 3) >> 2], (b + 3) & 3), sx, ox);
        hi.y = wideHi(wideByte(q[(b + 4) >> 2], (b + 4) & 3), sy, oy);
        hi.z = wideHi(wideByte(q[(b + 5) >> 2], (b + 5) & 3), sz, oz);
        lo.w = hi.w = 0.f;
        if (!pred.box(lo, hi))
          continue;
        if (ref[k] < 0)
        {
          unsigned const run = (unsigned)~ref[k]; // (first << 2) | (leaves - 1)
          myq[(cnt++) * stride] = run;
          tot += (int)(run & 3u) + 1;
        }
        else
        {
          if (next >= 0)
            stack[sp++] = next;
          next = ref[k];
        }
      }
      if (next >= 0)
        node = next;
      else if (sp == 0)
        active = false;
      else
        node = stack[--sp];
    }
    // the warp is converged here: every lane tests its queued leaves, one per iteration
    int const maxt = __reduce_max_sync(0xffffffffu, tot);
    int e = 0, o = 0;
    unsigned cur = myq[0];
    for (int k = 0; k < maxt; ++k)
    {
      if (k < tot)
      {
        int const j = (int)(cur >> 2) + o;
        if (o == (int)(cur & 3u))
        {
          ++e;
          o = 0;
          cur = myq[min(e, QCAP - 1) * stride];
        }
        else
          ++o;
        bool hit;
        unsigned orig;
        if (LEAF_F4 == 1)
        {
          float4 const p = __ldg(leaf_box + j);
          hit = pred.point(p);
          orig = __float_as_uint(p.w);
        }
        else
        {
          float4 const l = __ldg(leaf_box + 2 * (size_t)j), h = __ldg(leaf_box + 2 * (size_t)j + 1);
          hit = pred.box(l, h);
          orig = __float_as_uint(l.w);
        }
        if (hit && emit(orig, j))
        {
          active = false; // early exit (CountUpToN): drop the rest of the queue
          tot = 0;
        }
      }
    }
    cnt = 0;
    tot = 0;
    if (!__any_sync(0xffffffffu, active))
      return;
  }
}

// exact leaf test for triangle leaves (only sphere predicates are defined in 3-D:
// Intersects.hpp:118-126)
template <int PRED>
__device__ __forceinline__ bool triangleLeafTest(Pred<PRED> const &, float4 const *, int)
{
  return true;
}
template <>
__device__ __forceinline__ bool triangleLeafTest<ABX_PRED_SPHERE3F>(Pred<ABX_PRED_SPHERE3F> const &p,
                                                                     float4 const *__restrict__ leaf_tri, int pos)
{
  float4 A = __ldg(leaf_tri + 3 * (size_t)pos), B = __ldg(leaf_tri + 3 * (size_t)pos + 1),
         C = __ldg(leaf_tri + 3 * (size_t)pos + 2);
  return pointTriangleDist2(p.cx, p.cy, p.cz, A, B, C) <= p.t;
}

// ray-triangle (ArborX_Ray.hpp:266-427): Woop et al. watertight test with a double-precision
// fallback for zero barycentrics and edge hits for coplanar rays; mixed float/double expressions
// reproduced with explicit conversions, nothing contracted
__device__ __forceinline__ void rayRotate2D(float const p[3], float out[3])
{
  float const r = __fsqrt_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])));
  if (p[0] != 0.f)
    out[0] = p[0] > 0.f ? r : -r;
  else
    out[0] = p[1] > 0.f ? r : -r;
  out[1] = p[2];
  out[2] = 0.f;
}
__device__ __forceinline__ bool rayEdgeIntersect(float const v1[3], float const v2[3], float &t)
{
  float const x3 = v1[0], y3 = v1[1], x4 = v2[0], y4 = v2[1];
  float const y2 = fabsf(y3) > fabsf(y4) ? y3 : y4;
  float const det = __fmul_rn(y2, __fsub_rn(x3, x4));
  if (det == 0.f)
    return false;
  t = __fmul_rn(__fdiv_rn(__fsub_rn(__fmul_rn(x3, y4), __fmul_rn(x4, y3)), det), y2);
  float const u = __fdiv_rn(__fmul_rn(x3, y2), det);
  float const epsilon = 0.00001f;
  return (u >= __fsub_rn(0.f, epsilon) && u <= __fadd_rn(1.f, epsilon));
}
__device__ inline bool rayTriangleIntersects(Pred<ABX_PRED_RAY3F> const &ray, float4 TA, float4 TB, float4 TC)
{
  float const dir[3] = {ray.dx, ray.dy, ray.dz};
  float const ta[3] = {TA.x, TA.y, TA.z}, tb[3] = {TB.x, TB.y, TB.z}, tc[3] = {TC.x, TC.y, TC.z};
  float const o[3] = {ray.ox, ray.oy, ray.oz};
  int kz = 0;
  {
    float mx = fabsf(dir[0]);
    for (int i = 1; i < 3; ++i)
    {
      float const f = fabsf(dir[i]);
      if (f > mx)
      {
        mx = f;
        kz = i;
      }
    }
  }
  int kx = (kz + 1) % 3, ky = (kz + 2) % 3;
  if (dir[kz] < 0.f)
  {
    int const tmp = kx;
    kx = ky;
    ky = tmp;
  }
  float s[3];
  s[2] = __fdiv_rn(1.0f, dir[kz]);
  s[0] = __fmul_rn(dir[kx], s[2]);
  s[1] = __fmul_rn(dir[ky], s[2]);
  float oA[3], oB[3], oC[3];
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    oA[d] = __fsub_rn(ta[d], o[d]);
    oB[d] = __fsub_rn(tb[d], o[d]);
    oC[d] = __fsub_rn(tc[d], o[d]);
  }
  float const mag_oA = __fsqrt_rn(dot3(oA[0], oA[1], oA[2], oA[0], oA[1], oA[2]));
  float const mag_oB = __fsqrt_rn(dot3(oB[0], oB[1], oB[2], oB[0], oB[1], oB[2]));
  float const mag_oC = __fsqrt_rn(dot3(oC[0], oC[1], oC[2], oC[0], oC[1], oC[2]));
  double const mag_bar = __ddiv_rn(3.0, (double)__fadd_rn(__fadd_rn(mag_oA, mag_oB), mag_oC));
  float A[3], B[3], C[3];
  A[0] = (float)__dmul_rn((double)__fsub_rn(oA[kx], __fmul_rn(s[0], oA[kz])), mag_bar);
  A[1] = (float)__dmul_rn((double)__fsub_rn(oA[ky], __fmul_rn(s[1], oA[kz])), mag_bar);
  B[0] = (float)__dmul_rn((double)__fsub_rn(oB[kx], __fmul_rn(s[0], oB[kz])), mag_bar);
  B[1] = (float)__dmul_rn((double)__fsub_rn(oB[ky], __fmul_rn(s[1], oB[kz])), mag_bar);
  C[0] = (float)__dmul_rn((double)__fsub_rn(oC[kx], __fmul_rn(s[0], oC[kz])), mag_bar);
  C[1] = (float)__dmul_rn((double)__fsub_rn(oC[ky], __fmul_rn(s[1], oC[kz])), mag_bar);
  float u = __fsub_rn(__fmul_rn(C[0], B[1]), __fmul_rn(C[1], B[0]));
  float v = __fsub_rn(__fmul_rn(A[0], C[1]), __fmul_rn(A[1], C[0]));
  float w = __fsub_rn(__fmul_rn(B[0], A[1]), __fmul_rn(B[1], A[0]));
  if (u == 0.f || v == 0.f || w == 0.f)
  {
    u = (float)__dsub_rn(__dmul_rn((double)C[0], (double)B[1]), __dmul_rn((double)C[1], (double)B[0]));
    v = (float)__dsub_rn(__dmul_rn((double)A[0], (double)C[1]), __dmul_rn((double)A[1], (double)C[0]));
    w = (float)__dsub_rn(__dmul_rn((double)B[0], (double)A[1]), __dmul_rn((double)B[1], (double)A[0]));
  }
  float const inf = __int_as_float(0x7f800000);
  float tmin = inf, tmax = -inf;
  float const epsilon = 0.0000001f;
  if ((u < -epsilon || v < -epsilon || w < -epsilon) && (u > epsilon || v > epsilon || w > epsilon))
    return false;
  float const det = __fadd_rn(__fadd_rn(u, v), w);
  A[2] = __fmul_rn(s[2], oA[kz]);
  B[2] = __fmul_rn(s[2], oB[kz]);
  C[2] = __fmul_rn(s[2], oC[kz]);
  if (det < -epsilon || det > epsilon)
  {
    float const t = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(u, A[2]), __fmul_rn(v, B[2])), __fmul_rn(w, C[2])), det);
    return t >= 0.f; // tmax = tmin = t
  }
  float As[3], Bs[3], Cs[3];
  rayRotate2D(A, As);
  rayRotate2D(B, Bs);
  rayRotate2D(C, Cs);
  float t_ab = inf, t_bc = inf, t_ca = inf;
  bool const ab = rayEdgeIntersect(As, Bs, t_ab);
  if (ab)
  {
    tmin = t_ab;
    tmax = t_ab;
  }
  bool const bc = rayEdgeIntersect(Bs, Cs, t_bc);
  if (bc)
  {
    tmin = fminf(tmin, t_bc);
    tmax = fmaxf(tmax, t_bc);
  }
  bool const ca = rayEdgeIntersect(Cs, As, t_ca);
  if (ca)
  {
    tmin = fminf(tmin, t_ca);
    tmax = fmaxf(tmax, t_ca);
  }
  if (ab || bc || ca)
  {
    if (__fmul_rn(tmin, tmax) <= 0.f)
    {
      tmin = 0.f;
      tmax = 0.f;
    }
    else if (tmin < 0.f)
    {
      float const tmp = tmin;
      tmin = tmax;
      tmax = tmp;
    }
    return tmax >= 0.f;
  }
  return false;
}
template <>
__device__ __forceinline__ bool triangleLeafTest<ABX_PRED_RAY3F>(Pred<ABX_PRED_RAY3F> const &p,
                                                                  float4 const *__restrict__ leaf_tri, int pos)
{
  return rayTriangleIntersects(p, __ldg(leaf_tri + 3 * (size_t)pos), __ldg(leaf_tri + 3 * (size_t)pos + 1),
                               __ldg(leaf_tri + 3 * (size_t)pos + 2));
}

// ---- half traversal ---------------------------------------------------------------
// Leaf i (sorted position) pairs with every leaf j > i within r: the subtrees to
// the right of the root-to-leaf path, which is what starting at rope(i) visits in
// the reference (HalfTraversal.hpp:52-74).  Point leaves only.  emit(orig, pos).
template <class Emit>
__device__ __forceinline__ void traverseHalf(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                                             int i, Pred<ABX_PRED_SPHERE3F> const &pred, Emit &&emit)
{
  int stack[kStackSize];
  int sp = 0;
  int node = 0;
  auto emit_all = [&](unsigned orig, int pos) {
    emit(orig, pos);
    return false;
  };
  while (true)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
    float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
    int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
    int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
    int const l_hi = refIsLeaf(lref) ? rl : lref;
    int const r_lo = refIsLeaf(rref) ? rr : rref;
    // a subtree holds leaves > i iff its last position is > i
    bool hit_l = (l_hi > i) && pred.box(a0, a1);
    bool hit_r = (rr > i) && pred.box(a2, a3);
    if (hit_l)
    {
      if (refIsLeaf(lref))
      {
        emit(refOrig(lref), rl);
        hit_l = false;
      }
      else if (l_hi - rl < kBucket)
      {
        scanLeaves<1>(leaf_box, max(rl, i + 1), l_hi, pred, emit_all);
        hit_l = false;
      }
    }
    if (hit_r)
    {
      if (refIsLeaf(rref))
      {
        emit(refOrig(rref), rr);
        hit_r = false;
      }
      else if (rr - r_lo < kBucket)
      {
        scanLeaves<1>(leaf_box, max(r_lo, i + 1), rr, pred, emit_all);
        hit_r = false;
      }
    }
    if (hit_l)
    {
      if (hit_r)
        stack[sp++] = rref;
      node = lref;
    }
    else if (hit_r)
      node = rref;
    else
    {
      if (sp == 0)
        return;
      node = stack[--sp];
    }
  }
}

} // namespace abx
