/*
 * oracle/arborx_oracle.cpp -- CPU restatement of the ArborX hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the CPU
 * baseline of bench.py; it is never linked into, imported by or called from
 * the product (arborx_b200/, include/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  The real reference cannot be compiled in this image
 * (header-only over Kokkos >= 4.5, which is absent), so this is
 * a from-scratch restatement of the algorithms in the reference headers, and it
 * is pinned by the reference's own golden vectors (tests/test_oracle_golden.py;
 * SURVEY.md App. B).  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/src unless stated otherwise).
 *
 * Build: g++ -O3 -march=native -fopenmp -ffp-contract=off -shared -fPIC
 * (-ffp-contract=off matters: the reference's predicates are float compares
 * of unfused mul/add chains, geometry/algorithms/ArborX_Distance.hpp:54-70).
 */
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <parallel/algorithm>
#include <vector>

#include <omp.h>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace
{

// ---------------------------------------------------------------------------
// Geometry value types (geometry/ArborX_Point.hpp:24-38, ArborX_Box.hpp:32-67,
// ArborX_Sphere.hpp:25-47, ArborX_Triangle.hpp:21-26), all Point<3,float>.
// ---------------------------------------------------------------------------
struct P3
{
  float v[3];
  float operator[](int d) const { return v[d]; }
  float &operator[](int d) { return v[d]; }
};

struct Box3
{
  // "empty" box: min = +FLT_MAX, max = -FLT_MAX (ArborX_Box.hpp:35-44)
  P3 lo{{FLT_MAX, FLT_MAX, FLT_MAX}};
  P3 hi{{-FLT_MAX, -FLT_MAX, -FLT_MAX}};
};

struct Tri3
{
  P3 a, b, c;
};

enum PrimKind
{
  PRIM_POINT = 0, // 3 floats
  PRIM_BOX = 1,   // 6 floats: min xyz, max xyz
  PRIM_TRI = 2    // 9 floats: a, b, c
};
static int primStride(int kind) { return kind == PRIM_POINT ? 3 : kind == PRIM_BOX ? 6 : 9; }

enum PredKind
{
  PRED_SPHERE = 0, // 4 floats: centre xyz, radius       intersects(Sphere)
  PRED_BOX = 1,    // 6 floats                            intersects(Box)
  PRED_POINT = 2,  // 3 floats                            intersects(Point) / nearest(Point)
  PRED_RAY = 3     // 6 floats: origin xyz, direction xyz intersects(Experimental::Ray)
};
static int predStride(int kind) { return kind == PRED_SPHERE ? 4 : (kind == PRED_BOX || kind == PRED_RAY) ? 6 : 3; }

// geometry/algorithms/ArborX_Expand.hpp:44-107
static inline void expand(Box3 &b, P3 const &p)
{
  for (int d = 0; d < 3; ++d)
  {
    b.lo[d] = std::min(b.lo[d], p[d]);
    b.hi[d] = std::max(b.hi[d], p[d]);
  }
}
static inline void expand(Box3 &b, Box3 const &o)
{
  for (int d = 0; d < 3; ++d)
  {
    b.lo[d] = std::min(b.lo[d], o.lo[d]);
    b.hi[d] = std::max(b.hi[d], o.hi[d]);
  }
}
static inline void expand(Box3 &b, Tri3 const &t)
{
  expand(b, t.a);
  expand(b, t.b);
  expand(b, t.c);
}

// geometry/algorithms/ArborX_Centroid.hpp:41-82
static inline P3 centroid(P3 const &p) { return p; }
static inline P3 centroid(Box3 const &b)
{
  P3 c = b.lo;
  for (int d = 0; d < 3; ++d)
    c[d] = (c[d] + b.hi[d]) / 2;
  return c;
}
static inline P3 centroid(Tri3 const &t)
{
  P3 c = t.a;
  for (int d = 0; d < 3; ++d)
    c[d] = (c[d] + t.b[d] + t.c[d]) / 3;
  return c;
}

// geometry/algorithms/ArborX_Distance.hpp:54-70  (point-point)
static inline float distance(P3 const &a, P3 const &b)
{
  float d2 = 0;
  for (int d = 0; d < 3; ++d)
  {
    float tmp = b[d] - a[d];
    d2 += tmp * tmp;
  }
  return std::sqrt(d2);
}

// geometry/algorithms/ArborX_ClosestPoint.hpp:49-66 (point-box)
static inline P3 closestPoint(P3 const &p, Box3 const &b)
{
  P3 r;
  for (int d = 0; d < 3; ++d)
  {
    if (p[d] < b.lo[d])
      r[d] = b.lo[d];
    else if (p[d] > b.hi[d])
      r[d] = b.hi[d];
    else
      r[d] = p[d];
  }
  return r;
}

static inline P3 sub(P3 const &a, P3 const &b) { return P3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
// misc/ArborX_Vector.hpp dot(): accumulates from 0 in index order
static inline float dot(P3 const &a, P3 const &b)
{
  float r = 0;
  for (int d = 0; d < 3; ++d)
    r += a[d] * b[d];
  return r;
}

// geometry/algorithms/ArborX_ClosestPoint.hpp:69-153 (point-triangle, 7 zones)
static inline P3 combine(P3 const &a, P3 const &b, P3 const &c, float u, float v, float w)
{
  P3 r;
  for (int d = 0; d < 3; ++d)
    r[d] = u * a[d] + v * b[d] + w * c[d];
  return r;
}
static inline P3 closestPoint(P3 const &p, Tri3 const &t)
{
  P3 const &a = t.a, &b = t.b, &c = t.c;
  P3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
  float d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0)
    return a;
  P3 bp = sub(p, b);
  float d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3)
    return b;
  P3 cp = sub(p, c);
  float d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6)
    return c;
  float vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0)
  {
    float v = d1 / (d1 - d3);
    return combine(a, b, c, 1 - v, v, 0);
  }
  float vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0)
  {
    float v = d2 / (d2 - d6);
    return combine(a, b, c, 1 - v, 0, v);
  }
  float va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0)
  {
    float v = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    return combine(a, b, c, 0, 1 - v, v);
  }
  float denom = 1 / (va + vb + vc);
  float v = vb * denom;
  float w = vc * denom;
  return combine(a, b, c, 1 - v - w, v, w);
}

// geometry/algorithms/ArborX_Distance.hpp:72-80, 112-123
static inline float distance(P3 const &p, Box3 const &b) { return distance(p, closestPoint(p, b)); }
static inline float distance(P3 const &p, Tri3 const &t) { return distance(p, closestPoint(p, t)); }

// geometry/algorithms/ArborX_Intersects.hpp:53-65 (box-box), :69-80 (point-box)
static inline bool intersects(Box3 const &b, Box3 const &o)
{
  for (int d = 0; d < 3; ++d)
    if (b.lo[d] > o.hi[d] || b.hi[d] < o.lo[d])
      return false;
  return true;
}
static inline bool intersects(P3 const &p, Box3 const &b)
{
  for (int d = 0; d < 3; ++d)
    if (p[d] > b.hi[d] || p[d] < b.lo[d])
      return false;
  return true;
}

// ---------------------------------------------------------------------------
// Rays: geometry/ArborX_Ray.hpp.  Unfused float arithmetic; the mixed float/double
// expressions of the reference are reproduced with explicit conversions.
// ---------------------------------------------------------------------------
struct Ray3
{
  P3 o;
  P3 dir; // normalised in double at construction (ArborX_Ray.hpp:47-55, misc/ArborX_Vector.hpp:77-87)
};
static inline Ray3 makeRay(float const *g)
{
  Ray3 r;
  r.o = P3{{g[0], g[1], g[2]}};
  double m = 0;
  for (int d = 0; d < 3; ++d)
    m += (double)g[3 + d] * (double)g[3 + d];
  m = std::sqrt(m);
  for (int d = 0; d < 3; ++d)
    r.dir[d] = (float)((double)g[3 + d] / m);
  return r;
}
// :107-157
static inline bool rayBoxIntersection(Ray3 const &ray, Box3 const &box, float &tmin, float &tmax)
{
  float const inf = std::numeric_limits<float>::infinity();
  tmin = -inf;
  tmax = inf;
  for (int d = 0; d < 3; ++d)
  {
    float tdmin, tdmax;
    if (ray.dir[d] == 0)
    {
      float const min_orig = box.lo[d] - ray.o[d];
      if (min_orig == 0)
        continue;
      float const max_orig = box.hi[d] - ray.o[d];
      tdmin = std::signbit(ray.dir[d] * max_orig) ? inf : -inf;
      tdmax = std::signbit(ray.dir[d] * min_orig) ? inf : -inf;
    }
    else if (ray.dir[d] > 0)
    {
      tdmin = (box.lo[d] - ray.o[d]) / ray.dir[d];
      tdmax = (box.hi[d] - ray.o[d]) / ray.dir[d];
    }
    else
    {
      tdmin = (box.hi[d] - ray.o[d]) / ray.dir[d];
      tdmax = (box.lo[d] - ray.o[d]) / ray.dir[d];
    }
    if (tmin < tdmin)
      tmin = tdmin;
    if (tmax > tdmax)
      tmax = tdmax;
  }
  return tmin <= tmax;
}
// :159-167
static inline bool intersects(Ray3 const &ray, Box3 const &box)
{
  float tmin, tmax;
  return rayBoxIntersection(ray, box, tmin, tmax) && (tmax >= 0);
}
// :189-212
static inline void rotate2D(float const p[3], float out[3])
{
  float r = std::sqrt(p[0] * p[0] + p[1] * p[1]);
  if (p[0] != 0)
    out[0] = (p[0] > 0 ? 1 : -1) * r;
  else
    out[0] = (p[1] > 0 ? 1 : -1) * r;
  out[1] = p[2];
  out[2] = 0;
}
// :219-249
static inline bool rayEdgeIntersect(float const v1[3], float const v2[3], float &t)
{
  float x3 = v1[0], y3 = v1[1], x4 = v2[0], y4 = v2[1];
  float y2 = std::fabs(y3) > std::fabs(y4) ? y3 : y4;
  float det = y2 * (x3 - x4);
  if (det == 0)
    return false;
  t = (x3 * y4 - x4 * y3) / det * y2;
  float u = x3 * y2 / det;
  float const epsilon = 0.00001f;
  return (u >= 0 - epsilon && u <= 1 + epsilon);
}
// :266-417 (Woop et al. watertight test, with the coplanar case returning edge hits)
static inline bool rayTriIntersection(Ray3 const &ray, Tri3 const &tri, float &tmin, float &tmax)
{
  P3 const &dir = ray.dir;
  int kz = 0;
  {
    float mx = std::abs(dir[0]);
    for (int i = 1; i < 3; ++i)
    {
      float f = std::fabs(dir[i]);
      if (f > mx)
      {
        mx = f;
        kz = i;
      }
    }
  }
  int kx = (kz + 1) % 3, ky = (kz + 2) % 3;
  if (dir[kz] < 0)
    std::swap(kx, ky);
  float s[3];
  s[2] = 1.0f / dir[kz];
  s[0] = dir[kx] * s[2];
  s[1] = dir[ky] * s[2];
  P3 const oA = sub(tri.a, ray.o), oB = sub(tri.b, ray.o), oC = sub(tri.c, ray.o);
  float const mag_oA = std::sqrt(dot(oA, oA)), mag_oB = std::sqrt(dot(oB, oB)), mag_oC = std::sqrt(dot(oC, oC));
  double const mag_bar = 3.0 / (double)(mag_oA + mag_oB + mag_oC);
  float A[3], B[3], C[3];
  A[0] = (float)((double)(oA[kx] - s[0] * oA[kz]) * mag_bar);
  A[1] = (float)((double)(oA[ky] - s[1] * oA[kz]) * mag_bar);
  B[0] = (float)((double)(oB[kx] - s[0] * oB[kz]) * mag_bar);
  B[1] = (float)((double)(oB[ky] - s[1] * oB[kz]) * mag_bar);
  C[0] = (float)((double)(oC[kx] - s[0] * oC[kz]) * mag_bar);
  C[1] = (float)((double)(oC[ky] - s[1] * oC[kz]) * mag_bar);
  float u = C[0] * B[1] - C[1] * B[0];
  float v = A[0] * C[1] - A[1] * C[0];
  float w = B[0] * A[1] - B[1] * A[0];
  if (u == 0 || v == 0 || w == 0)
  {
    u = (float)((double)C[0] * B[1] - (double)C[1] * B[0]);
    v = (float)((double)A[0] * C[1] - (double)A[1] * C[0]);
    w = (float)((double)B[0] * A[1] - (double)B[1] * A[0]);
  }
  float const inf = std::numeric_limits<float>::infinity();
  tmin = inf;
  tmax = -inf;
  float const epsilon = 0.0000001f;
  if ((u < -epsilon || v < -epsilon || w < -epsilon) && (u > epsilon || v > epsilon || w > epsilon))
    return false;
  float const det = u + v + w;
  A[2] = s[2] * oA[kz];
  B[2] = s[2] * oB[kz];
  C[2] = s[2] * oC[kz];
  if (det < -epsilon || det > epsilon)
  {
    float t = (u * A[2] + v * B[2] + w * C[2]) / det;
    tmax = t;
    tmin = t;
    return true;
  }
  float As[3], Bs[3], Cs[3];
  rotate2D(A, As);
  rotate2D(B, Bs);
  rotate2D(C, Cs);
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
    tmin = std::min(tmin, t_bc);
    tmax = std::max(tmax, t_bc);
  }
  bool const ca = rayEdgeIntersect(Cs, As, t_ca);
  if (ca)
  {
    tmin = std::min(tmin, t_ca);
    tmax = std::max(tmax, t_ca);
  }
  if (ab || bc || ca)
  {
    if (tmin * tmax <= 0)
    {
      tmin = 0;
      tmax = 0;
    }
    else if (tmin < 0)
      std::swap(tmin, tmax);
    return true;
  }
  return false;
}
// :419-427
static inline bool intersects(Ray3 const &ray, Tri3 const &tri)
{
  float tmin, tmax;
  return rayTriIntersection(ray, tri, tmin, tmax) && (tmax >= 0);
}

// ---------------------------------------------------------------------------
// Morton codes: spatial/detail/ArborX_MortonCode.hpp
// ---------------------------------------------------------------------------
// :63-73  expandBitsBy<2>(unsigned int)
static inline unsigned expandBits2_32(unsigned x)
{
  x &= 0x000003ffu;
  x = (x ^ (x << 16)) & 0xff0000ffu;
  x = (x ^ (x << 8)) & 0x0300f00fu;
  x = (x ^ (x << 4)) & 0x030c30c3u;
  x = (x ^ (x << 2)) & 0x09249249u;
  return x;
}
// :187-197 expandBitsBy<2>(unsigned long long)
static inline unsigned long long expandBits2_64(unsigned long long x)
{
  x &= 0x1fffffllu;
  x = (x | x << 32) & 0x1f00000000ffffllu;
  x = (x | x << 16) & 0x1f0000ff0000ffllu;
  x = (x | x << 8) & 0x100f00f00f00f00fllu;
  x = (x | x << 4) & 0x10c30c30c30c30c3llu;
  x = (x | x << 2) & 0x1249249249249249llu;
  return x;
}
static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (hi < x ? hi : x); }
// :293-316 morton32, DIM=3 -> N = 1<<10
static inline unsigned morton32(P3 const &p)
{
  constexpr unsigned N = 1u << 10;
  unsigned r = 0;
  for (int d = 0; d < 3; ++d)
  {
    float x = clampf(p[d] * N, 0.f, float(N - 1));
    r += expandBits2_32((unsigned)x) << (3 - d - 1);
  }
  return r;
}
// :320-347 morton64, DIM=3 -> N = 1<<21, NeedDouble=false -> float arithmetic
static inline unsigned long long morton64(P3 const &p)
{
  constexpr unsigned long long N = 1llu << 21;
  unsigned long long r = 0;
  for (int d = 0; d < 3; ++d)
  {
    float x = clampf(p[d] * N, 0.f, float(N - 1));
    r += expandBits2_64((unsigned long long)x) << (3 - d - 1);
  }
  return r;
}
// geometry/algorithms/ArborX_TranslateAndScale.hpp:25-37
static inline P3 translateAndScale(P3 const &in, Box3 const &ref)
{
  P3 out;
  for (int d = 0; d < 3; ++d)
  {
    float a = ref.lo[d], b = ref.hi[d];
    out[d] = (a != b ? (in[d] - a) / (b - a) : 0);
  }
  return out;
}

// ---------------------------------------------------------------------------
// The tree, in the reference layout (spatial/detail/ArborX_Node.hpp:24-44,
// ArborX_HappyTreeFriends.hpp:26-85): leaves 0..n-1 in sorted order, internal
// nodes n..2n-2 stored at [i-n], root = n.
// ---------------------------------------------------------------------------
struct Tree
{
  int kind = PRIM_POINT;
  int n = 0;
  std::vector<float> prims; // copy of the user's primitives
  std::vector<unsigned long long> codes; // sorted Morton64 codes
  std::vector<unsigned> perm;            // leaf i holds original index perm[i]
  std::vector<int> leaf_rope;
  std::vector<int> left_child, rope;
  std::vector<Box3> box;
  Box3 bounds;

  P3 point(unsigned o) const { return P3{{prims[3 * o], prims[3 * o + 1], prims[3 * o + 2]}}; }
  Box3 pbox(unsigned o) const
  {
    Box3 b;
    for (int d = 0; d < 3; ++d)
    {
      b.lo[d] = prims[6 * o + d];
      b.hi[d] = prims[6 * o + 3 + d];
    }
    return b;
  }
  Tri3 tri(unsigned o) const
  {
    Tri3 t;
    for (int d = 0; d < 3; ++d)
    {
      t.a[d] = prims[9 * o + d];
      t.b[d] = prims[9 * o + 3 + d];
      t.c[d] = prims[9 * o + 6 + d];
    }
    return t;
  }
  Box3 primBox(unsigned o) const
  {
    Box3 b;
    if (kind == PRIM_POINT)
      expand(b, point(o));
    else if (kind == PRIM_BOX)
      expand(b, pbox(o));
    else
      expand(b, tri(o));
    return b;
  }
  P3 primCentroid(unsigned o) const
  {
    if (kind == PRIM_POINT)
      return centroid(point(o));
    if (kind == PRIM_BOX)
      return centroid(pbox(o));
    return centroid(tri(o));
  }
  bool isLeaf(int i) const { return i < n; }
  int getRope(int i) const { return isLeaf(i) ? leaf_rope[i] : rope[i - n]; }
};

// spatial/detail/ArborX_TreeConstruction.hpp:27-39
static Box3 sceneBounds(Tree const &t)
{
  int const n = t.n;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma omp parallel for reduction(min : lo[:3]) reduction(max : hi[:3])
  for (int i = 0; i < n; ++i)
  {
    Box3 b = t.primBox(i);
    for (int d = 0; d < 3; ++d)
    {
      lo[d] = std::min(lo[d], b.lo[d]);
      hi[d] = std::max(hi[d], b.hi[d]);
    }
  }
  Box3 r;
  for (int d = 0; d < 3; ++d)
  {
    r.lo[d] = lo[d];
    r.hi[d] = hi[d];
  }
  return r;
}

// spatial/detail/ArborX_SpaceFillingCurves.hpp:45-57 (Morton64::operator()) and
// :67-84 (projectOntoSpaceFillingCurve)
static void computeCodes(Tree &t, Box3 const &scene)
{
  int const n = t.n;
  t.codes.resize(n);
#pragma omp parallel for
  for (int i = 0; i < n; ++i)
    t.codes[i] = morton64(translateAndScale(t.primCentroid(i), scene));
}

// misc/ArborX_SortUtils.hpp:28-43 (sortObjects): iota + sortByKey.  Order among
// equal keys is unspecified in the reference (kokkos_ext/ArborX_KokkosExtSort.hpp
// :80-148); we fix STABLE (ties by original index), as the product does.
template <class Key>
static void sortObjects(std::vector<Key> &keys, std::vector<unsigned> &perm)
{
  size_t const n = keys.size();
  std::vector<std::pair<Key, unsigned>> kv(n);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i)
    kv[i] = {keys[i], (unsigned)i};
  __gnu_parallel::stable_sort(kv.begin(), kv.end(),
                              [](auto const &a, auto const &b) { return a.first < b.first; });
  perm.resize(n);
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i)
  {
    keys[i] = kv[i].first;
    perm[i] = kv[i].second;
  }
}

// spatial/detail/ArborX_TreeConstruction.hpp:74-322  (GenerateHierarchy)
struct Hierarchy
{
  Tree &t;
  int n_int;
  std::vector<int> ranges; // :97-102, UNTOUCHED_NODE = -1

  explicit Hierarchy(Tree &tree)
      : t(tree)
      , n_int(tree.n - 1)
      , ranges(tree.n - 1, -1)
  {}

  int internalIndex(int i) const { return i + n_int + 1; } // :128

  // :131-170
  long long delta(int i) const
  {
    if (i < 0 || i >= n_int)
      return LLONG_MAX;
    unsigned long long const x = t.codes[i] ^ t.codes[i + 1];
    return (long long)(x + (unsigned long long)(!x) * (unsigned long long)(LLONG_MIN + (long long)(i ^ (i + 1))) - 1);
  }

  // :173-195
  int ropeOf(int range_right, long long delta_right) const
  {
    if (range_right != n_int)
      return (delta_right < delta(range_right + 1) ? range_right + 1 : internalIndex(range_right + 1));
    return -1;
  }

  // :197-311
  void operator()(int i)
  {
    unsigned const original_index = t.perm[i];
    Box3 bv = t.primBox(original_index);

    int range_left = i, range_right = i;
    long long delta_left = delta(range_left - 1);
    long long delta_right = delta(range_right);

    t.leaf_rope[i] = ropeOf(range_right, delta_right);

    int const root = internalIndex(0);
    do
    {
      bool const is_left_child = (delta_right < delta_left);
      int left_child;
      if (is_left_child)
      {
        int const apetrei_parent = range_right;
        int expected = -1;
        // :241-249 atomic_compare_exchange returns the old value
        if (__atomic_compare_exchange_n(&ranges[apetrei_parent], &expected, range_left, false, __ATOMIC_ACQ_REL,
                                        __ATOMIC_ACQUIRE))
          break;
        range_right = expected;

        left_child = i;
        int const right_child = apetrei_parent + 1;
        bool const right_child_is_leaf = (right_child == range_right);
        delta_right = delta(range_right);
        if (right_child_is_leaf)
          expand(bv, t.primBox(t.perm[right_child]));
        else
          expand(bv, t.box[right_child]);
      }
      else
      {
        int const apetrei_parent = range_left - 1;
        int expected = -1;
        if (__atomic_compare_exchange_n(&ranges[apetrei_parent], &expected, range_right, false, __ATOMIC_ACQ_REL,
                                        __ATOMIC_ACQUIRE))
          break;
        range_left = expected;

        left_child = apetrei_parent;
        bool const left_child_is_leaf = (left_child == range_left);
        delta_left = delta(range_left - 1);
        if (left_child_is_leaf)
          expand(bv, t.primBox(t.perm[left_child]));
        else
          expand(bv, t.box[left_child]);
        if (!left_child_is_leaf)
          left_child = internalIndex(left_child);
      }

      int const karras_parent = delta_right < delta_left ? range_right : range_left;
      t.left_child[karras_parent] = left_child;
      t.rope[karras_parent] = ropeOf(range_right, delta_right);
      t.box[karras_parent] = bv;
      i = internalIndex(karras_parent);
    } while (i != root);
  }
};

// codes must be sorted and perm set.  Fills the node arrays.
static void generateHierarchy(Tree &t)
{
  int const n = t.n;
  t.leaf_rope.assign(n, -1);
  t.left_child.assign(n - 1, -1);
  t.rope.assign(n - 1, -1);
  t.box.assign(n - 1, Box3{});
  Hierarchy h(t);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i)
    h(i);
  t.bounds = t.box[0]; // :108-113 root box copied to host
}

// spatial/ArborX_LinearBVH.hpp:171-256
static void buildTree(Tree &t)
{
  int const n = t.n;
  t.bounds = Box3{};
  if (n == 0)
    return; // :203-206
  if (n == 1)
  { // :208-213, TreeConstruction.hpp:43-69
    t.perm.assign(1, 0u);
    t.codes.assign(1, 0ull);
    t.leaf_rope.assign(1, -1);
    t.bounds = t.primBox(0);
    return;
  }
  Box3 scene = sceneBounds(t);
  computeCodes(t, scene);
  sortObjects(t.codes, t.perm);
  generateHierarchy(t);
}

// ---------------------------------------------------------------------------
// Predicates
// ---------------------------------------------------------------------------
struct Pred
{
  int kind;
  float const *g;
  P3 center() const { return P3{{g[0], g[1], g[2]}}; }
  Box3 pbox() const
  {
    Box3 b;
    for (int d = 0; d < 3; ++d)
    {
      b.lo[d] = g[d];
      b.hi[d] = g[3 + d];
    }
    return b;
  }
  // Intersects<G>::operator()(Box)  (Predicates.hpp:83-101 -> Intersects.hpp)
  bool operator()(Box3 const &b) const
  {
    if (kind == PRED_RAY)
      return intersects(makeRay(g), b); // ArborX_Ray.hpp:159-167
    if (kind == PRED_SPHERE)
      return distance(center(), b) <= g[3]; // Intersects.hpp:84-91
    if (kind == PRED_BOX)
      return intersects(pbox(), b); // :53-65
    return intersects(center(), b); // :69-80
  }
  bool operator()(P3 const &p) const
  {
    if (kind == PRED_RAY)
    { // not defined by the reference for points; a point is the degenerate box [p, p]
      Box3 b;
      b.lo = p;
      b.hi = p;
      return intersects(makeRay(g), b);
    }
    if (kind == PRED_SPHERE)
      return distance(center(), p) <= g[3]; // Intersects.hpp:107-114
    if (kind == PRED_BOX)
      return intersects(p, pbox());
    return center()[0] == p[0] && center()[1] == p[1] && center()[2] == p[2];
  }
  bool operator()(Tri3 const &t) const
  {
    if (kind == PRED_RAY)
      return intersects(makeRay(g), t); // ArborX_Ray.hpp:419-427
    // sphere-triangle (Intersects.hpp:118-126)
    return distance(center(), t) <= g[3];
  }
  P3 centroidOf() const
  {
    if (kind == PRED_BOX)
      return centroid(pbox());
    return center();
  }
};

static inline bool leafTest(Tree const &t, Pred const &p, int leaf)
{
  unsigned o = t.perm[leaf];
  if (t.kind == PRIM_POINT)
    return p(t.point(o));
  if (t.kind == PRIM_BOX)
    return p(t.pbox(o));
  return p(t.tri(o));
}

struct Counters
{
  long long internal_tests = 0; // I_q summed
  long long leaf_tests = 0;     // L_q summed
};

// spatial/detail/ArborX_TreeTraversal.hpp:97-119 (rope traversal) and :80-90
// (one-leaf tree).  cb(leaf) returns true for early exit (Callbacks.hpp:24-28).
template <class CB>
static inline void traverseSpatial(Tree const &t, Pred const &p, CB &&cb, Counters *c)
{
  if (t.n == 0)
    return;
  if (t.n == 1)
  {
    if (c)
      c->leaf_tests++;
    if (leafTest(t, p, 0))
      cb(0);
    return;
  }
  int node = t.n;
  do
  {
    if (t.isLeaf(node))
    {
      if (c)
        c->leaf_tests++;
      if (leafTest(t, p, node) && cb(node))
        return;
      node = t.leaf_rope[node];
    }
    else
    {
      if (c)
        c->internal_tests++;
      node = p(t.box[node - t.n]) ? t.left_child[node - t.n] : t.rope[node - t.n];
    }
  } while (node != -1);
}

// LinearBVH.hpp:287-298 / CrsGraphWrapperImpl.hpp:407-419: Morton32 of the
// predicate centroid scaled into the tree's bounds, then sortObjects.
static std::vector<unsigned> predicatePermutation(Tree const &t, int kind, float const *preds, int q, bool sort)
{
  std::vector<unsigned> perm(q);
  if (!sort)
  {
    std::iota(perm.begin(), perm.end(), 0u);
    return perm;
  }
  Box3 scene;
  expand(scene, t.bounds);
  std::vector<unsigned> codes(q);
  int const s = predStride(kind);
#pragma omp parallel for
  for (int i = 0; i < q; ++i)
  {
    Pred p{kind, preds + (size_t)s * i};
    codes[i] = morton32(translateAndScale(p.centroidOf(), scene));
  }
  sortObjects(codes, perm);
  return perm;
}

// ---------------------------------------------------------------------------
// kNN: spatial/detail/ArborX_TreeTraversal.hpp:180-335, heap ops from
// misc/ArborX_Heap.hpp:42-128 and misc/ArborX_PriorityQueue.hpp:63-93.
// ---------------------------------------------------------------------------
struct PairID
{
  int first;
  float second;
};
static inline bool lessDist(PairID const &a, PairID const &b) { return a.second < b.second; }

static inline void bubbleUp(PairID *first, long pos, long top, PairID val)
{
  long parent = (pos - 1) / 2;
  while (pos > top && lessDist(first[parent], val))
  {
    first[pos] = first[parent];
    pos = parent;
    parent = (pos - 1) / 2;
  }
  first[pos] = val;
}
static inline void bubbleDown(PairID *first, long pos, long len, PairID val)
{
  long child = 2 * pos + 1;
  if (child + 1 < len && lessDist(first[child], first[child + 1]))
    ++child;
  while (child < len && lessDist(val, first[child]))
  {
    first[pos] = first[child];
    pos = child;
    child = 2 * pos + 1;
    if (child + 1 < len && lessDist(first[child], first[child + 1]))
      ++child;
  }
  first[pos] = val;
}
static inline void pushHeap(PairID *first, PairID *last)
{
  long n = last - first;
  if (n > 1)
  {
    PairID v = first[n - 1];
    bubbleUp(first, n - 1, 0, v);
  }
}
static inline void popHeap(PairID *first, PairID *last)
{
  long n = last - first;
  if (n > 1)
  {
    PairID v = first[0];
    bubbleDown(first, 0, n - 1, first[n - 1]);
    first[n - 1] = v;
  }
}
static inline void sortHeap(PairID *first, PairID *last)
{
  while (first != last)
    popHeap(first, last--);
}

// query geometries of nearest predicates other than a point (spatial/detail/ArborX_Predicates.hpp:58-80 with
// geometry/algorithms/ArborX_Distance.hpp:83-108,166-209, ArborX_Ray.hpp:433-444)
struct SphereQ
{
  P3 c;
  float r;
};
// distance box-box :166-197 (a = query, b = other)
static inline float distance(Box3 const &a, Box3 const &b)
{
  float d2 = 0;
  for (int d = 0; d < 3; ++d)
  {
    if (a.lo[d] > b.hi[d])
    {
      float const delta = a.lo[d] - b.hi[d];
      d2 += delta * delta;
    }
    else if (b.lo[d] > a.hi[d])
    {
      float const delta = b.lo[d] - a.hi[d];
      d2 += delta * delta;
    }
  }
  return std::sqrt(d2);
}
// distance sphere-box :199-209, point-sphere :83-94 (through ReverseDispatch)
static inline float distance(SphereQ const &s, Box3 const &b) { return std::max(distance(s.c, b) - s.r, 0.f); }
static inline float distance(SphereQ const &s, P3 const &p) { return std::max(distance(p, s.c) - s.r, 0.f); }
// distance ray-box ArborX_Ray.hpp:433-444
static inline float distance(Ray3 const &ray, Box3 const &box)
{
  float tmin, tmax;
  bool const hit = rayBoxIntersection(ray, box, tmin, tmax) && (tmax >= 0);
  return hit ? std::max(tmin, 0.f) : std::numeric_limits<float>::infinity();
}
static inline Box3 pointAsBox(P3 const &p)
{
  Box3 b;
  for (int d = 0; d < 3; ++d)
    b.lo[d] = b.hi[d] = p[d];
  return b;
}
static inline float nearestDistance(Tree const &t, Box3 const &q, int node)
{
  if (t.isLeaf(node))
  {
    unsigned o = t.perm[node];
    if (t.kind == PRIM_POINT)
      return distance(t.point(o), q); // distance(Box, Point) -> distance(Point, Box)
    return distance(q, t.pbox(o));
  }
  return distance(q, t.box[node - t.n]);
}
static inline float nearestDistance(Tree const &t, SphereQ const &q, int node)
{
  if (t.isLeaf(node))
  {
    unsigned o = t.perm[node];
    if (t.kind == PRIM_POINT)
      return distance(q, t.point(o));
    return distance(q, t.pbox(o));
  }
  return distance(q, t.box[node - t.n]);
}
static inline float nearestDistance(Tree const &t, Ray3 const &q, int node)
{
  if (t.isLeaf(node))
  {
    unsigned o = t.perm[node];
    return distance(q, t.kind == PRIM_POINT ? pointAsBox(t.point(o)) : t.pbox(o));
  }
  return distance(q, t.box[node - t.n]);
}
static inline float nearestDistance(Tree const &t, P3 const &q, int node)
{
  if (t.isLeaf(node))
  {
    unsigned o = t.perm[node];
    if (t.kind == PRIM_POINT)
      return distance(q, t.point(o));
    if (t.kind == PRIM_BOX)
      return distance(q, t.pbox(o));
    return distance(q, t.tri(o));
  }
  return distance(q, t.box[node - t.n]);
}

// returns number of results written to buf (sorted ascending by distance)
template <class Q>
static int traverseNearest(Tree const &t, Q const &q, int k, PairID *buf, std::vector<int> &stack,
                           std::vector<float> &stack_d, Counters *c)
{
  if (k < 1 || t.n == 0)
    return 0;
  if (t.n == 1)
  { // :168-178 (callback invoked unconditionally)
    buf[0] = PairID{0, nearestDistance(t, q, 0)};
    return 1;
  }
  int heap_size = 0;
  stack.clear();
  stack_d.clear();
  stack.push_back(-1);
  stack_d.push_back(0.f);
  int node = t.n;
  int left_child = 0, right_child = 0;
  float distance_left = 0, distance_right = 0, distance_node = 0;
  float radius = std::numeric_limits<float>::infinity();
  do
  {
    bool traverse_left = false, traverse_right = false;
    if (distance_node < radius)
    {
      left_child = t.left_child[node - t.n];
      right_child = t.getRope(left_child); // HappyTreeFriends.hpp:73-78
      distance_left = nearestDistance(t, q, left_child);
      distance_right = nearestDistance(t, q, right_child);
      if (c)
      {
        (t.isLeaf(left_child) ? c->leaf_tests : c->internal_tests)++;
        (t.isLeaf(right_child) ? c->leaf_tests : c->internal_tests)++;
      }
      if (distance_left < radius)
      {
        if (t.isLeaf(left_child))
        {
          PairID lp{left_child, distance_left};
          if (heap_size < k)
          {
            buf[heap_size++] = lp;
            pushHeap(buf, buf + heap_size);
          }
          else
            bubbleDown(buf, 0, heap_size, lp); // popPush
          if (heap_size == k)
            radius = buf[0].second;
        }
        else
          traverse_left = true;
      }
      if (distance_right < radius)
      {
        if (t.isLeaf(right_child))
        {
          PairID lp{right_child, distance_right};
          if (heap_size < k)
          {
            buf[heap_size++] = lp;
            pushHeap(buf, buf + heap_size);
          }
          else
            bubbleDown(buf, 0, heap_size, lp);
          if (heap_size == k)
            radius = buf[0].second;
        }
        else
          traverse_right = true;
      }
    }
    if (!traverse_left && !traverse_right)
    {
      node = stack.back();
      stack.pop_back();
      distance_node = stack_d.back();
      stack_d.pop_back();
    }
    else
    {
      node = (traverse_left && (distance_left <= distance_right || !traverse_right)) ? left_child : right_child;
      distance_node = (node == left_child ? distance_left : distance_right);
      if (traverse_left && traverse_right)
      {
        stack.push_back(node == left_child ? right_child : left_child);
        stack_d.push_back(node == left_child ? distance_right : distance_left);
      }
    }
  } while (node != -1);
  sortHeap(buf, buf + heap_size);
  return heap_size;
}

// ---------------------------------------------------------------------------
// Union-find: cluster/detail/ArborX_UnionFind.hpp:74-181 (ECL-CC)
// ---------------------------------------------------------------------------
struct UnionFind
{
  int *labels;
  int representative(int i) const
  { // :113-128
    int curr = __atomic_load_n(&labels[i], __ATOMIC_RELAXED);
    if (curr != i)
    {
      int next, prev = i;
      while (curr > (next = __atomic_load_n(&labels[curr], __ATOMIC_RELAXED)))
      {
        __atomic_store_n(&labels[prev], next, __ATOMIC_RELAXED);
        prev = curr;
        curr = next;
      }
    }
    return curr;
  }
  void merge_into(int i, int j) const { __atomic_store_n(&labels[i], representative(j), __ATOMIC_RELAXED); } // :134-135
  void merge(int i, int j) const
  { // :139-181
    int vstat = representative(i);
    int ostat = representative(j);
    while (vstat != ostat)
    {
      if (vstat < ostat)
      {
        int expected = ostat;
        __atomic_compare_exchange_n(&labels[ostat], &expected, vstat, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
        ostat = expected; // atomic_compare_exchange returns the old value
      }
      else
      {
        int expected = vstat;
        __atomic_compare_exchange_n(&labels[vstat], &expected, ostat, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
        vstat = expected;
      }
    }
  }
};

// ---------------------------------------------------------------------------
// DenseBox grid: cluster/detail/ArborX_CartesianGrid.hpp:27-142
// ---------------------------------------------------------------------------
struct CartesianGrid
{
  Box3 bounds;
  float h[3];
  size_t n[3];
  int error = 0; // 1: loss-of-precision guard (:121-136)
  CartesianGrid(Box3 const &b, float hh)
      : bounds(b)
  {
    for (int d = 0; d < 3; ++d)
      h[d] = hh;
    for (int d = 0; d < 3; ++d)
    { // :92-108
      float delta = bounds.hi[d] - bounds.lo[d];
      if (delta != 0)
        n[d] = (size_t)std::ceil(delta / h[d]);
      else
        n[d] = 1;
    }
    constexpr float eps = 5 * std::numeric_limits<float>::epsilon();
    for (int d = 0; d < 3; ++d)
      if (std::abs(h[d] / bounds.lo[d]) < eps)
        error = 1;
  }
  size_t cellIndex(P3 const &p) const
  { // :51-63
    size_t s = 0;
    for (int d = 2; d >= 0; --d)
    {
      int i = (int)std::floor((p[d] - bounds.lo[d]) / h[d]);
      s = s * n[d] + i;
    }
    return s;
  }
  Box3 cellBox(size_t cell) const
  { // :66-82
    Box3 r;
    P3 mn = bounds.lo, mx;
    for (int d = 0; d < 3; ++d)
    {
      size_t i = cell % n[d];
      cell /= n[d];
      mx[d] = mn[d] + (i + 1) * h[d];
      mn[d] += i * h[d];
    }
    r.lo = mn;
    r.hi = mx;
    return r;
  }
};

} // namespace

// ===========================================================================
// C API (ctypes-friendly)
// ===========================================================================

ORC_API int orc_num_threads() { return omp_get_max_threads(); }
ORC_API void orc_set_num_threads(int n) { omp_set_num_threads(n); }

ORC_API unsigned orc_expand_bits2_32(unsigned x) { return expandBits2_32(x); }
ORC_API unsigned long long orc_expand_bits2_64(unsigned long long x) { return expandBits2_64(x); }
ORC_API unsigned orc_morton32(float x, float y, float z) { return morton32(P3{{x, y, z}}); }
ORC_API unsigned long long orc_morton64(float x, float y, float z) { return morton64(P3{{x, y, z}}); }

// geometry known answers (test/tstGeometryDistance.cpp, tstGeometryIntersects.cpp)
ORC_API float orc_distance_point_point(float const *a, float const *b)
{
  return distance(P3{{a[0], a[1], a[2]}}, P3{{b[0], b[1], b[2]}});
}
ORC_API float orc_distance_point_box(float const *p, float const *b)
{
  Box3 bx;
  for (int d = 0; d < 3; ++d)
  {
    bx.lo[d] = b[d];
    bx.hi[d] = b[3 + d];
  }
  return distance(P3{{p[0], p[1], p[2]}}, bx);
}
ORC_API float orc_distance_point_triangle(float const *p, float const *t)
{
  Tri3 tr;
  for (int d = 0; d < 3; ++d)
  {
    tr.a[d] = t[d];
    tr.b[d] = t[3 + d];
    tr.c[d] = t[6 + d];
  }
  return distance(P3{{p[0], p[1], p[2]}}, tr);
}
ORC_API void orc_closest_point_triangle(float const *p, float const *t, float *out)
{
  Tri3 tr;
  for (int d = 0; d < 3; ++d)
  {
    tr.a[d] = t[d];
    tr.b[d] = t[3 + d];
    tr.c[d] = t[6 + d];
  }
  P3 r = closestPoint(P3{{p[0], p[1], p[2]}}, tr);
  for (int d = 0; d < 3; ++d)
    out[d] = r[d];
}
ORC_API int orc_intersects(int pred_kind, float const *pred, int prim_kind, float const *prim)
{
  Pred p{pred_kind, pred};
  if (prim_kind == PRIM_POINT)
    return p(P3{{prim[0], prim[1], prim[2]}});
  if (prim_kind == PRIM_BOX)
  {
    Box3 b;
    for (int d = 0; d < 3; ++d)
    {
      b.lo[d] = prim[d];
      b.hi[d] = prim[3 + d];
    }
    return p(b);
  }
  Tri3 tr;
  for (int d = 0; d < 3; ++d)
  {
    tr.a[d] = prim[d];
    tr.b[d] = prim[3 + d];
    tr.c[d] = prim[6 + d];
  }
  return p(tr);
}

// ---- stage-level entry points (used to localise GPU mismatches) ------------
ORC_API void orc_scene_bounds(int kind, float const *prims, int n, float *out6)
{
  Tree t;
  t.kind = kind;
  t.n = n;
  t.prims.assign(prims, prims + (size_t)primStride(kind) * n);
  Box3 b = sceneBounds(t);
  for (int d = 0; d < 3; ++d)
  {
    out6[d] = b.lo[d];
    out6[3 + d] = b.hi[d];
  }
}
ORC_API void orc_morton64_codes(int kind, float const *prims, int n, float const *bounds6, unsigned long long *codes)
{
  Tree t;
  t.kind = kind;
  t.n = n;
  t.prims.assign(prims, prims + (size_t)primStride(kind) * n);
  Box3 scene;
  for (int d = 0; d < 3; ++d)
  {
    scene.lo[d] = bounds6[d];
    scene.hi[d] = bounds6[3 + d];
  }
  computeCodes(t, scene);
  std::memcpy(codes, t.codes.data(), sizeof(unsigned long long) * n);
}
ORC_API void orc_morton32_codes(int pred_kind, float const *preds, int q, float const *bounds6, unsigned *codes)
{
  Box3 scene;
  for (int d = 0; d < 3; ++d)
  {
    scene.lo[d] = bounds6[d];
    scene.hi[d] = bounds6[3 + d];
  }
  int const s = predStride(pred_kind);
  for (int i = 0; i < q; ++i)
  {
    Pred p{pred_kind, preds + (size_t)s * i};
    codes[i] = morton32(translateAndScale(p.centroidOf(), scene));
  }
}
ORC_API void orc_sort_u64(unsigned long long *keys, int n, unsigned *perm)
{
  std::vector<unsigned long long> k(keys, keys + n);
  std::vector<unsigned> p;
  sortObjects(k, p);
  std::memcpy(keys, k.data(), sizeof(unsigned long long) * n);
  std::memcpy(perm, p.data(), sizeof(unsigned) * n);
}

// ---- tree ------------------------------------------------------------------
ORC_API void *orc_bvh_build(int kind, float const *prims, int n)
{
  Tree *t = new Tree;
  t->kind = kind;
  t->n = n;
  t->prims.assign(prims, prims + (size_t)primStride(kind) * n);
  buildTree(*t);
  return t;
}
// hierarchy from caller-supplied (already sorted) codes; primitives are taken in
// the given order (perm = iota).  Used for the Karras example
// (test/tstDetailsTreeConstruction.cpp:217-269).
ORC_API void *orc_bvh_from_sorted_codes(int kind, float const *prims, unsigned long long const *codes, int n)
{
  Tree *t = new Tree;
  t->kind = kind;
  t->n = n;
  t->prims.assign(prims, prims + (size_t)primStride(kind) * n);
  t->codes.assign(codes, codes + n);
  t->perm.resize(n);
  std::iota(t->perm.begin(), t->perm.end(), 0u);
  if (n >= 2)
    generateHierarchy(*t);
  else if (n == 1)
  {
    t->leaf_rope.assign(1, -1);
    t->bounds = t->primBox(0);
  }
  return t;
}
ORC_API void orc_bvh_destroy(void *h) { delete (Tree *)h; }
ORC_API int orc_bvh_size(void *h) { return ((Tree *)h)->n; }
ORC_API void orc_bvh_bounds(void *h, float *out6)
{
  Tree *t = (Tree *)h;
  for (int d = 0; d < 3; ++d)
  {
    out6[d] = t->bounds.lo[d];
    out6[3 + d] = t->bounds.hi[d];
  }
}
// reference-layout export: leaf_rope[n], leaf_index[n], left_child[n-1],
// rope[n-1], boxes[6(n-1)], codes[n]; any pointer may be NULL.
ORC_API void orc_bvh_export(void *h, int *leaf_rope, unsigned *leaf_index, int *left_child, int *rope, float *boxes6,
                            unsigned long long *codes)
{
  Tree *t = (Tree *)h;
  int const n = t->n;
  if (leaf_rope)
    std::memcpy(leaf_rope, t->leaf_rope.data(), sizeof(int) * n);
  if (leaf_index)
    std::memcpy(leaf_index, t->perm.data(), sizeof(unsigned) * n);
  if (codes)
    std::memcpy(codes, t->codes.data(), sizeof(unsigned long long) * n);
  if (n < 2)
    return;
  if (left_child)
    std::memcpy(left_child, t->left_child.data(), sizeof(int) * (n - 1));
  if (rope)
    std::memcpy(rope, t->rope.data(), sizeof(int) * (n - 1));
  if (boxes6)
    for (int i = 0; i < n - 1; ++i)
      for (int d = 0; d < 3; ++d)
      {
        boxes6[6 * (size_t)i + d] = t->box[i].lo[d];
        boxes6[6 * (size_t)i + 3 + d] = t->box[i].hi[d];
      }
}

// ---- spatial queries -------------------------------------------------------
// counts[q] = number of matches per ORIGINAL query index; limit>0 => stop at
// limit (CountUpToN, cluster/detail/ArborX_FDBSCAN.hpp:31-46).  counters2 (may
// be NULL) receives {sum internal box tests, sum leaf tests}.
ORC_API void orc_query_spatial_count(void *h, int pred_kind, float const *preds, int q, int limit, int *counts,
                                     long long *counters2)
{
  Tree const &t = *(Tree *)h;
  int const s = predStride(pred_kind);
  long long it = 0, lt = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : it, lt)
  for (int i = 0; i < q; ++i)
  {
    Pred p{pred_kind, preds + (size_t)s * i};
    int cnt = 0;
    Counters c;
    traverseSpatial(
        t, p,
        [&](int) {
          ++cnt;
          return limit > 0 && cnt >= limit;
        },
        counters2 ? &c : nullptr);
    counts[i] = cnt;
    it += c.internal_tests;
    lt += c.leaf_tests;
  }
  if (counters2)
  {
    counters2[0] = it;
    counters2[1] = lt;
  }
}

// Two-pass CRS query: spatial/detail/ArborX_CrsGraphWrapperImpl.hpp:148-446.
// offsets[q+1] is filled; returns nnz (>= 0), or -1 when buffer_size < 0 and a
// query overflowed (the reference throws SearchException, :263-268).  Call with
// indices == NULL to obtain nnz/offsets only, then again with storage.
// Row i = results of ORIGINAL query i, in traversal (DFS) order.
ORC_API long long orc_query_spatial_crs(void *h, int pred_kind, float const *preds, int q, int sort_predicates,
                                        int buffer_size, int *offsets, unsigned *indices)
{
  Tree const &t = *(Tree *)h;
  int const s = predStride(pred_kind);
  std::vector<unsigned> permute = predicatePermutation(t, pred_kind, preds, q, sort_predicates != 0);
  std::vector<int> counts(q, 0);
  // first pass (:176-222): count per permuted predicate
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < q; ++i)
  {
    unsigned const o = permute[i];
    Pred p{pred_kind, preds + (size_t)s * o};
    int cnt = 0;
    traverseSpatial(
        t, p,
        [&](int) {
          ++cnt;
          return false;
        },
        nullptr);
    counts[o] = cnt;
  }
  if (buffer_size < 0)
  { // hard preallocation: overflow throws (:263-268)
    int const b = -buffer_size;
    for (int i = 0; i < q; ++i)
      if (counts[i] > b)
        return -1;
  }
  // :235-248 exclusive scan
  long long total = 0;
  for (int i = 0; i < q; ++i)
  {
    offsets[i] = (int)total;
    total += counts[i];
  }
  offsets[q] = (int)total;
  if (!indices || total == 0)
    return total;
  // second pass (:263-295)
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < q; ++i)
  {
    unsigned const o = permute[i];
    Pred p{pred_kind, preds + (size_t)s * o};
    int pos = offsets[o];
    traverseSpatial(
        t, p,
        [&](int leaf) {
          indices[pos++] = t.perm[leaf];
          return false;
        },
        nullptr);
  }
  return total;
}

// ---- nearest queries -------------------------------------------------------
// CrsGraphWrapperImpl.hpp:353-374 + TreeTraversal.hpp:180-335.  k is the same
// for all queries (pass k_per_query != NULL for per-query k).  offsets[q+1],
// indices/distances sized sum(k); rows are compacted (rows shorter than k when
// n < k) and ascending by distance.  Returns nnz.
ORC_API long long orc_query_nearest_crs(void *h, float const *pts, int q, int k, int const *k_per_query,
                                        int sort_predicates, int *offsets, unsigned *indices, float *distances,
                                        long long *counters2)
{
  Tree const &t = *(Tree *)h;
  std::vector<unsigned> permute = predicatePermutation(t, PRED_POINT, pts, q, sort_predicates != 0);
  std::vector<long long> boff(q + 1, 0);
  for (int i = 0; i < q; ++i)
    boff[i + 1] = boff[i] + std::max(0, k_per_query ? k_per_query[i] : k);
  std::vector<PairID> buffer(boff[q]);
  std::vector<int> counts(q, 0);
  long long it = 0, lt = 0;
#pragma omp parallel reduction(+ : it, lt)
  {
    std::vector<int> stack;
    std::vector<float> stack_d;
    stack.reserve(128);
    stack_d.reserve(128);
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < q; ++i)
    {
      unsigned const o = permute[i];
      P3 p{{pts[3 * (size_t)o], pts[3 * (size_t)o + 1], pts[3 * (size_t)o + 2]}};
      Counters c;
      counts[o] = traverseNearest(t, p, k_per_query ? k_per_query[o] : k, buffer.data() + boff[o], stack, stack_d,
                                  counters2 ? &c : nullptr);
      it += c.internal_tests;
      lt += c.leaf_tests;
    }
  }
  if (counters2)
  {
    counters2[0] = it;
    counters2[1] = lt;
  }
  long long total = 0;
  for (int i = 0; i < q; ++i)
  {
    offsets[i] = (int)total;
    for (int j = 0; j < counts[i]; ++j)
    {
      PairID const &e = buffer[boff[i] + j];
      if (indices)
        indices[total + j] = t.perm[e.first];
      if (distances)
        distances[total + j] = e.second;
    }
    total += counts[i];
  }
  offsets[q] = (int)total;
  return total;
}

// nearest(Box | Sphere | Ray, k): the same traversal with the predicate geometry's distance
// (spatial/detail/ArborX_Predicates.hpp:58-80).  Point and box primitives.
ORC_API long long orc_query_nearest_geom_crs(void *h, int pred_kind, float const *preds, int q, int k,
                                             int sort_predicates, int *offsets, unsigned *indices, float *distances)
{
  Tree const &t = *(Tree *)h;
  if (t.kind == PRIM_TRI || (pred_kind != PRED_POINT && pred_kind != PRED_BOX && pred_kind != PRED_SPHERE && pred_kind != PRED_RAY))
    return -1;
  std::vector<unsigned> permute = predicatePermutation(t, pred_kind, preds, q, sort_predicates != 0);
  int const kk = std::max(0, k);
  std::vector<PairID> buffer((size_t)q * kk);
  std::vector<int> counts(q, 0);
#pragma omp parallel
  {
    std::vector<int> stack;
    std::vector<float> stack_d;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < q; ++i)
    {
      unsigned const o = permute[i];
      PairID *buf = buffer.data() + (size_t)o * kk;
      if (pred_kind == PRED_POINT)
        counts[o] = traverseNearest(t, P3{{preds[3 * (size_t)o], preds[3 * (size_t)o + 1], preds[3 * (size_t)o + 2]}}, k, buf,
                                    stack, stack_d, nullptr);
      else if (pred_kind == PRED_BOX)
      {
        Box3 b;
        for (int d = 0; d < 3; ++d)
        {
          b.lo[d] = preds[6 * (size_t)o + d];
          b.hi[d] = preds[6 * (size_t)o + 3 + d];
        }
        counts[o] = traverseNearest(t, b, k, buf, stack, stack_d, nullptr);
      }
      else if (pred_kind == PRED_SPHERE)
      {
        SphereQ sq{P3{{preds[4 * (size_t)o], preds[4 * (size_t)o + 1], preds[4 * (size_t)o + 2]}}, preds[4 * (size_t)o + 3]};
        counts[o] = traverseNearest(t, sq, k, buf, stack, stack_d, nullptr);
      }
      else
        counts[o] = traverseNearest(t, makeRay(preds + 6 * (size_t)o), k, buf, stack, stack_d, nullptr);
    }
  }
  long long total = 0;
  for (int i = 0; i < q; ++i)
  {
    offsets[i] = (int)total;
    for (int j = 0; j < counts[i]; ++j)
    {
      PairID const &e = buffer[(size_t)i * kk + j];
      if (indices)
        indices[total + j] = t.perm[e.first];
      if (distances)
        distances[total + j] = e.second;
    }
    total += counts[i];
  }
  offsets[q] = (int)total;
  return total;
}

// Experimental::ordered_intersects(ray): TreeTraversal<..., OrderedSpatialPredicateTag>
// (spatial/detail/ArborX_TreeTraversal.hpp:338-489): a priority queue of (node, distance(ray, box)) hands out the
// leaves in order of entry distance; the callback sees every leaf whose box the ray hits (distance < inf), nearest
// first, and may end the query.  Here every hit is recorded (limit > 0: the callback exits after `limit` hits).
// Returns the number of (index, distance) records; rows in predicate order.  pass 1: indices == NULL counts.
static int traverseOrderedRay(Tree const &t, Ray3 const &ray, int limit, std::vector<PairID> &out)
{
  float const inf = std::numeric_limits<float>::infinity();
  out.clear();
  if (t.n == 0)
    return 0;
  if (t.n == 1)
  {
    float const d = nearestDistance(t, ray, 0);
    if (d != inf)
      out.push_back(PairID{0, d});
    return (int)out.size();
  }
  // PriorityQueue with CompareDistance (lhs.second > rhs.second): min-heap on the distance
  std::vector<PairID> heap;
  auto cmp = [](PairID const &a, PairID const &b) { return a.second > b.second; };
  int node = t.n;
  while (true)
  {
    if (t.isLeaf(node))
    {
      out.push_back(PairID{node, nearestDistance(t, ray, node)});
      if (limit > 0 && (int)out.size() >= limit)
        return (int)out.size();
      if (heap.empty())
        return (int)out.size();
      std::pop_heap(heap.begin(), heap.end(), cmp);
      node = heap.back().first;
      heap.pop_back();
    }
    else
    {
      int const left_child = t.left_child[node - t.n];
      int const right_child = t.getRope(left_child);
      float const dl = nearestDistance(t, ray, left_child), dr = nearestDistance(t, ray, right_child);
      PairID const lp{left_child, dl}, rp{right_child, dr};
      PairID const &closer = dl < dr ? lp : rp;
      PairID const &further = dl < dr ? rp : lp;
      if (!heap.empty() && heap.front().second < closer.second)
      {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        node = heap.back().first;
        heap.pop_back();
        if (closer.second < inf)
        {
          heap.push_back(closer);
          std::push_heap(heap.begin(), heap.end(), cmp);
        }
      }
      else
        node = closer.first;
      if (further.second < inf)
      {
        heap.push_back(further);
        std::push_heap(heap.begin(), heap.end(), cmp);
      }
    }
  }
}
ORC_API long long orc_query_ordered_ray_crs(void *h, float const *rays, int q, int limit, int *offsets,
                                            unsigned *indices, float *distances)
{
  Tree const &t = *(Tree *)h;
  if (t.kind == PRIM_TRI)
    return -1;
  std::vector<std::vector<PairID>> rows(q);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < q; ++i)
    traverseOrderedRay(t, makeRay(rays + 6 * (size_t)i), limit, rows[i]);
  long long total = 0;
  for (int i = 0; i < q; ++i)
  {
    offsets[i] = (int)total;
    for (size_t j = 0; j < rows[i].size(); ++j)
    {
      if (indices)
        indices[total + j] = t.perm[rows[i][j].first];
      if (distances)
        distances[total + j] = rows[i][j].second;
    }
    total += (long long)rows[i].size();
  }
  offsets[q] = (int)total;
  return total;
}

// ---- half traversal (spatial/detail/ArborX_HalfTraversal.hpp:24-75) ---------
// Emits every unordered pair (original indices) of point leaves within r of
// each other exactly once.  pairs may be NULL to count only.  Returns #pairs.
ORC_API long long orc_half_traversal_pairs(void *h, float r, unsigned *pairs2, long long capacity)
{
  Tree const &t = *(Tree *)h;
  if (t.n < 2 || t.kind != PRIM_POINT)
    return 0;
  long long count = 0;
  for (int i = 0; i < t.n; ++i)
  {
    P3 c = t.point(t.perm[i]);
    float g[4] = {c[0], c[1], c[2], r};
    Pred p{PRED_SPHERE, g};
    int node = t.leaf_rope[i];
    while (node != -1)
    {
      if (t.isLeaf(node))
      {
        if (leafTest(t, p, node))
        {
          if (pairs2 && count < capacity)
          {
            pairs2[2 * count] = t.perm[i];
            pairs2[2 * count + 1] = t.perm[node];
          }
          ++count;
        }
        node = t.leaf_rope[node];
      }
      else
        node = p(t.box[node - t.n]) ? t.left_child[node - t.n] : t.rope[node - t.n];
    }
  }
  return count;
}

// ---- union-find known answers (test/tstUnionFind.cpp:69-125) ----------------
ORC_API void orc_union_find_merge(int *labels, int i, int j) { UnionFind{labels}.merge(i, j); }
ORC_API void orc_union_find_merge_into(int *labels, int i, int j) { UnionFind{labels}.merge_into(i, j); }
ORC_API int orc_union_find_representative(int *labels, int i) { return UnionFind{labels}.representative(i); }

// ---- DBSCAN (cluster/ArborX_DBSCAN.hpp:219-522) -----------------------------
// impl: 0 FDBSCAN, 1 FDBSCAN_DenseBox;  algo: 0 DBSCAN, 1 DBSCAN*.
// Returns 0, or 1 precondition failure (eps<=0, minpts<2: :240-241), or 2 for
// the DenseBox loss-of-precision guard (CartesianGrid.hpp:132-135).
// is_core_out (may be NULL) receives the core flag per point (minpts>2 only;
// all-ones for minpts==2).  stats (may be NULL): {#dense cells, #points in
// dense cells, sum internal tests, sum leaf tests}.
ORC_API int orc_dbscan(float const *xyz, int n, float eps, int minpts, int impl, int algo, int *labels,
                       int *is_core_out, long long *stats)
{
  if (!(eps > 0) || minpts < 2)
    return 1;
  bool const special = (minpts == 2);
  bool const star = (algo == 1);
  std::vector<int> num_neigh;
  auto isCore = [&](int i) { return special ? true : num_neigh[i] >= minpts; };
  UnionFind uf{labels};
  long long it = 0, lt = 0, n_dense_cells = 0, n_dense_pts = 0;

  if (impl == 0)
  {
    // :269-327
    Tree t;
    t.kind = PRIM_POINT;
    t.n = n;
    t.prims.assign(xyz, xyz + 3 * (size_t)n);
    buildTree(t);
    for (int i = 0; i < n; ++i)
      labels[i] = i;
    if (!special)
    {
      num_neigh.assign(n, 0);
      // CountUpToN (FDBSCAN.hpp:31-46), predicates = spheres around every point
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : it, lt)
      for (int i = 0; i < n; ++i)
      {
        float g[4] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], eps};
        Pred p{PRED_SPHERE, g};
        int cnt = 0;
        Counters c;
        traverseSpatial(
            t, p, [&](int) { return ++cnt >= minpts; }, &c);
        num_neigh[i] = cnt;
        it += c.internal_tests;
        lt += c.leaf_tests;
      }
    }
    // HalfTraversal + FDBSCANCallback (FDBSCAN.hpp:49-110); callback return
    // value is ignored by HalfTraversal (HalfTraversal.hpp:62-63)
    if (t.n >= 2)
    {
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : it, lt)
      for (int li = 0; li < n; ++li)
      {
        int const i = (int)t.perm[li];
        float g[4] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], eps};
        Pred p{PRED_SPHERE, g};
        int node = t.leaf_rope[li];
        while (node != -1)
        {
          if (t.isLeaf(node))
          {
            ++lt;
            if (leafTest(t, p, node))
            {
              int const j = (int)t.perm[node];
              bool const is_border = !isCore(i);
              bool const neigh_core = isCore(j);
              if (star && (is_border || !neigh_core))
              {
                // early_exit, ignored
              }
              else if (is_border)
              {
                if (neigh_core)
                  uf.merge_into(i, j);
              }
              else
              {
                if (neigh_core)
                  uf.merge(i, j);
                else
                  uf.merge_into(j, i);
              }
            }
            node = t.leaf_rope[node];
          }
          else
          {
            ++it;
            node = p(t.box[node - t.n]) ? t.left_child[node - t.n] : t.rope[node - t.n];
          }
        }
      }
    }
  }
  else
  {
    // :328-481
    Tree pts;
    pts.kind = PRIM_POINT;
    pts.n = n;
    pts.prims.assign(xyz, xyz + 3 * (size_t)n);
    Box3 bounds = sceneBounds(pts);
    float const h = eps / std::sqrt(3.f);
    CartesianGrid grid(bounds, h);
    if (grid.error)
      return 2;
    std::vector<size_t> cell(n);
#pragma omp parallel for
    for (int i = 0; i < n; ++i)
      cell[i] = grid.cellIndex(pts.point(i));
    std::vector<unsigned> permute;
    sortObjects(cell, permute);
    // computeOffsetsInOrderedView (misc/ArborX_Utils.hpp:25-51)
    std::vector<int> cell_offsets;
    for (int i = 0; i <= n; ++i)
      if (i == 0 || i == n || cell[i] != cell[i - 1])
        cell_offsets.push_back(i);
    int const n_cells = (int)cell_offsets.size() - 1;
    // reorderDenseAndSparseCells (FDBSCANDenseBox.hpp:215-283); cell order
    // inside each class is nondeterministic in the reference, we keep sorted order
    std::vector<size_t> rcell(n);
    std::vector<unsigned> rperm(n);
    int npd = 0;
    for (int c = 0; c < n_cells; ++c)
    {
      int sz = cell_offsets[c + 1] - cell_offsets[c];
      if (sz >= minpts)
        npd += sz;
    }
    {
      int doff = 0, soff = npd;
      for (int c = 0; c < n_cells; ++c)
      {
        int sz = cell_offsets[c + 1] - cell_offsets[c];
        int &off = (sz >= minpts) ? doff : soff;
        for (int j = cell_offsets[c]; j < cell_offsets[c + 1]; ++j, ++off)
        {
          rcell[off] = cell[j];
          rperm[off] = permute[j];
        }
      }
    }
    cell.swap(rcell);
    permute.swap(rperm);
    std::vector<int> dco; // dense_cell_offsets
    for (int i = 0; i <= npd; ++i)
      if (i == 0 || i == npd || cell[i] != cell[i - 1])
        dco.push_back(i);
    int const n_dense = (int)dco.size() - 1;
    n_dense_cells = n_dense;
    n_dense_pts = npd;
    for (int i = 0; i < n; ++i)
      labels[i] = i;
    // unionFindWithinEachDenseCell (:286-307)
    for (int i = 1; i < npd; ++i)
      if (cell[i] == cell[i - 1])
        uf.merge((int)permute[i], (int)permute[i - 1]);
    // MixedBoxPrimitives (ArborX_DBSCAN.hpp:73-175): dense-cell boxes then
    // degenerate boxes of sparse points
    int const n_sparse = n - npd;
    Tree t;
    t.kind = PRIM_BOX;
    t.n = n_dense + n_sparse;
    t.prims.resize(6 * (size_t)t.n);
    for (int c = 0; c < n_dense; ++c)
    {
      Box3 b = grid.cellBox(cell[dco[c]]);
      for (int d = 0; d < 3; ++d)
      {
        t.prims[6 * (size_t)c + d] = b.lo[d];
        t.prims[6 * (size_t)c + 3 + d] = b.hi[d];
      }
    }
    for (int s = 0; s < n_sparse; ++s)
    {
      P3 p = pts.point(permute[npd + s]);
      size_t const o = (size_t)n_dense + s;
      for (int d = 0; d < 3; ++d)
      {
        t.prims[6 * o + d] = p[d];
        t.prims[6 * o + 3 + d] = p[d];
      }
    }
    buildTree(t);

    if (!special)
    {
      num_neigh.assign(n, 0);
      for (int i = 0; i < npd; ++i)
        num_neigh[permute[i]] = INT_MAX; // :438-441
      // CountUpToN_DenseBox (FDBSCANDenseBox.hpp:32-96) for sparse points only
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : it, lt)
      for (int s = 0; s < n_sparse; ++s)
      {
        int const i = (int)permute[npd + s];
        P3 const qp = pts.point(i);
        float g[4] = {qp[0], qp[1], qp[2], eps};
        Pred p{PRED_SPHERE, g};
        int cnt = 0;
        Counters c;
        traverseSpatial(
            t, p,
            [&](int leaf) {
              int const k = (int)t.perm[leaf];
              if (k < n_dense)
              {
                for (int jj = dco[k]; jj < dco[k + 1]; ++jj)
                  if (distance(qp, pts.point(permute[jj])) <= eps)
                    if (++cnt >= minpts)
                      return true;
              }
              else if (++cnt >= minpts)
                return true;
              return false;
            },
            &c);
        num_neigh[i] = cnt;
        it += c.internal_tests;
        lt += c.leaf_tests;
      }
    }
    // FDBSCANDenseBoxCallback (:98-205): full traversal for every point
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : it, lt)
    for (int i = 0; i < n; ++i)
    {
      P3 const qp = pts.point(i);
      float g[4] = {qp[0], qp[1], qp[2], eps};
      Pred p{PRED_SPHERE, g};
      Counters c;
      traverseSpatial(
          t, p,
          [&](int leaf) {
            int const k = (int)t.perm[leaf];
            if (!isCore(i))
              return true; // early_exit
            if (k < n_dense)
            {
              int const cs = dco[k], ce = dco[k + 1];
              if (uf.representative(i) == uf.representative((int)permute[cs]))
                return false;
              for (int jj = cs; jj < ce; ++jj)
              {
                int const j = (int)permute[jj];
                if (uf.representative(i) == uf.representative(j))
                  break;
                if (distance(qp, pts.point(j)) <= eps)
                {
                  uf.merge(i, j);
                  break;
                }
              }
            }
            else
            {
              int const j = (int)permute[npd + (k - n_dense)];
              bool const neigh_core = isCore(j);
              if (neigh_core && i > j)
                uf.merge(i, j);
              else if (!star && !neigh_core)
                uf.merge_into(j, i);
            }
            return false;
          },
          &c);
      it += c.internal_tests;
      lt += c.leaf_tests;
    }
  }

  // finalize_labels (:489-506) + mark_noise (:512-518)
  std::vector<int> cluster_sizes(n, 0);
  for (int i = 0; i < n; ++i)
  {
    int next, vstat = labels[i];
    int const old = vstat;
    while (vstat > (next = labels[vstat]))
      vstat = next;
    if (vstat != old)
      labels[i] = vstat;
    cluster_sizes[labels[i]]++;
  }
  std::vector<int> lab(labels, labels + n);
  for (int i = 0; i < n; ++i)
    if (cluster_sizes[lab[i]] == 1 && (special || !isCore(i)))
      labels[i] = -1;
  if (is_core_out)
    for (int i = 0; i < n; ++i)
      is_core_out[i] = isCore(i) ? 1 : 0;
  if (stats)
  {
    stats[0] = n_dense_cells;
    stats[1] = n_dense_pts;
    stats[2] = it;
    stats[3] = lt;
  }
  return 0;
}

// ---- DBSCAN verifier (benchmarks/cluster/ArborX_DBSCANVerification.hpp:83-429)
// Returns a bit mask of FAILED checks (0 = labels accepted):
//  1 core point marked noise (:83)   2 connected cores differ (:108)
//  4 noise point not -1 (:141)       8 border check (:175 DBSCAN / :219 DBSCAN*)
// 16 cluster ids not unique / bridged (:257)
ORC_API int orc_dbscan_verify(float const *xyz, int n, float eps, int minpts, int const *labels, int algo)
{
  Tree t;
  t.kind = PRIM_POINT;
  t.n = n;
  t.prims.assign(xyz, xyz + 3 * (size_t)n);
  buildTree(t);
  // neighbour graph (:403-405), self placed first (:407-424)
  std::vector<int> offset(n + 1, 0);
  std::vector<int> cnt(n, 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < n; ++i)
  {
    float g[4] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], eps};
    Pred p{PRED_SPHERE, g};
    int c = 0;
    traverseSpatial(
        t, p,
        [&](int) {
          ++c;
          return false;
        },
        nullptr);
    cnt[i] = c;
  }
  for (int i = 0; i < n; ++i)
    offset[i + 1] = offset[i] + cnt[i];
  std::vector<int> nidx(offset[n]);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < n; ++i)
  {
    float g[4] = {xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], eps};
    Pred p{PRED_SPHERE, g};
    int pos = offset[i];
    traverseSpatial(
        t, p,
        [&](int leaf) {
          nidx[pos++] = (int)t.perm[leaf];
          return false;
        },
        nullptr);
    for (int jj = offset[i]; jj < offset[i + 1]; ++jj)
      if (nidx[jj] == i)
      {
        std::swap(nidx[offset[i]], nidx[jj]);
        break;
      }
  }
  auto isCore = [&](int j) { return offset[j + 1] - offset[j] >= minpts; };
  int fail = 0;
  for (int i = 0; i < n; ++i)
  {
    bool const self_core = isCore(i);
    int const self_label = labels[i];
    if (self_core && self_label < 0)
      fail |= 1;
    if (self_core)
    {
      for (int j = offset[i] + 1; j < offset[i + 1]; ++j)
        if (isCore(nidx[j]) && labels[nidx[j]] != self_label)
          fail |= 2;
      continue;
    }
    bool is_border = false, shared = false;
    for (int j = offset[i] + 1; j < offset[i + 1]; ++j)
      if (isCore(nidx[j]))
      {
        is_border = true;
        if (labels[nidx[j]] == self_label)
          shared = true;
      }
    if (!is_border && self_label != -1)
      fail |= 4;
    if (algo == 0)
    {
      if (is_border && !shared)
        fail |= 8;
    }
    else if (is_border && self_label != -1)
      fail |= 8;
  }
  // verifyClustersAreUnique (:257-346)
  std::vector<int> lab(labels, labels + n);
  for (int i = 0; i < n; ++i)
    if (!isCore(i))
      for (int j = offset[i]; j < offset[i + 1]; ++j)
        if (isCore(nidx[j]))
        {
          lab[i] = -1;
          break;
        }
  std::vector<int> uniq;
  for (int i = 0; i < n; ++i)
    if (lab[i] != -1)
      uniq.push_back(lab[i]);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  size_t num_clusters = 0;
  std::vector<int> cluster_sets;
  std::vector<int> stack;
  for (int i = 0; i < n; ++i)
    if (lab[i] >= 0)
    {
      int const id = lab[i];
      cluster_sets.push_back(id);
      ++num_clusters;
      stack.assign(1, i);
      while (!stack.empty())
      {
        int k = stack.back();
        stack.pop_back();
        if (lab[k] >= 0)
        {
          lab[k] = -1;
          for (int j = offset[k]; j < offset[k + 1]; ++j)
            if (isCore(nidx[j]) || lab[nidx[j]] == id)
              stack.push_back(nidx[j]);
        }
      }
    }
  std::sort(cluster_sets.begin(), cluster_sets.end());
  cluster_sets.erase(std::unique(cluster_sets.begin(), cluster_sets.end()), cluster_sets.end());
  if (cluster_sets.size() != uniq.size() || num_clusters != uniq.size())
    fail |= 16;
  return fail;
}

// ---------------------------------------------------------------------------
// Euclidean / mutual-reachability minimum spanning tree (Boruvka over the BVH) and the dendrogram of its edges:
// cluster/ArborX_MinimumSpanningTree.hpp:46-258 (MST mode), cluster/detail/ArborX_BoruvkaHelpers.hpp,
// cluster/detail/ArborX_MutualReachabilityDistance.hpp, spatial/detail/ArborX_TreeNodeLabeling.hpp,
// cluster/ArborX_Dendrogram.hpp:47-76 + cluster/detail/ArborX_DendrogramHelpers.hpp:31-80 (UNION_FIND).
// ---------------------------------------------------------------------------
namespace
{
// BoruvkaHelpers.hpp:37-105: (weight, unordered pair of leaf positions, direction flag)
struct DirectedEdge
{
  unsigned long long key = ~0ull;
  float weight = std::numeric_limits<float>::infinity();
  DirectedEdge() = default;
  DirectedEdge(int source, int target, float w) : weight(w)
  {
    unsigned long long const lo = (unsigned)std::min(source, target), hi = (unsigned)std::max(source, target);
    key = (lo << 32) | (hi << 1) | (source < target ? 0ull : 1ull);
  }
  bool reverse() const { return key & 1ull; }
  int lo() const { return (int)((key >> 32) & 0x7fffffffull); }
  int hi() const { return (int)((key >> 1) & 0x7fffffffull); }
  int source() const { return reverse() ? hi() : lo(); }
  int target() const { return reverse() ? lo() : hi(); }
  bool operator<(DirectedEdge const &o) const { return weight != o.weight ? weight < o.weight : key < o.key; }
};
} // namespace

// edges2: (n - 1) x (source, target) in ORIGINAL indices; weights: n - 1.  k > 1: mutual reachability with core
// distance = distance to the k-th nearest point, the point itself included (MinimumSpanningTree.hpp:70-88).
// Returns the number of Boruvka iterations.
// dendrogram_parents != nullptr: BoruvkaMode::HDBSCAN (MinimumSpanningTree.hpp:176-185,214-230,266-297): the rounds
// also record which edge every component picked, the edges come out ordered by (chain, weight) and
// dendrogram_parents [2 n - 1] / dendrogram_heights [n - 1] describe the dendrogram in that edge order.
static int mstImpl(float const *xyz, int n, int k, int *edges2, float *weights_out, int *dendrogram_parents,
                   float *dendrogram_heights)
{
  if (n < 2)
    return 0;
  bool const hdbscan = dendrogram_parents != nullptr;
  constexpr int ROOT_CHAIN_VALUE = -2, FOLLOW_CHAIN_VALUE = -3; // BoruvkaHelpers.hpp:34-35
  std::vector<int> edges_mapping, sided_parents;
  if (hdbscan)
  {
    edges_mapping.assign(n, -1);
    sided_parents.assign(n - 1, ROOT_CHAIN_VALUE);
  }
  int edges_start = 0, edges_end = 0;
  Tree t;
  t.kind = PRIM_POINT;
  t.n = n;
  t.prims.assign(xyz, xyz + 3 * (size_t)n);
  buildTree(t);
  float const inf = std::numeric_limits<float>::infinity();
  // core distances by original index (MaxDistance over nearest(point, k))
  std::vector<float> core;
  if (k > 1)
  {
    core.assign(n, 0.f);
#pragma omp parallel
    {
      std::vector<int> stack;
      std::vector<float> stack_d;
      std::vector<PairID> buf(k);
#pragma omp for schedule(dynamic, 256)
      for (int pos = 0; pos < n; ++pos)
      {
        unsigned const o = t.perm[pos];
        int const m = traverseNearest(t, t.point(o), k, buf.data(), stack, stack_d, nullptr);
        float d = 0.f;
        for (int j = 0; j < m; ++j)
          d = std::max(d, buf[j].second);
        core[o] = d;
      }
    }
  }
  auto metric = [&](unsigned oi, unsigned oj, float d) {
    return k > 1 ? std::max(std::max(core[oi], core[oj]), d) : d;
  };
  // TreeNodeLabeling.hpp:27-42
  std::vector<int> parents(2 * (size_t)n - 1, -1);
  for (int i = n; i < 2 * n - 1; ++i)
  {
    int const l = t.left_child[i - n];
    parents[l] = i;
    parents[t.getRope(l)] = i;
  }
  std::vector<int> labels(2 * (size_t)n - 1);
  for (int i = 0; i < n; ++i)
    labels[i] = i;
  std::vector<DirectedEdge> out_edges(n);
  std::vector<float> comp_weight(n), radii(n);
  int num_edges = 0, iterations = 0;
  std::vector<int> e_src(n - 1), e_dst(n - 1);
  do
  {
    ++iterations;
    // reduceLabels (TreeNodeLabeling.hpp:44-93): an internal node carries a label iff its whole subtree does
    for (int i = 2 * n - 2; i >= n; --i)
      labels[i] = -2;
    {
      // children before parents: in this numbering a child's internal index is not ordered with its parent's, so
      // walk up from the leaves like the reference does (second arriver continues)
      for (int leaf = 0; leaf < n; ++leaf)
      {
        int i = leaf;
        do
        {
          int const label = labels[i];
          int const parent = parents[i];
          int const parent_label = labels[parent];
          if (parent_label == -2)
          {
            labels[parent] = label;
            break;
          }
          if (parent_label != label)
            labels[parent] = -1;
          i = parent;
        } while (i != n);
      }
    }
    std::fill(out_edges.begin(), out_edges.end(), DirectedEdge());
    std::fill(comp_weight.begin(), comp_weight.end(), inf);
    std::fill(radii.begin(), radii.end(), inf);
    // resetSharedRadii (BoruvkaHelpers.hpp:735-778): Morton neighbours in different components bound both
    for (int i = 0; i + 1 < n; ++i)
      if (labels[i] != labels[i + 1])
      {
        float const r = metric(t.perm[i], t.perm[i + 1], distance(t.point(t.perm[i]), t.point(t.perm[i + 1])));
        radii[labels[i]] = std::min(radii[labels[i]], r);
        radii[labels[i + 1]] = std::min(radii[labels[i + 1]], r);
      }
    // FindComponentNearestNeighbors (BoruvkaHelpers.hpp:160-312), private radius copy
    std::vector<DirectedEdge> best(n);
#pragma omp parallel
    {
      std::vector<int> stack;
      std::vector<float> stack_d;
#pragma omp for schedule(dynamic, 256)
      for (int i = 0; i < n; ++i)
      {
        int const component = labels[i];
        unsigned const oi = t.perm[i];
        P3 const p = t.point(oi);
        DirectedEdge current_best;
        float radius = radii[component];
        stack.assign(1, -1);
        stack_d.assign(1, 0.f);
        int node = n;
        float distance_node = 0.f;
        do
        {
          bool traverse_left = false, traverse_right = false;
          int left_child = 0, right_child = 0;
          float distance_left = inf, distance_right = inf;
          if (distance_node <= radius)
          {
            left_child = t.left_child[node - n];
            right_child = t.getRope(left_child);
            distance_left = nearestDistance(t, p, left_child);
            distance_right = nearestDistance(t, p, right_child);
            if (labels[left_child] != component && distance_left <= radius)
            {
              if (t.isLeaf(left_child))
              {
                DirectedEdge const cand(i, left_child, metric(oi, t.perm[left_child], distance_left));
                if (cand < current_best)
                {
                  current_best = cand;
                  radius = cand.weight;
                }
              }
              else
                traverse_left = true;
            }
            if (labels[right_child] != component && distance_right <= radius)
            {
              if (t.isLeaf(right_child))
              {
                DirectedEdge const cand(i, right_child, metric(oi, t.perm[right_child], distance_right));
                if (cand < current_best)
                {
                  current_best = cand;
                  radius = cand.weight;
                }
              }
              else
                traverse_right = true;
            }
          }
          if (!traverse_left && !traverse_right)
          {
            node = stack.back();
            stack.pop_back();
            distance_node = stack_d.back();
            stack_d.pop_back();
          }
          else
          {
            node = (traverse_left && (distance_left <= distance_right || !traverse_right)) ? left_child : right_child;
            distance_node = node == left_child ? distance_left : distance_right;
            if (traverse_left && traverse_right)
            {
              stack.push_back(node == left_child ? right_child : left_child);
              stack_d.push_back(node == left_child ? distance_right : distance_left);
            }
          }
        } while (node != -1);
        best[i] = current_best;
      }
    }
    // component minimum (the reference's atomic_min on the weight, then retrieveEdges :336-372 on the pair key)
    for (int i = 0; i < n; ++i)
      if (best[i].weight < inf && best[i] < out_edges[labels[i]])
        out_edges[labels[i]] = best[i];
    // UpdateComponentsAndEdges (BoruvkaHelpers.hpp:381-470)
    auto nextComponent = [&](int component) {
      int const next = labels[out_edges[component].target()];
      int const next_next = labels[out_edges[next].target()];
      if (next_next != component)
        return next;
      return std::min(component, next);
    };
    for (int i = 0; i < n; ++i)
    {
      if (labels[i] != i || nextComponent(i) == i)
        continue;
      e_src[num_edges] = out_edges[i].source();
      e_dst[num_edges] = out_edges[i].target();
      weights_out[num_edges] = out_edges[i].weight;
      if (hdbscan)
        edges_mapping[i] = num_edges;
      ++num_edges;
    }
    if (hdbscan)
    {
      // the (weight, smaller position, larger position) order of WeightedEdge.hpp:30-50
      auto edgeLess = [&](int a, int b) {
        if (weights_out[a] != weights_out[b])
          return weights_out[a] < weights_out[b];
        int const amin = std::min(e_src[a], e_dst[a]), bmin = std::min(e_src[b], e_dst[b]);
        if (amin != bmin)
          return amin < bmin;
        return std::max(e_src[a], e_dst[a]) < std::max(e_src[b], e_dst[b]);
      };
      // BidirectionalEdgesTag (BoruvkaHelpers.hpp:439-447): the smaller label of a mutual pair did not append
      for (int i = 0; i < n; ++i)
        if (labels[i] == i && nextComponent(i) == i)
          edges_mapping[i] = edges_mapping[labels[out_edges[i].target()]];
      if (iterations > 1)
      {
        // updateSidedParents (:490-525): the edges of the previous round hang below the edge their merged
        // component picks now, on its source or target side, or follow its chain upwards
        for (int e = edges_start; e < edges_end; ++e)
        {
          int const component = labels[e_src[e]];
          int const alpha = edges_mapping[component];
          if (edgeLess(e, alpha))
            sided_parents[e] = 2 * alpha + (labels[e_src[alpha]] == component ? 1 : 0);
          else
            sided_parents[e] = FOLLOW_CHAIN_VALUE - alpha;
        }
      }
      else
      {
        // assignVertexParents (:527-548): in the first round every vertex is a component
        for (int e = 0; e < n; ++e)
          dendrogram_parents[(int)t.perm[labels[out_edges[e].source()]] + (n - 1)] =
              edges_mapping[labels[out_edges[e].source()]];
      }
    }
    std::vector<int> new_labels(n);
    for (int i = 0; i < n; ++i)
    {
      int prev = labels[i], next;
      while ((next = nextComponent(prev)) != prev)
        prev = next;
      new_labels[i] = next;
    }
    std::copy(new_labels.begin(), new_labels.end(), labels.begin());
    edges_start = edges_end;
    edges_end = num_edges;
  } while (n - num_edges > 1);
  if (hdbscan)
  {
    // computeParentsAndReorderEdges (BoruvkaHelpers.hpp:550-733)
    int const m = n - 1;
    for (int e = edges_start; e < edges_end; ++e)
      sided_parents[e] = ROOT_CHAIN_VALUE; // MinimumSpanningTree.hpp:271-275
    auto edgeLess = [&](int a, int b) {
      if (weights_out[a] != weights_out[b])
        return weights_out[a] < weights_out[b];
      int const amin = std::min(e_src[a], e_dst[a]), bmin = std::min(e_src[b], e_dst[b]);
      if (amin != bmin)
        return amin < bmin;
      return std::max(e_src[a], e_dst[a]) < std::max(e_src[b], e_dst[b]);
    };
    std::vector<long long> keys(m);
    for (int e = 0; e < m; ++e)
    {
      long long key = sided_parents[e];
      if (key <= FOLLOW_CHAIN_VALUE)
      {
        int next = FOLLOW_CHAIN_VALUE - (int)key;
        while (true)
        {
          key = sided_parents[next];
          if (key <= FOLLOW_CHAIN_VALUE)
            next = FOLLOW_CHAIN_VALUE - (int)key;
          else if (key >= 0)
          {
            next = (int)(key / 2);
            if (edgeLess(e, next))
              break;
          }
          else if (key == ROOT_CHAIN_VALUE)
            break;
        }
      }
      if (key == ROOT_CHAIN_VALUE)
        key = INT_MAX;
      int wbits;
      std::memcpy(&wbits, &weights_out[e], sizeof(int));
      keys[e] = (key << 32) + wbits;
    }
    std::vector<unsigned> permute(m);
    for (int e = 0; e < m; ++e)
      permute[e] = e;
    std::stable_sort(permute.begin(), permute.end(), [&](unsigned a, unsigned b) { return keys[a] < keys[b]; });
    {
      std::vector<long long> sorted(m);
      for (int e = 0; e < m; ++e)
        sorted[e] = keys[permute[e]];
      keys.swap(sorted);
    }
    // the smallest edge of a chain goes first even among equal weights (:625-660)
    for (int i = 0; i + 1 < m; ++i)
      if (i == 0 || (keys[i - 1] >> 32) != (keys[i] >> 32))
      {
        int mm = i;
        for (int kk = i + 1; kk < m && keys[kk] == keys[i]; ++kk)
          if (edgeLess((int)permute[kk], (int)permute[mm]))
            mm = kk;
        if (mm != i)
          std::swap(permute[i], permute[mm]);
      }
    std::vector<int> rev(m);
    for (int i = 0; i < m; ++i)
      rev[permute[i]] = i;
    for (int i = m; i < 2 * n - 1; ++i)
      dendrogram_parents[i] = rev[dendrogram_parents[i]];
    for (int i = 0; i < m; ++i)
    {
      if (i == m - 1)
        dendrogram_parents[i] = -1;
      else if ((keys[i] >> 32) == (keys[i + 1] >> 32))
        dendrogram_parents[i] = i + 1;
      else
        dendrogram_parents[i] = rev[(int)((keys[i] >> 32) / 2)];
    }
    std::vector<int> s2(m), d2(m);
    std::vector<float> w2(m);
    for (int i = 0; i < m; ++i)
    {
      s2[i] = e_src[permute[i]];
      d2[i] = e_dst[permute[i]];
      w2[i] = weights_out[permute[i]];
    }
    e_src.swap(s2);
    e_dst.swap(d2);
    for (int i = 0; i < m; ++i)
    {
      weights_out[i] = w2[i];
      dendrogram_heights[i] = w2[i];
    }
  }
  // finalizeEdges (BoruvkaHelpers.hpp:472-488)
  for (int e = 0; e < n - 1; ++e)
  {
    edges2[2 * e] = (int)t.perm[e_src[e]];
    edges2[2 * e + 1] = (int)t.perm[e_dst[e]];
  }
  return iterations;
}

ORC_API int orc_mst(float const *xyz, int n, int k, int *edges2, float *weights_out)
{
  return mstImpl(xyz, n, k, edges2, weights_out, nullptr, nullptr);
}

// MinimumSpanningTree<MemorySpace, BoruvkaMode::HDBSCAN>: edges in (chain, weight) order + the dendrogram over them
ORC_API int orc_mst_hdbscan(float const *xyz, int n, int k, int *edges2, float *weights_out, int *parents,
                            float *heights)
{
  if (n == 1 && parents)
    parents[0] = -1;
  return mstImpl(xyz, n, k, edges2, weights_out, parents, heights);
}

// Dendrogram of weighted edges (Dendrogram.hpp:47-76): edges sorted by weight, then the sequential union-find of
// DendrogramHelpers.hpp:31-80.  parents: 2 * num_edges + 1 entries (edges first, then vertices); heights: num_edges.
ORC_API void orc_dendrogram_union_find(int const *edges2, float const *weights, int num_edges, int *parents,
                                       float *heights)
{
  int const num_vertices = num_edges + 1;
  std::vector<int> order(num_edges);
  for (int e = 0; e < num_edges; ++e)
    order[e] = e;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weights[a] < weights[b]; });
  std::vector<int> labels(num_vertices), set_edges(num_vertices, -1);
  for (int i = 0; i < num_vertices; ++i)
    labels[i] = i;
  UnionFind uf{labels.data()};
  for (int e = 0; e < num_edges; ++e)
  {
    heights[e] = weights[order[e]];
    int const i = uf.representative(edges2[2 * order[e]]);
    int const j = uf.representative(edges2[2 * order[e] + 1]);
    for (int k : {i, j})
    {
      int const child = set_edges[k];
      if (child != -1)
        parents[child] = e;
      else
        parents[num_edges + k] = e;
    }
    uf.merge(i, j);
    set_edges[uf.representative(i)] = e;
  }
  if (num_edges > 0)
    parents[num_edges - 1] = -1;
}
