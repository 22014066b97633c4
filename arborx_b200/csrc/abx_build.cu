// abx_build.cu -- BVH construction: scene bounds, Morton64 codes, Apetrei hierarchy.
//
// Follows the behaviour of spatial/ArborX_LinearBVH.hpp:171-256 and
// spatial/detail/ArborX_TreeConstruction.hpp:27-39,74-322, but emits the compact
// two-children-per-record Node64 layout (abx_common.cuh) instead of the
// reference's {left_child, rope, box} nodes.
#include "abx_common.cuh"

namespace abx
{

namespace
{

constexpr int kThreads = 256;

// ---- primitive accessors -----------------------------------------------------
template <int KIND>
__device__ __forceinline__ Box primBox(float const *__restrict__ prims, int64_t i);

template <>
__device__ __forceinline__ Box primBox<ABX_PRIM_POINT3F>(float const *__restrict__ p, int64_t i)
{
  Box b;
  float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
  b.lo[0] = b.hi[0] = x;
  b.lo[1] = b.hi[1] = y;
  b.lo[2] = b.hi[2] = z;
  return b;
}
template <>
__device__ __forceinline__ Box primBox<ABX_PRIM_BOX3F>(float const *__restrict__ p, int64_t i)
{
  Box b;
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    b.lo[d] = p[6 * i + d];
    b.hi[d] = p[6 * i + 3 + d];
  }
  return b;
}
template <>
__device__ __forceinline__ Box primBox<ABX_PRIM_TRI3F>(float const *__restrict__ p, int64_t i)
{
  // expand(Box, Triangle): geometry/algorithms/ArborX_Expand.hpp:100-107
  Box b;
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    float a = p[9 * i + d], bb = p[9 * i + 3 + d], c = p[9 * i + 6 + d];
    b.lo[d] = fminf(fminf(a, bb), c);
    b.hi[d] = fmaxf(fmaxf(a, bb), c);
  }
  return b;
}

// returnCentroid: geometry/algorithms/ArborX_Centroid.hpp:41-82
template <int KIND>
__device__ __forceinline__ void primCentroid(float const *__restrict__ p, int64_t i, float c[3])
{
  if (KIND == ABX_PRIM_POINT3F)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      c[d] = p[3 * i + d];
  }
  else if (KIND == ABX_PRIM_BOX3F)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      c[d] = __fdiv_rn(__fadd_rn(p[6 * i + d], p[6 * i + 3 + d]), 2.f);
  }
  else
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      c[d] = __fdiv_rn(__fadd_rn(__fadd_rn(p[9 * i + d], p[9 * i + 3 + d]), p[9 * i + 6 + d]), 3.f);
  }
}

// ---- scene bounds: TreeConstruction.hpp:27-39 --------------------------------
template <int KIND>
__global__ void __launch_bounds__(kThreads)
    sceneBoundsKernel(float const *__restrict__ prims, int64_t n, unsigned *__restrict__ bounds_enc)
{
  Box b = emptyBox();
  int64_t const stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
  {
    Box p = primBox<KIND>(prims, i);
    boxUnion(b, p);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      b.lo[d] = fminf(b.lo[d], __shfl_xor_sync(0xffffffffu, b.lo[d], o));
      b.hi[d] = fmaxf(b.hi[d], __shfl_xor_sync(0xffffffffu, b.hi[d], o));
    }
  }
  __shared__ float sh[kThreads / 32][6];
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      sh[warp][d] = b.lo[d];
      sh[warp][3 + d] = b.hi[d];
    }
  }
  __syncthreads();
  if (threadIdx.x < 6)
  {
    int const c = threadIdx.x;
    float v = sh[0][c];
    for (int w = 1; w < kThreads / 32; ++w)
      v = c < 3 ? fminf(v, sh[w][c]) : fmaxf(v, sh[w][c]);
    if (c < 3)
      atomicMin(&bounds_enc[c], floatToOrdered(v));
    else
      atomicMax(&bounds_enc[c], floatToOrdered(v));
  }
}

__global__ void initBoundsKernel(unsigned *bounds_enc)
{
  if (threadIdx.x < 3)
    bounds_enc[threadIdx.x] = floatToOrdered(FLT_MAX);
  else if (threadIdx.x < 6)
    bounds_enc[threadIdx.x] = floatToOrdered(-FLT_MAX);
}

__global__ void decodeBoundsKernel(unsigned const *bounds_enc, float *bounds6)
{
  if (threadIdx.x < 6)
    bounds6[threadIdx.x] = orderedToFloat(bounds_enc[threadIdx.x]);
}

// ---- Morton64: SpaceFillingCurves.hpp:45-57, TranslateAndScale.hpp:25-37,
//      MortonCode.hpp:320-347 ----------------------------------------------------
__device__ __forceinline__ unsigned long long morton64Of(float const c[3], float const *__restrict__ bounds6)
{
  unsigned long long r = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    float const a = bounds6[d], b = bounds6[3 + d];
    float t = (a != b) ? __fdiv_rn(__fsub_rn(c[d], a), __fsub_rn(b, a)) : 0.f;
    float x = __fmul_rn(t, 2097152.f);
    x = x < 0.f ? 0.f : (2097151.f < x ? 2097151.f : x); // Kokkos::clamp
    r += expandBits2_64((unsigned long long)x) << (2 - d);
  }
  return r;
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) morton64Kernel(float const *__restrict__ prims, int64_t n,
                                                          float const *__restrict__ bounds6,
                                                          unsigned long long *__restrict__ codes)
{
  int64_t const i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n)
    return;
  float c[3];
  primCentroid<KIND>(prims, i, c);
  codes[i] = morton64Of(c, bounds6);
}

// Morton32 of predicate centroids: MortonCode.hpp:293-316, LinearBVH.hpp:287-298
template <int PRED>
__global__ void __launch_bounds__(kThreads) morton32Kernel(float const *__restrict__ preds, int64_t q,
                                                          float const *__restrict__ bounds6,
                                                          unsigned *__restrict__ codes)
{
  int64_t const i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= q)
    return;
  float c[3];
  if (PRED == ABX_PRED_SPHERE3F)
  {
    c[0] = preds[4 * i];
    c[1] = preds[4 * i + 1];
    c[2] = preds[4 * i + 2];
  }
  else if (PRED == ABX_PRED_BOX3F)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
      c[d] = __fdiv_rn(__fadd_rn(preds[6 * i + d], preds[6 * i + 3 + d]), 2.f);
  }
  else if (PRED == ABX_PRED_RAY3F)
  {
    // returnCentroid(ray) = origin (geometry/ArborX_Ray.hpp)
    c[0] = preds[6 * i];
    c[1] = preds[6 * i + 1];
    c[2] = preds[6 * i + 2];
  }
  else
  {
    c[0] = preds[3 * i];
    c[1] = preds[3 * i + 1];
    c[2] = preds[3 * i + 2];
  }
  unsigned r = 0;
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    float const a = bounds6[d], b = bounds6[3 + d];
    float t = (a != b) ? __fdiv_rn(__fsub_rn(c[d], a), __fsub_rn(b, a)) : 0.f;
    float x = __fmul_rn(t, 1024.f);
    x = x < 0.f ? 0.f : (1023.f < x ? 1023.f : x);
    r += expandBits2_32((unsigned)x) << (2 - d);
  }
  codes[i] = r;
}

// ---- hierarchy: TreeConstruction.hpp:74-322 ------------------------------------
// delta(i): :131-170.  Equal codes fall back to the index bits and always compare
// below any real difference.
__device__ __forceinline__ long long deltaOf(unsigned long long const *__restrict__ codes, int i, int n_int)
{
  if (i < 0 || i >= n_int)
    return LLONG_MAX;
  unsigned long long const x = codes[i] ^ codes[i + 1];
  unsigned long long const fallback = (unsigned long long)LLONG_MIN + (unsigned long long)(i ^ (i + 1));
  return (long long)((x ? x : fallback) - 1ull);
}

template <int KIND>
__device__ __forceinline__ void loadLeafBox(float4 const *leaf_box, int pos, Box &b, int &ref)
{
  if (KIND == ABX_PRIM_POINT3F)
  {
    float4 v = ldcg4(leaf_box + pos);
    b.lo[0] = b.hi[0] = v.x;
    b.lo[1] = b.hi[1] = v.y;
    b.lo[2] = b.hi[2] = v.z;
    ref = refLeaf(__float_as_uint(v.w));
  }
  else
  {
    float4 lo = ldcg4(leaf_box + 2 * (size_t)pos), hi = ldcg4(leaf_box + 2 * (size_t)pos + 1);
    b.lo[0] = lo.x;
    b.lo[1] = lo.y;
    b.lo[2] = lo.z;
    b.hi[0] = hi.x;
    b.hi[1] = hi.y;
    b.hi[2] = hi.z;
    ref = refLeaf(__float_as_uint(lo.w));
  }
}

__device__ __forceinline__ void loadNodeBox(Node64 const *nodes, int k, Box &b)
{
  float4 const *f = reinterpret_cast<float4 const *>(nodes + k);
  float4 a0 = ldcg4(f), a1 = ldcg4(f + 1), a2 = ldcg4(f + 2), a3 = ldcg4(f + 3);
  b.lo[0] = fminf(a0.x, a2.x);
  b.lo[1] = fminf(a0.y, a2.y);
  b.lo[2] = fminf(a0.z, a2.z);
  b.hi[0] = fmaxf(a1.x, a3.x);
  b.hi[1] = fmaxf(a1.y, a3.y);
  b.hi[2] = fmaxf(a1.z, a3.z);
}

// ---- hierarchy kernels ---------------------------------------------------------------
// Apetrei's bottom-up construction (TreeConstruction.hpp:197-311): a node is finished
// by the second of its two children to arrive at an atomic flag.
//
// B200 shape.  The LBVH is the Cartesian tree of the delta array (key = (delta, index),
// larger key = closer to the root).  A block owns a chunk [a, b] of kHierThreads
// consecutive sorted leaves; node p lies entirely inside the chunk iff a larger key
// exists on both sides of p inside [a-1, b]:
//     max(delta[a-1 .. p-1]) > delta(p)   and   max(delta[p+1 .. b]) >= delta(p).
// hierarchyLocalKernel builds these "local" nodes (> 95 % of all nodes) in ROUNDS over a
// shared-memory work queue: round 0 holds the chunk's leaves, every finished node is
// pushed for the next round, so all rounds run with compact, fully populated warps
// (the plain one-thread-per-leaf walk leaves 4-8 of 32 lanes alive after two levels).
// Flags, boxes, ranges and deltas live in shared memory; no global atomics, no
// device-scope fences.  A subtree whose parent is not local is appended to a pending
// list; hierarchyGlobalKernel finishes those few nodes with the global CAS +
// __threadfence protocol.
constexpr int kHierThreads = 512;
constexpr int kHierWarps = kHierThreads / 32;

template <int KIND>
struct LeafFloats
{
  static constexpr int value = (KIND == ABX_PRIM_POINT3F) ? 3 : 6;
};

// a finished subtree waiting for its (non-local) parent
struct PendingNode
{
  int range_left, range_right, ref;
  float box[6];
};

__device__ __forceinline__ long long shflUp64(long long v, int o)
{
  return __shfl_up_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ long long shflDown64(long long v, int o)
{
  return __shfl_down_sync(0xffffffffu, v, o);
}

__device__ __forceinline__ void writeNode(Node64 *nodes, int k, Box const &L, int lref, Box const &R, int rref,
                                          int range_left, int range_right)
{
  float4 *f = reinterpret_cast<float4 *>(nodes + k);
  f[0] = make_float4(L.lo[0], L.lo[1], L.lo[2], __int_as_float(lref));
  f[1] = make_float4(L.hi[0], L.hi[1], L.hi[2], __int_as_float(rref));
  f[2] = make_float4(R.lo[0], R.lo[1], R.lo[2], __int_as_float(range_left));
  f[3] = make_float4(R.hi[0], R.hi[1], R.hi[2], __int_as_float(range_right));
}

template <int KIND>
__global__ void __launch_bounds__(kHierThreads)
    hierarchyLocalKernel(int n, unsigned long long const *__restrict__ codes, unsigned const *__restrict__ perm,
                         float const *__restrict__ prims, Node64 *nodes, float4 *leaf_box, float4 *leaf_tri,
                         PendingNode *pending, unsigned *pending_count, float *bounds6)
{
  constexpr int LF = LeafFloats<KIND>::value;
  constexpr int T = kHierThreads;
  __shared__ long long sdelta[T + 1]; // sdelta[j] = delta(a - 1 + j)
  __shared__ int sflag[T];            // per parent p-a: -1 untouched, -2 not local, else a range end
  __shared__ float snode[T][6];       // box of a finished local node (by Karras index - a)
  __shared__ short srl[T], srr[T];    // its range, relative to a
  __shared__ float sleaf[T][LF];      // leaf boxes of the chunk
  __shared__ unsigned sperm[T];
  __shared__ unsigned short squeue[2][T];
  __shared__ unsigned short spend[T]; // items handed to the global kernel
  __shared__ int sqn[2];
  __shared__ int spn;
  __shared__ unsigned spbase;
  __shared__ long long swarp[2][kHierWarps];

  int const tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int const a = blockIdx.x * T;
  int const cn = min(T, n - a); // leaves in this chunk
  int const n_int = n - 1;
  int const i = a + tid;
  bool const active = tid < cn;

  if (active)
  {
    unsigned const orig = perm[i];
    Box const box = primBox<KIND>(prims, orig);
    if (KIND == ABX_PRIM_POINT3F)
      leaf_box[i] = make_float4(box.lo[0], box.lo[1], box.lo[2], __uint_as_float(orig));
    else
    {
      leaf_box[2 * (size_t)i] = make_float4(box.lo[0], box.lo[1], box.lo[2], __uint_as_float(orig));
      leaf_box[2 * (size_t)i + 1] = make_float4(box.hi[0], box.hi[1], box.hi[2], 0.f);
    }
    if (KIND == ABX_PRIM_TRI3F)
    {
      float const *t = prims + 9 * (size_t)orig;
      leaf_tri[3 * (size_t)i] = make_float4(t[0], t[1], t[2], 0.f);
      leaf_tri[3 * (size_t)i + 1] = make_float4(t[3], t[4], t[5], 0.f);
      leaf_tri[3 * (size_t)i + 2] = make_float4(t[6], t[7], t[8], 0.f);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
      sleaf[tid][d] = box.lo[d];
    if (LF == 6)
    {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        sleaf[tid][LF == 6 ? 3 + d : d] = box.hi[d];
    }
    sperm[tid] = orig;
    sdelta[tid + 1] = deltaOf(codes, i, n_int);
  }
  else
    sdelta[tid + 1] = LLONG_MAX;
  if (tid == 0)
  {
    sdelta[0] = deltaOf(codes, a - 1, n_int);
    sqn[0] = 0;
    sqn[1] = 0;
    spn = 0;
  }
  __syncthreads();

  // locality of parent p = a + tid (needs leaves p and p+1 in the chunk)
  {
    long long const mine = sdelta[tid + 1];
    long long pre = sdelta[tid]; // inclusive prefix max of sdelta[0..tid]
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      long long t = shflUp64(pre, o);
      if (lane >= o)
        pre = max(pre, t);
    }
    long long suf = (tid + 2 <= cn) ? sdelta[tid + 2] : LLONG_MIN; // inclusive suffix max of sdelta[t+2], t >= tid
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      long long t = shflDown64(suf, o);
      if (lane + o < 32)
        suf = max(suf, t);
    }
    if (lane == 31)
      swarp[0][warp] = pre;
    if (lane == 0)
      swarp[1][warp] = suf;
    __syncthreads();
    for (int w = 0; w < warp; ++w)
      pre = max(pre, swarp[0][w]);
    for (int w = warp + 1; w < kHierWarps; ++w)
      suf = max(suf, swarp[1][w]);
    bool const in_chunk = tid + 1 < cn;
    bool const local = in_chunk && (pre > mine) && (suf >= mine);
    sflag[tid] = local ? -1 : -2;
  }
  __syncthreads();

  // rounds over the work queue; item < T: leaf (chunk position), item >= T: finished
  // local node (Karras index - a + T)
  int cur = 0;
  int count = cn; // round 0: one leaf per thread
  bool first_round = true;
  while (count > 0)
  {
    if (tid < count)
    {
      int const item = first_round ? tid : (int)squeue[cur][tid];
      int range_left, range_right, ref;
      Box box;
      if (item < T)
      {
        range_left = range_right = a + item;
        ref = refLeaf(sperm[item]);
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          box.lo[d] = sleaf[item][d];
          box.hi[d] = LF == 6 ? sleaf[item][LF == 6 ? 3 + d : d] : sleaf[item][d];
        }
      }
      else
      {
        int const kk = item - T;
        range_left = a + srl[kk];
        range_right = a + srr[kk];
        ref = a + kk;
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          box.lo[d] = snode[kk][d];
          box.hi[d] = snode[kk][3 + d];
        }
      }
      long long delta_left = sdelta[range_left - a];       // delta(range_left - 1)
      long long delta_right = sdelta[range_right - a + 1]; // delta(range_right)
      bool const is_left_child = delta_right < delta_left;
      int const apetrei_parent = is_left_child ? range_right : range_left - 1;
      int const lp = apetrei_parent - a;
      if (lp < 0 || lp >= cn - 1 || sflag[lp] == -2)
      {
        // parent straddles the chunk boundary: hand the subtree to the global kernel
        // (written out at the end, one global atomic per block)
        spend[atomicAdd(&spn, 1)] = (unsigned short)item;
      }
      else
      {
        int const old = atomicCAS(&sflag[lp], -1, is_left_child ? range_left : range_right);
        if (old != -1)
        {
          // second to arrive: the sibling's record was written in an earlier round
          int sib_pos;
          bool sib_is_leaf;
          if (is_left_child)
          {
            range_right = old;
            sib_pos = apetrei_parent + 1;
            sib_is_leaf = (sib_pos == range_right);
            delta_right = sdelta[range_right - a + 1];
          }
          else
          {
            range_left = old;
            sib_pos = apetrei_parent;
            sib_is_leaf = (sib_pos == range_left);
            delta_left = sdelta[range_left - a];
          }
          int const sl = sib_pos - a;
          Box sib;
          int sib_ref;
          if (sib_is_leaf)
          {
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
              sib.lo[d] = sleaf[sl][d];
              sib.hi[d] = LF == 6 ? sleaf[sl][LF == 6 ? 3 + d : d] : sleaf[sl][d];
            }
            sib_ref = refLeaf(sperm[sl]);
          }
          else
          {
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
              sib.lo[d] = snode[sl][d];
              sib.hi[d] = snode[sl][3 + d];
            }
            sib_ref = sib_pos;
          }
          int const karras_parent = delta_right < delta_left ? range_right : range_left;
          if (is_left_child)
            writeNode(nodes, karras_parent, box, ref, sib, sib_ref, range_left, range_right);
          else
            writeNode(nodes, karras_parent, sib, sib_ref, box, ref, range_left, range_right);
          boxUnion(box, sib);
          if (karras_parent == 0)
          {
            // root (the whole tree fits in one chunk): TreeConstruction.hpp:108-113
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
              bounds6[d] = box.lo[d];
              bounds6[3 + d] = box.hi[d];
            }
          }
          else
          {
            int const kk = karras_parent - a; // a local node's Karras index is an end of its range
#pragma unroll
            for (int d = 0; d < 3; ++d)
            {
              snode[kk][d] = box.lo[d];
              snode[kk][3 + d] = box.hi[d];
            }
            srl[kk] = (short)(range_left - a);
            srr[kk] = (short)(range_right - a);
            int const slot = atomicAdd(&sqn[cur ^ 1], 1);
            squeue[cur ^ 1][slot] = (unsigned short)(T + kk);
          }
        }
      }
    }
    __syncthreads();
    cur ^= 1;
    first_round = false;
    count = sqn[cur];
    __syncthreads();
    if (tid == 0)
      sqn[cur ^ 1] = 0;
    __syncthreads();
  }

  // append this chunk's maximal local subtrees to the pending list
  if (tid == 0)
    spbase = spn ? atomicAdd(pending_count, (unsigned)spn) : 0u;
  __syncthreads();
  if (tid < spn)
  {
    int const item = spend[tid];
    PendingNode pn;
    if (item < T)
    {
      pn.range_left = pn.range_right = a + item;
      pn.ref = refLeaf(sperm[item]);
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        pn.box[d] = sleaf[item][d];
        pn.box[3 + d] = LF == 6 ? sleaf[item][LF == 6 ? 3 + d : d] : sleaf[item][d];
      }
    }
    else
    {
      int const kk = item - T;
      pn.range_left = a + srl[kk];
      pn.range_right = a + srr[kk];
      pn.ref = a + kk;
#pragma unroll
      for (int d = 0; d < 6; ++d)
        pn.box[d] = snode[kk][d];
    }
    pending[spbase + tid] = pn;
  }
}

// finishes the nodes that straddle chunk boundaries: global CAS flags in `ranges`,
// records published with __threadfence before the flag, read with __ldcg after it
// (the reference's CAS + load_fence, TreeConstruction.hpp:241-270)
template <int KIND>
__global__ void __launch_bounds__(256)
    hierarchyGlobalKernel(int n, unsigned long long const *__restrict__ codes, Node64 *nodes,
                          float4 const *leaf_box, int *ranges, PendingNode const *__restrict__ pending,
                          unsigned const *__restrict__ pending_count, float *bounds6)
{
  int const n_int = n - 1;
  unsigned const total = *pending_count;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x)
  {
    PendingNode const pn = pending[it];
    int range_left = pn.range_left, range_right = pn.range_right;
    int cur_ref = pn.ref;
    Box box;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      box.lo[d] = pn.box[d];
      box.hi[d] = pn.box[3 + d];
    }
    long long delta_left = deltaOf(codes, range_left - 1, n_int);
    long long delta_right = deltaOf(codes, range_right, n_int);
    // the subtree's own record (leaf_box / Node64) was written by the local kernel: visible
    while (true)
    {
      bool const is_left_child = delta_right < delta_left;
      Box sib;
      int sib_ref;
      if (is_left_child)
      {
        int const apetrei_parent = range_right;
        int const old = atomicCAS(&ranges[apetrei_parent], -1, range_left);
        if (old == -1)
          break; // first to arrive: the sibling's thread finishes this node
        range_right = old;
        int const right_child = apetrei_parent + 1;
        bool const right_is_leaf = (right_child == range_right);
        delta_right = deltaOf(codes, range_right, n_int);
        __threadfence(); // acquire: the sibling published its record before its CAS
        if (right_is_leaf)
          loadLeafBox<KIND>(leaf_box, right_child, sib, sib_ref);
        else
        {
          loadNodeBox(nodes, right_child, sib);
          sib_ref = right_child;
        }
      }
      else
      {
        int const apetrei_parent = range_left - 1;
        int const old = atomicCAS(&ranges[apetrei_parent], -1, range_right);
        if (old == -1)
          break;
        range_left = old;
        int const left_child = apetrei_parent;
        bool const left_is_leaf = (left_child == range_left);
        delta_left = deltaOf(codes, range_left - 1, n_int);
        __threadfence();
        if (left_is_leaf)
          loadLeafBox<KIND>(leaf_box, left_child, sib, sib_ref);
        else
        {
          loadNodeBox(nodes, left_child, sib);
          sib_ref = left_child;
        }
      }
      int const karras_parent = delta_right < delta_left ? range_right : range_left;
      if (is_left_child)
        writeNode(nodes, karras_parent, box, cur_ref, sib, sib_ref, range_left, range_right);
      else
        writeNode(nodes, karras_parent, sib, sib_ref, box, cur_ref, range_left, range_right);
      boxUnion(box, sib);
      cur_ref = karras_parent;
      if (karras_parent == 0)
      {
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          bounds6[d] = box.lo[d];
          bounds6[3 + d] = box.hi[d];
        }
        break;
      }
      __threadfence(); // release the finished node before signalling its parent
    }
  }
}

template <int KIND>
__global__ void singleLeafKernel(float const *__restrict__ prims, float4 *leaf_box, float4 *leaf_tri, unsigned *perm,
                                 unsigned long long *codes, float *bounds6)
{
  // TreeConstruction.hpp:43-69
  Box box = primBox<KIND>(prims, 0);
  if (KIND == ABX_PRIM_POINT3F)
    leaf_box[0] = make_float4(box.lo[0], box.lo[1], box.lo[2], __uint_as_float(0u));
  else
  {
    leaf_box[0] = make_float4(box.lo[0], box.lo[1], box.lo[2], __uint_as_float(0u));
    leaf_box[1] = make_float4(box.hi[0], box.hi[1], box.hi[2], 0.f);
  }
  if (KIND == ABX_PRIM_TRI3F)
  {
    leaf_tri[0] = make_float4(prims[0], prims[1], prims[2], 0.f);
    leaf_tri[1] = make_float4(prims[3], prims[4], prims[5], 0.f);
    leaf_tri[2] = make_float4(prims[6], prims[7], prims[8], 0.f);
  }
  perm[0] = 0;
  codes[0] = 0;
  for (int d = 0; d < 3; ++d)
  {
    bounds6[d] = box.lo[d];
    bounds6[3 + d] = box.hi[d];
  }
}

__global__ void iotaKernel(unsigned *p, int64_t n)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    p[i] = (unsigned)i;
}

__global__ void emptyBoundsKernel(float *bounds6)
{
  if (threadIdx.x < 3)
    bounds6[threadIdx.x] = FLT_MAX;
  else if (threadIdx.x < 6)
    bounds6[threadIdx.x] = -FLT_MAX;
}

// ---- export in the reference's node layout (tests) -----------------------------
// rope rule: TreeConstruction.hpp:173-195
__device__ __forceinline__ int ropeOf(unsigned long long const *__restrict__ codes, int range_right, int n)
{
  int const n_int = n - 1;
  if (range_right == n_int)
    return -1;
  long long const dr = deltaOf(codes, range_right, n_int);
  return dr < deltaOf(codes, range_right + 1, n_int) ? range_right + 1 : (range_right + 1) + n;
}

__global__ void __launch_bounds__(kThreads)
    exportKernel(int n, Node64 const *__restrict__ nodes, unsigned const *__restrict__ perm,
                 unsigned long long const *__restrict__ codes, int *leaf_rope, unsigned *leaf_index, int *left_child,
                 int *rope, float *boxes6, unsigned long long *codes_out)
{
  int const i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n)
    return;
  if (leaf_rope)
    leaf_rope[i] = n > 1 ? ropeOf(codes, i, n) : -1;
  if (leaf_index)
    leaf_index[i] = perm[i];
  if (codes_out)
    codes_out[i] = codes[i];
  if (i < n - 1)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + i);
    float4 a0 = f[0], a1 = f[1], a2 = f[2], a3 = f[3];
    int const lref = __float_as_int(a0.w);
    int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
    if (left_child)
      left_child[i] = refIsLeaf(lref) ? rl : lref + n;
    if (rope)
      rope[i] = ropeOf(codes, rr, n);
    if (boxes6)
    {
      boxes6[6 * (size_t)i + 0] = fminf(a0.x, a2.x);
      boxes6[6 * (size_t)i + 1] = fminf(a0.y, a2.y);
      boxes6[6 * (size_t)i + 2] = fminf(a0.z, a2.z);
      boxes6[6 * (size_t)i + 3] = fmaxf(a1.x, a3.x);
      boxes6[6 * (size_t)i + 4] = fmaxf(a1.y, a3.y);
      boxes6[6 * (size_t)i + 5] = fmaxf(a1.z, a3.z);
    }
  }
}

#define ABX_DISPATCH_PRIM(kind, CALL)                                                                                 \
  switch (kind)                                                                                                        \
  {                                                                                                                    \
  case ABX_PRIM_POINT3F:                                                                                               \
  {                                                                                                                    \
    constexpr int K = ABX_PRIM_POINT3F;                                                                                \
    CALL;                                                                                                              \
    break;                                                                                                             \
  }                                                                                                                    \
  case ABX_PRIM_BOX3F:                                                                                                 \
  {                                                                                                                    \
    constexpr int K = ABX_PRIM_BOX3F;                                                                                  \
    CALL;                                                                                                              \
    break;                                                                                                             \
  }                                                                                                                    \
  case ABX_PRIM_TRI3F:                                                                                                 \
  {                                                                                                                    \
    constexpr int K = ABX_PRIM_TRI3F;                                                                                  \
    CALL;                                                                                                              \
    break;                                                                                                             \
  }                                                                                                                    \
  default:                                                                                                             \
    setError("unknown primitive kind");                                                                                \
    return ABX_ERR_ARG;                                                                                                \
  }

} // namespace

abx_status sceneBounds(cudaStream_t s, int kind, void const *prims, int64_t n, unsigned *bounds_enc6)
{
  ABX_LAUNCH(initBoundsKernel, 1, 32, 0, s, bounds_enc6);
  if (n <= 0)
    return ABX_OK;
  int const grid = (int)std::min<int64_t>(divUp(n, kThreads), kNumSMs * 8);
  ABX_DISPATCH_PRIM(kind, ABX_LAUNCH((sceneBoundsKernel<K>), grid, kThreads, 0, s, (float const *)prims, n, bounds_enc6));
  return ABX_OK;
}

abx_status decodeBounds(cudaStream_t s, unsigned const *bounds_enc6, float *bounds6)
{
  ABX_LAUNCH(decodeBoundsKernel, 1, 32, 0, s, bounds_enc6, bounds6);
  return ABX_OK;
}

abx_status morton64(cudaStream_t s, int kind, void const *prims, int64_t n, float const *bounds6, uint64_t *codes)
{
  if (n <= 0)
    return ABX_OK;
  ABX_DISPATCH_PRIM(kind, ABX_LAUNCH((morton64Kernel<K>), divUp(n, kThreads), kThreads, 0, s, (float const *)prims, n,
                                     bounds6, (unsigned long long *)codes));
  return ABX_OK;
}

abx_status morton32(cudaStream_t s, int pred_kind, void const *preds, int64_t q, float const *bounds6, uint32_t *codes)
{
  if (q <= 0)
    return ABX_OK;
  int const grid = divUp(q, kThreads);
  switch (pred_kind)
  {
  case ABX_PRED_SPHERE3F:
    ABX_LAUNCH((morton32Kernel<ABX_PRED_SPHERE3F>), grid, kThreads, 0, s, (float const *)preds, q, bounds6, codes);
    break;
  case ABX_PRED_BOX3F:
    ABX_LAUNCH((morton32Kernel<ABX_PRED_BOX3F>), grid, kThreads, 0, s, (float const *)preds, q, bounds6, codes);
    break;
  case ABX_PRED_POINT3F:
    ABX_LAUNCH((morton32Kernel<ABX_PRED_POINT3F>), grid, kThreads, 0, s, (float const *)preds, q, bounds6, codes);
    break;
  case ABX_PRED_RAY3F:
    ABX_LAUNCH((morton32Kernel<ABX_PRED_RAY3F>), grid, kThreads, 0, s, (float const *)preds, q, bounds6, codes);
    break;
  default:
    setError("unknown predicate kind");
    return ABX_ERR_ARG;
  }
  return ABX_OK;
}

static void destroyTree(abx_bvh *t)
{
  if (!t)
    return;
  cudaStream_t s = t->stream;
  deviceFree(t->nodes, s);
  deviceFree(t->leaf_box, s);
  deviceFree(t->leaf_tri, s);
  deviceFree(t->perm, s);
  deviceFree(t->codes, s);
  deviceFree(t->bounds_dev, s);
  delete t;
}

// codes (sorted) and perm must already be in bvh; fills nodes / leaf arrays / bounds
abx_status buildHierarchy(cudaStream_t s, abx_bvh *t, void const *prims)
{
  int const n = (int)t->n;
  TempBuffer<int> ranges;
  TempBuffer<PendingNode> pending;
  TempBuffer<unsigned> pending_count;
  ABX_TRY(ranges.alloc(n - 1, s));
  ABX_TRY(pending.alloc(n, s)); // every maximal chunk-local subtree, at most one per leaf
  ABX_TRY(pending_count.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(ranges.ptr, 0xff, sizeof(int) * (size_t)(n - 1), s));
  ABX_CUDA_TRY(cudaMemsetAsync(pending_count.ptr, 0, sizeof(unsigned), s));
  ABX_DISPATCH_PRIM(t->kind, ABX_LAUNCH((hierarchyLocalKernel<K>), divUp(n, kHierThreads), kHierThreads, 0, s, n,
                                        (unsigned long long const *)t->codes, t->perm, (float const *)prims, t->nodes,
                                        t->leaf_box, t->leaf_tri, pending.ptr, pending_count.ptr, t->bounds_dev));
  // the local kernel's records are complete at the kernel boundary; the global kernel
  // only orders its own writes
  int const grid = std::min(divUp(n, 256 * 8), kNumSMs * 8);
  ABX_DISPATCH_PRIM(t->kind, ABX_LAUNCH((hierarchyGlobalKernel<K>), std::max(grid, 1), 256, 0, s, n,
                                        (unsigned long long const *)t->codes, t->nodes, t->leaf_box, ranges.ptr,
                                        pending.ptr, pending_count.ptr, t->bounds_dev));
  return ABX_OK;
}

abx_status buildTree(cudaStream_t s, int kind, void const *prims, int64_t n, uint64_t const *sorted_codes,
                     abx_bvh **out)
{
  if (kind != ABX_PRIM_POINT3F && kind != ABX_PRIM_BOX3F && kind != ABX_PRIM_TRI3F)
  {
    setError("unknown primitive kind");
    return ABX_ERR_ARG;
  }
  if (n < 0 || n >= (int64_t)1 << 30)
  {
    setError("number of primitives must be in [0, 2^30)");
    return ABX_ERR_ARG;
  }
  if (n > 0 && !prims)
  {
    setError("null primitives");
    return ABX_ERR_ARG;
  }
  abx_bvh *t = new abx_bvh;
  t->kind = kind;
  t->n = n;
  t->stream = s;
  cudaGetDevice(&t->device);
  auto fail = [&](abx_status st) {
    destroyTree(t);
    return st;
  };
#define ABX_TRY_T(expr)                                                                                               \
  do                                                                                                                   \
  {                                                                                                                    \
    abx_status _s = (expr);                                                                                            \
    if (_s != ABX_OK)                                                                                                  \
      return fail(_s);                                                                                                 \
  } while (0)

  ABX_TRY_T(deviceAlloc((void **)&t->bounds_dev, 6 * sizeof(float), s));
  t->bytes += 6 * sizeof(float);
  if (n == 0)
  {
    // LinearBVH.hpp:203-206: bounds() stays the default (empty) box
    ABX_LAUNCH(emptyBoundsKernel, 1, 32, 0, s, t->bounds_dev);
    *out = t;
    return ABX_OK;
  }
  size_t const leaf_f4 = (kind == ABX_PRIM_POINT3F) ? 1 : 2;
  ABX_TRY_T(deviceAlloc((void **)&t->leaf_box, sizeof(float4) * leaf_f4 * n, s));
  t->bytes += sizeof(float4) * leaf_f4 * n;
  if (kind == ABX_PRIM_TRI3F)
  {
    ABX_TRY_T(deviceAlloc((void **)&t->leaf_tri, sizeof(float4) * 3 * n, s));
    t->bytes += sizeof(float4) * 3 * n;
  }
  ABX_TRY_T(deviceAlloc((void **)&t->perm, sizeof(uint32_t) * n, s));
  ABX_TRY_T(deviceAlloc((void **)&t->codes, sizeof(uint64_t) * n, s));
  t->bytes += 12 * n;
  if (n == 1)
  {
    ABX_DISPATCH_PRIM(kind, ABX_LAUNCH((singleLeafKernel<K>), 1, 1, 0, s, (float const *)prims, t->leaf_box,
                                       t->leaf_tri, t->perm, (unsigned long long *)t->codes, t->bounds_dev));
    *out = t;
    return ABX_OK;
  }
  ABX_TRY_T(deviceAlloc((void **)&t->nodes, sizeof(Node64) * (n - 1), s));
  t->bytes += sizeof(Node64) * (n - 1);

  if (sorted_codes)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(t->codes, sorted_codes, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s));
    ABX_LAUNCH(iotaKernel, divUp(n, kThreads), kThreads, 0, s, t->perm, n);
  }
  else
  {
    TempBuffer<unsigned> enc;
    ABX_TRY_T(enc.alloc(6, s));
    ABX_TRY_T(sceneBounds(s, kind, prims, n, enc.ptr));
    ABX_TRY_T(decodeBounds(s, enc.ptr, t->bounds_dev));
    ABX_TRY_T(morton64(s, kind, prims, n, t->bounds_dev, t->codes));
    // Morton64 codes use 63 bits (3 x 21).  Double-buffered: the tree keeps whichever pair of
    // buffers the sort finished in
    TempBuffer<uint64_t> codes_alt;
    TempBuffer<uint32_t> perm_alt;
    ABX_TRY_T(codes_alt.alloc(n, s));
    ABX_TRY_T(perm_alt.alloc(n, s));
    uint64_t *kb[2] = {t->codes, codes_alt.ptr};
    uint32_t *vb[2] = {t->perm, perm_alt.ptr};
    int cur = 0;
    ABX_TRY_T(sortPairsU64DB(s, kb, vb, &cur, n, /*iota_vals=*/true, 63));
    if (cur != 0)
    {
      std::swap(t->codes, codes_alt.ptr);
      std::swap(t->perm, perm_alt.ptr);
    }
  }
  ABX_TRY_T(buildHierarchy(s, t, prims));
  *out = t;
  return ABX_OK;
#undef ABX_TRY_T
}

abx_status exportReference(cudaStream_t s, abx_bvh *t, int32_t *leaf_rope, uint32_t *leaf_index, int32_t *left_child,
                           int32_t *rope, float *boxes6, uint64_t *codes)
{
  int const n = (int)t->n;
  if (n == 0)
    return ABX_OK;
  ABX_LAUNCH(exportKernel, divUp(n, kThreads), kThreads, 0, s, n, t->nodes, t->perm,
             (unsigned long long const *)t->codes, leaf_rope, leaf_index, left_child, rope, boxes6,
             (unsigned long long *)codes);
  return ABX_OK;
}

} // namespace abx

extern "C" abx_status abx_bvh_destroy(abx_bvh *bvh)
{
  abx::destroyTree(bvh);
  return ABX_OK;
}
