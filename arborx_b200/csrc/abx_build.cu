// abx_build.cu -- BVH construction: scene bounds, Morton64 codes, Apetrei hierarchy.
//
// Follows the behaviour of spatial/ArborX_LinearBVH.hpp:171-256 and
// spatial/detail/ArborX_TreeConstruction.hpp:27-39,74-322, but emits the compact
// two-children-per-record Node64 layout (abx_common.cuh) instead of the
// reference's {left_child, rope, box} nodes.
#include "abx_common.cuh"

#include <mutex>

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
// B200 shape.  A node whose leaf range lies inside a window of consecutive sorted leaves
// can be built from that window alone, and both of its children arrive at its flag from
// inside the window; a node whose range leaves the window sees at most one arrival.  So
// windows are processed independently and whatever is left with ONE arrival (or whose
// parent slot is outside the window) is handed to the next, wider window:
//   stage 1  one warp per 64 leaves, rounds over a warp-private work queue in shared
//            memory, __syncwarp only: ~87 % of the nodes
//   stage 2  warp 0 of the block over the block's W * 64 leaves: the maximal subtrees
//            the warps left over (~12 per warp)
//   stage 3  hierarchyGlobalKernel over the maximal subtrees of every block, with the
//            reference's global CAS + fence protocol: ~4 % at W = 4
// Measured at 10M points (scripts/time_build.py): W = 2: 0.59 + 0.25 ms (local + global),
// W = 4: 0.57 + 0.16, W = 8: 0.67 + 0.10, W = 16: 0.75 + 0.09.
// Rounds keep the warps fully populated (round r holds the nodes of height r; the plain
// one-thread-per-leaf walk leaves 4-8 of 32 lanes alive after two levels), flags, boxes,
// ranges and deltas live in shared memory, and stages 1-2 need no device-scope fence and
// only two block barriers.
#ifndef ABX_HIER_MINB
#define ABX_HIER_MINB 1
#endif
#ifndef ABX_HIER_WARP_LEAVES
#define ABX_HIER_WARP_LEAVES 64
#endif
constexpr int kHierWarpLeaves = ABX_HIER_WARP_LEAVES; // stage-1 window
constexpr int kHierWarpsDefault = 4; // warps per block: the stage-2 window is W * 64 leaves (ABX_HIER_WARPS overrides)
constexpr int kFlagFree = -1;                            // no child has arrived
constexpr int kFlagDone = -3;                            // both children arrived, node written

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

__device__ __forceinline__ void writeNode(Node64 *nodes, int k, Box const &L, int lref, Box const &R, int rref,
                                          int range_left, int range_right)
{
  float4 *f = reinterpret_cast<float4 *>(nodes + k);
  f[0] = make_float4(L.lo[0], L.lo[1], L.lo[2], __int_as_float(lref));
  f[1] = make_float4(L.hi[0], L.hi[1], L.hi[2], __int_as_float(rref));
  f[2] = make_float4(R.lo[0], R.lo[1], R.lo[2], __int_as_float(range_left));
  f[3] = make_float4(R.hi[0], R.hi[1], R.hi[2], __int_as_float(range_right));
}

template <int KIND, int W>
struct HierSmem
{
  static constexpr int LF = LeafFloats<KIND>::value;
  static constexpr int T = W * kHierWarpLeaves;
  long long delta[T + 1]; // delta[j] = delta(a - 1 + j)
  int flag[T];            // per parent p - a: kFlagFree, kFlagDone, or the range end of the one child so far
  float node[T][6];       // box of a finished local node (by Karras index - a)
  short rl[T], rr[T];     // its range, relative to a
  float leaf[T][LF];      // leaf boxes of the chunk
  unsigned perm[T];
  unsigned short wqueue[W][2][kHierWarpLeaves];
  unsigned short bqueue[2][T];
  unsigned short pend[T]; // items handed to the global kernel
  int bqn;
  int pn;
  unsigned pbase;
};

// item < T: leaf (chunk position); item >= T: finished local node (Karras index - a + T)
template <int KIND, int W>
__device__ __forceinline__ void loadItem(HierSmem<KIND, W> const &sm, int a, int item, int &range_left,
                                         int &range_right, int &ref, Box &box)
{
  constexpr int LF = LeafFloats<KIND>::value;
  constexpr int T = W * kHierWarpLeaves;
  if (item < T)
  {
    range_left = range_right = a + item;
    ref = refLeaf(sm.perm[item]);
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      box.lo[d] = sm.leaf[item][d];
      box.hi[d] = sm.leaf[item][LF == 6 ? 3 + d : d];
    }
  }
  else
  {
    int const kk = item - T;
    range_left = a + sm.rl[kk];
    range_right = a + sm.rr[kk];
    ref = a + kk;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      box.lo[d] = sm.node[kk][d];
      box.hi[d] = sm.node[kk][3 + d];
    }
  }
}

// One Apetrei step for `item` against the parent slots [slot_lo, slot_hi) (chunk-relative).
// Returns 0: nothing to forward (first child to arrive, or the root was written);
//         1: second child: the parent was written and is returned in `next` for the next round;
//         2: the parent's slot is outside the window: the caller hands `item` to the wider window.
template <int KIND, int W>
__device__ __forceinline__ int apetreiStep(HierSmem<KIND, W> &sm, int a, int item, int slot_lo, int slot_hi,
                                           Node64 *nodes, float *bounds6, int &next)
{
  constexpr int T = W * kHierWarpLeaves;
  // the first child to arrive only needs its range; boxes are loaded by the second
  int range_left, range_right;
  if (item < T)
    range_left = range_right = a + item;
  else
  {
    range_left = a + sm.rl[item - T];
    range_right = a + sm.rr[item - T];
  }
  long long delta_left = sm.delta[range_left - a];       // delta(range_left - 1)
  long long delta_right = sm.delta[range_right - a + 1]; // delta(range_right)
  bool const is_left_child = delta_right < delta_left;
  int const apetrei_parent = is_left_child ? range_right : range_left - 1;
  int const lp = apetrei_parent - a;
  if (lp < slot_lo || lp >= slot_hi)
    return 2;
  int const old = atomicCAS(&sm.flag[lp], kFlagFree, is_left_child ? range_left : range_right);
  if (old == kFlagFree)
    return 0;
  // second to arrive: the sibling's record was written in an earlier round
  sm.flag[lp] = kFlagDone;
  int sib_item;
  if (is_left_child)
  {
    range_right = old;
    int const sib_pos = apetrei_parent + 1;
    sib_item = (sib_pos == range_right) ? sib_pos - a : T + sib_pos - a;
    delta_right = sm.delta[range_right - a + 1];
  }
  else
  {
    range_left = old;
    int const sib_pos = apetrei_parent;
    sib_item = (sib_pos == range_left) ? sib_pos - a : T + sib_pos - a;
    delta_left = sm.delta[range_left - a];
  }
  Box box, sib;
  int ref, sib_ref, unused_l, unused_r;
  loadItem<KIND, W>(sm, a, item, unused_l, unused_r, ref, box);
  loadItem<KIND, W>(sm, a, sib_item, unused_l, unused_r, sib_ref, sib);
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
    return 0;
  }
  int const kk = karras_parent - a; // a local node's Karras index is an end of its range
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    sm.node[kk][d] = box.lo[d];
    sm.node[kk][3 + d] = box.hi[d];
  }
  sm.rl[kk] = (short)(range_left - a);
  sm.rr[kk] = (short)(range_right - a);
  next = T + kk;
  return 1;
}

// the child that arrived alone at parent slot s (flag value f): its item id
template <int W>
__device__ __forceinline__ int loneChildItem(int a, int s, int f)
{
  constexpr int T = W * kHierWarpLeaves;
  int const p = a + s;
  if (f <= p) // left child, range [f, p]: leaf p or internal node p
    return f == p ? s : T + s;
  return f == p + 1 ? s + 1 : T + s + 1; // right child, range [p + 1, f]
}

// Rounds of one warp over `queue` (two buffers of `cap` items) against the slots [slot_lo, slot_hi).
// count0 items are taken from first0 + idx in the first round (leaves) when queue_init is false.
// Items whose parent slot is outside the window go to out[] (counter *out_n, shared by several warps).
template <int KIND, int W>
__device__ __forceinline__ void warpRounds(HierSmem<KIND, W> &sm, int a, unsigned short *q0, unsigned short *q1,
                                           int count, bool leaves_first, int first0, int slot_lo, int slot_hi,
                                           unsigned short *out, int *out_n, Node64 *nodes, float *bounds6)
{
  int const lane = threadIdx.x & 31;
  unsigned const lt = (1u << lane) - 1u;
  unsigned short *cur = q0, *nxt = q1;
  while (count > 0)
  {
    int produced = 0;
    for (int base = 0; base < count; base += 32)
    {
      int const idx = base + lane;
      int r = 0, next = 0, item = 0;
      if (idx < count)
      {
        item = leaves_first ? first0 + idx : (int)cur[idx];
        r = apetreiStep<KIND, W>(sm, a, item, slot_lo, slot_hi, nodes, bounds6, next);
      }
      unsigned const m1 = __ballot_sync(0xffffffffu, r == 1);
      if (r == 1)
        nxt[produced + __popc(m1 & lt)] = (unsigned short)next;
      produced += __popc(m1);
      unsigned const m2 = __ballot_sync(0xffffffffu, r == 2);
      if (m2)
      {
        int ob = 0;
        if (lane == 0)
          ob = atomicAdd(out_n, __popc(m2));
        ob = __shfl_sync(0xffffffffu, ob, 0);
        if (r == 2)
          out[ob + __popc(m2 & lt)] = (unsigned short)item;
      }
    }
    __syncwarp();
    unsigned short *t = cur;
    cur = nxt;
    nxt = t;
    count = produced;
    leaves_first = false;
  }
}

// after the rounds: slots [slot_lo, slot_hi) that saw exactly one child hand that child to out[]
template <int KIND, int W>
__device__ __forceinline__ void collectLoneChildren(HierSmem<KIND, W> &sm, int a, int slot_lo, int slot_hi, bool reset,
                                                    unsigned short *out, int *out_n)
{
  int const lane = threadIdx.x & 31;
  unsigned const lt = (1u << lane) - 1u;
  for (int base = slot_lo; base < slot_hi; base += 32)
  {
    int const sidx = base + lane;
    int f = kFlagFree;
    if (sidx < slot_hi)
      f = sm.flag[sidx];
    bool const lone = f != kFlagFree && f != kFlagDone;
    unsigned const m = __ballot_sync(0xffffffffu, lone);
    if (m)
    {
      int ob = 0;
      if (lane == 0)
        ob = atomicAdd(out_n, __popc(m));
      ob = __shfl_sync(0xffffffffu, ob, 0);
      if (lone)
      {
        out[ob + __popc(m & lt)] = (unsigned short)loneChildItem<W>(a, sidx, f);
        if (reset)
          sm.flag[sidx] = kFlagFree;
      }
    }
  }
}

template <int KIND, int W>
__global__ void __launch_bounds__(W * 32, ABX_HIER_MINB)
    hierarchyLocalKernel(int n, unsigned long long const *__restrict__ codes, unsigned const *__restrict__ perm,
                         float const *__restrict__ prims, Node64 *nodes, float4 *leaf_box, float4 *leaf_tri,
                         PendingNode *pending, unsigned *pending_count, float *bounds6)
{
  constexpr int LF = LeafFloats<KIND>::value;
  constexpr int T = W * kHierWarpLeaves;
  constexpr int kHierThreads = W * 32;
  extern __shared__ __align__(16) unsigned char hier_smem_raw[];
  HierSmem<KIND, W> &sm = *reinterpret_cast<HierSmem<KIND, W> *>(hier_smem_raw);

  int const tid = threadIdx.x, warp = tid >> 5;
  int const a = blockIdx.x * T;
  int const cn = min(T, n - a); // leaves in this chunk
  int const n_int = n - 1;

  // leaf records out, leaf boxes / permutation / deltas / flags into shared memory
#pragma unroll
  for (int k = 0; k < T / kHierThreads; ++k)
  {
    int const j = k * kHierThreads + tid;
    int const i = a + j;
    if (j < cn)
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
        sm.leaf[j][d] = box.lo[d];
      if (LF == 6)
      {
#pragma unroll
        for (int d = 0; d < 3; ++d)
          sm.leaf[j][LF == 6 ? 3 + d : d] = box.hi[d];
      }
      sm.perm[j] = orig;
      sm.delta[j + 1] = deltaOf(codes, i, n_int);
    }
    else
      sm.delta[j + 1] = LLONG_MAX;
    sm.flag[j] = kFlagFree;
  }
  if (tid == 0)
  {
    sm.delta[0] = deltaOf(codes, a - 1, n_int);
    sm.bqn = 0;
    sm.pn = 0;
  }
  __syncthreads();

  // stage 1: every warp builds what lies inside its 64 leaves
  {
    int const w0 = warp * kHierWarpLeaves;
    int const wcount = max(0, min(kHierWarpLeaves, cn - w0));
    // parent slot s needs leaves s and s + 1 in the window
    int const slot_lo = w0, slot_hi = w0 + max(wcount - 1, 0);
    warpRounds<KIND, W>(sm, a, sm.wqueue[warp][0], sm.wqueue[warp][1], wcount, true, w0, slot_lo, slot_hi, sm.bqueue[0],
                     &sm.bqn, nodes, bounds6);
    collectLoneChildren<KIND, W>(sm, a, slot_lo, slot_hi, true, sm.bqueue[0], &sm.bqn);
  }
  __syncthreads();

  // stage 2: warp 0 joins the warps' leftovers inside the block's chunk
  if (warp == 0)
  {
    // the first round reads bqueue[0] and fills bqueue[1]; its out-of-chunk items go to pend[]
    warpRounds<KIND, W>(sm, a, sm.bqueue[0], sm.bqueue[1], sm.bqn, false, 0, 0, cn - 1, sm.pend, &sm.pn, nodes,
                     bounds6);
    collectLoneChildren<KIND, W>(sm, a, 0, cn - 1, false, sm.pend, &sm.pn);
    __syncwarp();
    // stage 3 input: append this chunk's maximal subtrees to the pending list
    int const pn_total = sm.pn;
    unsigned pbase = 0;
    if ((tid & 31) == 0 && pn_total)
      pbase = atomicAdd(pending_count, (unsigned)pn_total);
    pbase = __shfl_sync(0xffffffffu, pbase, 0);
    for (int k = tid; k < pn_total; k += 32)
    {
      PendingNode pnode;
      Box box;
      loadItem<KIND, W>(sm, a, sm.pend[k], pnode.range_left, pnode.range_right, pnode.ref, box);
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        pnode.box[d] = box.lo[d];
        pnode.box[3 + d] = box.hi[d];
      }
      pending[pbase + k] = pnode;
    }
  }
}

// Measured and rejected in round 2 (profiles/r02_hierarchy_chunk_experiment.log, bit-exact on the parity suite):
// a round-free form of this kernel.  Apetrei's merge order is a function of the delta sequence alone -- the node
// that splits at boundary s owns the range between the nearest boundary to the left with a larger delta and the
// nearest one to the right with a delta that is not smaller -- so every boundary of a chunk can be finished by its
// own thread: range ends by binary lifting over a sparse table of delta maxima, the child boxes as range unions from
// a sparse table of leaf boxes, all lanes busy, no flags.  With 256-leaf chunks (69 KB of shared memory, 3 blocks
// per SM) it took 0.77 ms against the 0.58 ms of the rounds below, 0.65 ms with 128-leaf chunks (and 0.20 instead of
// 0.13 ms in the global kernel), 0.53 / 0.47 ms with the leaves gathered by a separate streaming kernel (+0.15 ms):
// eight block-wide barriers for the tables at a third of the occupancy cost more than the idle lanes of the rounds.

// finishes the nodes that straddle chunk boundaries: global acquire-release CAS flags in `ranges`,
// sibling records read with __ldcg after it (the reference's CAS + load_fence,
// TreeConstruction.hpp:241-270)
// acquire-release CAS at device scope: the first child to arrive releases its finished record with
// it, the second acquires the sibling's record with it (TreeConstruction.hpp:241-270: CAS + load_fence)
__device__ __forceinline__ int casAcqRel(int *addr, int compare, int value)
{
  int old;
  asm volatile("atom.acq_rel.gpu.global.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old)
               : "l"(addr), "r"(compare), "r"(value)
               : "memory");
  return old;
}

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
        int const old = casAcqRel(&ranges[apetrei_parent], -1, range_left);
        if (old == -1)
          break; // first to arrive: the sibling's thread finishes this node
        range_right = old;
        int const right_child = apetrei_parent + 1;
        bool const right_is_leaf = (right_child == range_right);
        delta_right = deltaOf(codes, range_right, n_int);
        // the acquire half of the CAS orders the loads below after it; they go to L2 (ld.cg), where
        // the sibling's record was released by its own CAS
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
        int const old = casAcqRel(&ranges[apetrei_parent], -1, range_right);
        if (old == -1)
          break;
        range_left = old;
        int const left_child = apetrei_parent;
        bool const left_is_leaf = (left_child == range_left);
        delta_left = deltaOf(codes, range_left - 1, n_int);
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
      // the node written above is released by the next iteration's CAS (acq_rel)
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
  deviceFree(t->wide, s);
  deviceFree(t->wide_bad, s);
  if (t->wide_ready)
    cudaEventDestroy(t->wide_ready);
  deviceFree(t->bounds_dev, s);
  delete t;
}

template <int K, int W>
abx_status launchHierarchyLocalW(cudaStream_t s, abx_bvh *t, void const *prims, PendingNode *pending,
                                 unsigned *pending_count)
{
  static PerDeviceOnce attr; // function attributes are per device
  if (attr.needed())
  {
    ABX_CUDA_TRY(cudaFuncSetAttribute(hierarchyLocalKernel<K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(HierSmem<K, W>)));
    attr.done();
  }
  int const n = (int)t->n;
  ABX_LAUNCH_TAGGED("hierarchyLocalKernel", (hierarchyLocalKernel<K, W>), divUp(n, W * kHierWarpLeaves), W * 32,
                    sizeof(HierSmem<K, W>), s, n, (unsigned long long const *)t->codes, t->perm, (float const *)prims,
                    t->nodes, t->leaf_box, t->leaf_tri, pending, pending_count, t->bounds_dev);
  return ABX_OK;
}

template <int K>
abx_status launchHierarchyLocal(cudaStream_t s, abx_bvh *t, void const *prims, PendingNode *pending,
                                unsigned *pending_count)
{
#ifdef ABX_TUNING
  switch (ABX_TUNE_INT("ABX_HIER_WARPS", kHierWarpsDefault))
  {
  case 2: return launchHierarchyLocalW<K, 2>(s, t, prims, pending, pending_count);
  case 4: return launchHierarchyLocalW<K, 4>(s, t, prims, pending, pending_count);
  case 16: return launchHierarchyLocalW<K, 16>(s, t, prims, pending, pending_count);
  default: return launchHierarchyLocalW<K, 8>(s, t, prims, pending, pending_count);
  }
#else
  return launchHierarchyLocalW<K, kHierWarpsDefault>(s, t, prims, pending, pending_count);
#endif
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
  ABX_DISPATCH_PRIM(t->kind, ABX_TRY(launchHierarchyLocal<K>(s, t, prims, pending.ptr, pending_count.ptr)));
  // the local kernel's records are complete at the kernel boundary; the global kernel
  // only orders its own writes
  int const grid = std::min(divUp(n, 256 * 8), kNumSMs * 8);
  ABX_DISPATCH_PRIM(t->kind, ABX_LAUNCH((hierarchyGlobalKernel<K>), std::max(grid, 1), 256, 0, s, n,
                                        (unsigned long long const *)t->codes, t->nodes, t->leaf_box, ranges.ptr,
                                        pending.ptr, pending_count.ptr, t->bounds_dev));
  return ABX_OK;
}

// ---- 4-wide nodes (Wide64, abx_common.cuh) -----------------------------------------------------------------
// One thread per binary internal node: its children, with internal children of more than kWideRun leaves replaced
// by their own two children.  Slots 0-1 belong to the left child, 2-3 to the right child; an unexpanded side fills
// its first slot only.  All indices are compile-time constants, so the twelve child boxes stay in registers.
// Quantisation: q = floor / ceil of (coordinate - origin) * (1 / scale) with directed rounding, one correction step
// against the DECODER's arithmetic, and a final containment check with that same arithmetic: a record that is not
// conservative is counted in *violations and the tree keeps the exact Node64 walk (non-finite boxes end up there).
__global__ void __launch_bounds__(256) wideConvertKernel(int n, Node64 const *__restrict__ nodes, Wide64 *__restrict__ wide,
                                                       unsigned *__restrict__ violations)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;
  float const inf = __int_as_float(0x7f800000);
  float lo[4][3], hi[4][3];
  int ref[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    ref[k] = kWideEmpty;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      lo[k][d] = inf;
      hi[k][d] = -inf;
    }
  }
#define ABX_WIDE_SET(K, BL, BH, CHILD_REF, FIRST, LAST)                                                               \
  do                                                                                                                   \
  {                                                                                                                    \
    int const leaves_ = (LAST) - (FIRST) + 1;                                                                          \
    lo[K][0] = (BL).x, lo[K][1] = (BL).y, lo[K][2] = (BL).z;                                                           \
    hi[K][0] = (BH).x, hi[K][1] = (BH).y, hi[K][2] = (BH).z;                                                           \
    ref[K] = leaves_ <= kWideRun ? ~(((FIRST) << 2) | (leaves_ - 1)) : (CHILD_REF);                                    \
  } while (0)
#define ABX_WIDE_SIDE(K0, K1, BL, BH, CHILD_REF, FIRST, LAST)                                                          \
  do                                                                                                                   \
  {                                                                                                                    \
    if ((LAST) - (FIRST) + 1 <= kWideRun)                                                                              \
      ABX_WIDE_SET(K0, BL, BH, CHILD_REF, FIRST, LAST);                                                                \
    else                                                                                                               \
    {                                                                                                                  \
      float4 const *g_ = reinterpret_cast<float4 const *>(nodes + (CHILD_REF));                                        \
      float4 const c0_ = __ldg(g_), c1_ = __ldg(g_ + 1), c2_ = __ldg(g_ + 2), c3_ = __ldg(g_ + 3);                     \
      int const clref_ = __float_as_int(c0_.w), crref_ = __float_as_int(c1_.w);                                        \
      int const crl_ = __float_as_int(c2_.w), crr_ = __float_as_int(c3_.w);                                            \
      int const cl_hi_ = refIsLeaf(clref_) ? crl_ : clref_;                                                            \
      int const cr_lo_ = refIsLeaf(crref_) ? crr_ : crref_;                                                            \
      ABX_WIDE_SET(K0, c0_, c1_, clref_, crl_, cl_hi_);                                                                \
      ABX_WIDE_SET(K1, c2_, c3_, crref_, cr_lo_, crr_);                                                                \
    }                                                                                                                  \
  } while (0)
  float4 const *f = reinterpret_cast<float4 const *>(nodes + i);
  float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
  int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
  int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
  int const l_hi = refIsLeaf(lref) ? rl : lref;
  int const r_lo = refIsLeaf(rref) ? rr : rref;
  ABX_WIDE_SIDE(0, 1, a0, a1, lref, rl, l_hi);
  ABX_WIDE_SIDE(2, 3, a2, a3, rref, r_lo, rr);
#undef ABX_WIDE_SIDE
#undef ABX_WIDE_SET

  float origin[3], scale[3], inv_dn[3], inv_up[3];
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    float const mn = fminf(fminf(lo[0][d], lo[1][d]), fminf(lo[2][d], lo[3][d]));
    float const mx = fmaxf(fmaxf(hi[0][d], hi[1][d]), fmaxf(hi[2][d], hi[3][d]));
    origin[d] = mn;
    scale[d] = __fdiv_ru(__fsub_ru(mx, mn), 255.0f);
    bool const pos = scale[d] > 0.f;
    inv_dn[d] = pos ? __fdiv_rd(1.0f, scale[d]) : 0.f;
    inv_up[d] = pos ? __fdiv_ru(1.0f, scale[d]) : 0.f;
  }
  bool bad = !(isfinite(scale[0]) && isfinite(scale[1]) && isfinite(scale[2]) && isfinite(origin[0]) &&
               isfinite(origin[1]) && isfinite(origin[2]));
  unsigned q[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k)
  {
    if (ref[k] == kWideEmpty)
      continue;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      // estimates from below / above, then one step against the decoder's own arithmetic
      int ql = min(255, max(0, (int)__fmul_rd(__fsub_rd(lo[k][d], origin[d]), inv_dn[d])));
      int qh = min(255, max(0, (int)ceilf(__fmul_ru(__fsub_ru(hi[k][d], origin[d]), inv_up[d]))));
      if (ql > 0 && wideLo((float)ql, scale[d], origin[d]) > lo[k][d])
        --ql;
      if (qh < 255 && wideHi((float)qh, scale[d], origin[d]) < hi[k][d])
        ++qh;
      bad |= wideLo((float)ql, scale[d], origin[d]) > lo[k][d] || wideHi((float)qh, scale[d], origin[d]) < hi[k][d];
      constexpr int kB[4] = {0, 6, 12, 18};
      int const bl = kB[k] + d, bh = kB[k] + 3 + d;
      q[bl >> 2] |= (unsigned)ql << (8 * (bl & 3));
      q[bh >> 2] |= (unsigned)qh << (8 * (bh & 3));
    }
  }
  if (bad)
    atomicAdd(violations, 1u);
  uint4 *o = wide[i].w;
  o[0] = make_uint4(__float_as_uint(origin[0]), __float_as_uint(origin[1]), __float_as_uint(origin[2]),
                    __float_as_uint(scale[0]));
  o[1] = make_uint4(__float_as_uint(scale[1]), __float_as_uint(scale[2]), q[0], q[1]);
  o[2] = make_uint4(q[2], q[3], q[4], q[5]);
  o[3] = make_uint4((unsigned)ref[0], (unsigned)ref[1], (unsigned)ref[2], (unsigned)ref[3]);
}

// Written by the first spatial query of a tree that was built with want_wide (kNN-only users never pay for the
// records), on that query's stream; nothing is read back -- a tree whose boxes cannot be quantised is flagged on the
// device and the kernels fall back by themselves.  Queries on other streams wait for the conversion through an event.
abx_status ensureWide(cudaStream_t s, abx_bvh *t)
{
  if (!t->want_wide)
    return ABX_OK;
  static std::mutex mtx;
  std::lock_guard<std::mutex> lock(mtx);
  if (!t->wide)
  {
    int const n = (int)t->n;
    ABX_TRY(deviceAlloc((void **)&t->wide_bad, sizeof(unsigned), s));
    ABX_CUDA_TRY(cudaMemsetAsync(t->wide_bad, 0, sizeof(unsigned), s));
    Wide64 *w = nullptr;
    ABX_TRY(deviceAlloc((void **)&w, sizeof(Wide64) * (size_t)(n - 1), s));
    ABX_LAUNCH(wideConvertKernel, divUp(n - 1, 256), 256, 0, s, n, t->nodes, w, t->wide_bad);
    ABX_CUDA_TRY(cudaEventCreateWithFlags(&t->wide_ready, cudaEventDisableTiming));
    ABX_CUDA_TRY(cudaEventRecord(t->wide_ready, s));
    t->wide_stream = s;
    t->bytes += sizeof(Wide64) * (size_t)(n - 1);
    t->wide = w; // published last
  }
  else if (s != t->wide_stream)
    ABX_CUDA_TRY(cudaStreamWaitEvent(s, t->wide_ready, 0));
  return ABX_OK;
}

// fills a freshly allocated tree; on failure the caller destroys it (every early return below is safe)
static abx_status buildTreeInto(cudaStream_t s, abx_bvh *t, void const *prims, uint64_t const *sorted_codes,
                                bool want_wide)
{
  int const kind = t->kind;
  int64_t const n = t->n;
  ABX_TRY(deviceAlloc((void **)&t->bounds_dev, 6 * sizeof(float), s));
  t->bytes += 6 * sizeof(float);
  if (n == 0)
  {
    // LinearBVH.hpp:203-206: bounds() stays the default (empty) box
    ABX_LAUNCH(emptyBoundsKernel, 1, 32, 0, s, t->bounds_dev);
    return ABX_OK;
  }
  size_t const leaf_f4 = (kind == ABX_PRIM_POINT3F) ? 1 : 2;
  ABX_TRY(deviceAlloc((void **)&t->leaf_box, sizeof(float4) * leaf_f4 * n, s));
  t->bytes += sizeof(float4) * leaf_f4 * n;
  if (kind == ABX_PRIM_TRI3F)
  {
    ABX_TRY(deviceAlloc((void **)&t->leaf_tri, sizeof(float4) * 3 * n, s));
    t->bytes += sizeof(float4) * 3 * n;
  }
  ABX_TRY(deviceAlloc((void **)&t->perm, sizeof(uint32_t) * n, s));
  ABX_TRY(deviceAlloc((void **)&t->codes, sizeof(uint64_t) * n, s));
  t->bytes += 12 * n;
  if (n == 1)
  {
    ABX_DISPATCH_PRIM(kind, ABX_LAUNCH((singleLeafKernel<K>), 1, 1, 0, s, (float const *)prims, t->leaf_box,
                                       t->leaf_tri, t->perm, (unsigned long long *)t->codes, t->bounds_dev));
    return ABX_OK;
  }
  ABX_TRY(deviceAlloc((void **)&t->nodes, sizeof(Node64) * (n - 1), s));
  t->bytes += sizeof(Node64) * (n - 1);

  if (sorted_codes)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(t->codes, sorted_codes, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s));
    ABX_LAUNCH(iotaKernel, divUp(n, kThreads), kThreads, 0, s, t->perm, n);
  }
  else
  {
    TempBuffer<unsigned> enc;
    ABX_TRY(enc.alloc(6, s));
    ABX_TRY(sceneBounds(s, kind, prims, n, enc.ptr));
    ABX_TRY(decodeBounds(s, enc.ptr, t->bounds_dev));
    ABX_TRY(morton64(s, kind, prims, n, t->bounds_dev, t->codes));
    // Morton64 codes use 63 bits (3 x 21).  Double-buffered: the tree keeps whichever pair of
    // buffers the sort finished in
    TempBuffer<uint64_t> codes_alt;
    TempBuffer<uint32_t> perm_alt;
    ABX_TRY(codes_alt.alloc(n, s));
    ABX_TRY(perm_alt.alloc(n, s));
    uint64_t *kb[2] = {t->codes, codes_alt.ptr};
    uint32_t *vb[2] = {t->perm, perm_alt.ptr};
    int cur = 0;
    ABX_TRY(sortPairsU64DB(s, kb, vb, &cur, n, /*iota_vals=*/true, 63));
    if (cur != 0)
    {
      std::swap(t->codes, codes_alt.ptr);
      std::swap(t->perm, perm_alt.ptr);
    }
  }
  ABX_TRY(buildHierarchy(s, t, prims));
  // trees of more than 64 leaves may get 4-wide records on their first spatial query (n < 2^29: run encoding)
  t->want_wide = want_wide && n > 64 && n < ((int64_t)1 << 29);
  return ABX_OK;
}

abx_status buildTree(cudaStream_t s, int kind, void const *prims, int64_t n, uint64_t const *sorted_codes,
                     abx_bvh **out, bool want_wide)
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
  abx_status const st = buildTreeInto(s, t, prims, sorted_codes, want_wide);
  if (st != ABX_OK)
  {
    destroyTree(t);
    return st;
  }
  *out = t;
  return ABX_OK;
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
