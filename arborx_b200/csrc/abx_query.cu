// abx_query.cu -- tree traversals: spatial (count / fill), nearest (kNN), half.
//
// Behavioural contract: spatial/detail/ArborX_TreeTraversal.hpp:34-120 (spatial),
// :122-336 (nearest), spatial/detail/ArborX_HalfTraversal.hpp:24-75 (half).  The
// reference walks {left_child, rope} nodes one box test per step; here one
// 64-byte Node64 load tests both children, leaf boxes live in the parent record
// (a point leaf is never dereferenced), and a short per-thread stack replaces the
// ropes.  Result sets are identical; order inside a row is traversal order, which
// the reference does not specify either.
#include "abx_traverse.cuh"

namespace abx
{

namespace
{

enum
{
  MODE_COUNT = 0,
  MODE_FILL = 1
};

template <int PRED, int MODE, bool TRI>
__global__ void __launch_bounds__(kThreads)
    spatialKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                  float4 const *__restrict__ leaf_tri, int n, float const *__restrict__ preds, int64_t q,
                  unsigned const *__restrict__ qperm, int limit, int32_t *__restrict__ counts,
                  int32_t const *__restrict__ offsets, uint32_t *__restrict__ indices)
{
  int64_t const t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= q)
    return;
  int64_t const qi = qperm ? (int64_t)qperm[t] : t;
  Pred<PRED> pred;
  pred.load(preds, qi);
  int count = 0;
  int64_t const base = (MODE == MODE_FILL) ? (int64_t)offsets[qi] : 0;
  if (n == 1)
    return; // handled by singleLeafKernel
  traverseSpatial(nodes, pred, [&](int ref, int pos) {
    if (TRI && !triangleLeafTest<PRED>(pred, leaf_tri, pos))
      return false;
    if (MODE == MODE_FILL)
      indices[base + count] = refOrig(ref);
    ++count;
    return limit > 0 && count >= limit;
  });
  if (MODE == MODE_COUNT)
    counts[qi] = count;
}

// n == 1: test the predicate against the single leaf (TreeTraversal.hpp:80-90)
template <int PRED, int MODE>
__global__ void __launch_bounds__(kThreads)
    spatialSingleLeafKernel(float4 const *__restrict__ leaf_box, float4 const *__restrict__ leaf_tri, int prim_kind,
                            float const *__restrict__ preds, int64_t q, int32_t *__restrict__ counts,
                            int32_t const *__restrict__ offsets, uint32_t *__restrict__ indices)
{
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (qi >= q)
    return;
  Pred<PRED> pred;
  pred.load(preds, qi);
  float4 lo = __ldg(leaf_box);
  float4 hi = prim_kind == ABX_PRIM_POINT3F ? lo : __ldg(leaf_box + 1);
  bool hit = pred.box(lo, hi);
  if (hit && prim_kind == ABX_PRIM_TRI3F)
    hit = triangleLeafTest<PRED>(pred, leaf_tri, 0);
  if (MODE == MODE_COUNT)
    counts[qi] = hit ? 1 : 0;
  else if (hit)
    indices[offsets[qi]] = 0u;
}

// ---- nearest ---------------------------------------------------------------------
// distance(Point, Box) as a float, sqrt included (Distance.hpp:72-80): kNN reports
// distances, so the root is taken (correctly rounded) rather than skipped.
__device__ __forceinline__ float pointBoxDist(float px, float py, float pz, float4 lo, float4 hi)
{
  return __fsqrt_rn(pointBoxDist2(px, py, pz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z));
}

// Bounded candidate list kept sorted ascending in registers.  Acceptance rule is
// the reference's: a leaf enters iff distance < radius, radius = k-th distance
// once k candidates are known (TreeTraversal.hpp:255-290).
template <int KCAP>
struct RegList
{
  float d[KCAP];
  unsigned id[KCAP];
  __device__ __forceinline__ void init()
  {
#pragma unroll
    for (int i = 0; i < KCAP; ++i)
    {
      d[i] = __int_as_float(0x7f800000); // +inf
      id[i] = 0xffffffffu;
    }
  }
  // insert (dist, idx) knowing dist < d[k-1]; entries beyond k-1 stay +inf
  __device__ __forceinline__ void insert(float dist, unsigned idx, int k)
  {
    // place at slot k-1 then bubble toward the front; strict < keeps earlier
    // arrivals ahead of later ones at equal distance
#pragma unroll
    for (int i = KCAP - 1; i >= 0; --i)
    {
      if (i == k - 1)
      {
        d[i] = dist;
        id[i] = idx;
      }
    }
#pragma unroll
    for (int i = KCAP - 1; i >= 1; --i)
    {
      if (i <= k - 1 && d[i] < d[i - 1])
      {
        float td = d[i];
        d[i] = d[i - 1];
        d[i - 1] = td;
        unsigned ti = id[i];
        id[i] = id[i - 1];
        id[i - 1] = ti;
      }
    }
  }
  __device__ __forceinline__ float radius(int k) const
  {
    float r = d[0];
#pragma unroll
    for (int i = 1; i < KCAP; ++i)
      if (i == k - 1)
        r = d[i];
    return r;
  }
};

// max-heap in global scratch for large k (reference: NearestBufferProvider.hpp:24-72,
// misc/ArborX_Heap.hpp).  Entries are (distance, index) pairs.
struct GlobalHeap
{
  float2 *h; // x = distance, y = bits(index)
  int size;
  __device__ __forceinline__ void push(float dist, unsigned idx)
  {
    int pos = size++;
    while (pos > 0)
    {
      int parent = (pos - 1) / 2;
      float2 pv = h[parent];
      if (!(pv.x < dist))
        break;
      h[pos] = pv;
      pos = parent;
    }
    h[pos] = make_float2(dist, __uint_as_float(idx));
  }
  __device__ __forceinline__ void replaceTop(float dist, unsigned idx)
  {
    int pos = 0;
    int const len = size;
    while (true)
    {
      int child = 2 * pos + 1;
      if (child >= len)
        break;
      float2 cv = h[child];
      if (child + 1 < len)
      {
        float2 c2 = h[child + 1];
        if (cv.x < c2.x)
        {
          cv = c2;
          ++child;
        }
      }
      if (!(dist < cv.x))
        break;
      h[pos] = cv;
      pos = child;
    }
    h[pos] = make_float2(dist, __uint_as_float(idx));
  }
  __device__ __forceinline__ float top() const { return h[0].x; }
  // in-place heap sort -> ascending by distance
  __device__ __forceinline__ void sortAscending()
  {
    int const total = size;
    while (size > 1)
    {
      float2 last = h[size - 1];
      float2 top = h[0];
      --size;
      replaceTop(last.x, __float_as_uint(last.y));
      h[size] = top;
    }
    size = total;
  }
};

template <int KCAP, bool TRI>
__global__ void __launch_bounds__(kThreads)
    nearestKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                  float4 const *__restrict__ leaf_tri, int n, int prim_kind, float const *__restrict__ pts, int64_t q,
                  unsigned const *__restrict__ qperm, int k_uniform, int row_stride,
                  int32_t const *__restrict__ k_per_query, int32_t const *__restrict__ offsets,
                  int32_t *__restrict__ counts, uint32_t *__restrict__ indices, float *__restrict__ distances,
                  float2 *__restrict__ scratch)
{
  int64_t const t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= q)
    return;
  int64_t const qi = qperm ? (int64_t)qperm[t] : t;
  int k = k_per_query ? k_per_query[qi] : k_uniform;
  // rows are compact: row_stride = min(k, n) for uniform k, CRS offsets otherwise
  int64_t const base = offsets ? (int64_t)offsets[qi] : qi * (int64_t)row_stride;
  if (k < 1)
  {
    if (counts)
      counts[qi] = 0;
    return;
  }
  float const px = pts[3 * qi], py = pts[3 * qi + 1], pz = pts[3 * qi + 2];

  if (n == 1)
  {
    // TreeTraversal.hpp:168-178: the single value is reported unconditionally
    float4 lo = __ldg(leaf_box);
    float4 hi = prim_kind == ABX_PRIM_POINT3F ? lo : __ldg(leaf_box + 1);
    float dist = TRI ? __fsqrt_rn(pointTriangleDist2(px, py, pz, __ldg(leaf_tri), __ldg(leaf_tri + 1), __ldg(leaf_tri + 2)))
                     : pointBoxDist(px, py, pz, lo, hi);
    indices[base] = 0u;
    if (distances)
      distances[base] = dist;
    if (counts)
      counts[qi] = 1;
    return;
  }

  constexpr bool USE_REGS = KCAP > 0;
  RegList<USE_REGS ? KCAP : 1> list;
  GlobalHeap heap;
  if (USE_REGS)
    list.init();
  else
  {
    heap.h = scratch + base;
    heap.size = 0;
  }
  float radius = __int_as_float(0x7f800000);
  int found = 0;

  auto offer = [&](float dist, int ref, int pos) {
    // leaf candidate with (box) distance < radius
    if (TRI)
    {
      dist = __fsqrt_rn(pointTriangleDist2(px, py, pz, __ldg(leaf_tri + 3 * (size_t)pos),
                                            __ldg(leaf_tri + 3 * (size_t)pos + 1), __ldg(leaf_tri + 3 * (size_t)pos + 2)));
      if (!(dist < radius))
        return;
    }
    unsigned const idx = refOrig(ref);
    if (USE_REGS)
    {
      list.insert(dist, idx, k);
      if (found < k)
        ++found;
      if (found == k)
        radius = list.radius(k);
    }
    else
    {
      if (heap.size < k)
        heap.push(dist, idx);
      else
        heap.replaceTop(dist, idx);
      found = heap.size;
      if (found == k)
        radius = heap.top();
    }
  };

  int stack[kStackSize];
  float stack_d[kStackSize];
  int sp = 0;
  int node = 0;
  while (true)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
    float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
    int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
    float const dl = pointBoxDist(px, py, pz, a0, a1);
    float const dr = pointBoxDist(px, py, pz, a2, a3);
    bool go_l = false, go_r = false;
    if (dl < radius)
    {
      if (refIsLeaf(lref))
        offer(dl, lref, __float_as_int(a2.w));
      else
        go_l = true;
    }
    if (dr < radius) // radius may already have shrunk (TreeTraversal.hpp:273-274)
    {
      if (refIsLeaf(rref))
        offer(dr, rref, __float_as_int(a3.w));
      else
        go_r = true;
    }
    if (go_l || go_r)
    {
      // nearer child first; left on ties (TreeTraversal.hpp:310-313)
      bool const left_first = go_l && (dl <= dr || !go_r);
      if (go_l && go_r)
      {
        stack[sp] = left_first ? rref : lref;
        stack_d[sp] = left_first ? dr : dl;
        ++sp;
      }
      node = left_first ? lref : rref;
      continue;
    }
    // pop until a node that can still contain a closer leaf
    bool popped = false;
    while (sp > 0)
    {
      --sp;
      if (stack_d[sp] < radius)
      {
        node = stack[sp];
        popped = true;
        break;
      }
    }
    if (!popped)
      break;
  }

  if (USE_REGS)
  {
#pragma unroll
    for (int i = 0; i < (USE_REGS ? KCAP : 1); ++i)
      if (i < found)
      {
        indices[base + i] = list.id[i];
        if (distances)
          distances[base + i] = list.d[i];
      }
  }
  else
  {
    heap.sortAscending();
    for (int i = 0; i < found; ++i)
    {
      float2 e = heap.h[i];
      indices[base + i] = __float_as_uint(e.y);
      if (distances)
        distances[base + i] = e.x;
    }
  }
  if (counts)
    counts[qi] = found;
}

__global__ void __launch_bounds__(kThreads)
    halfPairsKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n, float r,
                    uint32_t *__restrict__ pairs, unsigned long long capacity, unsigned long long *count)
{
  int const i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n)
    return;
  float4 const p = __ldg(leaf_box + i);
  Pred<ABX_PRED_SPHERE3F> pred;
  pred.cx = p.x, pred.cy = p.y, pred.cz = p.z, pred.r = r;
  pred.t = sqrtThreshold(r);
  unsigned const me = __float_as_uint(p.w);
  traverseHalf(nodes, i, pred, [&](int ref, int) {
    unsigned long long slot = atomicAdd(count, 1ull);
    if (pairs && slot < capacity)
    {
      pairs[2 * slot] = me;
      pairs[2 * slot + 1] = refOrig(ref);
    }
  });
}

// k per query clipped to the tree size (rows are shorter than k when n < k)
__global__ void clipKKernel(int32_t const *__restrict__ k_per_query, int k_uniform, int n, int64_t q,
                            int32_t *__restrict__ out)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q)
  {
    int k = k_per_query ? k_per_query[i] : k_uniform;
    out[i] = max(0, min(k, n));
  }
}

#define ABX_DISPATCH_PRED(kind, CALL)                                                                                 \
  switch (kind)                                                                                                        \
  {                                                                                                                    \
  case ABX_PRED_SPHERE3F:                                                                                              \
  {                                                                                                                    \
    constexpr int P = ABX_PRED_SPHERE3F;                                                                               \
    CALL;                                                                                                              \
    break;                                                                                                             \
  }                                                                                                                    \
  case ABX_PRED_BOX3F:                                                                                                 \
  {                                                                                                                    \
    constexpr int P = ABX_PRED_BOX3F;                                                                                  \
    CALL;                                                                                                              \
    break;                                                                                                             \
  }                                                                                                                    \
  case ABX_PRED_POINT3F:                                                                                               \
  {                                                                                                                    \
    constexpr int P = ABX_PRED_POINT3F;                                                                                \
    CALL;                                                                                                              \
    break;                                                                                                             \
  }                                                                                                                    \
  default:                                                                                                             \
    setError("unknown predicate kind");                                                                                \
    return ABX_ERR_ARG;                                                                                                \
  }

template <int MODE>
abx_status spatialLaunch(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q,
                         uint32_t const *qperm, int32_t limit, int32_t *counts, int32_t const *offsets,
                         uint32_t *indices)
{
  if (q <= 0)
    return ABX_OK;
  int const grid = divUp(q, kThreads);
  int const n = (int)t->n;
  char const *tag = MODE == MODE_COUNT ? "spatialKernel<count>" : "spatialKernel<fill>";
  if (t->kind == ABX_PRIM_TRI3F && pred_kind != ABX_PRED_SPHERE3F)
  {
    setError("only intersects(Sphere) is defined for triangle primitives");
    return ABX_ERR_ARG;
  }
  if (n == 0)
  {
    if (MODE == MODE_COUNT)
      ABX_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * q, s));
    return ABX_OK;
  }
  if (n == 1)
  {
    ABX_DISPATCH_PRED(pred_kind, ABX_LAUNCH((spatialSingleLeafKernel<P, MODE>), grid, kThreads, 0, s, t->leaf_box,
                                            t->leaf_tri, t->kind, (float const *)preds, q, counts, offsets, indices));
    return ABX_OK;
  }
  if (t->kind == ABX_PRIM_TRI3F)
  {
    ABX_DISPATCH_PRED(pred_kind, ABX_LAUNCH_TAGGED(tag, (spatialKernel<P, MODE, true>), grid, kThreads, 0, s, t->nodes,
                                            t->leaf_box, t->leaf_tri, n, (float const *)preds, q, qperm, limit, counts,
                                            offsets, indices));
  }
  else
  {
    ABX_DISPATCH_PRED(pred_kind, ABX_LAUNCH_TAGGED(tag, (spatialKernel<P, MODE, false>), grid, kThreads, 0, s, t->nodes,
                                            t->leaf_box, t->leaf_tri, n, (float const *)preds, q, qperm, limit, counts,
                                            offsets, indices));
  }
  return ABX_OK;
}

} // namespace

// Morton32 permutation of the predicates (CrsGraphWrapperImpl.hpp:407-419)
abx_status predicatePermutation(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q,
                                TempBuffer<uint32_t> &perm)
{
  TempBuffer<uint32_t> codes;
  ABX_TRY(codes.alloc(q, s));
  ABX_TRY(perm.alloc(q, s));
  ABX_TRY(morton32(s, pred_kind, preds, q, t->bounds_dev, codes.ptr));
  ABX_TRY(sortPairsU32(s, codes.ptr, perm.ptr, q, true));
  return ABX_OK;
}

abx_status spatialCount(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q, uint32_t const *qperm,
                        int32_t limit, int32_t *counts)
{
  return spatialLaunch<MODE_COUNT>(s, t, pred_kind, preds, q, qperm, limit, counts, nullptr, nullptr);
}

abx_status spatialFill(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q, uint32_t const *qperm,
                       int32_t const *offsets, uint32_t *indices)
{
  return spatialLaunch<MODE_FILL>(s, t, pred_kind, preds, q, qperm, 0, nullptr, offsets, indices);
}

// uniform k: offsets == nullptr and rows start at i * min(k, n); per-query k:
// offsets = CRS offsets of min(k_i, n).  total_rows = size of indices.
abx_status nearestQuery(cudaStream_t s, abx_bvh *t, float const *pts, int64_t q, int32_t k, int32_t const *k_per_query,
                        uint32_t const *qperm, int32_t const *offsets, int64_t total_rows, int32_t *counts,
                        uint32_t *indices, float *distances)
{
  if (q <= 0)
    return ABX_OK;
  int const n = (int)t->n;
  if (n == 0)
  {
    if (counts)
      ABX_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * q, s));
    return ABX_OK;
  }
  int const grid = divUp(q, kThreads);
  bool const tri = t->kind == ABX_PRIM_TRI3F;
  int const kmax = k_per_query ? INT_MAX : k; // per-query k: general path
  int const row_stride = std::max(0, std::min(k, n));
#define ABX_NEAREST(KCAP, SCRATCH)                                                                                    \
  do                                                                                                                   \
  {                                                                                                                    \
    if (tri)                                                                                                           \
      ABX_LAUNCH_TAGGED("nearestKernel<" #KCAP ",tri>", (nearestKernel<KCAP, true>), grid, kThreads, 0, s, t->nodes, t->leaf_box, t->leaf_tri, n, t->kind,    \
                 pts, q, qperm, k, row_stride, k_per_query, offsets, counts, indices, distances, SCRATCH);                         \
    else                                                                                                               \
      ABX_LAUNCH_TAGGED("nearestKernel<" #KCAP ">", (nearestKernel<KCAP, false>), grid, kThreads, 0, s, t->nodes, t->leaf_box, t->leaf_tri, n, t->kind,   \
                 pts, q, qperm, k, row_stride, k_per_query, offsets, counts, indices, distances, SCRATCH);                         \
  } while (0)
  if (kmax <= 1)
    ABX_NEAREST(1, nullptr);
  else if (kmax <= 4)
    ABX_NEAREST(4, nullptr);
  else if (kmax <= 8)
    ABX_NEAREST(8, nullptr);
  else if (kmax <= 12)
    ABX_NEAREST(12, nullptr);
  else if (kmax <= 16)
    ABX_NEAREST(16, nullptr);
  else if (kmax <= 32)
    ABX_NEAREST(32, nullptr);
  else
  {
    // heap in global scratch, one slot range per query laid out like the output rows
    TempBuffer<float2> scratch;
    ABX_TRY(scratch.alloc((size_t)std::max<int64_t>(total_rows, 1), s));
    ABX_NEAREST(0, scratch.ptr);
  }
#undef ABX_NEAREST
  return ABX_OK;
}

abx_status halfTraversalPairs(cudaStream_t s, abx_bvh *t, float r, uint32_t *pairs, int64_t capacity,
                              unsigned long long *count_dev)
{
  ABX_CUDA_TRY(cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), s));
  if (t->n < 2)
    return ABX_OK;
  if (t->kind != ABX_PRIM_POINT3F)
  {
    setError("half traversal is defined over point primitives");
    return ABX_ERR_ARG;
  }
  ABX_LAUNCH(halfPairsKernel, divUp(t->n, kThreads), kThreads, 0, s, t->nodes, t->leaf_box, (int)t->n, r, pairs,
             (unsigned long long)capacity, count_dev);
  return ABX_OK;
}

abx_status clipK(cudaStream_t s, int32_t const *k_per_query, int k, int n, int64_t q, int32_t *out)
{
  if (q <= 0)
    return ABX_OK;
  ABX_LAUNCH(clipKKernel, divUp(q, 256), 256, 0, s, k_per_query, k, n, q, out);
  return ABX_OK;
}

} // namespace abx
