/* ArborX_B200_Callbacks.cuh -- user callbacks on the device (CUDA, sm_100a).
 *
 * The reference's query(space, predicates, callback) instantiates the user's functor inside its traversal
 * templates (spatial/ArborX_LinearBVH.hpp:84-88,112-122, spatial/detail/ArborX_Callbacks.hpp:79-150,
 * spatial/detail/ArborX_TreeTraversal.hpp:97-119,180-335).  libabx.so is a C ABI, so the generic form lives in
 * this header instead: it is compiled in the USER's translation unit (nvcc, -gencode arch=compute_100a,code=sm_100a
 * --extended-lambda) and runs the same traversal cores as the library kernels (arborx_b200/csrc/abx_traverse.cuh)
 * over the device view exported by abx_bvh_device_view().  The library must be linked for everything else (build,
 * CRS queries, the nearest search behind nearest callbacks).
 *
 *   struct Count { int *n; __device__ void operator()(int64_t query, unsigned value) const { atomicAdd(n + query, 1); } };
 *   abx::cb::query(bvh, stream, abx::cb::intersects_spheres(spheres_dev, q), Count{counts_dev});
 *
 * Callback forms (Callbacks.hpp:79-150):
 *   void f(int64_t query, unsigned value)                        every match
 *   abx::cb::Control f(int64_t query, unsigned value)            return Control::early_exit to end that query
 *                                                                (CallbackTreeTraversalControl, :24-28)
 *   nearest: void f(int64_t query, unsigned value, float distance), called for the k nearest in ascending order
 *   output form (Callbacks.hpp:86-110, CrsGraphWrapperImpl.hpp:86-110): f(int64_t query, unsigned value, Out &out)
 *       with out(x) emitting zero or more results of any trivially copyable type per match; query_crs() returns
 *       them as CRS rows in the original predicate order (count pass, scan, fill pass, like the reference's 2P path)
 *   attach(predicates, data): the callback gets data[query] instead of the position (Predicates.hpp:221-238)
 *   query_per_thread(tree, predicate, callback): one query from inside the caller's kernel (LinearBVH.hpp:112-122)
 *   ordered_intersects_rays: leaves handed out nearest first along a ray, with early exit (TreeTraversal.hpp:338-489)
 * `query` is the position of the predicate in the batch (what the reference passes through attach()/getData()),
 * `value` the index of the primitive the tree was built on.  Predicates are visited one per thread in batch
 * order; pass a Morton-ordered batch (abx_morton32 + abx_sort_u32) for coherent warps. */
#pragma once
#include "../arborx_b200/csrc/abx_traverse.cuh"

#include <type_traits>

namespace abx
{
namespace cb
{

enum class Control
{
  normal_continuation,
  early_exit
};

struct SpatialPredicates
{
  int kind; // ABX_PRED_*
  float const *data;
  int64_t q;
};
inline SpatialPredicates intersects_spheres(float const *spheres4_dev, int64_t q)
{
  return {ABX_PRED_SPHERE3F, spheres4_dev, q};
}
inline SpatialPredicates intersects_boxes(float const *boxes6_dev, int64_t q) { return {ABX_PRED_BOX3F, boxes6_dev, q}; }
inline SpatialPredicates intersects_points(float const *points3_dev, int64_t q)
{
  return {ABX_PRED_POINT3F, points3_dev, q};
}
inline SpatialPredicates intersects_rays(float const *rays6_dev, int64_t q) { return {ABX_PRED_RAY3F, rays6_dev, q}; }

struct NearestPredicates
{
  float const *points;
  int64_t q;
  int k;
};
inline NearestPredicates nearest(float const *points3_dev, int64_t q, int k) { return {points3_dev, q, k}; }

namespace detail
{
template <class Callback>
__device__ __forceinline__ bool invoke(Callback const &cb, int64_t query, unsigned value)
{
  if constexpr (std::is_same_v<decltype(cb(query, value)), Control>)
    return cb(query, value) == Control::early_exit;
  else
  {
    cb(query, value);
    return false;
  }
}

template <int PRED, int LEAF_F4, bool TRI, class Callback>
__global__ void __launch_bounds__(kThreads)
    spatialCallbackKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                          float4 const *__restrict__ leaf_tri, int64_t n, float const *__restrict__ preds, int64_t q,
                          Callback cb)
{
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (qi >= q)
    return;
  Pred<PRED> pred;
  pred.load(preds, qi);
  if (n == 1)
  {
    // TreeTraversal.hpp:80-90: the predicate against the single leaf
    float4 const lo = __ldg(leaf_box);
    float4 const hi = LEAF_F4 == 1 ? lo : __ldg(leaf_box + 1);
    if (TRI ? triangleLeafTest<PRED>(pred, leaf_tri, 0) : pred.box(lo, hi))
      invoke(cb, qi, 0u);
    return;
  }
  traverseSpatial<LEAF_F4, (TRI ? 0 : kBucket)>(nodes, leaf_box, pred, [&](unsigned orig, int pos) {
    if (TRI && !triangleLeafTest<PRED>(pred, leaf_tri, pos))
      return false;
    return invoke(cb, qi, orig);
  });
}

template <class Callback>
__global__ void nearestCallbackKernel(int64_t q, int32_t const *__restrict__ offsets, uint32_t const *__restrict__ indices,
                                      float const *__restrict__ distances, Callback cb)
{
  int64_t const qi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= q)
    return;
  for (int j = offsets[qi]; j < offsets[qi + 1]; ++j)
    cb(qi, indices[j], distances[j]);
}

template <int PRED, class Callback>
inline abx_status launchSpatial(abx_device_view const &v, cudaStream_t s, SpatialPredicates const &p, Callback const &cb)
{
  int const grid = (int)((p.q + kThreads - 1) / kThreads);
  auto const *nodes = (Node64 const *)v.nodes;
  auto const *lb = (float4 const *)v.leaf_box;
  auto const *lt = (float4 const *)v.leaf_tri;
  if (v.prim_kind == ABX_PRIM_TRI3F)
    spatialCallbackKernel<PRED, 2, true><<<grid, kThreads, 0, s>>>(nodes, lb, lt, v.n, p.data, p.q, cb);
  else if (v.prim_kind == ABX_PRIM_BOX3F)
    spatialCallbackKernel<PRED, 2, false><<<grid, kThreads, 0, s>>>(nodes, lb, lt, v.n, p.data, p.q, cb);
  else
    spatialCallbackKernel<PRED, 1, false><<<grid, kThreads, 0, s>>>(nodes, lb, lt, v.n, p.data, p.q, cb);
  return cudaGetLastError() == cudaSuccess ? ABX_OK : ABX_ERR_CUDA;
}
} // namespace detail

/* bvh.query(space, intersects(...), callback): the callback runs on the device for every match. */
template <class Callback>
inline abx_status query(abx_bvh const *bvh, cudaStream_t stream, SpatialPredicates const &predicates,
                        Callback const &callback)
{
  abx_device_view v;
  abx_status st = abx_bvh_device_view(bvh, &v);
  if (st != ABX_OK)
    return st;
  if (v.n == 0 || predicates.q <= 0)
    return ABX_OK; // empty tree: no callback is ever called (TreeTraversal.hpp:48-51)
  switch (predicates.kind)
  {
  case ABX_PRED_SPHERE3F: return detail::launchSpatial<ABX_PRED_SPHERE3F>(v, stream, predicates, callback);
  case ABX_PRED_BOX3F: return detail::launchSpatial<ABX_PRED_BOX3F>(v, stream, predicates, callback);
  case ABX_PRED_POINT3F: return detail::launchSpatial<ABX_PRED_POINT3F>(v, stream, predicates, callback);
  case ABX_PRED_RAY3F: return detail::launchSpatial<ABX_PRED_RAY3F>(v, stream, predicates, callback);
  default: return ABX_ERR_ARG;
  }
}

/* bvh.query(space, nearest(points, k), callback): callback(query, value, distance) for the k nearest of every
 * point, nearest first (the library's kNN search produces the rows; TreeTraversal.hpp:320-334 calls the
 * callback on the sorted heap the same way). */
template <class Callback>
inline abx_status query(abx_bvh *bvh, cudaStream_t stream, NearestPredicates const &predicates, Callback const &callback)
{
  if (predicates.q <= 0)
    return ABX_OK;
  int32_t *offsets = nullptr;
  uint32_t *indices = nullptr;
  float *distances = nullptr;
  int64_t nnz = 0;
  abx_status st = abx_query_nearest_crs(bvh, stream, predicates.points, predicates.q, predicates.k, nullptr, nullptr,
                                        nullptr, nullptr, &offsets, &indices, &distances, &nnz);
  if (st != ABX_OK)
    return st;
  int const grid = (int)((predicates.q + 255) / 256);
  detail::nearestCallbackKernel<<<grid, 256, 0, stream>>>(predicates.q, offsets, indices, distances, callback);
  st = cudaGetLastError() == cudaSuccess ? ABX_OK : ABX_ERR_CUDA;
  abx_free(stream, offsets);
  abx_free(stream, indices);
  abx_free(stream, distances);
  return st;
}

namespace detail
{
// the `out` object handed to an output-form callback: counts in the first pass, stores in the second
template <class T>
struct OutputFunctor
{
  T *row;     // nullptr while counting
  int count;
  __device__ void operator()(T const &x)
  {
    if (row)
      row[count] = x;
    ++count;
  }
};

template <class T, class Callback>
struct OutputAdapter
{
  Callback cb;
  int32_t *counts;          // pass 1: results per query (written at the end of the query's traversal)
  int32_t const *offsets;   // pass 2
  T *values;
};

template <int PRED, int LEAF_F4, bool TRI, class T, class Callback>
__global__ void __launch_bounds__(kThreads)
    spatialOutputKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                        float4 const *__restrict__ leaf_tri, int64_t n, float const *__restrict__ preds, int64_t q,
                        Callback cb, int32_t *counts, int32_t const *__restrict__ offsets, T *values)
{
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (qi >= q)
    return;
  Pred<PRED> pred;
  pred.load(preds, qi);
  OutputFunctor<T> out{offsets ? values + offsets[qi] : nullptr, 0};
  if (n == 1)
  {
    float4 const lo = __ldg(leaf_box);
    float4 const hi = LEAF_F4 == 1 ? lo : __ldg(leaf_box + 1);
    if (TRI ? triangleLeafTest<PRED>(pred, leaf_tri, 0) : pred.box(lo, hi))
      cb(qi, 0u, out);
  }
  else
    traverseSpatial<LEAF_F4, (TRI ? 0 : kBucket)>(nodes, leaf_box, pred, [&](unsigned orig, int pos) {
      if (TRI && !triangleLeafTest<PRED>(pred, leaf_tri, pos))
        return false;
      cb(qi, orig, out);
      return false;
    });
  if (!offsets)
    counts[qi] = out.count;
}

template <int PRED, class T, class Callback>
inline abx_status launchOutput(abx_device_view const &v, cudaStream_t s, SpatialPredicates const &p, Callback const &cb,
                               int32_t *counts, int32_t const *offsets, T *values)
{
  int const grid = (int)((p.q + kThreads - 1) / kThreads);
  auto const *nodes = (Node64 const *)v.nodes;
  auto const *lb = (float4 const *)v.leaf_box;
  auto const *lt = (float4 const *)v.leaf_tri;
  if (v.prim_kind == ABX_PRIM_TRI3F)
    spatialOutputKernel<PRED, 2, true, T><<<grid, kThreads, 0, s>>>(nodes, lb, lt, v.n, p.data, p.q, cb, counts, offsets, values);
  else if (v.prim_kind == ABX_PRIM_BOX3F)
    spatialOutputKernel<PRED, 2, false, T><<<grid, kThreads, 0, s>>>(nodes, lb, lt, v.n, p.data, p.q, cb, counts, offsets, values);
  else
    spatialOutputKernel<PRED, 1, false, T><<<grid, kThreads, 0, s>>>(nodes, lb, lt, v.n, p.data, p.q, cb, counts, offsets, values);
  return cudaGetLastError() == cudaSuccess ? ABX_OK : ABX_ERR_CUDA;
}

template <class T, class Callback>
inline abx_status dispatchOutput(abx_device_view const &v, cudaStream_t s, SpatialPredicates const &p, Callback const &cb,
                                 int32_t *counts, int32_t const *offsets, T *values)
{
  switch (p.kind)
  {
  case ABX_PRED_SPHERE3F: return launchOutput<ABX_PRED_SPHERE3F, T>(v, s, p, cb, counts, offsets, values);
  case ABX_PRED_BOX3F: return launchOutput<ABX_PRED_BOX3F, T>(v, s, p, cb, counts, offsets, values);
  case ABX_PRED_POINT3F: return launchOutput<ABX_PRED_POINT3F, T>(v, s, p, cb, counts, offsets, values);
  case ABX_PRED_RAY3F: return launchOutput<ABX_PRED_RAY3F, T>(v, s, p, cb, counts, offsets, values);
  default: return ABX_ERR_ARG;
  }
}
} // namespace detail

/* bvh.query(space, predicates, callback, out, offsets) with an output-form callback
 * callback(query, value, out): *offsets_dev (q + 1 ints) and *values_dev (nnz elements of T) are allocated
 * with cudaMallocAsync on `stream` (release with cudaFreeAsync); *nnz = number of emitted values.  Two
 * traversals (count, fill), the reference's default two-pass CRS path; blocks once for nnz. */
template <class T, class Callback>
inline abx_status query_crs(abx_bvh const *bvh, cudaStream_t stream, SpatialPredicates const &predicates,
                            Callback const &callback, int32_t **offsets_dev, T **values_dev, int64_t *nnz)
{
  static_assert(std::is_trivially_copyable_v<T>, "output values are copied bytewise");
  abx_device_view v;
  abx_status st = abx_bvh_device_view(bvh, &v);
  if (st != ABX_OK)
    return st;
  int64_t const q = predicates.q < 0 ? 0 : predicates.q;
  *values_dev = nullptr;
  *nnz = 0;
  if (cudaMallocAsync((void **)offsets_dev, sizeof(int32_t) * (size_t)(q + 1), stream) != cudaSuccess)
    return ABX_ERR_CUDA;
  cudaMemsetAsync(*offsets_dev, 0, sizeof(int32_t) * (size_t)(q + 1), stream);
  if (v.n == 0 || q == 0)
    return ABX_OK;
  st = detail::dispatchOutput<T>(v, stream, predicates, callback, *offsets_dev, nullptr, (T *)nullptr);
  if (st != ABX_OK)
    return st;
  st = abx_exclusive_scan_i32(stream, *offsets_dev, *offsets_dev, q + 1);
  if (st != ABX_OK)
    return st;
  int32_t total = 0;
  cudaMemcpyAsync(&total, *offsets_dev + q, sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
  if (cudaStreamSynchronize(stream) != cudaSuccess)
    return ABX_ERR_CUDA;
  *nnz = total;
  if (total == 0)
    return ABX_OK;
  if (cudaMallocAsync((void **)values_dev, sizeof(T) * (size_t)total, stream) != cudaSuccess)
    return ABX_ERR_CUDA;
  return detail::dispatchOutput<T>(v, stream, predicates, callback, nullptr, *offsets_dev, *values_dev);
}

// ---- attach(predicates, data) (detail/ArborX_Predicates.hpp:221-238, getData) -----------------------------------
// The callback receives the datum attached to its predicate instead of the predicate's position in the batch:
//   abx::cb::query(bvh, stream, abx::cb::attach(abx::cb::intersects_spheres(s, q), data_dev), Callback{});
//   struct Callback { __device__ void operator()(MyData const &d, unsigned value) const; };   // or -> Control
template <class T>
struct AttachedPredicates
{
  SpatialPredicates predicates;
  T const *data;
};
template <class T>
inline AttachedPredicates<T> attach(SpatialPredicates const &predicates, T const *data_dev)
{
  return {predicates, data_dev};
}
namespace detail
{
template <class T, class Callback>
struct AttachAdapter
{
  T const *data;
  Callback cb;
  __device__ auto operator()(int64_t query, unsigned value) const { return cb(data[query], value); }
};
} // namespace detail
template <class T, class Callback>
inline abx_status query(abx_bvh const *bvh, cudaStream_t stream, AttachedPredicates<T> const &attached,
                        Callback const &callback)
{
  return query(bvh, stream, attached.predicates, detail::AttachAdapter<T, Callback>{attached.data, callback});
}

// ---- query(Experimental::PerThread{}, predicate, callback) (spatial/ArborX_LinearBVH.hpp:112-122) ----------------
// A single query issued from inside the caller's own kernel by the calling thread.  DeviceTree is a plain struct
// the caller passes to its kernel; the predicate is one of abx::Pred<ABX_PRED_*> (make_sphere / make_box /
// make_point / make_ray), the callback takes the value (index of the primitive) and may return Control.
struct DeviceTree
{
  Node64 const *nodes;
  float4 const *leaf_box;
  float4 const *leaf_tri;
  int64_t n;
  int prim_kind;
};
inline abx_status device_tree(abx_bvh const *bvh, DeviceTree *out)
{
  abx_device_view v;
  abx_status const st = abx_bvh_device_view(bvh, &v);
  if (st == ABX_OK)
    *out = DeviceTree{(Node64 const *)v.nodes, (float4 const *)v.leaf_box, (float4 const *)v.leaf_tri, v.n, v.prim_kind};
  return st;
}
__device__ inline Pred<ABX_PRED_SPHERE3F> make_sphere(float x, float y, float z, float r)
{
  Pred<ABX_PRED_SPHERE3F> p;
  p.cx = x, p.cy = y, p.cz = z, p.r = r;
  p.t = sqrtThreshold(r);
  return p;
}
__device__ inline Pred<ABX_PRED_BOX3F> make_box(float lx, float ly, float lz, float hx, float hy, float hz)
{
  Pred<ABX_PRED_BOX3F> p;
  p.lx = lx, p.ly = ly, p.lz = lz, p.hx = hx, p.hy = hy, p.hz = hz;
  return p;
}
__device__ inline Pred<ABX_PRED_POINT3F> make_point(float x, float y, float z)
{
  Pred<ABX_PRED_POINT3F> p;
  p.x = x, p.y = y, p.z = z;
  return p;
}
__device__ inline Pred<ABX_PRED_RAY3F> make_ray(float ox, float oy, float oz, float dx, float dy, float dz)
{
  float const g[6] = {ox, oy, oz, dx, dy, dz};
  Pred<ABX_PRED_RAY3F> p;
  p.load(g, 0); // normalises the direction in double like the Ray constructor
  return p;
}
namespace detail
{
template <class Callback>
__device__ __forceinline__ bool invokeValue(Callback const &cb, unsigned value)
{
  if constexpr (std::is_same_v<decltype(cb(value)), Control>)
    return cb(value) == Control::early_exit;
  else
  {
    cb(value);
    return false;
  }
}
} // namespace detail
template <int PRED, class Callback>
__device__ inline void query_per_thread(DeviceTree const &tree, Pred<PRED> const &pred, Callback const &callback)
{
  if (tree.n == 0)
    return;
  bool const tri = tree.prim_kind == ABX_PRIM_TRI3F;
  int const leaf_f4 = tree.prim_kind == ABX_PRIM_POINT3F ? 1 : 2;
  if (tree.n == 1)
  {
    float4 const lo = __ldg(tree.leaf_box);
    float4 const hi = leaf_f4 == 1 ? lo : __ldg(tree.leaf_box + 1);
    if (tri ? triangleLeafTest<PRED>(pred, tree.leaf_tri, 0) : pred.box(lo, hi))
      detail::invokeValue(callback, 0u);
    return;
  }
  auto emit = [&](unsigned orig, int pos) {
    if (tri && !triangleLeafTest<PRED>(pred, tree.leaf_tri, pos))
      return false;
    return detail::invokeValue(callback, orig);
  };
  if (leaf_f4 == 1)
    traverseSpatial<1>(tree.nodes, tree.leaf_box, pred, emit);
  else if (tri) // triangles: a leaf is tested whenever its parent is visited (the reference's leaves have no box)
    traverseSpatial<2, 0>(tree.nodes, tree.leaf_box, pred, emit);
  else
    traverseSpatial<2>(tree.nodes, tree.leaf_box, pred, emit);
}

// ---- Experimental::ordered_intersects(ray) (detail/ArborX_Predicates.hpp:103-127,148-156;
//      TreeTraversal<..., OrderedSpatialPredicateTag>, detail/ArborX_TreeTraversal.hpp:338-489) --------------------
// The leaves whose box the ray hits are handed to the callback nearest first (by where the ray enters the box,
// distance(Ray, Box), geometry/ArborX_Ray.hpp:433-444); the callback may end the query (first-hit ray casting).
//   callback(int64_t query, unsigned value, float distance)  -> void or Control
// Point and box primitives.  Like the reference, the traversal keeps a priority queue of 64 (node, distance)
// entries, and it shares the reference's corner case: when the queue is empty and the ray misses both children of
// the node at hand, the walk still ends on that node's last leaf and calls the callback for it (with an infinite
// distance here; the reference's callback sees the leaf without a distance).
struct OrderedRayPredicates
{
  float const *rays;
  int64_t q;
};
inline OrderedRayPredicates ordered_intersects_rays(float const *rays6_dev, int64_t q) { return {rays6_dev, q}; }

namespace detail
{
struct OrderedHeap // min-heap on the distance; item >= 0: internal node, item < 0: leaf at sorted position ~item
{
  static constexpr int kCapacity = 64;
  int item[kCapacity];
  float dist[kCapacity];
  int size = 0;
  __device__ void push(int it, float d)
  {
    int pos = size++;
    while (pos > 0)
    {
      int const parent = (pos - 1) / 2;
      if (!(d < dist[parent]))
        break;
      item[pos] = item[parent];
      dist[pos] = dist[parent];
      pos = parent;
    }
    item[pos] = it;
    dist[pos] = d;
  }
  __device__ void pop(int &it, float &d)
  {
    it = item[0];
    d = dist[0];
    int const last_item = item[size - 1];
    float const last_d = dist[size - 1];
    --size;
    int pos = 0;
    while (true)
    {
      int child = 2 * pos + 1;
      if (child >= size)
        break;
      if (child + 1 < size && dist[child + 1] < dist[child])
        ++child;
      if (!(dist[child] < last_d))
        break;
      item[pos] = item[child];
      dist[pos] = dist[child];
      pos = child;
    }
    if (size > 0)
    {
      item[pos] = last_item;
      dist[pos] = last_d;
    }
  }
};

template <class Callback>
__device__ __forceinline__ bool invokeOrdered(Callback const &cb, int64_t query, unsigned value, float distance)
{
  if constexpr (std::is_same_v<decltype(cb(query, value, distance)), Control>)
    return cb(query, value, distance) == Control::early_exit;
  else
  {
    cb(query, value, distance);
    return false;
  }
}

template <int LEAF_F4, class Callback>
__global__ void __launch_bounds__(kThreads)
    orderedRayKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int64_t n,
                     float const *__restrict__ rays, int64_t q, Callback cb)
{
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (qi >= q)
    return;
  Pred<ABX_PRED_RAY3F> ray;
  ray.load(rays, qi);
  float const inf = __int_as_float(0x7f800000);
  auto leafValue = [&](int pos) { return __float_as_uint(__ldg(leaf_box + (size_t)LEAF_F4 * pos).w); };
  if (n == 1)
  {
    float4 const lo = __ldg(leaf_box);
    float4 const hi = LEAF_F4 == 1 ? lo : __ldg(leaf_box + 1);
    float const d = ray.distance(lo, hi);
    if (d != inf)
      invokeOrdered(cb, qi, 0u, d);
    return;
  }
  OrderedHeap heap;
  int cur = 0; // the root
  float cur_d = 0.f;
  while (true)
  {
    if (cur < 0)
    {
      if (invokeOrdered(cb, qi, leafValue(~cur), cur_d))
        return;
      if (heap.size == 0)
        return;
      heap.pop(cur, cur_d);
      continue;
    }
    float4 const *f = reinterpret_cast<float4 const *>(nodes + cur);
    float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
    int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
    int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
    float const dl = ray.distance(a0, a1), dr = ray.distance(a2, a3);
    int const litem = refIsLeaf(lref) ? ~rl : lref, ritem = refIsLeaf(rref) ? ~rr : rref;
    bool const left_closer = dl < dr;
    int const c_item = left_closer ? litem : ritem, f_item = left_closer ? ritem : litem;
    float const c_d = left_closer ? dl : dr, f_d = left_closer ? dr : dl;
    if (heap.size > 0 && heap.dist[0] < c_d)
    {
      heap.pop(cur, cur_d);
      if (c_d < inf && heap.size < OrderedHeap::kCapacity)
        heap.push(c_item, c_d);
    }
    else if (c_d < inf)
    {
      cur = c_item;
      cur_d = c_d;
    }
    else
    {
      // queue empty and the ray misses both children: the reference walks down the right children and calls the
      // callback for the last leaf of this node (TreeTraversal.hpp:474-482 with both distances infinite)
      invokeOrdered(cb, qi, leafValue(rr), inf);
      return;
    }
    if (f_d < inf && heap.size < OrderedHeap::kCapacity)
      heap.push(f_item, f_d);
  }
}
} // namespace detail

template <class Callback>
inline abx_status query(abx_bvh const *bvh, cudaStream_t stream, OrderedRayPredicates const &predicates,
                        Callback const &callback)
{
  abx_device_view v;
  abx_status st = abx_bvh_device_view(bvh, &v);
  if (st != ABX_OK)
    return st;
  if (v.n == 0 || predicates.q <= 0)
    return ABX_OK;
  if (v.prim_kind == ABX_PRIM_TRI3F)
    return ABX_ERR_ARG; // distance(Ray, Triangle) is not defined in the reference either
  int const grid = (int)((predicates.q + kThreads - 1) / kThreads);
  if (v.prim_kind == ABX_PRIM_POINT3F)
    detail::orderedRayKernel<1><<<grid, kThreads, 0, stream>>>((Node64 const *)v.nodes, (float4 const *)v.leaf_box, v.n,
                                                             predicates.rays, predicates.q, callback);
  else
    detail::orderedRayKernel<2><<<grid, kThreads, 0, stream>>>((Node64 const *)v.nodes, (float4 const *)v.leaf_box, v.n,
                                                             predicates.rays, predicates.q, callback);
  return cudaGetLastError() == cudaSuccess ? ABX_OK : ABX_ERR_CUDA;
}

} // namespace cb
} // namespace abx
