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
    if (pred.box(lo, hi) && (!TRI || triangleLeafTest<PRED>(pred, leaf_tri, 0)))
      invoke(cb, qi, 0u);
    return;
  }
  traverseSpatial<LEAF_F4>(nodes, leaf_box, pred, [&](unsigned orig, int pos) {
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
    if (pred.box(lo, hi) && (!TRI || triangleLeafTest<PRED>(pred, leaf_tri, 0)))
      cb(qi, 0u, out);
  }
  else
    traverseSpatial<LEAF_F4>(nodes, leaf_box, pred, [&](unsigned orig, int pos) {
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

} // namespace cb
} // namespace abx
