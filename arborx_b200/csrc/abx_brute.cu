// abx_brute.cu -- ArborX::BruteForce: the same query interface as the BVH, answered by testing every predicate
// against every primitive.
//
// Behavioural contract: spatial/ArborX_BruteForce.hpp:42-160 (ctor, size / empty / bounds, query),
// spatial/detail/ArborX_BruteForceImpl.hpp:40-233 (tiles of predicates x tiles of indexables staged in team scratch
// memory for spatial predicates; one thread per predicate with a k-slot heap for nearest predicates).
// Shape here: a block owns 128 predicates (one per thread, in registers) and streams the primitives through shared
// memory in tiles of kTile; every thread tests its predicate against the whole tile with broadcast reads.  CRS output
// like the tree: count pass, scan, fill pass (rows in ascending primitive order).  Point and box primitives.
#include "abx_traverse.cuh"

#include <algorithm>

struct abx_brute
{
  int kind = 0;
  int64_t n = 0;
  cudaStream_t stream = nullptr;
  float4 *lo = nullptr; // (min corner, bits(index))
  float4 *hi = nullptr; // boxes only: (max corner, -)
  float *bounds_dev = nullptr;
  float bounds_host[6];
  bool bounds_host_valid = false;
};

namespace abx
{
namespace
{

constexpr int kTile = 512;

template <int KIND>
__global__ void bruteInitKernel(float const *__restrict__ prims, int64_t n, float4 *__restrict__ lo,
                                float4 *__restrict__ hi)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  if (KIND == ABX_PRIM_POINT3F)
    lo[i] = make_float4(prims[3 * i], prims[3 * i + 1], prims[3 * i + 2], __uint_as_float((unsigned)i));
  else
  {
    lo[i] = make_float4(prims[6 * i], prims[6 * i + 1], prims[6 * i + 2], __uint_as_float((unsigned)i));
    hi[i] = make_float4(prims[6 * i + 3], prims[6 * i + 4], prims[6 * i + 5], 0.f);
  }
}

// FILL = false: counts[i] = matches of predicate i; FILL = true: indices[offsets[i] ...] = the matching primitives
template <int PRED, bool BOXES, bool FILL>
__global__ void __launch_bounds__(kThreads)
    bruteSpatialKernel(float4 const *__restrict__ lo, float4 const *__restrict__ hi, int64_t n,
                       float const *__restrict__ preds, int64_t q, int32_t *__restrict__ counts,
                       int32_t const *__restrict__ offsets, uint32_t *__restrict__ indices)
{
  __shared__ float4 s_lo[kTile];
  __shared__ float4 s_hi[BOXES ? kTile : 1];
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  bool const active = qi < q;
  Pred<PRED> pred;
  pred.load(preds, active ? qi : 0);
  int count = 0;
  int64_t const base = (FILL && active) ? (int64_t)offsets[qi] : 0;
  for (int64_t t0 = 0; t0 < n; t0 += kTile)
  {
    int const m = (int)min((int64_t)kTile, n - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += kThreads)
    {
      s_lo[j] = lo[t0 + j];
      if (BOXES)
        s_hi[j] = hi[t0 + j];
    }
    __syncthreads();
    if (active)
      for (int j = 0; j < m; ++j)
      {
        float4 const l = s_lo[j];
        bool const hit = BOXES ? pred.box(l, s_hi[j]) : pred.point(l);
        if (hit)
        {
          if (FILL)
            indices[base + count] = __float_as_uint(l.w);
          ++count;
        }
      }
  }
  if (!FILL && active)
    counts[qi] = count;
}

// nearest(Point, k): one thread per predicate, max-heap of k (squared distance, index) slots in global scratch
// (BruteForceImpl.hpp:152-229: the first k primitives are pushed unconditionally, later ones replace the top when
// strictly closer); rows ascending by distance
template <bool BOXES>
__global__ void __launch_bounds__(kThreads)
    bruteNearestKernel(float4 const *__restrict__ lo, float4 const *__restrict__ hi, int64_t n,
                       float const *__restrict__ pts, int64_t q, int k, int row, float2 *__restrict__ scratch,
                       uint32_t *__restrict__ indices, float *__restrict__ distances)
{
  __shared__ float4 s_lo[kTile];
  __shared__ float4 s_hi[BOXES ? kTile : 1];
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  bool const active = qi < q;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (active)
    px = pts[3 * qi], py = pts[3 * qi + 1], pz = pts[3 * qi + 2];
  float2 *const heap = scratch + (active ? qi : 0) * (int64_t)row;
  int size = 0;
  auto siftDown = [&](int pos, float d, float id) {
    while (true)
    {
      int child = 2 * pos + 1;
      if (child >= size)
        break;
      float2 cv = heap[child];
      if (child + 1 < size)
      {
        float2 const c2 = heap[child + 1];
        if (cv.x < c2.x)
        {
          cv = c2;
          ++child;
        }
      }
      if (!(d < cv.x))
        break;
      heap[pos] = cv;
      pos = child;
    }
    heap[pos] = make_float2(d, id);
  };
  for (int64_t t0 = 0; t0 < n; t0 += kTile)
  {
    int const m = (int)min((int64_t)kTile, n - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += kThreads)
    {
      s_lo[j] = lo[t0 + j];
      if (BOXES)
        s_hi[j] = hi[t0 + j];
    }
    __syncthreads();
    if (active)
      for (int j = 0; j < m; ++j)
      {
        float4 const l = s_lo[j];
        float4 const h = BOXES ? s_hi[j] : l;
        float const d2 = pointBoxDist2(px, py, pz, l.x, l.y, l.z, h.x, h.y, h.z);
        if (size < row)
        {
          // push (sift up)
          int pos = size++;
          while (pos > 0)
          {
            int const parent = (pos - 1) / 2;
            float2 const pv = heap[parent];
            if (!(pv.x < d2))
              break;
            heap[pos] = pv;
            pos = parent;
          }
          heap[pos] = make_float2(d2, l.w);
        }
        else if (d2 < heap[0].x)
          siftDown(0, d2, l.w);
      }
  }
  if (!active)
    return;
  // heap sort -> ascending
  int const total = size;
  while (size > 1)
  {
    float2 const last = heap[size - 1], top = heap[0];
    --size;
    siftDown(0, last.x, last.y);
    heap[size] = top;
  }
  for (int i = 0; i < total; ++i)
  {
    float2 const e = heap[i];
    indices[qi * (int64_t)row + i] = __float_as_uint(e.y);
    if (distances)
      distances[qi * (int64_t)row + i] = __fsqrt_rn(e.x);
  }
}

__global__ void bruteStrideOffsetsKernel(int32_t *offsets, int64_t q_plus_1, int stride)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q_plus_1)
    offsets[i] = (int32_t)(i * stride);
}

abx_status bruteAlloc(abx_alloc_fn alloc, void *user, int which, size_t bytes, cudaStream_t s, void **out)
{
  if (alloc)
  {
    *out = alloc(user, which, bytes);
    if (!*out && bytes)
    {
      setError("output allocator returned NULL");
      return ABX_ERR_ARG;
    }
    return ABX_OK;
  }
  return deviceAlloc(out, std::max<size_t>(bytes, 4), s);
}

template <int PRED, bool FILL>
abx_status launchSpatial(abx_brute *b, cudaStream_t s, float const *preds, int64_t q, int32_t *counts,
                         int32_t const *offsets, uint32_t *indices)
{
  int const grid = divUp(q, kThreads);
  if (b->kind == ABX_PRIM_BOX3F)
    ABX_LAUNCH((bruteSpatialKernel<PRED, true, FILL>), grid, kThreads, 0, s, b->lo, b->hi, b->n, preds, q, counts, offsets,
               indices);
  else
    ABX_LAUNCH((bruteSpatialKernel<PRED, false, FILL>), grid, kThreads, 0, s, b->lo, b->hi, b->n, preds, q, counts,
               offsets, indices);
  return ABX_OK;
}

template <bool FILL>
abx_status dispatchSpatial(abx_brute *b, cudaStream_t s, int pred_kind, float const *preds, int64_t q, int32_t *counts,
                           int32_t const *offsets, uint32_t *indices)
{
  switch (pred_kind)
  {
  case ABX_PRED_SPHERE3F: return launchSpatial<ABX_PRED_SPHERE3F, FILL>(b, s, preds, q, counts, offsets, indices);
  case ABX_PRED_BOX3F: return launchSpatial<ABX_PRED_BOX3F, FILL>(b, s, preds, q, counts, offsets, indices);
  case ABX_PRED_POINT3F: return launchSpatial<ABX_PRED_POINT3F, FILL>(b, s, preds, q, counts, offsets, indices);
  case ABX_PRED_RAY3F:
    if (b->kind == ABX_PRIM_BOX3F)
      return launchSpatial<ABX_PRED_RAY3F, FILL>(b, s, preds, q, counts, offsets, indices);
    setError("intersects(Ray) is defined for box primitives");
    return ABX_ERR_ARG;
  default: setError("unknown predicate kind"); return ABX_ERR_ARG;
  }
}

} // namespace
} // namespace abx

using namespace abx;

extern "C"
{

abx_status abx_brute_destroy(abx_brute *b)
{
  if (!b)
    return ABX_OK;
  deviceFree(b->lo, b->stream);
  deviceFree(b->hi, b->stream);
  deviceFree(b->bounds_dev, b->stream);
  delete b;
  return ABX_OK;
}

abx_status abx_brute_create(void *stream, int prim_kind, const void *prims_dev, int64_t n, abx_brute **out)
{
  if (!out)
  {
    setError("null output handle");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  if ((prim_kind != ABX_PRIM_POINT3F && prim_kind != ABX_PRIM_BOX3F) || n < 0 || n >= (int64_t)1 << 30 ||
      (n > 0 && !prims_dev))
  {
    setError("BruteForce: point or box primitives, 0 <= n < 2^30");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  abx_brute *b = new abx_brute;
  b->kind = prim_kind;
  b->n = n;
  b->stream = s;
  auto fail = [&](abx_status st) {
    abx_brute_destroy(b);
    return st;
  };
  abx_status st = deviceAlloc((void **)&b->bounds_dev, 6 * sizeof(float), s);
  if (st == ABX_OK)
    st = deviceAlloc((void **)&b->lo, sizeof(float4) * (size_t)std::max<int64_t>(n, 1), s);
  if (st == ABX_OK && prim_kind == ABX_PRIM_BOX3F)
    st = deviceAlloc((void **)&b->hi, sizeof(float4) * (size_t)std::max<int64_t>(n, 1), s);
  if (st != ABX_OK)
    return fail(st);
  // bounds of the scene (BruteForceImpl.hpp:33-50): the same reduction the tree build uses
  TempBuffer<unsigned> enc;
  st = enc.alloc(6, s);
  if (st == ABX_OK)
    st = sceneBounds(s, prim_kind, prims_dev, n, enc.ptr);
  if (st == ABX_OK)
    st = decodeBounds(s, enc.ptr, b->bounds_dev);
  if (st != ABX_OK)
    return fail(st);
  if (n > 0)
  {
    if (prim_kind == ABX_PRIM_POINT3F)
      bruteInitKernel<ABX_PRIM_POINT3F><<<divUp(n, 256), 256, 0, s>>>((float const *)prims_dev, n, b->lo, b->hi);
    else
      bruteInitKernel<ABX_PRIM_BOX3F><<<divUp(n, 256), 256, 0, s>>>((float const *)prims_dev, n, b->lo, b->hi);
    ++g_launch_count;
    if (cudaGetLastError() != cudaSuccess)
    {
      setError("bruteInitKernel launch failed");
      return fail(ABX_ERR_CUDA);
    }
  }
  *out = b;
  return ABX_OK;
}

int64_t abx_brute_size(const abx_brute *b) { return b ? b->n : 0; }

abx_status abx_brute_bounds(abx_brute *b, float out6[6])
{
  if (!b || !out6)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  if (!b->bounds_host_valid)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(b->bounds_host, b->bounds_dev, 6 * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    ABX_CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->bounds_host_valid = true;
  }
  for (int d = 0; d < 6; ++d)
    out6[d] = b->bounds_host[d];
  return ABX_OK;
}

abx_status abx_brute_query_spatial_crs(abx_brute *b, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                       abx_alloc_fn alloc, void *user, int32_t **offsets_dev, uint32_t **indices_dev,
                                       int64_t *nnz)
{
  if (!b || !offsets_dev || !indices_dev || !nnz || q < 0 || q >= (int64_t)1 << 30 || (q > 0 && !preds_dev))
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  if (pred_kind == ABX_PRED_SPHERE3F && (reinterpret_cast<uintptr_t>(preds_dev) & 15u))
  {
    setError("sphere predicates must be 16-byte aligned");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  void *off_v = nullptr, *idx_v = nullptr;
  ABX_TRY(bruteAlloc(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off_v));
  int32_t *offsets = (int32_t *)off_v;
  *offsets_dev = offsets;
  *indices_dev = nullptr;
  *nnz = 0;
  ABX_CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)(q + 1), s));
  if (q == 0 || b->n == 0)
  {
    ABX_TRY(bruteAlloc(alloc, user, 1, 0, s, &idx_v));
    *indices_dev = (uint32_t *)idx_v;
    return ABX_OK;
  }
  ABX_TRY(dispatchSpatial<false>(b, s, pred_kind, (float const *)preds_dev, q, offsets, nullptr, nullptr));
  TempBuffer<unsigned long long> total64;
  ABX_TRY(total64.alloc(1, s));
  ABX_TRY(exclusiveScanI32(s, offsets, offsets, q + 1, total64.ptr));
  unsigned long long total = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&total, total64.ptr, sizeof(total), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  if (total >= (1ull << 31))
  {
    setError("BruteForce: more than 2^31 results");
    return ABX_ERR_ARG;
  }
  *nnz = (int64_t)total;
  ABX_TRY(bruteAlloc(alloc, user, 1, sizeof(uint32_t) * (size_t)total, s, &idx_v));
  *indices_dev = (uint32_t *)idx_v;
  if (total > 0)
    ABX_TRY(dispatchSpatial<true>(b, s, pred_kind, (float const *)preds_dev, q, nullptr, offsets, *indices_dev));
  return ABX_OK;
}

abx_status abx_brute_query_nearest_crs(abx_brute *b, void *stream, const void *points_dev, int64_t q, int32_t k,
                                       abx_alloc_fn alloc, void *user, int32_t **offsets_dev, uint32_t **indices_dev,
                                       float **distances_dev, int64_t *nnz)
{
  if (!b || !offsets_dev || !indices_dev || !nnz || q < 0 || q >= (int64_t)1 << 30 || (q > 0 && !points_dev))
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int const row = (int)std::max<int64_t>(0, std::min<int64_t>(k, b->n));
  int64_t const total = (int64_t)row * q;
  if (total >= (int64_t)1 << 31)
  {
    setError("BruteForce: more than 2^31 results");
    return ABX_ERR_ARG;
  }
  void *off_v = nullptr, *idx_v = nullptr, *d_v = nullptr;
  ABX_TRY(bruteAlloc(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off_v));
  ABX_TRY(bruteAlloc(alloc, user, 1, sizeof(uint32_t) * (size_t)total, s, &idx_v));
  if (distances_dev)
    ABX_TRY(bruteAlloc(alloc, user, 2, sizeof(float) * (size_t)total, s, &d_v));
  *offsets_dev = (int32_t *)off_v;
  *indices_dev = (uint32_t *)idx_v;
  if (distances_dev)
    *distances_dev = (float *)d_v;
  *nnz = total;
  ABX_LAUNCH(bruteStrideOffsetsKernel, divUp(q + 1, 256), 256, 0, s, (int32_t *)off_v, q + 1, row);
  if (total == 0)
    return ABX_OK;
  TempBuffer<float2> scratch;
  ABX_TRY(scratch.alloc((size_t)total, s));
  int const grid = divUp(q, kThreads);
  if (b->kind == ABX_PRIM_BOX3F)
    ABX_LAUNCH((bruteNearestKernel<true>), grid, kThreads, 0, s, b->lo, b->hi, b->n, (float const *)points_dev, q, k, row,
               scratch.ptr, (uint32_t *)idx_v, (float *)d_v);
  else
    ABX_LAUNCH((bruteNearestKernel<false>), grid, kThreads, 0, s, b->lo, b->hi, b->n, (float const *)points_dev, q, k,
               row, scratch.ptr, (uint32_t *)idx_v, (float *)d_v);
  return ABX_OK;
}

} // extern "C"
