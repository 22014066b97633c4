// abx_capi.cu -- the C ABI (include/abx.h): argument checking, orchestration of
// the kernels in abx_{sort,build,query,dbscan}.cu, output allocation.
#include "abx_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace abx
{

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launch_count{0};

void setError(std::string const &msg) { t_last_error = msg; }

#ifdef ABX_TUNING
int tuneIntEnv(char const *name, int dflt)
{
  char const *e = getenv(name);
  return e ? atoi(e) : dflt;
}
#endif

// ---- per-kernel timing -------------------------------------------------------------
bool g_profile = false;
namespace
{
struct ProfileRecord
{
  char const *name;
  cudaEvent_t start, stop;
};
std::vector<ProfileRecord> g_records;
std::vector<cudaEvent_t> g_event_pool;
std::mutex g_profile_mutex;
cudaEvent_t takeEvent()
{
  if (!g_event_pool.empty())
  {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
} // namespace
int profileBegin(char const *name, cudaStream_t s)
{
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  ProfileRecord r{name, takeEvent(), takeEvent()};
  cudaEventRecord(r.start, s);
  g_records.push_back(r);
  return (int)g_records.size() - 1;
}
void profileEnd(int handle, cudaStream_t s)
{
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  if (handle >= 0 && handle < (int)g_records.size())
    cudaEventRecord(g_records[handle].stop, s);
}

abx_status ensureDevice()
{
  static std::once_flag once;
  static cudaError_t init_err = cudaSuccess;
  std::call_once(once, [] {
    int count = 0;
    init_err = cudaGetDeviceCount(&count);
    if (init_err == cudaSuccess && count == 0)
      init_err = cudaErrorNoDevice;
    if (init_err != cudaSuccess)
      return;
  });
  if (init_err != cudaSuccess)
  {
    setError(std::string("no usable CUDA device (libabx has no CPU fallback): ") + cudaGetErrorString(init_err));
    return ABX_ERR_CUDA;
  }
  return ABX_OK;
}

// ---- device memory: a small caching allocator -------------------------------------------
// Temporaries and tree buffers are recycled through per-(device, stream, size class) free
// lists, like the reference's memory pools keep Views cheap: a block released on a stream
// is handed to the next request of the same class on the SAME stream, where stream order
// already guarantees that its previous users are done.  After the first call of a given
// shape no driver allocation happens at all.  (cudaMallocAsync's pool was measured to keep
// re-growing for ~10 steps under this library's mix of 4 KB ... 1.3 GB requests.)
namespace
{
struct AllocKey
{
  int device;
  cudaStream_t stream;
  size_t cls;
  bool operator<(AllocKey const &o) const
  {
    if (device != o.device)
      return device < o.device;
    if (stream != o.stream)
      return stream < o.stream;
    return cls < o.cls;
  }
};
std::mutex g_alloc_mutex;
std::map<AllocKey, std::vector<void *>> g_free_blocks;
struct BlockInfo
{
  int device;
  size_t cls;
  cudaStream_t owner; // stream of the request that handed the block out (its users are ordered there)
};
std::unordered_map<void *, BlockInfo> g_block_info;
int64_t g_cached_bytes = 0;

size_t sizeClass(size_t bytes)
{
  if (bytes < 512)
    return 512;
  // 1/8-octave classes: at most 12.5 % slack
  int const msb = 63 - __builtin_clzll((unsigned long long)bytes);
  size_t const step = (size_t)1 << (msb >= 3 ? msb - 3 : 0);
  return (bytes + step - 1) / step * step;
}
} // namespace

abx_status deviceAlloc(void **p, size_t bytes, cudaStream_t s)
{
  *p = nullptr;
  size_t const cls = sizeClass(bytes);
  int dev = 0;
  ABX_CUDA_TRY(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    auto it = g_free_blocks.find(AllocKey{dev, s, cls});
    if (it != g_free_blocks.end() && !it->second.empty())
    {
      *p = it->second.back();
      it->second.pop_back();
      g_cached_bytes -= (int64_t)cls;
      g_block_info[*p].owner = s;
      return ABX_OK;
    }
  }
  cudaError_t e = cudaMalloc(p, cls);
  if (e != cudaSuccess)
  {
    // give cached blocks back to the driver and retry once
    abx_trim();
    e = cudaMalloc(p, cls);
  }
  if (e != cudaSuccess)
  {
    setError(std::string("cudaMalloc(") + std::to_string(cls) + "): " + cudaGetErrorString(e));
    return ABX_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lock(g_alloc_mutex);
  g_block_info[*p] = BlockInfo{dev, cls, s};
  return ABX_OK;
}

void deviceFree(void *p, cudaStream_t s)
{
  if (!p)
    return;
  BlockInfo info;
  {
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    auto it = g_block_info.find(p);
    if (it == g_block_info.end())
      return; // not ours
    info = it->second;
  }
  if (info.owner != s)
  {
    // Released under another stream than the one its users were enqueued on (abx_free of a result
    // from a different execution space): the block is filed under `s`, so `s` must first wait for
    // everything the owner stream has been given so far.
    cudaEvent_t ev = nullptr;
    bool ordered = false;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess)
    {
      ordered = cudaEventRecord(ev, info.owner) == cudaSuccess && cudaStreamWaitEvent(s, ev, 0) == cudaSuccess;
      cudaEventDestroy(ev); // released once the wait has been satisfied
    }
    if (!ordered)
    {
      cudaGetLastError();      // e.g. the owner stream no longer exists
      cudaDeviceSynchronize(); // nothing can still be using the block after this
    }
  }
  std::lock_guard<std::mutex> lock(g_alloc_mutex);
  g_free_blocks[AllocKey{info.device, s, info.cls}].push_back(p);
  g_cached_bytes += (int64_t)info.cls;
}

static abx_policy defaultPolicy()
{
  abx_policy p;
  p.buffer_size = 0;
  p.sort_predicates = 1;
  return p;
}

static int predStride(int kind)
{
  return kind == ABX_PRED_SPHERE3F ? 4 : (kind == ABX_PRED_BOX3F || kind == ABX_PRED_RAY3F) ? 6 : 3;
}
static int primStride(int kind) { return kind == ABX_PRIM_POINT3F ? 3 : kind == ABX_PRIM_BOX3F ? 6 : 9; }

// output buffer from the caller's allocator or from the stream-ordered pool
static abx_status allocOut(abx_alloc_fn alloc, void *user, int which, size_t bytes, cudaStream_t s, void **out)
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
  return deviceAlloc(out, bytes, s);
}

__global__ void fillStrideKernel(int32_t *offsets, int64_t q_plus_1, int stride)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q_plus_1)
    offsets[i] = (int32_t)(i * stride);
}

// overflow[0] = 1 if any count exceeds |buffer|
__global__ void overflowKernel(int32_t const *__restrict__ counts, int64_t q, int buffer, int *overflow)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q && counts[i] > buffer)
    *overflow = 1;
}

static abx_status checkPredPointer(int pred_kind, void const *preds, int64_t q)
{
  if (q > 0 && !preds)
  {
    setError("null predicates");
    return ABX_ERR_ARG;
  }
  if (pred_kind == ABX_PRED_SPHERE3F && (reinterpret_cast<uintptr_t>(preds) & 15u))
  {
    setError("sphere predicates must be 16-byte aligned");
    return ABX_ERR_ARG;
  }
  return ABX_OK;
}

// spatial CRS: CrsGraphWrapperImpl.hpp:148-446.  The reference counts, scans and
// traverses again to fill (buffer_size = 0), or stores into per-query buffers in the
// first pass and compacts.  Here ONE traversal counts and stages the first kStage
// results of every query in a slot-major buffer; after the scan a compaction kernel
// writes the CRS rows and re-traverses only queries with more than kStage results.
// buffer_size never changes the result, so the policy is honoured for its error
// contract only (hard preallocation overflow throws, :263-268).
// The CRS driver in two halves, so that a caller (DistributedTree) can put its own host work between the enqueue
// of the traversal and the one blocking point of the query.
//   spatialCrsBegin  enqueues the predicate ordering, the traversal (count + stage), the scan and the read-back of nnz
//   spatialCrsEnd    blocks for nnz (the reference blocks there too: lastElement, CrsGraphWrapperImpl.hpp:248),
//                    allocates the indices and enqueues the compaction
abx_status spatialCrsBegin(SpatialCrsCall &c, abx_bvh *bvh, cudaStream_t s, int pred_kind, void const *preds, int64_t q,
                           abx_policy const &policy, abx_alloc_fn alloc, void *user)
{
  c.bvh = bvh, c.s = s, c.pred_kind = pred_kind, c.preds = preds, c.q = q, c.policy = policy, c.alloc = alloc, c.user = user;
  ABX_TRY(checkPredPointer(pred_kind, preds, q));
  if (q < 0 || q >= (int64_t)1 << 30)
  {
    setError("number of predicates must be in [0, 2^30)");
    return ABX_ERR_ARG;
  }
  void *offsets_v = nullptr;
  ABX_TRY(allocOut(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &offsets_v));
  c.offsets = (int32_t *)offsets_v;
  c.trivial = q == 0 || bvh->n == 0;
  if (c.trivial)
  {
    ABX_CUDA_TRY(cudaMemsetAsync(c.offsets, 0, sizeof(int32_t) * (size_t)(q + 1), s));
    return ABX_OK;
  }
  if (policy.sort_predicates && bvh->n > 1)
    ABX_TRY(predicatePermutation(s, bvh, pred_kind, preds, q, c.qperm));
  // the traversal: counts land in offsets[0..q) in ORIGINAL query order
  c.staged = bvh->n > 1;
  if (c.staged)
  {
    ABX_TRY(c.staging.alloc((size_t)spatialStageSlots() * (size_t)q, s));
    ABX_TRY(spatialStage(s, bvh, pred_kind, preds, q, c.qperm.ptr, c.offsets, c.staging.ptr));
  }
  else
    ABX_TRY(spatialCount(s, bvh, pred_kind, preds, q, c.qperm.ptr, 0, c.offsets));
  if (policy.buffer_size < 0)
  {
    ABX_TRY(c.overflow.alloc(1, s));
    ABX_CUDA_TRY(cudaMemsetAsync(c.overflow.ptr, 0, sizeof(int), s));
    ABX_LAUNCH(overflowKernel, divUp(q, 256), 256, 0, s, c.offsets, q, -policy.buffer_size, c.overflow.ptr);
    ABX_CUDA_TRY(cudaMemcpyAsync(c.overflow_out, c.overflow.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  // the int32 scan wraps silently past 2^31 results: the total is accumulated in 64 bits next to it
  ABX_TRY(c.total64.alloc(1, s));
  ABX_TRY(exclusiveScanI32(s, c.offsets, c.offsets, q + 1, c.total64.ptr));
  ABX_CUDA_TRY(cudaMemcpyAsync(c.total_out, c.total64.ptr, sizeof(c.h_total), cudaMemcpyDeviceToHost, s));
  return ABX_OK;
}

abx_status spatialCrsEnd(SpatialCrsCall &c, int32_t **offsets_out, uint32_t **indices_out, int64_t *nnz_out, bool sync)
{
  cudaStream_t const s = c.s;
  *offsets_out = c.offsets;
  *indices_out = nullptr;
  *nnz_out = 0;
  void *idx = nullptr;
  if (c.trivial)
  {
    ABX_TRY(allocOut(c.alloc, c.user, 1, 0, s, &idx));
    *indices_out = (uint32_t *)idx;
    if (sync)
      ABX_CUDA_TRY(cudaStreamSynchronize(s));
    return ABX_OK;
  }
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  if (*c.total_out >= (1ull << 31))
  {
    setError("spatial query: more than 2^31 results (CRS offsets are 32-bit like the reference's)");
    return ABX_ERR_ARG;
  }
  int64_t const total = (int64_t)*c.total_out;
  *nnz_out = total;
  if (total == 0)
  {
    ABX_TRY(allocOut(c.alloc, c.user, 1, 0, s, &idx));
    *indices_out = (uint32_t *)idx;
    return ABX_OK; // :252-261
  }
  if (*c.overflow_out)
  {
    setError("SearchException: hard preallocation buffer_size is too small for the results");
    return ABX_ERR_SEARCH;
  }
  ABX_TRY(allocOut(c.alloc, c.user, 1, sizeof(uint32_t) * (size_t)total, s, &idx));
  *indices_out = (uint32_t *)idx;
  if (c.staged)
    // original query order: coalesced staging-row reads and CRS writes (see kStage in abx_query.cu)
    ABX_TRY(spatialCompact(s, c.bvh, c.pred_kind, c.preds, c.q, nullptr, c.offsets, *indices_out, c.staging.ptr));
  else
    ABX_TRY(spatialFill(s, c.bvh, c.pred_kind, c.preds, c.q, c.qperm.ptr, c.offsets, *indices_out));
  return ABX_OK;
}

abx_status spatialCrsWait(SpatialCrsCall &c, int64_t *nnz_out)
{
  *nnz_out = 0;
  ABX_CUDA_TRY(cudaStreamSynchronize(c.s));
  if (c.trivial)
    return ABX_OK;
  if (*c.total_out >= (1ull << 31))
  {
    setError("spatial query: more than 2^31 results (CRS offsets are 32-bit like the reference's)");
    return ABX_ERR_ARG;
  }
  if (*c.overflow_out && *c.total_out > 0)
  {
    setError("SearchException: hard preallocation buffer_size is too small for the results");
    return ABX_ERR_SEARCH;
  }
  *nnz_out = (int64_t)*c.total_out;
  return ABX_OK;
}

abx_status spatialCrsFillInto(SpatialCrsCall &c, int64_t nnz, int32_t const *out_offsets, void *values, int pair_rank)
{
  if (c.trivial || nnz == 0)
    return ABX_OK;
  if (c.staged)
    return spatialCompact(c.s, c.bvh, c.pred_kind, c.preds, c.q, nullptr, c.offsets, (uint32_t *)values, c.staging.ptr,
                          out_offsets, pair_rank);
  return spatialFill(c.s, c.bvh, c.pred_kind, c.preds, c.q, c.qperm.ptr, c.offsets, (uint32_t *)values, out_offsets,
                     pair_rank);
}

// before_sync (optional): enqueues more work / read-backs on `s` that the call's one blocking point
// should cover as well (the DistributedTree exchange piggy-backs its count matrix on it).
abx_status spatialCrs(abx_bvh *bvh, cudaStream_t s, int pred_kind, void const *preds, int64_t q,
                      abx_policy const &policy, abx_alloc_fn alloc, void *user, int32_t **offsets_out,
                      uint32_t **indices_out, int64_t *nnz_out, std::function<abx_status()> const &before_sync)
{
  SpatialCrsCall c;
  ABX_TRY(spatialCrsBegin(c, bvh, s, pred_kind, preds, q, policy, alloc, user));
  *offsets_out = c.offsets; // the hook may want to read the scanned offsets
  if (before_sync)
    ABX_TRY(before_sync());
  return spatialCrsEnd(c, offsets_out, indices_out, nnz_out, before_sync != nullptr);
}

abx_status nearestCrs(abx_bvh *bvh, cudaStream_t s, void const *pts, int64_t q, int32_t k,
                             int32_t const *k_per_query, abx_policy const &policy, abx_alloc_fn alloc, void *user,
                             int32_t **offsets_out, uint32_t **indices_out, float **distances_out, int64_t *nnz_out,
                             int pred_kind)
{
  if (q > 0 && !pts)
  {
    setError("null query points");
    return ABX_ERR_ARG;
  }
  if (pred_kind != ABX_PRED_POINT3F && pred_kind != ABX_PRED_BOX3F && pred_kind != ABX_PRED_SPHERE3F &&
      pred_kind != ABX_PRED_RAY3F)
  {
    setError("unknown predicate kind");
    return ABX_ERR_ARG;
  }
  if (pred_kind != ABX_PRED_POINT3F && k_per_query)
  {
    setError("per-query k is implemented for nearest(Point, k)");
    return ABX_ERR_ARG;
  }
  if (pred_kind != ABX_PRED_POINT3F && bvh->kind == ABX_PRIM_TRI3F)
  {
    setError("nearest(Box | Sphere | Ray, k) is defined for point and box primitives");
    return ABX_ERR_ARG;
  }
  ABX_TRY(checkPredPointer(pred_kind, pts, q));
  // nearest(Sphere, k): distance(Sphere, X) = max(distance(centre, X) - r, 0) ranks like the centre's distance
  void const *const preds_in = pts;
  TempBuffer<float> centres;
  if (pred_kind == ABX_PRED_SPHERE3F)
  {
    ABX_TRY(centres.alloc(3 * (size_t)std::max<int64_t>(q, 1), s));
    ABX_TRY(sphereCentres(s, (float const *)preds_in, q, centres.ptr));
    pts = centres.ptr;
  }
  bool const geom = pred_kind == ABX_PRED_BOX3F || pred_kind == ABX_PRED_RAY3F;
  if (q < 0 || q >= (int64_t)1 << 30)
  {
    setError("number of predicates must be in [0, 2^30)");
    return ABX_ERR_ARG;
  }
  void *offsets_v = nullptr;
  ABX_TRY(allocOut(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &offsets_v));
  int32_t *offsets = (int32_t *)offsets_v;
  *offsets_out = offsets;
  *indices_out = nullptr;
  if (distances_out)
    *distances_out = nullptr;
  *nnz_out = 0;
  int const n = (int)bvh->n;
  int64_t total = 0;
  bool const uniform = (k_per_query == nullptr);
  if (uniform)
  {
    // allocateAndInitializeStorage(Nearest) (CrsGraphWrapperImpl.hpp:353-374) with
    // rows already compacted to min(k, n)
    int const row = std::max(0, std::min(k, n));
    total = (int64_t)row * q;
    if (total >= (int64_t)1 << 31)
    {
      setError("nearest query: more than 2^31 results");
      return ABX_ERR_ARG;
    }
    ABX_LAUNCH(fillStrideKernel, divUp(q + 1, 256), 256, 0, s, offsets, q + 1, row);
  }
  else
  {
    ABX_TRY(clipK(s, k_per_query, k, n, q, offsets));
    TempBuffer<unsigned long long> total64;
    ABX_TRY(total64.alloc(1, s));
    ABX_TRY(exclusiveScanI32(s, offsets, offsets, q + 1, total64.ptr));
    unsigned long long t64 = 0;
    ABX_CUDA_TRY(cudaMemcpyAsync(&t64, total64.ptr, sizeof(t64), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    if (t64 >= (1ull << 31))
    {
      setError("nearest query: more than 2^31 results");
      return ABX_ERR_ARG;
    }
    total = (int64_t)t64;
  }
  *nnz_out = total;
  void *idx = nullptr, *dist = nullptr;
  ABX_TRY(allocOut(alloc, user, 1, sizeof(uint32_t) * (size_t)total, s, &idx));
  *indices_out = (uint32_t *)idx;
  if (distances_out)
  {
    ABX_TRY(allocOut(alloc, user, 2, sizeof(float) * (size_t)total, s, &dist));
    *distances_out = (float *)dist;
  }
  if (total == 0)
    return ABX_OK;
  TempBuffer<uint32_t> qperm;
  if (policy.sort_predicates && n > 1)
    ABX_TRY(predicatePermutation(s, bvh, geom ? pred_kind : ABX_PRED_POINT3F, pts, q, qperm));
  // Rows are laid out for min(k, n) results, but a leaf at infinite (or NaN) distance is never
  // accepted (`distance < radius` with radius = +inf, TreeTraversal.hpp:255): such rows come out
  // short and the reference compacts them (CrsGraphWrapperImpl.hpp:296-318).  The kernel counts
  // per row and adds up the missing entries; the compaction below only runs when there are any.
  TempBuffer<int32_t> counts;
  TempBuffer<unsigned long long> missing;
  ABX_TRY(counts.alloc((size_t)q + 1, s));
  ABX_TRY(missing.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(missing.ptr, 0, sizeof(unsigned long long), s));
  if (geom)
    ABX_TRY(nearestGeomQuery(s, bvh, pred_kind, (float const *)pts, q, k, qperm.ptr, total, counts.ptr, *indices_out,
                             (float *)dist, missing.ptr));
  else
    ABX_TRY(nearestQuery(s, bvh, (float const *)pts, q, k, k_per_query, qperm.ptr, uniform ? nullptr : offsets, total,
                         counts.ptr, *indices_out, (float *)dist, missing.ptr));
  if (pred_kind == ABX_PRED_SPHERE3F && dist) // rows still have their uniform stride here
    ABX_TRY(sphereDistances(s, total, std::max(1, std::min(k, n)), (float const *)preds_in, (float *)dist));
  unsigned long long h_missing = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&h_missing, missing.ptr, sizeof(h_missing), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  if (h_missing == 0)
    return ABX_OK;
  // underflow: compact into exactly sized outputs
  int64_t const new_total = total - (int64_t)h_missing;
  TempBuffer<int32_t> old_offsets;
  ABX_TRY(old_offsets.alloc((size_t)q + 1, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(old_offsets.ptr, offsets, sizeof(int32_t) * (size_t)(q + 1), cudaMemcpyDeviceToDevice, s));
  ABX_TRY(exclusiveScanI32(s, counts.ptr, offsets, q + 1));
  uint32_t *old_idx = *indices_out;
  float *old_dist = (float *)dist;
  TempBuffer<uint32_t> keep_idx;   // library-allocated outputs must outlive the copy
  TempBuffer<float> keep_dist;
  if (alloc)
  {
    // the caller's allocator hands out new views; stage the old rows first
    ABX_TRY(keep_idx.alloc((size_t)total, s));
    ABX_CUDA_TRY(cudaMemcpyAsync(keep_idx.ptr, old_idx, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToDevice, s));
    old_idx = keep_idx.ptr;
    if (dist)
    {
      ABX_TRY(keep_dist.alloc((size_t)total, s));
      ABX_CUDA_TRY(cudaMemcpyAsync(keep_dist.ptr, old_dist, sizeof(float) * (size_t)total, cudaMemcpyDeviceToDevice, s));
      old_dist = keep_dist.ptr;
    }
    ABX_CUDA_TRY(cudaStreamSynchronize(s)); // the allocator may free the previous views
  }
  void *nidx = nullptr, *ndist = nullptr;
  ABX_TRY(allocOut(alloc, user, 1, sizeof(uint32_t) * (size_t)new_total, s, &nidx));
  if (distances_out)
    ABX_TRY(allocOut(alloc, user, 2, sizeof(float) * (size_t)new_total, s, &ndist));
  ABX_TRY(compactRows(s, q, old_offsets.ptr, offsets, old_idx, old_dist, (uint32_t *)nidx, (float *)ndist));
  if (!alloc)
  {
    deviceFree(*indices_out, s);
    deviceFree(dist, s);
  }
  *indices_out = (uint32_t *)nidx;
  if (distances_out)
    *distances_out = (float *)ndist;
  *nnz_out = new_total;
  return ABX_OK;
}

} // namespace abx

// ---- Experimental::findHalfNeighborList / findFullNeighborList (spatial/detail/ArborX_NeighborList.hpp:47-192,
//      ArborX_ExpandHalfToFull.hpp:24-72): CRS neighbour lists from the half traversal's pair list ----
namespace abx
{
namespace
{
// pass 1: rows[e] = row of entry e, counts[row] += 1; entry e < m: pair e as (second -> first); full lists add the
// mirrored entries e >= m as (first -> second)
__global__ void neighborRowsKernel(uint32_t const *__restrict__ pairs, int64_t m, bool full, uint32_t *__restrict__ rows,
                                   uint32_t *__restrict__ vals, int32_t *__restrict__ counts)
{
  int64_t const e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t const total = full ? 2 * m : m;
  if (e >= total)
    return;
  int64_t const p = e < m ? e : e - m;
  uint32_t const a = pairs[2 * p], b = pairs[2 * p + 1];
  uint32_t const row = e < m ? b : a, val = e < m ? a : b;
  rows[e] = row;
  vals[e] = val;
  atomicAdd(counts + row, 1);
}
__global__ void gatherU32Kernel(uint32_t const *__restrict__ src, uint32_t const *__restrict__ perm, int64_t m,
                                uint32_t *__restrict__ dst)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
    dst[i] = src[perm[i]];
}
} // namespace

abx_status neighborList(cudaStream_t s, float const *xyz, int64_t n, float radius, bool full, abx_alloc_fn alloc,
                               void *user, int32_t **offsets_out, uint32_t **indices_out, int64_t *nnz_out)
{
  *nnz_out = 0;
  *indices_out = nullptr;
  void *off_v = nullptr, *idx_v = nullptr;
  ABX_TRY(allocOut(alloc, user, 0, sizeof(int32_t) * (size_t)(n + 1), s, &off_v));
  int32_t *offsets = (int32_t *)off_v;
  *offsets_out = offsets;
  ABX_CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)(n + 1), s));
  abx_bvh *tree = nullptr;
  ABX_TRY(buildTree(s, ABX_PRIM_POINT3F, xyz, n, nullptr, &tree));
  struct Guard
  {
    abx_bvh *t;
    ~Guard() { abx_bvh_destroy(t); }
  } guard{tree};
  // the pairs: count, then fill
  TempBuffer<unsigned long long> cnt;
  ABX_TRY(cnt.alloc(1, s));
  ABX_TRY(halfTraversalPairs(s, tree, radius, nullptr, 0, cnt.ptr));
  unsigned long long m = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&m, cnt.ptr, sizeof(m), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  int64_t const total = (int64_t)(full ? 2 * m : m);
  if (total >= (int64_t)1 << 30)
  {
    setError("neighbour list: more than 2^30 entries");
    return ABX_ERR_ARG;
  }
  *nnz_out = total;
  ABX_TRY(allocOut(alloc, user, 1, sizeof(uint32_t) * (size_t)total, s, &idx_v));
  *indices_out = (uint32_t *)idx_v;
  if (total == 0)
    return ABX_OK;
  TempBuffer<uint32_t> pairs, rows, vals, perm;
  ABX_TRY(pairs.alloc(2 * (size_t)m, s));
  ABX_TRY(rows.alloc((size_t)total, s));
  ABX_TRY(vals.alloc((size_t)total, s));
  ABX_TRY(perm.alloc((size_t)total, s));
  ABX_TRY(halfTraversalPairs(s, tree, radius, pairs.ptr, (int64_t)m, cnt.ptr));
  ABX_LAUNCH(neighborRowsKernel, divUp(total, 256), 256, 0, s, pairs.ptr, (int64_t)m, full, rows.ptr, vals.ptr, offsets);
  ABX_TRY(exclusiveScanI32(s, offsets, offsets, n + 1));
  // group the entries by row: stable sort of (row, entry), then one gather
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < n)
    ++bits;
  ABX_TRY(sortPairsU32(s, rows.ptr, perm.ptr, total, true, bits, /*fixup=*/false));
  ABX_LAUNCH(gatherU32Kernel, divUp(total, 256), 256, 0, s, vals.ptr, perm.ptr, total, *indices_out);
  return ABX_OK;
}
} // namespace abx

using namespace abx;

extern "C"
{

const char *abx_last_error(void) { return t_last_error.c_str(); }
int abx_version(void) { return ABX_VERSION; }
int64_t abx_launch_count(void) { return g_launch_count; }

abx_status abx_profile_enable(int on)
{
  ABX_TRY(ensureDevice());
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  for (auto &r : g_records)
  {
    g_event_pool.push_back(r.start);
    g_event_pool.push_back(r.stop);
  }
  g_records.clear();
  g_profile = on != 0;
  return ABX_OK;
}

// "name\tlaunches\ttotal_ms\tmax_ms\n" per kernel, most expensive first; returns the number
// of bytes needed (including the terminating NUL)
int64_t abx_profile_report(char *buf, int64_t capacity)
{
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lock(g_profile_mutex);
  struct Agg
  {
    std::string name;
    int64_t count = 0;
    double ms = 0, max_ms = 0;
  };
  std::vector<Agg> aggs;
  for (auto &r : g_records)
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess)
      continue;
    std::string name(r.name);
    auto it = std::find_if(aggs.begin(), aggs.end(), [&](Agg const &a) { return a.name == name; });
    if (it == aggs.end())
    {
      aggs.push_back(Agg{name, 0, 0, 0});
      it = aggs.end() - 1;
    }
    it->count++;
    it->ms += ms;
    it->max_ms = std::max(it->max_ms, (double)ms);
  }
  std::sort(aggs.begin(), aggs.end(), [](Agg const &a, Agg const &b) { return a.ms > b.ms; });
  std::string out;
  for (auto &a : aggs)
  {
    char line[512];
    snprintf(line, sizeof line, "%s\t%lld\t%.6f\t%.6f\n", a.name.c_str(), (long long)a.count, a.ms, a.max_ms);
    out += line;
  }
  if (buf && capacity > 0)
  {
    size_t const c = std::min<size_t>(out.size(), (size_t)capacity - 1);
    memcpy(buf, out.data(), c);
    buf[c] = 0;
  }
  return (int64_t)out.size() + 1;
}

// releases every cached (currently unused) device block; returns the bytes released
int64_t abx_trim(void)
{
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lock(g_alloc_mutex);
  int64_t released = 0;
  for (auto &kv : g_free_blocks)
  {
    for (void *p : kv.second)
    {
      cudaFree(p);
      g_block_info.erase(p);
      released += (int64_t)kv.first.cls;
    }
    kv.second.clear();
  }
  g_cached_bytes = 0;
  return released;
}

abx_status abx_free(void *stream, void *ptr_dev)
{
  ABX_TRY(ensureDevice());
  deviceFree(ptr_dev, (cudaStream_t)stream);
  return ABX_OK;
}

abx_status abx_bvh_build(void *stream, int prim_kind, const void *prims_dev, int64_t n, abx_bvh **out)
{
  if (!out)
  {
    setError("null output handle");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  return buildTree((cudaStream_t)stream, prim_kind, prims_dev, n, nullptr, out, /*want_wide=*/true);
}

abx_status abx_bvh_build_host(void *stream, int prim_kind, const void *prims_host, int64_t n, abx_bvh **out)
{
  if (!out)
  {
    setError("null output handle");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  if (prim_kind < 0 || prim_kind > ABX_PRIM_TRI3F || n < 0)
  {
    setError("bad primitive kind or count");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> dev;
  ABX_TRY(dev.alloc((size_t)primStride(prim_kind) * (size_t)std::max<int64_t>(n, 1), s));
  if (n > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(dev.ptr, prims_host, sizeof(float) * primStride(prim_kind) * (size_t)n,
                                 cudaMemcpyHostToDevice, s));
  return buildTree(s, prim_kind, dev.ptr, n, nullptr, out, /*want_wide=*/true);
}

// triangles given as vertex-index triples (the Triangles AccessTraits of
// benchmarks/triangulated_surface_distance/triangulated_surface_distance.cpp:34-58): gathered into the flat form
__global__ void gatherTrianglesKernel(float const *__restrict__ vertices, int64_t n_vertices,
                                      int32_t const *__restrict__ tri3, int64_t n, float *__restrict__ flat9,
                                      int *__restrict__ bad)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
#pragma unroll
  for (int c = 0; c < 3; ++c)
  {
    int64_t v = tri3[3 * i + c];
    if (v < 0 || v >= n_vertices)
    {
      *bad = 1;
      v = 0;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
      flat9[9 * i + 3 * c + d] = vertices[3 * v + d];
  }
}

abx_status abx_bvh_build_indexed_triangles(void *stream, const float *vertices_dev, int64_t n_vertices,
                                           const int32_t *triangles_dev, int64_t n, abx_bvh **out)
{
  if (!out || n < 0 || n_vertices < 0 || (n > 0 && (!vertices_dev || !triangles_dev || n_vertices == 0)))
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> flat;
  TempBuffer<int> bad;
  ABX_TRY(flat.alloc(9 * (size_t)std::max<int64_t>(n, 1), s));
  ABX_TRY(bad.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(bad.ptr, 0, sizeof(int), s));
  if (n > 0)
    ABX_LAUNCH(gatherTrianglesKernel, divUp(n, 256), 256, 0, s, vertices_dev, n_vertices, triangles_dev, n, flat.ptr,
               bad.ptr);
  int h_bad = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  if (h_bad)
  {
    setError("triangle vertex index out of range");
    return ABX_ERR_ARG;
  }
  return buildTree(s, ABX_PRIM_TRI3F, flat.ptr, n, nullptr, out, /*want_wide=*/true);
}

abx_status abx_bvh_build_from_sorted_codes(void *stream, int prim_kind, const void *prims_dev,
                                           const uint64_t *sorted_codes_dev, int64_t n, abx_bvh **out)
{
  if (!out || (n > 0 && !sorted_codes_dev))
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  return buildTree((cudaStream_t)stream, prim_kind, prims_dev, n, sorted_codes_dev, out, /*want_wide=*/true);
}

int64_t abx_bvh_size(const abx_bvh *bvh) { return bvh ? bvh->n : 0; }
int abx_bvh_empty(const abx_bvh *bvh) { return !bvh || bvh->n == 0; }
int64_t abx_bvh_memory_bytes(const abx_bvh *bvh) { return bvh ? bvh->bytes : 0; }

abx_status abx_bvh_bounds(abx_bvh *bvh, float out6[6])
{
  if (!bvh || !out6)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  if (!bvh->bounds_host_valid)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(bvh->bounds_host, bvh->bounds_dev, 6 * sizeof(float), cudaMemcpyDeviceToHost,
                                 bvh->stream));
    ABX_CUDA_TRY(cudaStreamSynchronize(bvh->stream));
    bvh->bounds_host_valid = true;
  }
  for (int d = 0; d < 6; ++d)
    out6[d] = bvh->bounds_host[d];
  return ABX_OK;
}

abx_status abx_bvh_export_reference_layout(abx_bvh *bvh, void *stream, int32_t *leaf_rope, uint32_t *leaf_index,
                                           int32_t *left_child, int32_t *rope, float *boxes6, uint64_t *sorted_codes)
{
  if (!bvh)
  {
    setError("null tree");
    return ABX_ERR_ARG;
  }
  return exportReference((cudaStream_t)stream, bvh, leaf_rope, leaf_index, left_child, rope, boxes6, sorted_codes);
}

abx_status abx_query_spatial_crs(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                 const abx_policy *policy, abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                 uint32_t **indices_dev, int64_t *nnz)
{
  if (!bvh || !offsets_dev || !indices_dev || !nnz)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  abx_policy const p = policy ? *policy : defaultPolicy();
  return spatialCrs(bvh, (cudaStream_t)stream, pred_kind, preds_dev, q, p, alloc, user, offsets_dev, indices_dev, nnz);
}

abx_status abx_query_spatial_count(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                   int sort_predicates, int32_t limit, int32_t *counts_dev)
{
  if (!bvh || (q > 0 && !counts_dev))
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  ABX_TRY(checkPredPointer(pred_kind, preds_dev, q));
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<uint32_t> qperm;
  if (sort_predicates && bvh->n > 1 && q > 0)
    ABX_TRY(predicatePermutation(s, bvh, pred_kind, preds_dev, q, qperm));
  return spatialCount(s, bvh, pred_kind, preds_dev, q, qperm.ptr, limit, counts_dev);
}

abx_status abx_query_nearest_crs(abx_bvh *bvh, void *stream, const void *points_dev, int64_t q, int32_t k,
                                 const int32_t *k_per_query_dev, const abx_policy *policy, abx_alloc_fn alloc,
                                 void *user, int32_t **offsets_dev, uint32_t **indices_dev, float **distances_dev,
                                 int64_t *nnz)
{
  if (!bvh || !offsets_dev || !indices_dev || !nnz)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  abx_policy const p = policy ? *policy : defaultPolicy();
  return nearestCrs(bvh, (cudaStream_t)stream, points_dev, q, k, k_per_query_dev, p, alloc, user, offsets_dev,
                    indices_dev, distances_dev, nnz);
}

abx_status abx_query_nearest_geom_crs(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                      int32_t k, const abx_policy *policy, abx_alloc_fn alloc, void *user,
                                      int32_t **offsets_dev, uint32_t **indices_dev, float **distances_dev, int64_t *nnz)
{
  if (!bvh || !offsets_dev || !indices_dev || !nnz)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  abx_policy const p = policy ? *policy : defaultPolicy();
  return nearestCrs(bvh, (cudaStream_t)stream, preds_dev, q, k, nullptr, p, alloc, user, offsets_dev, indices_dev,
                    distances_dev, nnz, pred_kind);
}

// ---- host-buffer variants: H2D of the predicates and D2H of the CRS arrays are
// part of the call (the end-to-end path timed by bench.py) -------------------------
abx_status abx_query_spatial_crs_host(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_host, int64_t q,
                                      const abx_policy *policy, abx_alloc_fn alloc_host, void *user,
                                      int32_t **offsets_host, uint32_t **indices_host, int64_t *nnz)
{
  if (!bvh || !offsets_host || !indices_host || !nnz || !alloc_host)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  if (pred_kind < 0 || pred_kind > ABX_PRED_RAY3F || q < 0)
  {
    setError("bad predicate kind or count");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  abx_policy const p = policy ? *policy : defaultPolicy();
  TempBuffer<float> preds;
  ABX_TRY(preds.alloc((size_t)predStride(pred_kind) * (size_t)std::max<int64_t>(q, 1), s));
  if (q > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(preds.ptr, preds_host, sizeof(float) * predStride(pred_kind) * (size_t)q,
                                 cudaMemcpyHostToDevice, s));
  int32_t *off_dev = nullptr;
  uint32_t *idx_dev = nullptr;
  abx_status st = spatialCrs(bvh, s, pred_kind, preds.ptr, q, p, nullptr, nullptr, &off_dev, &idx_dev, nnz);
  if (st == ABX_OK)
  {
    *offsets_host = (int32_t *)alloc_host(user, 0, sizeof(int32_t) * (size_t)(q + 1));
    *indices_host = (uint32_t *)alloc_host(user, 1, sizeof(uint32_t) * (size_t)*nnz);
    cudaError_t e = cudaMemcpyAsync(*offsets_host, off_dev, sizeof(int32_t) * (size_t)(q + 1), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && *nnz > 0)
      e = cudaMemcpyAsync(*indices_host, idx_dev, sizeof(uint32_t) * (size_t)*nnz, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
    {
      setError(std::string("result copy: ") + cudaGetErrorString(e));
      st = ABX_ERR_CUDA;
    }
  }
  deviceFree(off_dev, s);
  deviceFree(idx_dev, s);
  return st;
}

abx_status abx_query_nearest_crs_host(abx_bvh *bvh, void *stream, const void *points_host, int64_t q, int32_t k,
                                      const abx_policy *policy, abx_alloc_fn alloc_host, void *user,
                                      int32_t **offsets_host, uint32_t **indices_host, float **distances_host,
                                      int64_t *nnz)
{
  if (!bvh || !offsets_host || !indices_host || !nnz || !alloc_host || q < 0)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  abx_policy const p = policy ? *policy : defaultPolicy();
  TempBuffer<float> pts;
  ABX_TRY(pts.alloc(3 * (size_t)std::max<int64_t>(q, 1), s));
  if (q > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(pts.ptr, points_host, sizeof(float) * 3 * (size_t)q, cudaMemcpyHostToDevice, s));
  int32_t *off_dev = nullptr;
  uint32_t *idx_dev = nullptr;
  float *dist_dev = nullptr;
  abx_status st = nearestCrs(bvh, s, pts.ptr, q, k, nullptr, p, nullptr, nullptr, &off_dev, &idx_dev,
                             distances_host ? &dist_dev : nullptr, nnz);
  if (st == ABX_OK)
  {
    *offsets_host = (int32_t *)alloc_host(user, 0, sizeof(int32_t) * (size_t)(q + 1));
    *indices_host = (uint32_t *)alloc_host(user, 1, sizeof(uint32_t) * (size_t)*nnz);
    cudaError_t e = cudaMemcpyAsync(*offsets_host, off_dev, sizeof(int32_t) * (size_t)(q + 1), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && *nnz > 0)
      e = cudaMemcpyAsync(*indices_host, idx_dev, sizeof(uint32_t) * (size_t)*nnz, cudaMemcpyDeviceToHost, s);
    if (distances_host)
    {
      *distances_host = (float *)alloc_host(user, 2, sizeof(float) * (size_t)*nnz);
      if (e == cudaSuccess && *nnz > 0)
        e = cudaMemcpyAsync(*distances_host, dist_dev, sizeof(float) * (size_t)*nnz, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
    {
      setError(std::string("result copy: ") + cudaGetErrorString(e));
      st = ABX_ERR_CUDA;
    }
  }
  deviceFree(off_dev, s);
  deviceFree(idx_dev, s);
  deviceFree(dist_dev, s);
  return st;
}

abx_status abx_half_traversal_pairs(abx_bvh *bvh, void *stream, float r, uint32_t *pairs_dev, int64_t capacity,
                                    int64_t *count)
{
  if (!bvh || !count)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<unsigned long long> cnt;
  ABX_TRY(cnt.alloc(1, s));
  ABX_TRY(halfTraversalPairs(s, bvh, r, pairs_dev, capacity, cnt.ptr));
  unsigned long long h = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&h, cnt.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  *count = (int64_t)h;
  return ABX_OK;
}

abx_status abx_find_half_neighbor_list(void *stream, const float *xyz_dev, int64_t n, float radius, abx_alloc_fn alloc,
                                       void *user, int32_t **offsets_dev, uint32_t **indices_dev, int64_t *nnz)
{
  ABX_TRY(ensureDevice());
  if (!offsets_dev || !indices_dev || !nnz || n < 0 || (n > 0 && !xyz_dev))
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  return neighborList((cudaStream_t)stream, xyz_dev, n, radius, false, alloc, user, offsets_dev, indices_dev, nnz);
}
abx_status abx_find_full_neighbor_list(void *stream, const float *xyz_dev, int64_t n, float radius, abx_alloc_fn alloc,
                                       void *user, int32_t **offsets_dev, uint32_t **indices_dev, int64_t *nnz)
{
  ABX_TRY(ensureDevice());
  if (!offsets_dev || !indices_dev || !nnz || n < 0 || (n > 0 && !xyz_dev))
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  return neighborList((cudaStream_t)stream, xyz_dev, n, radius, true, alloc, user, offsets_dev, indices_dev, nnz);
}

abx_status abx_dbscan(void *stream, const float *xyz_dev, int64_t n, float eps, int32_t minpts, int implementation,
                      int algorithm, int32_t *labels_dev)
{
  ABX_TRY(ensureDevice());
  if (n > 0 && (!xyz_dev || !labels_dev))
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  return dbscan((cudaStream_t)stream, xyz_dev, n, eps, minpts, implementation, algorithm, labels_dev);
}

abx_status abx_dbscan_host(void *stream, const float *xyz_host, int64_t n, float eps, int32_t minpts,
                           int implementation, int algorithm, int32_t *labels_host)
{
  ABX_TRY(ensureDevice());
  if (n > 0 && (!xyz_host || !labels_host))
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  if (n < 0)
  {
    setError("negative point count");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> xyz;
  TempBuffer<int32_t> labels;
  ABX_TRY(xyz.alloc(3 * (size_t)std::max<int64_t>(n, 1), s));
  ABX_TRY(labels.alloc((size_t)std::max<int64_t>(n, 1), s));
  if (n > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(xyz.ptr, xyz_host, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  ABX_TRY(dbscan(s, xyz.ptr, n, eps, minpts, implementation, algorithm, labels.ptr));
  if (n > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(labels_host, labels.ptr, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  return ABX_OK;
}

abx_status abx_dist_merge_crs(void *stream, int64_t q, const int32_t *local_offsets_dev, const int32_t *local_indices_dev,
                              int32_t rank, const int32_t *remote_offsets_dev, const int32_t *remote_values2_dev,
                              int32_t *out_offsets_dev, int32_t *out_values2_dev)
{
  ABX_TRY(ensureDevice());
  return mergeCrs((cudaStream_t)stream, q, local_offsets_dev, local_indices_dev, rank, remote_offsets_dev,
                  remote_values2_dev, out_offsets_dev, out_values2_dev);
}

abx_status abx_bvh_device_view(const abx_bvh *bvh, abx_device_view *view)
{
  if (!bvh || !view)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  view->nodes = bvh->nodes;
  view->leaf_box = bvh->leaf_box;
  view->leaf_tri = bvh->leaf_tri;
  view->n = bvh->n;
  view->prim_kind = bvh->kind;
  return ABX_OK;
}

abx_status abx_dist_merge_sorted(void *stream, int64_t q, const int32_t *local_offsets_dev,
                                 const int32_t *local_indices_dev, int32_t rank, int64_t n_remote,
                                 const int32_t *remote_query_ids_dev, const int32_t *remote_values2_dev,
                                 int32_t *out_offsets_dev, int32_t *out_values2_dev)
{
  ABX_TRY(ensureDevice());
  return mergeSorted((cudaStream_t)stream, q, local_offsets_dev, local_indices_dev, rank, n_remote,
                     remote_query_ids_dev, remote_values2_dev, out_offsets_dev, out_values2_dev);
}

abx_status abx_dist_route_count(void *stream, int pred_kind, const void *preds_dev, int64_t q, const float *radius_dev,
                                int64_t radius_stride, const float *rank_boxes6_dev, int32_t n_ranks, int32_t self_rank,
                                uint32_t *counts_dev)
{
  ABX_TRY(ensureDevice());
  cudaStream_t s = (cudaStream_t)stream;
  ABX_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, sizeof(uint32_t) * (size_t)n_ranks, s));
  return routeLaunch(s, false, pred_kind, preds_dev, q, radius_dev, radius_stride, rank_boxes6_dev, n_ranks, self_rank,
                     counts_dev, nullptr, nullptr, nullptr);
}

abx_status abx_dist_route_fill(void *stream, int pred_kind, const void *preds_dev, int64_t q, const float *radius_dev,
                               int64_t radius_stride, const float *rank_boxes6_dev, int32_t n_ranks, int32_t self_rank,
                               const uint32_t *base_dev, uint32_t *cursors_dev, int32_t *query_ids_dev)
{
  ABX_TRY(ensureDevice());
  cudaStream_t s = (cudaStream_t)stream;
  ABX_CUDA_TRY(cudaMemsetAsync(cursors_dev, 0, sizeof(uint32_t) * (size_t)n_ranks, s));
  return routeLaunch(s, true, pred_kind, preds_dev, q, radius_dev, radius_stride, rank_boxes6_dev, n_ranks, self_rank,
                     nullptr, base_dev, cursors_dev, query_ids_dev);
}

abx_status abx_dist_pair_with_rank(void *stream, const int32_t *indices_dev, int64_t n, int32_t rank,
                                   int32_t *values2_dev)
{
  ABX_TRY(ensureDevice());
  return pairWithRank((cudaStream_t)stream, indices_dev, n, rank, values2_dev);
}

abx_status abx_dist_knn_merge(void *stream, int64_t n_candidates, const int32_t *query_ids_dev,
                              const int32_t *cand_values2_dev, const float *cand_distances_dev, int32_t k,
                              int32_t *values2_dev, float *distances_dev)
{
  ABX_TRY(ensureDevice());
  return knnMerge((cudaStream_t)stream, n_candidates, query_ids_dev, cand_values2_dev, cand_distances_dev, k,
                  values2_dev, distances_dev);
}

// ---- stage-level entry points -------------------------------------------------------
abx_status abx_scene_bounds(void *stream, int prim_kind, const void *prims_dev, int64_t n, float *bounds6_dev)
{
  ABX_TRY(ensureDevice());
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<unsigned> enc;
  ABX_TRY(enc.alloc(6, s));
  ABX_TRY(sceneBounds(s, prim_kind, prims_dev, n, enc.ptr));
  return decodeBounds(s, enc.ptr, bounds6_dev);
}

abx_status abx_morton64(void *stream, int prim_kind, const void *prims_dev, int64_t n, const float *bounds6_dev,
                        uint64_t *codes_dev)
{
  ABX_TRY(ensureDevice());
  return morton64((cudaStream_t)stream, prim_kind, prims_dev, n, bounds6_dev, codes_dev);
}

abx_status abx_morton32(void *stream, int pred_kind, const void *preds_dev, int64_t q, const float *bounds6_dev,
                        uint32_t *codes_dev)
{
  ABX_TRY(ensureDevice());
  return morton32((cudaStream_t)stream, pred_kind, preds_dev, q, bounds6_dev, codes_dev);
}

abx_status abx_sort_u64(void *stream, uint64_t *keys_dev, uint32_t *perm_dev, int64_t n)
{
  ABX_TRY(ensureDevice());
  return sortPairsU64((cudaStream_t)stream, keys_dev, perm_dev, n, true);
}

abx_status abx_sort_u32(void *stream, uint32_t *keys_dev, uint32_t *perm_dev, int64_t n)
{
  ABX_TRY(ensureDevice());
  return sortPairsU32((cudaStream_t)stream, keys_dev, perm_dev, n, true);
}

abx_status abx_exclusive_scan_i32(void *stream, const int32_t *in_dev, int32_t *out_dev, int64_t n_plus_1)
{
  ABX_TRY(ensureDevice());
  return exclusiveScanI32((cudaStream_t)stream, in_dev, out_dev, n_plus_1);
}

} // extern "C"
