// abx_dist.cu -- ArborX::DistributedTree: communicators and the host side of the sharded search.
//
// Behavioural contract: distributed/ArborX_DistributedTree.hpp:33-252 (ctor: bottom tree, all-gather of
// rank boxes and sizes, replicated top tree), detail/ArborX_DistributedTreeSpatial.hpp:31-60,
// detail/ArborX_DistributedTreeNearest.hpp:41-261 (two-phase kNN), detail/ArborX_DistributedTreeUtils.hpp
// (forwardQueries :52-115, communicateResultsBack :153-224, countResults / sort :229-263, filterResults
// :267-342), detail/ArborX_Distributor.hpp:40-541 (MPI point-to-point, three messages each way).
//
// Shape here.  The reference forwards EVERY query through its exchange, including the vast majority
// that only concern the rank they live on.  Here the local tree answers all local predicates directly;
// a routing kernel tests each predicate against the R rank boxes (the top tree for R <= 64, conservative
// for spheres) and only predicates that also touch OTHER ranks are forwarded, queried there and merged
// back per query.  For kNN the local k-th distance bounds the true one, so the reference's phase I needs
// no exchange at all: rows short of k locally carry an infinite bound and are forwarded everywhere.
// One schedule of collectives for every input (no rank-local fall-back decision):
//   spatial   all-gather(R x R counts) | send/recv(predicates, ids) | all-gather(R x R counts) | send/recv(results)
//   nearest   the same, with (index, id, distance) results
// and two blocking points per call (each covers one count matrix; the first one is the local query's own).
// Columns of one exchange travel in ONE NCCL group (a single fused kernel on NVLink / NVSwitch).
#include "abx_common.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

// ---------------------------------------------------------------- communicators ----
struct ExchangeColumn
{
  void const *send;
  void *recv;
  size_t elem_bytes;
};

struct abx_comm
{
  int rank = 0, size = 1;
  virtual ~abx_comm() {}
  // every rank contributes `bytes` bytes; recv holds size * bytes
  virtual abx_status allGather(void const *send, void *recv, size_t bytes, cudaStream_t s) = 0;
  // element ranges [off[r], off[r + 1]) of every column go to / come from rank r (host arrays, R + 1 entries)
  virtual abx_status allToAllV(ExchangeColumn const *cols, int ncols, int64_t const *send_off, int64_t const *recv_off,
                               cudaStream_t s) = 0;
};

namespace abx
{
namespace
{

// ---- NCCL, resolved at run time: the library loads (and the single-GPU path works) without it ----
struct NcclApi
{
  void *handle = nullptr;
  std::string error;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *ncclApi()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy the process already has (e.g. the one torch loaded) wins: same soname
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle)
      api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle)
    {
      api.error = std::string("libnccl.so.2 not found: ") + dlerror();
      return;
    }
#define ABX_NCCL_SYM(field, name)                                                                                     \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));                                          \
  if (!api.field)                                                                                                      \
    api.error = std::string("missing NCCL symbol ") + name;
    ABX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    ABX_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    ABX_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    ABX_NCCL_SYM(CommCount, "ncclCommCount")
    ABX_NCCL_SYM(CommUserRank, "ncclCommUserRank")
    ABX_NCCL_SYM(AllGather, "ncclAllGather")
    ABX_NCCL_SYM(Send, "ncclSend")
    ABX_NCCL_SYM(Recv, "ncclRecv")
    ABX_NCCL_SYM(GroupStart, "ncclGroupStart")
    ABX_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    ABX_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ABX_NCCL_SYM
  });
  if (!api.error.empty())
  {
    setError("NCCL unavailable: " + api.error);
    return nullptr;
  }
  return &api;
}

#define ABX_NCCL_TRY(api, expr)                                                                                       \
  do                                                                                                                   \
  {                                                                                                                    \
    ncclResult_t _r = (expr);                                                                                          \
    if (_r != ncclSuccess)                                                                                             \
    {                                                                                                                  \
      ::abx::setError(std::string(#expr) + ": " + (api)->GetErrorString(_r));                                          \
      return ABX_ERR_CUDA;                                                                                             \
    }                                                                                                                  \
  } while (0)

struct NcclComm : abx_comm
{
  NcclApi *api = nullptr;
  ncclComm_t comm = nullptr;
  bool owned = false;
  ~NcclComm() override
  {
    if (owned && comm && api)
      api->CommDestroy(comm);
  }
  abx_status allGather(void const *send, void *recv, size_t bytes, cudaStream_t s) override
  {
    ABX_NCCL_TRY(api, api->AllGather(send, recv, bytes, ncclInt8, comm, s));
    ++g_launch_count; // the collective's kernel
    return ABX_OK;
  }
  abx_status allToAllV(ExchangeColumn const *cols, int ncols, int64_t const *send_off, int64_t const *recv_off,
                       cudaStream_t s) override
  {
    bool any = false;
    for (int r = 0; r < size; ++r)
      any |= r != rank && (send_off[r + 1] > send_off[r] || recv_off[r + 1] > recv_off[r]);
    for (int c = 0; c < ncols; ++c)
    {
      // a rank is never its own destination in the DistributedTree exchanges, but keep the operation complete
      size_t const eb = cols[c].elem_bytes;
      int64_t const cnt = send_off[rank + 1] - send_off[rank];
      if (cnt > 0)
        ABX_CUDA_TRY(cudaMemcpyAsync((char *)cols[c].recv + recv_off[rank] * eb,
                                     (char const *)cols[c].send + send_off[rank] * eb, cnt * eb,
                                     cudaMemcpyDeviceToDevice, s));
    }
    if (!any)
      return ABX_OK;
    ABX_NCCL_TRY(api, api->GroupStart());
    for (int r = 0; r < size; ++r)
    {
      if (r == rank)
        continue;
      int64_t const ns = send_off[r + 1] - send_off[r], nr = recv_off[r + 1] - recv_off[r];
      for (int c = 0; c < ncols; ++c)
      {
        size_t const eb = cols[c].elem_bytes;
        if (ns > 0)
          ABX_NCCL_TRY(api, api->Send((char const *)cols[c].send + send_off[r] * eb, (size_t)ns * eb, ncclInt8, r, comm, s));
        if (nr > 0)
          ABX_NCCL_TRY(api, api->Recv((char *)cols[c].recv + recv_off[r] * eb, (size_t)nr * eb, ncclInt8, r, comm, s));
      }
    }
    ABX_NCCL_TRY(api, api->GroupEnd());
    ++g_launch_count;
    return ABX_OK;
  }
};

// ---- in-process group: R host threads, one GPU.  The collectives are device-to-device copies between
// the threads' buffers around a thread barrier -- the same protocol code runs on a single-GPU box.
struct LocalGroup
{
  int size = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<void const *> ptr;
  std::vector<ExchangeColumn const *> cols;
  std::vector<int64_t const *> off;
  void barrier()
  {
    std::unique_lock<std::mutex> lock(m);
    long const g = generation;
    if (++arrived == size)
    {
      arrived = 0;
      ++generation;
      cv.notify_all();
    }
    else
      cv.wait(lock, [&] { return generation != g; });
  }
};

struct LocalComm : abx_comm
{
  std::shared_ptr<LocalGroup> g;
  abx_status allGather(void const *send, void *recv, size_t bytes, cudaStream_t s) override
  {
    ABX_CUDA_TRY(cudaStreamSynchronize(s)); // the contribution is complete
    g->ptr[rank] = send;
    g->barrier();
    for (int r = 0; r < size; ++r)
      ABX_CUDA_TRY(cudaMemcpyAsync((char *)recv + (size_t)r * bytes, g->ptr[r], bytes, cudaMemcpyDeviceToDevice, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    g->barrier(); // nobody reuses its send buffer before every reader is done
    return ABX_OK;
  }
  abx_status allToAllV(ExchangeColumn const *cols, int ncols, int64_t const *send_off, int64_t const *recv_off,
                       cudaStream_t s) override
  {
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    g->cols[rank] = cols;
    g->off[rank] = send_off;
    g->barrier();
    for (int r = 0; r < size; ++r)
    {
      int64_t const first = g->off[r][rank], cnt = g->off[r][rank + 1] - first;
      if (cnt != recv_off[r + 1] - recv_off[r])
      {
        setError("local communicator: send / receive counts disagree");
        g->barrier();
        return ABX_ERR_ARG;
      }
      for (int c = 0; c < ncols && cnt > 0; ++c)
      {
        size_t const eb = cols[c].elem_bytes;
        ABX_CUDA_TRY(cudaMemcpyAsync((char *)cols[c].recv + recv_off[r] * eb, (char const *)g->cols[r][c].send + first * eb,
                                     cnt * eb, cudaMemcpyDeviceToDevice, s));
      }
    }
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    g->barrier();
    return ABX_OK;
  }
};

// ------------------------------------------------------------------ small kernels ----
// send_preds[j] = preds[qids[j]] (W 32-bit words per predicate)
__global__ void gatherWordsKernel(uint32_t const *__restrict__ src, int32_t const *__restrict__ ids, int64_t m, int W,
                                  uint32_t *__restrict__ dst)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * W)
    return;
  int64_t const j = i / W;
  int const w = (int)(i - j * W);
  dst[i] = src[(int64_t)ids[j] * W + w];
}

// out[r] = off[starts[r + 1]] - off[starts[r]] for r < R (results per source rank of the forwarded queries)
__global__ void segmentTotalsKernel(int32_t const *__restrict__ off, int32_t const *__restrict__ starts, int R,
                                    uint32_t *__restrict__ out)
{
  int const r = threadIdx.x;
  if (r < R)
    out[r] = (uint32_t)(off[starts[r + 1]] - off[starts[r]]);
}

// ids_out[e] = ids[row of result e] for a CRS with offsets off[0 .. g]
__global__ void expandRowIdsKernel(int32_t const *__restrict__ off, int32_t const *__restrict__ ids, int g, int64_t nnz,
                                   int32_t *__restrict__ ids_out)
{
  int64_t const e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz)
    return;
  int lo = 0, hi = g; // largest row with off[row] <= e
  while (hi - lo > 1)
  {
    int const mid = (lo + hi) >> 1;
    if (off[mid] <= e)
      lo = mid;
    else
      hi = mid;
  }
  ids_out[e] = ids[lo];
}

// kNN results of the forwarded queries: rows of `stride` slots with counts[j] valid entries -> contiguous
// records (index, distance, query id) at off[j]
__global__ void packKnnResultsKernel(int g, int stride, int32_t const *__restrict__ counts, int32_t const *__restrict__ off,
                                     uint32_t const *__restrict__ idx, float const *__restrict__ dist,
                                     int32_t const *__restrict__ ids, int32_t *__restrict__ out_idx,
                                     float *__restrict__ out_dist, int32_t *__restrict__ out_ids)
{
  int const j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g)
    return;
  int const c = counts[j], o = off[j], id = ids[j];
  for (int e = 0; e < c; ++e)
  {
    out_idx[o + e] = (int32_t)idx[(int64_t)j * stride + e];
    out_dist[o + e] = dist[(int64_t)j * stride + e];
    out_ids[o + e] = id;
  }
}

// received records in query-id order: position j takes record perm[j]; its owner rank is the segment of the
// receive buffer the record arrived in (seg_off: R + 1 entries)
__global__ void finishRemoteKernel(int64_t m, uint32_t const *__restrict__ perm, int32_t const *__restrict__ got_idx,
                                   float const *__restrict__ got_dist, int32_t const *__restrict__ seg_off, int R,
                                   int2 *__restrict__ vals2, float *__restrict__ dist_out)
{
  int64_t const j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m)
    return;
  int const src = (int)perm[j];
  int rank = 0;
  while (rank + 1 < R && seg_off[rank + 1] <= src)
    ++rank;
  vals2[j] = make_int2(got_idx[src], rank);
  if (dist_out)
    dist_out[j] = got_dist[src];
}

__global__ void fillPaddedRowsKernel(int64_t n, int2 *__restrict__ vals2, float *__restrict__ dist)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  vals2[i] = make_int2(-1, -1);
  if (dist)
    dist[i] = __int_as_float(0x7f800000);
}

__global__ void fillStrideOffsetsKernel(int32_t *offsets, int64_t q_plus_1, int stride)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q_plus_1)
    offsets[i] = (int32_t)(i * stride);
}

// valid (index >= 0) entries per padded kNN row
__global__ void countValidKernel(int64_t q, int k, int2 const *__restrict__ vals2, int32_t *__restrict__ counts)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q)
    return;
  int c = 0;
  for (int j = 0; j < k; ++j)
    c += vals2[i * k + j].x >= 0 ? 1 : 0;
  counts[i] = c;
}
__global__ void compactPaddedRowsKernel(int64_t q, int k, int32_t const *__restrict__ off, int2 const *__restrict__ vals2,
                                        float const *__restrict__ dist, int2 *__restrict__ out_vals2,
                                        float *__restrict__ out_dist)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q)
    return;
  int const o = off[i], c = off[i + 1] - o;
  for (int j = 0; j < c; ++j)
  {
    out_vals2[o + j] = vals2[i * k + j];
    if (out_dist)
      out_dist[o + j] = dist[i * k + j];
  }
}

// ---- compact (host) result form: indices + the list of entries owned by other ranks ----
// remote records (query id ascending) -> indices behind the local part of their rows + (position, rank) list
__global__ void compactRemoteRowsKernel(int64_t m, int32_t const *__restrict__ ids, int2 const *__restrict__ vals,
                                        int32_t const *__restrict__ local_off, int32_t const *__restrict__ out_off,
                                        uint32_t *__restrict__ out_idx, int32_t *__restrict__ remote_pos,
                                        int32_t *__restrict__ remote_rank)
{
  int64_t const c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m)
    return;
  int const qid = ids[c];
  if (c > 0 && ids[c - 1] == qid)
    return;
  int dst = out_off[qid] + (local_off[qid + 1] - local_off[qid]);
  for (int64_t e = c; e < m && ids[e] == qid; ++e, ++dst)
  {
    int2 const v = vals[e];
    out_idx[dst] = (uint32_t)v.x;
    remote_pos[e] = dst; // ids ascending => positions ascending
    remote_rank[e] = v.y;
  }
}
// kNN: (index, rank) rows -> index rows
__global__ void splitPairsKernel(int64_t n, int2 const *__restrict__ vals2, uint32_t *__restrict__ idx)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    idx[i] = (uint32_t)vals2[i].x;
}
// kNN: the rows that received candidates (segment leaders of ids) list their entries owned by other ranks
__global__ void listRemoteInRowsKernel(int64_t m, int32_t const *__restrict__ ids, int k, int2 const *__restrict__ vals2,
                                       int32_t const *__restrict__ row_off /* null: i * k */, int self_rank,
                                       unsigned *__restrict__ counter, uint32_t *__restrict__ pos_out,
                                       uint32_t *__restrict__ rank_out)
{
  int64_t const c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m)
    return;
  int const qid = ids[c];
  if (c > 0 && ids[c - 1] == qid)
    return;
  int64_t const out_base = row_off ? (int64_t)row_off[qid] : (int64_t)qid * k;
  int slot = 0;
  for (int j = 0; j < k; ++j)
  {
    int2 const v = vals2[(int64_t)qid * k + j];
    if (v.x < 0)
      break; // padding
    if (v.y != self_rank)
    {
      unsigned const o = atomicAdd(counter, 1u);
      pos_out[o] = (uint32_t)(out_base + slot);
      rank_out[o] = (uint32_t)v.y;
    }
    ++slot;
  }
}

// pinned host scratch of a tree (count matrices): cudaHostAlloc costs a fraction of a millisecond, and
// applications build trees inside their time step -- recycled through a small free list
constexpr size_t kPinnedScratchWords = 64 * 64 + 8;
std::mutex g_pinned_mutex;
std::vector<uint32_t *> g_pinned_free;
abx_status takePinnedScratch(uint32_t **out)
{
  {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    if (!g_pinned_free.empty())
    {
      *out = g_pinned_free.back();
      g_pinned_free.pop_back();
      return ABX_OK;
    }
  }
  ABX_CUDA_TRY(cudaHostAlloc((void **)out, sizeof(uint32_t) * kPinnedScratchWords, cudaHostAllocPortable));
  return ABX_OK;
}
void returnPinnedScratch(uint32_t *p)
{
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  g_pinned_free.push_back(p);
}

// side streams are recycled too: the device-memory cache files blocks per stream, and a tree built inside a time
// step must find the blocks its predecessor released
std::vector<std::pair<int, cudaStream_t>> g_side_free; // (device, stream), under g_pinned_mutex
abx_status takeSideStream(cudaStream_t *out)
{
  int dev = 0;
  ABX_CUDA_TRY(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    for (size_t i = 0; i < g_side_free.size(); ++i)
      if (g_side_free[i].first == dev)
      {
        *out = g_side_free[i].second;
        g_side_free.erase(g_side_free.begin() + i);
        return ABX_OK;
      }
  }
  int prio_lo = 0, prio_hi = 0;
  ABX_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  ABX_CUDA_TRY(cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, prio_hi));
  return ABX_OK;
}
void returnSideStream(cudaStream_t s)
{
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  g_side_free.push_back({dev, s});
}

// forwarded batches below this size are traversed in arrival order (already grouped by source rank and, inside a
// source, by the sender's routing order): six launches of Morton ordering cost more than they return
constexpr int64_t kSortForwardedAbove = 32768;
// distributed kNN with at least this many points per call goes in two stages (near-boundary points first, their
// exchange hidden behind the interior points' traversal)
constexpr int64_t kTwoStageKnnAbove = 262144;

int predWords(int kind) { return kind == ABX_PRED_SPHERE3F ? 4 : kind == ABX_PRED_BOX3F ? 6 : 3; }
int primWords(int kind) { return kind == ABX_PRIM_POINT3F ? 3 : kind == ABX_PRIM_BOX3F ? 6 : 9; }

int bitsFor(int64_t n)
{
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < n)
    ++b;
  return b;
}

abx_status allocOutDev(abx_alloc_fn alloc, void *user, int which, size_t bytes, cudaStream_t s, void **out)
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

} // namespace
} // namespace abx

using namespace abx;

// ----------------------------------------------------------------------- the tree ----
struct abx_dist_tree
{
  abx_comm *comm = nullptr;
  abx_bvh *bottom = nullptr;
  int kind = 0, R = 1, rank = 0;
  std::vector<float> boxes;   // R x 6 (rank boxes = leaves of the top tree)
  std::vector<int64_t> sizes; // primitives per rank
  int64_t total = 0;
  float *boxes_dev = nullptr;
  uint32_t *h_pin = nullptr; // pinned scratch: R x R count matrix, then one 64-bit word (8-byte aligned)
  cudaStream_t side = nullptr; // high-priority stream of the exchange (overlaps the local traversal)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  float bounds[6];
};

namespace abx
{
namespace
{

// Phase timeline of a distributed query (tuning build only, ABX_DIST_TRACE=1): every mark drains the streams it is
// given, so the phases are serialised and the numbers attribute the time instead of reproducing the overlapped call.
struct PhaseTrace
{
#ifdef ABX_TUNING
  char const *what;
  int rank;
  int on; // 1: drain the streams at every mark (attribution); 2: host clock only (the overlapped call as it runs)
  std::vector<std::pair<char const *, double>> marks;
  static double now()
  {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  }
  PhaseTrace(char const *w, int r) : what(w), rank(r), on(ABX_TUNE_INT("ABX_DIST_TRACE", 0))
  {
    if (on)
      marks.emplace_back("start", now());
  }
  void mark(char const *name, cudaStream_t a, cudaStream_t b = nullptr)
  {
    if (!on)
      return;
    if (on == 1)
    {
      cudaStreamSynchronize(a);
      if (b)
        cudaStreamSynchronize(b);
    }
    marks.emplace_back(name, now());
  }
  ~PhaseTrace()
  {
    if (!on)
      return;
    cudaDeviceSynchronize();
    marks.emplace_back("rest", now());
    if (rank != 0)
      return;
    std::string line = std::string("[abx trace] ") + what + ":";
    for (size_t i = 1; i < marks.size(); ++i)
    {
      char buf[96];
      snprintf(buf, sizeof buf, " %s %.3f", marks[i].first, marks[i].second - marks[i - 1].second);
      line += buf;
    }
    char buf[64];
    snprintf(buf, sizeof buf, " | total %.3f", marks.back().second - marks.front().second);
    fprintf(stderr, "%s%s\n", line.c_str(), buf);
  }
#else
  PhaseTrace(char const *, int) {}
  void mark(char const *, cudaStream_t, cudaStream_t = nullptr) {}
#endif
};

// counts_dev[R] of every rank -> pinned host matrix M[src][dst]; the caller's next blocking point covers it
abx_status gatherCountMatrix(abx_dist_tree *t, cudaStream_t s, uint32_t const *counts_dev, uint32_t *matrix_dev)
{
  int const R = t->R;
  ABX_TRY(t->comm->allGather(counts_dev, matrix_dev, sizeof(uint32_t) * R, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(t->h_pin, matrix_dev, sizeof(uint32_t) * R * R, cudaMemcpyDeviceToHost, s));
  return ABX_OK;
}

struct ExchangePlan
{
  std::vector<int64_t> send_off, recv_off; // R + 1
  int64_t n_send = 0, n_recv = 0, global = 0;
  void fromMatrix(uint32_t const *M, int R, int rank)
  {
    send_off.assign(R + 1, 0);
    recv_off.assign(R + 1, 0);
    global = 0;
    for (int r = 0; r < R; ++r)
    {
      send_off[r + 1] = send_off[r] + M[rank * R + r];
      recv_off[r + 1] = recv_off[r] + M[r * R + rank];
    }
    for (int i = 0; i < R * R; ++i)
      global += M[i];
    n_send = send_off[R];
    n_recv = recv_off[R];
  }
};

// Forward the predicates listed by the routing pass (counts already on the host in `plan`), returning the
// received predicates and their query ids on the source rank.
abx_status forwardPredicates(abx_dist_tree *t, cudaStream_t s, int route_kind, void const *preds, int words, int64_t q,
                             float const *radius, int64_t radius_stride, ExchangePlan const &plan,
                             TempBuffer<uint32_t> &fwd_preds, TempBuffer<int32_t> &fwd_ids,
                             uint32_t const *subset = nullptr /* q predicate ids to route, or all */)
{
  int const R = t->R;
  int64_t const F = plan.n_send, G = plan.n_recv;
  TempBuffer<int32_t> qids;
  TempBuffer<uint32_t> send_preds, base, cursors;
  ABX_TRY(qids.alloc((size_t)std::max<int64_t>(F, 1), s));
  ABX_TRY(send_preds.alloc((size_t)std::max<int64_t>(F, 1) * words, s));
  ABX_TRY(fwd_preds.alloc((size_t)std::max<int64_t>(G, 1) * words, s));
  ABX_TRY(fwd_ids.alloc((size_t)std::max<int64_t>(G, 1), s));
  if (F > 0)
  {
    ABX_TRY(base.alloc(R, s));
    ABX_TRY(cursors.alloc(R, s));
    std::vector<uint32_t> h_base(R);
    for (int r = 0; r < R; ++r)
      h_base[r] = (uint32_t)plan.send_off[r];
    // pageable source: staged by the runtime before the call returns
    ABX_CUDA_TRY(cudaMemcpyAsync(base.ptr, h_base.data(), sizeof(uint32_t) * R, cudaMemcpyHostToDevice, s));
    ABX_CUDA_TRY(cudaMemsetAsync(cursors.ptr, 0, sizeof(uint32_t) * R, s));
    ABX_TRY(routeLaunch(s, true, route_kind, preds, q, radius, radius_stride, t->boxes_dev, R, t->rank, nullptr,
                        base.ptr, cursors.ptr, qids.ptr, subset));
    ABX_LAUNCH(gatherWordsKernel, divUp(F * words, 256), 256, 0, s, (uint32_t const *)preds, qids.ptr, F, words,
               send_preds.ptr);
  }
  ExchangeColumn cols[2] = {{send_preds.ptr, fwd_preds.ptr, sizeof(uint32_t) * (size_t)words},
                            {qids.ptr, fwd_ids.ptr, sizeof(int32_t)}};
  return t->comm->allToAllV(cols, 2, plan.send_off.data(), plan.recv_off.data(), s);
}

// received (index, id[, distance]) records -> query-id order with owner ranks
abx_status sortReceived(abx_dist_tree *t, cudaStream_t s, int64_t M, int64_t q, ExchangePlan const &back,
                        TempBuffer<int32_t> &got_ids /* in: ids, out: sorted */, int32_t const *got_idx,
                        float const *got_dist, TempBuffer<int32_t> &vals2, TempBuffer<float> &dist_sorted)
{
  int const R = t->R;
  ABX_TRY(vals2.alloc((size_t)std::max<int64_t>(M, 1) * 2, s));
  if (got_dist)
    ABX_TRY(dist_sorted.alloc((size_t)std::max<int64_t>(M, 1), s));
  if (M == 0)
    return ABX_OK;
  TempBuffer<uint32_t> perm;
  TempBuffer<int32_t> seg;
  ABX_TRY(perm.alloc((size_t)M, s));
  ABX_TRY(seg.alloc(R + 1, s));
  std::vector<int32_t> h_seg(R + 1);
  for (int r = 0; r <= R; ++r)
    h_seg[r] = (int32_t)back.recv_off[r];
  ABX_CUDA_TRY(cudaMemcpyAsync(seg.ptr, h_seg.data(), sizeof(int32_t) * (R + 1), cudaMemcpyHostToDevice, s));
  // stable: records of one query keep their (source rank, traversal) order
  // (plain LSD passes: no fix-up stage, hence no blocking point)
  ABX_TRY(sortPairsU32(s, (uint32_t *)got_ids.ptr, perm.ptr, M, true, bitsFor(q), /*fixup=*/false));
  ABX_LAUNCH(finishRemoteKernel, divUp(M, 256), 256, 0, s, M, perm.ptr, got_idx, got_dist, seg.ptr, R, (int2 *)vals2.ptr,
             got_dist ? dist_sorted.ptr : nullptr);
  return ABX_OK;
}

} // namespace

abx_status localKnnPairs(abx_bvh *bvh, cudaStream_t s, void const *pts, int64_t q, int32_t k, int rank, int32_t *vals2,
                         float *dist, unsigned long long *missing_dev);

// ---- spatial -------------------------------------------------------------------------------
// compact = false: (index, rank) pairs in *values_out (8 bytes per result)
// compact = true : indices in *values_out (4 bytes per result) + remote_pos / remote_rank lists (device, library-owned)
abx_status distSpatial(abx_dist_tree *t, cudaStream_t s, int pred_kind, void const *preds, int64_t q, bool compact,
                       abx_alloc_fn alloc, void *user, int32_t **offsets_out, void **values_out, int64_t *nnz_out,
                       TempBuffer<int32_t> *remote_pos, TempBuffer<int32_t> *remote_rank, int64_t *n_remote)
{
  if (pred_kind != ABX_PRED_SPHERE3F && pred_kind != ABX_PRED_BOX3F && pred_kind != ABX_PRED_POINT3F)
  {
    setError("DistributedTree: spatial predicates are intersects(Sphere | Box | Point)");
    return ABX_ERR_ARG;
  }
  if (q < 0 || q >= (int64_t)1 << 30 || (q > 0 && !preds))
  {
    setError("DistributedTree: bad predicate array");
    return ABX_ERR_ARG;
  }
  int const R = t->R, W = predWords(pred_kind);
  *nnz_out = 0;
  if (n_remote)
    *n_remote = 0;
  abx_policy policy;
  policy.buffer_size = 0;
  policy.sort_predicates = 1;
  if (t->total == 0)
  {
    // DistributedTreeSpatial.hpp:44-50: nothing to search anywhere (every rank takes this branch)
    void *off = nullptr, *vals = nullptr;
    ABX_TRY(allocOutDev(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off));
    ABX_CUDA_TRY(cudaMemsetAsync(off, 0, sizeof(int32_t) * (size_t)(q + 1), s));
    ABX_TRY(allocOutDev(alloc, user, 1, 0, s, &vals));
    *offsets_out = (int32_t *)off;
    *values_out = vals;
    return ABX_OK;
  }
  if (R == 1)
  {
    // one rank: the bottom tree is the whole tree
    if (compact)
    {
      uint32_t *idx = nullptr;
      abx_status const st = spatialCrs(t->bottom, s, pred_kind, preds, q, policy, alloc, user, offsets_out, &idx, nnz_out);
      *values_out = idx;
      return st;
    }
    int32_t *off1 = nullptr;
    uint32_t *idx1 = nullptr;
    int64_t nnz1 = 0;
    ABX_TRY(spatialCrs(t->bottom, s, pred_kind, preds, q, policy, nullptr, nullptr, &off1, &idx1, &nnz1));
    void *off_o = nullptr, *vals_o = nullptr;
    abx_status st = allocOutDev(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off_o);
    if (st == ABX_OK)
      st = allocOutDev(alloc, user, 1, 2 * sizeof(int32_t) * (size_t)nnz1, s, &vals_o);
    if (st == ABX_OK && cudaMemcpyAsync(off_o, off1, sizeof(int32_t) * (size_t)(q + 1), cudaMemcpyDeviceToDevice, s) !=
                            cudaSuccess)
      st = ABX_ERR_CUDA;
    if (st == ABX_OK)
      st = pairWithRank(s, (int32_t const *)idx1, nnz1, t->rank, (int32_t *)vals_o);
    deviceFree(off1, s);
    deviceFree(idx1, s);
    *offsets_out = (int32_t *)off_o;
    *values_out = vals_o;
    *nnz_out = nnz1;
    return st;
  }
  // The exchange (routing, both count matrices, both NCCL exchanges, the query for other ranks' predicates) runs on
  // the tree's side stream while the big local traversal runs on `s`: its host round trips and small kernels hide
  // behind the local traversal instead of stretching the call.
  cudaStream_t const x = t->side;
  PhaseTrace trace("spatial", t->rank);
  ABX_CUDA_TRY(cudaEventRecord(t->ev[0], s)); // the predicates may be produced on s
  ABX_CUDA_TRY(cudaStreamWaitEvent(x, t->ev[0], 0));
  // 1. routing counts of every rank -> count matrix on its way to the host
  TempBuffer<uint32_t> counts, matrix;
  ABX_TRY(counts.alloc(R, x));
  ABX_TRY(matrix.alloc((size_t)R * R, x));
  ABX_CUDA_TRY(cudaMemsetAsync(counts.ptr, 0, sizeof(uint32_t) * R, x));
  ABX_TRY(routeLaunch(x, false, pred_kind, preds, q, nullptr, 0, t->boxes_dev, R, t->rank, counts.ptr, nullptr, nullptr,
                      nullptr));
  ABX_TRY(gatherCountMatrix(t, x, counts.ptr, matrix.ptr));
  ABX_CUDA_TRY(cudaEventRecord(t->ev[1], x));
  trace.mark("route_count+matrix", x);
  // 2. the local tree answers every local predicate (enqueued now, collected after the exchange)
  SpatialCrsCall local;
  // pinned words for its read-backs: Begin must return while the traversal runs, the exchange is enqueued next
  local.total_out = reinterpret_cast<unsigned long long *>(t->h_pin + kPinnedScratchWords - 6);
  local.overflow_out = reinterpret_cast<int *>(t->h_pin + kPinnedScratchWords - 4);
  *local.total_out = 0;
  *local.overflow_out = 0;
  ABX_TRY(spatialCrsBegin(local, t->bottom, s, pred_kind, preds, q, policy, nullptr, nullptr));
  trace.mark("local_begin", s);
  struct Guard
  {
    cudaStream_t s;
    void *a = nullptr, *b = nullptr;
    ~Guard()
    {
      deviceFree(a, s);
      deviceFree(b, s);
    }
  };
  Guard local_guard{s, local.offsets, nullptr}; // the local offsets are a library buffer (no allocator was passed)
  ABX_CUDA_TRY(cudaEventSynchronize(t->ev[1])); // the count matrix is on the host
  ExchangePlan fwd;
  fwd.fromMatrix(t->h_pin, R, t->rank);

  TempBuffer<int32_t> got_ids, got_idx, rvals2;
  TempBuffer<float> unused;
  int64_t M = 0;
  if (fwd.global > 0)
  {
    // 3. forward, query the bottom tree with what arrived; its blocking point covers the back-count matrix
    TempBuffer<uint32_t> fwd_preds;
    TempBuffer<int32_t> fwd_ids;
    ABX_TRY(forwardPredicates(t, x, pred_kind, preds, W, q, nullptr, 0, fwd, fwd_preds, fwd_ids));
    trace.mark("forward", x);
    int64_t const G = fwd.n_recv;
    TempBuffer<int32_t> starts;
    ABX_TRY(starts.alloc(R + 1, x));
    std::vector<int32_t> h_starts(R + 1);
    for (int r = 0; r <= R; ++r)
      h_starts[r] = (int32_t)fwd.recv_off[r];
    ABX_CUDA_TRY(cudaMemcpyAsync(starts.ptr, h_starts.data(), sizeof(int32_t) * (R + 1), cudaMemcpyHostToDevice, x));
    int32_t *off_r = nullptr;
    uint32_t *idx_r = nullptr;
    int64_t nnz_r = 0;
    abx_policy remote_policy = policy;
    remote_policy.sort_predicates = G >= kSortForwardedAbove;
    ABX_TRY(spatialCrs(t->bottom, x, pred_kind, fwd_preds.ptr, G, remote_policy, nullptr, nullptr, &off_r, &idx_r, &nnz_r,
                       [&]() -> abx_status {
                         // runs right after the scan of the remote query's offsets
                         ABX_LAUNCH(segmentTotalsKernel, 1, 64, 0, x, off_r, starts.ptr, R, counts.ptr);
                         return gatherCountMatrix(t, x, counts.ptr, matrix.ptr);
                       }));
    Guard remote_guard{x, off_r, idx_r};
    trace.mark("remote_query", x);
    ExchangePlan back;
    back.fromMatrix(t->h_pin, R, t->rank);
    M = back.n_recv;
    // 4. results back: (index, query id) columns; the indices travel straight out of the CRS array
    TempBuffer<int32_t> res_ids;
    ABX_TRY(res_ids.alloc((size_t)std::max<int64_t>(nnz_r, 1), x));
    if (nnz_r > 0)
      ABX_LAUNCH(expandRowIdsKernel, divUp(nnz_r, 256), 256, 0, x, off_r, fwd_ids.ptr, (int)G, nnz_r, res_ids.ptr);
    ABX_TRY(got_idx.alloc((size_t)std::max<int64_t>(M, 1), x));
    ABX_TRY(got_ids.alloc((size_t)std::max<int64_t>(M, 1), x));
    ExchangeColumn cols[2] = {{idx_r, got_idx.ptr, sizeof(int32_t)}, {res_ids.ptr, got_ids.ptr, sizeof(int32_t)}};
    ABX_TRY(t->comm->allToAllV(cols, 2, back.send_off.data(), back.recv_off.data(), x));
    trace.mark("back", x);
    ABX_TRY(sortReceived(t, x, M, q, back, got_ids, got_idx.ptr, nullptr, rvals2, unused));
    trace.mark("sort_received", x);
  }
  ABX_CUDA_TRY(cudaEventRecord(t->ev[2], x));
  // the local query's own blocking point (nnz); its rows are written straight into the merged result below
  int32_t *const off_l = local.offsets;
  int64_t nnz_l = 0;
  ABX_TRY(spatialCrsWait(local, &nnz_l));
  trace.mark("local_end", s);
  ABX_CUDA_TRY(cudaStreamWaitEvent(s, t->ev[2], 0)); // the remote records are in place
  // buffers taken under the side stream go back to it when this frame unwinds: not before the merge below (on s)
  // has read them
  struct Rejoin
  {
    abx_dist_tree *t;
    cudaStream_t s, x;
    ~Rejoin()
    {
      if (cudaEventRecord(t->ev[0], s) == cudaSuccess)
        cudaStreamWaitEvent(x, t->ev[0], 0);
    }
  } rejoin{t, s, x};
  // 5. merge per query: local results first, then the remote ones.  The merged offsets come from the local counts
  // and the remote ids; the local traversal's compaction writes its rows at those offsets (as pairs or as indices),
  // the remote records are scattered behind them: no intermediate local CRS, no copy pass.
  int64_t const nnz = nnz_l + M;
  if (nnz >= (int64_t)1 << 31)
  {
    setError("DistributedTree: more than 2^31 results on one rank");
    return ABX_ERR_ARG;
  }
  void *off_v = nullptr, *vals_v = nullptr;
  ABX_TRY(allocOutDev(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off_v));
  ABX_TRY(allocOutDev(alloc, user, 1, (compact ? sizeof(uint32_t) : 2 * sizeof(int32_t)) * (size_t)nnz, s, &vals_v));
  *offsets_out = (int32_t *)off_v;
  *values_out = vals_v;
  *nnz_out = nnz;
  if (n_remote)
    *n_remote = M;
  int const pair_rank = compact ? -1 : t->rank;
  if (M == 0)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(off_v, off_l, sizeof(int32_t) * (size_t)(q + 1), cudaMemcpyDeviceToDevice, s));
    return spatialCrsFillInto(local, nnz_l, nullptr, vals_v, pair_rank);
  }
  ABX_TRY(mergeCounts(s, q, off_l, M, got_ids.ptr, (int32_t *)off_v));
  ABX_TRY(spatialCrsFillInto(local, nnz_l, (int32_t const *)off_v, vals_v, pair_rank));
  if (!compact)
    return mergeRemoteRows(s, M, got_ids.ptr, rvals2.ptr, off_l, (int32_t const *)off_v, (int32_t *)vals_v);
  ABX_TRY(remote_pos->alloc((size_t)M, s));
  ABX_TRY(remote_rank->alloc((size_t)M, s));
  ABX_LAUNCH(compactRemoteRowsKernel, divUp(M, 256), 256, 0, s, M, got_ids.ptr, (int2 const *)rvals2.ptr, off_l,
             (int32_t const *)off_v, (uint32_t *)vals_v, remote_pos->ptr, remote_rank->ptr);
  return ABX_OK;
}

// One exchange of the distributed kNN for the points listed in `subset` (or all q points): forward those whose
// (point, local k-th distance) sphere reaches other ranks (`fwd` = the count matrix of the routing pass), k nearest
// in the receivers' bottom trees, candidates back, re-ranking of the affected rows in place.  All on stream s,
// one blocking point (the back-count matrix).
struct KnnRound
{
  TempBuffer<int32_t> got_ids; // query ids of the candidates received, ascending
  int64_t M = 0;
};
abx_status knnExchangeRound(abx_dist_tree *t, cudaStream_t s, void const *pts, int64_t q, int32_t k,
                            uint32_t const *subset, int64_t subset_n, float const *radius, ExchangePlan const &fwd,
                            TempBuffer<uint32_t> &counts, TempBuffer<uint32_t> &matrix, int32_t *rows_p, float *rowsd_p,
                            KnnRound &out, PhaseTrace &trace)
{
  int const R = t->R;
  TempBuffer<int32_t> rvals2;
  TempBuffer<float> rdist;
  TempBuffer<uint32_t> fwd_pts;
  TempBuffer<int32_t> fwd_ids;
  ABX_TRY(forwardPredicates(t, s, ABX_PRED_SPHERE3F, pts, 3, subset_n, radius, k, fwd, fwd_pts, fwd_ids, subset));
  trace.mark("forward", s);
  int64_t const G = fwd.n_recv;
  int const nloc = (int)t->bottom->n;
  int const stride = std::max(1, std::min(k, nloc));
  // 3. k nearest of the forwarded points in the bottom tree
  TempBuffer<uint32_t> r_idx, qperm;
  TempBuffer<float> r_dist;
  TempBuffer<int32_t> r_counts, r_off, starts;
  ABX_TRY(r_idx.alloc((size_t)std::max<int64_t>(G, 1) * stride, s));
  ABX_TRY(r_dist.alloc((size_t)std::max<int64_t>(G, 1) * stride, s));
  ABX_TRY(r_counts.alloc((size_t)G + 1, s));
  ABX_TRY(r_off.alloc((size_t)G + 1, s));
  ABX_TRY(starts.alloc(R + 1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(r_counts.ptr, 0, sizeof(int32_t) * ((size_t)G + 1), s));
  if (G > 0 && nloc > 0)
  {
    if (nloc > 1 && G >= kSortForwardedAbove)
      ABX_TRY(predicatePermutation(s, t->bottom, ABX_PRED_POINT3F, fwd_pts.ptr, G, qperm));
    ABX_TRY(nearestQuery(s, t->bottom, (float const *)fwd_pts.ptr, G, k, nullptr, qperm.ptr, nullptr, G * stride,
                         r_counts.ptr, r_idx.ptr, r_dist.ptr));
  }
  trace.mark("remote_knn", s);
  ABX_TRY(exclusiveScanI32(s, r_counts.ptr, r_off.ptr, G + 1));
  std::vector<int32_t> h_starts(R + 1);
  for (int r = 0; r <= R; ++r)
    h_starts[r] = (int32_t)fwd.recv_off[r];
  ABX_CUDA_TRY(cudaMemcpyAsync(starts.ptr, h_starts.data(), sizeof(int32_t) * (R + 1), cudaMemcpyHostToDevice, s));
  ABX_LAUNCH(segmentTotalsKernel, 1, 64, 0, s, r_off.ptr, starts.ptr, R, counts.ptr);
  ABX_TRY(gatherCountMatrix(t, s, counts.ptr, matrix.ptr));
  ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point
  trace.mark("scan+matrix", s);
  ExchangePlan back;
  back.fromMatrix(t->h_pin, R, t->rank);
  int64_t const M = back.n_recv;
  int64_t const nnz_r = back.n_send;
  // 4. candidates back as (index, distance, query id) columns
  TempBuffer<int32_t> s_idx, s_ids, got_idx;
  TempBuffer<float> s_dist, got_dist;
  ABX_TRY(s_idx.alloc((size_t)std::max<int64_t>(nnz_r, 1), s));
  ABX_TRY(s_ids.alloc((size_t)std::max<int64_t>(nnz_r, 1), s));
  ABX_TRY(s_dist.alloc((size_t)std::max<int64_t>(nnz_r, 1), s));
  if (G > 0)
    ABX_LAUNCH(packKnnResultsKernel, divUp(G, 128), 128, 0, s, (int)G, stride, r_counts.ptr, r_off.ptr, r_idx.ptr,
               r_dist.ptr, fwd_ids.ptr, s_idx.ptr, s_dist.ptr, s_ids.ptr);
  ABX_TRY(got_idx.alloc((size_t)std::max<int64_t>(M, 1), s));
  ABX_TRY(out.got_ids.alloc((size_t)std::max<int64_t>(M, 1), s));
  ABX_TRY(got_dist.alloc((size_t)std::max<int64_t>(M, 1), s));
  ExchangeColumn cols[3] = {{s_idx.ptr, got_idx.ptr, sizeof(int32_t)},
                            {s_ids.ptr, out.got_ids.ptr, sizeof(int32_t)},
                            {s_dist.ptr, got_dist.ptr, sizeof(float)}};
  ABX_TRY(t->comm->allToAllV(cols, 3, back.send_off.data(), back.recv_off.data(), s));
  trace.mark("back", s);
  ABX_TRY(sortReceived(t, s, M, q, back, out.got_ids, got_idx.ptr, got_dist.ptr, rvals2, rdist));
  trace.mark("sort_received", s);
  // 5. final ranking (DistributedTreeNearest.hpp:178-233): the k smallest of local row + candidates
  ABX_TRY(knnMerge(s, M, out.got_ids.ptr, rvals2.ptr, rdist.ptr, k, rows_p, rowsd_p));
  trace.mark("merge", s);
  out.M = M;
  return ABX_OK;
}

// ---- nearest -------------------------------------------------------------------------------
// pairs rows (index, rank) x k per query are produced on the device in both forms; compact = true then
// splits them into index rows + the list of entries owned by other ranks.
abx_status distNearest(abx_dist_tree *t, cudaStream_t s, void const *pts, int64_t q, int32_t k, bool compact,
                       bool want_dist, abx_alloc_fn alloc, void *user, int32_t **offsets_out, void **values_out,
                       float **dist_out, int64_t *nnz_out, TempBuffer<int32_t> *remote_pos,
                       TempBuffer<int32_t> *remote_rank, int64_t *n_remote)
{
  if (q < 0 || q >= (int64_t)1 << 30 || (q > 0 && !pts))
  {
    setError("DistributedTree: bad predicate array");
    return ABX_ERR_ARG;
  }
  int const R = t->R;
  *nnz_out = 0;
  if (n_remote)
    *n_remote = 0;
  if (dist_out)
    *dist_out = nullptr;
  size_t const val_bytes = compact ? sizeof(uint32_t) : 2 * sizeof(int32_t);
  if (t->total == 0 || k < 1)
  {
    void *off = nullptr, *vals = nullptr, *d = nullptr;
    ABX_TRY(allocOutDev(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off));
    ABX_CUDA_TRY(cudaMemsetAsync(off, 0, sizeof(int32_t) * (size_t)(q + 1), s));
    ABX_TRY(allocOutDev(alloc, user, 1, 0, s, &vals));
    if (want_dist)
      ABX_TRY(allocOutDev(alloc, user, 2, 0, s, &d));
    *offsets_out = (int32_t *)off;
    *values_out = vals;
    if (dist_out)
      *dist_out = (float *)d;
    return ABX_OK;
  }
  if ((int64_t)k * q >= (int64_t)1 << 31)
  {
    setError("DistributedTree: more than 2^31 results on one rank");
    return ABX_ERR_ARG;
  }
  int64_t const slots = (int64_t)k * q;
  // 1. local k nearest of every point, rows of k slots in (index, rank) form, padded when short.  In the pairs
  // form the rows ARE the caller's output arrays (rows are full unless a row stays short: handled at the end).
  TempBuffer<int32_t> rows;
  TempBuffer<float> rows_d;
  void *off_v = nullptr, *vals_v = nullptr, *d_v = nullptr;
  int32_t *rows_p = nullptr;
  float *rowsd_p = nullptr;
  if (!compact)
  {
    ABX_TRY(allocOutDev(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off_v));
    ABX_TRY(allocOutDev(alloc, user, 1, val_bytes * (size_t)slots, s, &vals_v));
    if (want_dist)
      ABX_TRY(allocOutDev(alloc, user, 2, sizeof(float) * (size_t)slots, s, &d_v));
    rows_p = (int32_t *)vals_v;
    rowsd_p = (float *)d_v;
  }
  else
  {
    ABX_TRY(rows.alloc((size_t)std::max<int64_t>(slots, 1) * 2, s));
    rows_p = rows.ptr;
  }
  if (!rowsd_p)
  {
    ABX_TRY(rows_d.alloc((size_t)std::max<int64_t>(slots, 1), s));
    rowsd_p = rows_d.ptr;
  }
  TempBuffer<unsigned long long> missing;
  ABX_TRY(missing.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(missing.ptr, 0, sizeof(unsigned long long), s));
  PhaseTrace trace("nearest", t->rank);
  trace.mark("alloc", s);
  float const *radius = rowsd_p + (k - 1);
  unsigned long long *h_missing = reinterpret_cast<unsigned long long *>(t->h_pin + (((size_t)R * R + 1) & ~(size_t)1));
  KnnRound round_a, round_b;
  cudaStream_t const x = t->side;
  // Two stages when there are other ranks and enough queries to hide an exchange behind: the points close to another
  // rank's box go first, and their exchange (routing, both count matrices, both NCCL exchanges, the remote kNN, the
  // re-ranking of their rows) runs on the side stream while the caller's stream walks the interior points.  `near`
  // is a guess (2.4 x the k-neighbour radius of a uniform cloud with the bottom tree's density): interior points
  // are routed as well once their k-th distances are known, and the rare one that does reach another rank takes a
  // second exchange -- the result does not depend on the guess.
  // Every rank runs the same two rounds of collectives (the split is a local choice: a rank with few points, or a
  // degenerate box, puts all its points in the first stage).
  float near = 0.f;
  if (R > 1 && q >= kTwoStageKnnAbove && t->bottom->n > (int64_t)k)
  {
    float const *b = t->boxes.data() + 6 * (size_t)t->rank;
    double const vol = (double)(b[3] - b[0]) * (double)(b[4] - b[1]) * (double)(b[5] - b[2]);
    if (vol > 0 && std::isfinite(vol))
      near = (float)(1.5 * std::cbrt((double)k * vol / (double)t->bottom->n));
  }
  // Measured on 2 and 8 B200s (10M points + 10M queries per rank, profiles/r02_knn_two_stage_experiment.log) and NOT
  // the default: the near launch and the exchange do disappear behind the interior launch, but that launch runs
  // 11.3-11.7 ms instead of 10.85 next to them, and the split itself costs 0.2 ms: kNN phase 12.5 ms against 12.1 ms
  // (N = 2), 12.9 against 12.7 (N = 8).  Kept behind the tuning library's ABX_KNN_TWO_STAGE=1 (every rank must agree:
  // the two forms issue different collectives); the release library always takes the single exchange below.
  bool const two_stage = R > 1 && ABX_TUNE_INT("ABX_KNN_TWO_STAGE", 0) != 0;
  bool const split = near > 0.f && std::isfinite(near);
  TempBuffer<uint32_t> qperm;
  TempBuffer<int32_t> row_found;
  TempBuffer<uint32_t> counts, matrix;
  ABX_TRY(counts.alloc(R, s));
  ABX_TRY(matrix.alloc((size_t)R * R, s));
  bool maybe_short = false;
  if (!two_stage)
  {
    ABX_TRY(localKnnPairs(t->bottom, s, pts, q, k, t->rank, rows_p, rowsd_p, missing.ptr));
    trace.mark("local_knn", s);
    // 2. phase II routing: sphere (point, local k-th distance); an infinite bound reaches every rank
    ABX_CUDA_TRY(cudaMemsetAsync(counts.ptr, 0, sizeof(uint32_t) * R, s));
    ABX_TRY(routeLaunch(s, false, ABX_PRED_SPHERE3F, pts, q, radius, k, t->boxes_dev, R, t->rank, counts.ptr, nullptr,
                        nullptr, nullptr));
    ABX_TRY(gatherCountMatrix(t, s, counts.ptr, matrix.ptr));
    ABX_CUDA_TRY(cudaMemcpyAsync(h_missing, missing.ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 1
    trace.mark("route_count+matrix", s);
    maybe_short = *h_missing != 0;
    ExchangePlan fwd;
    fwd.fromMatrix(t->h_pin, R, t->rank);
    if (fwd.global > 0)
      ABX_TRY(knnExchangeRound(t, s, pts, q, k, nullptr, q, radius, fwd, counts, matrix, rows_p, rowsd_p, round_a, trace));
  }
  else
  {
    // 1a. near-first order, number of near points to the host
    int64_t nb = q;
    ABX_TRY(row_found.alloc((size_t)std::max<int64_t>(q, 1), s));
    if (split)
    {
      TempBuffer<unsigned long long> n_near_dev;
      ABX_TRY(n_near_dev.alloc(1, s));
      ABX_TRY(pointPermutationNearFirst(s, t->bottom, (float const *)pts, q, t->boxes_dev, R, t->rank, near, qperm,
                                        n_near_dev.ptr));
      unsigned long long *h_near = reinterpret_cast<unsigned long long *>(t->h_pin + kPinnedScratchWords - 2);
      ABX_CUDA_TRY(cudaMemcpyAsync(h_near, n_near_dev.ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
      ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 0 (a few hundred microseconds into the call)
      nb = (int64_t)*h_near;
    }
    else if (q > 0 && t->bottom->n > 1)
      ABX_TRY(predicatePermutation(s, t->bottom, ABX_PRED_POINT3F, pts, q, qperm));
    int64_t const ni = q - nb;
    trace.mark("near_first_order", s);
    if (t->bottom->n == 0)
    {
      // nothing local: every row is padding until candidates arrive
      unsigned long long const m = (unsigned long long)slots;
      ABX_CUDA_TRY(cudaMemcpyAsync(missing.ptr, &m, sizeof(m), cudaMemcpyHostToDevice, s));
    }
    // 1b. the near points on the (high-priority) side stream, the interior ones on the caller's stream, at the
    // same time: the short near launch does not leave the device idle while its last warps finish
    ABX_CUDA_TRY(cudaEventRecord(t->ev[0], s));
    ABX_CUDA_TRY(cudaStreamWaitEvent(x, t->ev[0], 0));
    ABX_TRY(nearestQuery(x, t->bottom, (float const *)pts, nb, k, nullptr, qperm.ptr, nullptr, slots, row_found.ptr,
                         (uint32_t *)rows_p, rowsd_p, missing.ptr, t->rank, /*pad_pairs=*/false));
    ABX_TRY(padShortRows(x, nb, k, row_found.ptr, rows_p, rowsd_p, qperm.ptr));
    trace.mark("near_knn", x);
    ABX_TRY(nearestQuery(s, t->bottom, (float const *)pts, ni, k, nullptr, qperm.ptr + nb, nullptr, slots, row_found.ptr,
                         (uint32_t *)rows_p, rowsd_p, missing.ptr, t->rank, /*pad_pairs=*/false));
    ABX_TRY(padShortRows(s, ni, k, row_found.ptr, rows_p, rowsd_p, qperm.ptr + nb));
    ABX_CUDA_TRY(cudaMemsetAsync(counts.ptr, 0, sizeof(uint32_t) * R, s));
    ABX_TRY(routeLaunch(s, false, ABX_PRED_SPHERE3F, pts, ni, radius, k, t->boxes_dev, R, t->rank, counts.ptr, nullptr,
                        nullptr, nullptr, qperm.ptr + nb));
    trace.mark("interior_knn", s);
    // 2a. the near points' exchange on the side stream (host waits are on that stream only)
    {
      TempBuffer<uint32_t> counts_a, matrix_a;
      ABX_TRY(counts_a.alloc(R, x));
      ABX_TRY(matrix_a.alloc((size_t)R * R, x));
      ABX_CUDA_TRY(cudaMemsetAsync(counts_a.ptr, 0, sizeof(uint32_t) * R, x));
      ABX_TRY(routeLaunch(x, false, ABX_PRED_SPHERE3F, pts, nb, radius, k, t->boxes_dev, R, t->rank, counts_a.ptr,
                          nullptr, nullptr, nullptr, qperm.ptr));
      ABX_TRY(gatherCountMatrix(t, x, counts_a.ptr, matrix_a.ptr));
      ABX_CUDA_TRY(cudaStreamSynchronize(x)); // blocking point 1 (side stream)
      trace.mark("route_count+matrix", x);
      ExchangePlan fwd;
      fwd.fromMatrix(t->h_pin, R, t->rank);
      if (fwd.global > 0)
        ABX_TRY(knnExchangeRound(t, x, pts, q, k, qperm.ptr, nb, radius, fwd, counts_a, matrix_a, rows_p, rowsd_p, round_a,
                                 trace));
      ABX_CUDA_TRY(cudaEventRecord(t->ev[2], x));
    }
    // 2b. the interior points: normally none of them reaches another rank.  The communicator is used from one stream
    // at a time: the caller's stream joins the side stream before its collective.
    ABX_CUDA_TRY(cudaStreamWaitEvent(s, t->ev[2], 0));
    ABX_TRY(gatherCountMatrix(t, s, counts.ptr, matrix.ptr));
    ABX_CUDA_TRY(cudaMemcpyAsync(h_missing, missing.ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 2
    trace.mark("interior_route+matrix", s);
    maybe_short = *h_missing != 0;
    ExchangePlan fwd;
    fwd.fromMatrix(t->h_pin, R, t->rank);
    if (fwd.global > 0)
      ABX_TRY(knnExchangeRound(t, s, pts, q, k, qperm.ptr + nb, ni, radius, fwd, counts, matrix, rows_p, rowsd_p, round_b,
                               trace));
  }
  // buffers taken under the side stream go back to it when this frame unwinds: not before the work on s that reads
  // them (this guard is declared after them, so it runs first)
  struct Rejoin
  {
    abx_dist_tree *t;
    cudaStream_t s, x;
    bool on;
    ~Rejoin()
    {
      if (on && cudaEventRecord(t->ev[0], s) == cudaSuccess)
        cudaStreamWaitEvent(x, t->ev[0], 0);
    }
  } rejoin{t, s, x, two_stage};
  int64_t const M = round_a.M + round_b.M;
  // 6. outputs.  Rows are full (k entries) unless some local row was short and stayed short.
  TempBuffer<int32_t> row_counts, row_off;
  int64_t nnz = slots;
  bool short_rows = false;
  if (maybe_short)
  {
    ABX_TRY(row_counts.alloc((size_t)q + 1, s));
    ABX_TRY(row_off.alloc((size_t)q + 1, s));
    TempBuffer<unsigned long long> total64;
    ABX_TRY(total64.alloc(1, s));
    if (q > 0)
      ABX_LAUNCH(countValidKernel, divUp(q, 256), 256, 0, s, q, k, (int2 const *)rows_p, row_counts.ptr);
    ABX_TRY(exclusiveScanI32(s, row_counts.ptr, row_off.ptr, q + 1, total64.ptr));
    unsigned long long h_total = 0;
    ABX_CUDA_TRY(cudaMemcpyAsync(&h_total, total64.ptr, sizeof(h_total), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    nnz = (int64_t)h_total;
    short_rows = nnz != slots;
  }
  TempBuffer<int32_t> keep_rows;
  TempBuffer<float> keep_d;
  if (!compact && short_rows)
  {
    // the padded rows sit in the caller's arrays, which the allocator is about to replace: set them aside
    ABX_TRY(keep_rows.alloc((size_t)slots * 2, s));
    ABX_TRY(keep_d.alloc((size_t)slots, s));
    ABX_CUDA_TRY(cudaMemcpyAsync(keep_rows.ptr, rows_p, 2 * sizeof(int32_t) * (size_t)slots, cudaMemcpyDeviceToDevice, s));
    ABX_CUDA_TRY(cudaMemcpyAsync(keep_d.ptr, rowsd_p, sizeof(float) * (size_t)slots, cudaMemcpyDeviceToDevice, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s)); // the allocator may release the previous views
    if (!alloc)
    {
      deviceFree(off_v, s);
      deviceFree(vals_v, s);
      deviceFree(d_v, s);
    }
    rows_p = keep_rows.ptr;
    rowsd_p = keep_d.ptr;
    off_v = vals_v = d_v = nullptr;
  }
  if (!off_v)
  {
    ABX_TRY(allocOutDev(alloc, user, 0, sizeof(int32_t) * (size_t)(q + 1), s, &off_v));
    ABX_TRY(allocOutDev(alloc, user, 1, val_bytes * (size_t)nnz, s, &vals_v));
    if (want_dist)
      ABX_TRY(allocOutDev(alloc, user, 2, sizeof(float) * (size_t)nnz, s, &d_v));
  }
  *offsets_out = (int32_t *)off_v;
  *values_out = vals_v;
  if (dist_out)
    *dist_out = (float *)d_v;
  *nnz_out = nnz;
  if (short_rows)
    ABX_CUDA_TRY(cudaMemcpyAsync(off_v, row_off.ptr, sizeof(int32_t) * (size_t)(q + 1), cudaMemcpyDeviceToDevice, s));
  else
    ABX_LAUNCH(fillStrideOffsetsKernel, divUp(q + 1, 256), 256, 0, s, (int32_t *)off_v, q + 1, k);
  // rows -> output values: nothing to do for full rows in the pairs form (they were written in place)
  int2 const *src_rows = (int2 const *)rows_p;
  float const *src_d = rowsd_p;
  TempBuffer<int32_t> packed;
  TempBuffer<float> packed_d;
  if (short_rows && compact)
  {
    ABX_TRY(packed.alloc((size_t)std::max<int64_t>(nnz, 1) * 2, s));
    ABX_TRY(packed_d.alloc((size_t)std::max<int64_t>(nnz, 1), s));
    if (q > 0)
      ABX_LAUNCH(compactPaddedRowsKernel, divUp(q, 256), 256, 0, s, q, k, row_off.ptr, src_rows, src_d,
                 (int2 *)packed.ptr, packed_d.ptr);
    src_rows = (int2 const *)packed.ptr;
    src_d = packed_d.ptr;
  }
  else if (short_rows && q > 0)
    ABX_LAUNCH(compactPaddedRowsKernel, divUp(q, 256), 256, 0, s, q, k, row_off.ptr, src_rows, src_d, (int2 *)vals_v,
               (float *)d_v);
  if (compact && nnz > 0)
  {
    ABX_LAUNCH(splitPairsKernel, divUp(nnz, 256), 256, 0, s, nnz, src_rows, (uint32_t *)vals_v);
    if (want_dist)
      ABX_CUDA_TRY(cudaMemcpyAsync(d_v, src_d, sizeof(float) * (size_t)nnz, cudaMemcpyDeviceToDevice, s));
  }
  if (compact && M > 0)
  {
    // entries owned by other ranks can only sit in rows that received candidates
    TempBuffer<unsigned> counter;
    TempBuffer<uint32_t> pos, rk;
    ABX_TRY(counter.alloc(1, s));
    ABX_TRY(pos.alloc((size_t)M, s));
    ABX_TRY(rk.alloc((size_t)M, s));
    ABX_CUDA_TRY(cudaMemsetAsync(counter.ptr, 0, sizeof(unsigned), s));
    for (KnnRound const *r : {&round_a, &round_b})
      if (r->M > 0)
        ABX_LAUNCH(listRemoteInRowsKernel, divUp(r->M, 256), 256, 0, s, r->M, r->got_ids.ptr, k, (int2 const *)rows_p,
                   short_rows ? row_off.ptr : (int32_t const *)nullptr, t->rank, counter.ptr, pos.ptr, rk.ptr);
    unsigned h_count = 0;
    ABX_CUDA_TRY(cudaMemcpyAsync(&h_count, counter.ptr, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    if (h_count > 0)
    {
      ABX_TRY(sortPairsU32(s, pos.ptr, rk.ptr, h_count, false, bitsFor(std::max<int64_t>(nnz, 2)), /*fixup=*/false));
      ABX_TRY(remote_pos->alloc(h_count, s));
      ABX_TRY(remote_rank->alloc(h_count, s));
      ABX_CUDA_TRY(cudaMemcpyAsync(remote_pos->ptr, pos.ptr, sizeof(int32_t) * h_count, cudaMemcpyDeviceToDevice, s));
      ABX_CUDA_TRY(cudaMemcpyAsync(remote_rank->ptr, rk.ptr, sizeof(int32_t) * h_count, cudaMemcpyDeviceToDevice, s));
    }
    *n_remote = h_count;
  }
  return ABX_OK;
}

// local phase of the distributed kNN: rows of k (index, rank) slots, padded
abx_status localKnnPairs(abx_bvh *bvh, cudaStream_t s, void const *pts, int64_t q, int32_t k, int rank, int32_t *vals2,
                         float *dist, unsigned long long *missing_dev)
{
  int64_t const slots = (int64_t)std::max(k, 0) * q;
  if (slots == 0)
    return ABX_OK;
  if (bvh->n == 0)
  {
    ABX_LAUNCH(fillPaddedRowsKernel, divUp(slots, 256), 256, 0, s, slots, (int2 *)vals2, dist);
    if (missing_dev)
    {
      unsigned long long const m = (unsigned long long)slots;
      ABX_CUDA_TRY(cudaMemcpyAsync(missing_dev, &m, sizeof(m), cudaMemcpyHostToDevice, s));
    }
    return ABX_OK;
  }
  TempBuffer<uint32_t> qperm;
  if (bvh->n > 1)
    ABX_TRY(predicatePermutation(s, bvh, ABX_PRED_POINT3F, pts, q, qperm));
  return nearestQuery(s, bvh, (float const *)pts, q, k, nullptr, qperm.ptr, nullptr, slots, nullptr, (uint32_t *)vals2,
                      dist, missing_dev, rank);
}


// ---- distributed DBSCAN ----------------------------------------------------------------------------------------
// ArborX::Experimental::dbscan(comm, space, primitives, eps, core_min_size, labels, params)
// (cluster/ArborX_DistributedDBSCAN.hpp:29-190, cluster/detail/ArborX_DistributedDBSCANHelpers.hpp):
//   1. halo out   all-gather of the rank boxes; the points within the ghost distance of another rank's box (eps for
//                 core_min_size == 2, nextafter(2 eps) otherwise: a ghost's own core status must be decidable from
//                 what its host sees, DistributedDBSCAN.hpp:77-84) travel to that rank
//   2. local      abx::dbscan on local + ghost points (the hot kernels, unchanged)
//   3. labels     local labels -> global ids (rank offset + index, Helpers.hpp:132-175); ghost labels go back to
//                 their owners, which derive merge pairs (label -> smaller label) for core points that carry several
//                 labels (computeMergePairs, :420-499); all-gather-v of the pairs, sorted and filtered
//                 (sortAndFilterMergePairs, :501-561); every rank flattens its labels through the table (relabel,
//                 :619-660).
// Collectives: two all-gathers of R words, two grouped send/recvs, two all-gathers for the pairs; four blocking
// points.  Everything between them is kernels on the caller's stream.
namespace
{
__global__ void globalLabelsKernel(int64_t n_all, int64_t n_local, int32_t const *__restrict__ local_labels,
                                   int32_t const *__restrict__ ghost_ids, int32_t const *__restrict__ seg_off, int R,
                                   long long const *__restrict__ rank_off, int rank, long long *__restrict__ out)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_all)
    return;
  int const l = local_labels[i];
  long long g = -1;
  if (l >= 0)
  {
    if (l < n_local)
      g = rank_off[rank] + l;
    else
    {
      int const gi = l - (int)n_local; // label = a ghost: its owner's numbering
      int r = 0;
      while (r + 1 < R && seg_off[r + 1] <= gi)
        ++r;
      g = rank_off[r] + ghost_ids[gi];
    }
  }
  out[i] = g;
}
__global__ void flagLabelledKernel(int64_t n, long long const *__restrict__ labels, int32_t *__restrict__ flags)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    flags[i] = labels[i] != -1 ? 1 : 0;
}
__global__ void scatterLabelledKernel(int64_t n, long long const *__restrict__ labels, int32_t const *__restrict__ ids,
                                      int32_t const *__restrict__ pos, long long *__restrict__ out_labels,
                                      int32_t *__restrict__ out_ids)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && labels[i] != -1)
  {
    out_labels[pos[i]] = labels[i];
    out_ids[pos[i]] = ids[i];
  }
}
__global__ void gatherI64Kernel(long long const *__restrict__ src, uint32_t const *__restrict__ perm, int64_t m,
                                long long *__restrict__ dst)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
    dst[i] = src[perm[i]];
}
// computeMergePairs: one thread per point that received labels (ids ascending; the thread at the start of a segment)
__global__ void mergePairsKernel(int64_t m, int32_t const *__restrict__ ids, long long const *__restrict__ got,
                                 int32_t const *__restrict__ core, long long *labels, long long *__restrict__ pairs2,
                                 unsigned *__restrict__ n_pairs)
{
  int64_t const c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m)
    return;
  int const id = ids[c];
  if (c > 0 && ids[c - 1] == id)
    return;
  int64_t e = c;
  long long gmin = got[c];
  for (; e < m && ids[e] == id; ++e)
    gmin = min(gmin, got[e]);
  int64_t const cnt = e - c;
  long long const local = labels[id];
  bool const valid = local != -1;
  if (cnt + (valid ? 1 : 0) < 2)
    return; // a noise point or a point with a single label (Helpers.hpp:449-454)
  if (!core[id])
  {
    if (!valid)
      labels[id] = got[c];
    return;
  }
  long long const min_label = valid ? min(local, gmin) : gmin;
  auto emit = [&](long long from, long long to) {
    unsigned const o = atomicAdd(n_pairs, 1u);
    pairs2[2 * (size_t)o] = from;
    pairs2[2 * (size_t)o + 1] = to;
  };
  if (valid && local != min_label)
    emit(local, min_label);
  if (valid)
    labels[id] = min_label;
  for (int64_t k = c; k < e; ++k)
    if (got[k] != min_label)
      emit(got[k], min_label);
}
__global__ void splitPairsI64Kernel(int64_t m, long long const *__restrict__ pairs2, unsigned long long *__restrict__ from,
                                    unsigned long long *__restrict__ to)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
  {
    from[i] = (unsigned long long)pairs2[2 * i];
    to[i] = (unsigned long long)pairs2[2 * i + 1];
  }
}
// rows sorted by (from, to): keep[i] = 1 for the first of equal rows
__global__ void flagUniqueRowsKernel(int64_t m, unsigned long long const *__restrict__ from,
                                     unsigned long long const *__restrict__ to, int32_t *__restrict__ keep)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
    keep[i] = (i == 0 || from[i] != from[i - 1] || to[i] != to[i - 1]) ? 1 : 0;
}
__global__ void compactRowsI64Kernel(int64_t m, unsigned long long const *__restrict__ from,
                                     unsigned long long const *__restrict__ to, int32_t const *__restrict__ keep,
                                     int32_t const *__restrict__ pos, unsigned long long *__restrict__ out_from,
                                     unsigned long long *__restrict__ out_to)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m && keep[i])
  {
    out_from[pos[i]] = from[i];
    out_to[pos[i]] = to[i];
  }
}
// sortAndFilterMergePairs on unique rows sorted by (from, to): the first row of a `from` keeps its (lowest) `to`,
// every other `to` of that `from` is linked to the lowest one
__global__ void relinkRowsKernel(int64_t m, unsigned long long *__restrict__ from, unsigned long long *__restrict__ to)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m || i == 0 || from[i] != from[i - 1])
    return;
  int64_t first = i - 1;
  while (first > 0 && from[first - 1] == from[i])
    --first;
  // (from, to_i) -> (to_i, lowest): written after every thread has read its neighbours (separate arrays below)
  from[m + i] = to[i];
  to[m + i] = to[first];
}
__global__ void applyRelinkKernel(int64_t m, unsigned long long *__restrict__ from, unsigned long long *__restrict__ to)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  if (from[m + i] != ~0ull)
  {
    from[i] = from[m + i];
    to[i] = to[m + i];
  }
}
// relabel: follow label -> to while the label is a `from` (table sorted by from: the first row of a from holds its
// lowest to)
__global__ void relabelKernel(int64_t n, long long *labels, int64_t m, unsigned long long const *__restrict__ from,
                              unsigned long long const *__restrict__ to)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  long long l = labels[i];
  if (l < 0)
    return;
  for (int guard = 0; guard < 64; ++guard)
  {
    int64_t lo = 0, hi = m; // first row with from >= l
    while (lo < hi)
    {
      int64_t const mid = (lo + hi) >> 1;
      if (from[mid] < (unsigned long long)l)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo >= m || from[lo] != (unsigned long long)l)
      break;
    l = (long long)to[lo];
  }
  labels[i] = l;
}

// rows (from, to) -> sorted by (from, to), duplicates dropped; arrays have room for `m` rows, *m_out rows survive
abx_status sortUniqueRows(cudaStream_t s, int64_t m, TempBuffer<unsigned long long> &from,
                          TempBuffer<unsigned long long> &to, int64_t *m_out)
{
  *m_out = 0;
  if (m == 0)
    return ABX_OK;
  // stable LSD over the two columns: by `to`, then by `from`, carrying a permutation
  TempBuffer<uint32_t> perm;
  TempBuffer<unsigned long long> key, tmp_from, tmp_to;
  ABX_TRY(perm.alloc((size_t)m, s));
  ABX_TRY(key.alloc((size_t)m, s));
  ABX_TRY(tmp_from.alloc((size_t)m, s));
  ABX_TRY(tmp_to.alloc((size_t)m, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(key.ptr, to.ptr, sizeof(unsigned long long) * m, cudaMemcpyDeviceToDevice, s));
  ABX_TRY(sortPairsU64(s, (uint64_t *)key.ptr, perm.ptr, m, true, 64, /*fixup=*/false));
  ABX_LAUNCH(gatherI64Kernel, divUp(m, 256), 256, 0, s, (long long const *)from.ptr, perm.ptr, m, (long long *)tmp_from.ptr);
  ABX_CUDA_TRY(cudaMemcpyAsync(tmp_to.ptr, key.ptr, sizeof(unsigned long long) * m, cudaMemcpyDeviceToDevice, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(key.ptr, tmp_from.ptr, sizeof(unsigned long long) * m, cudaMemcpyDeviceToDevice, s));
  ABX_TRY(sortPairsU64(s, (uint64_t *)key.ptr, perm.ptr, m, true, 64, /*fixup=*/false));
  ABX_LAUNCH(gatherI64Kernel, divUp(m, 256), 256, 0, s, (long long const *)tmp_to.ptr, perm.ptr, m, (long long *)to.ptr);
  ABX_CUDA_TRY(cudaMemcpyAsync(from.ptr, key.ptr, sizeof(unsigned long long) * m, cudaMemcpyDeviceToDevice, s));
  // unique
  TempBuffer<int32_t> keep, pos;
  TempBuffer<unsigned long long> total64;
  ABX_TRY(keep.alloc((size_t)m + 1, s));
  ABX_TRY(pos.alloc((size_t)m + 1, s));
  ABX_TRY(total64.alloc(1, s));
  ABX_LAUNCH(flagUniqueRowsKernel, divUp(m, 256), 256, 0, s, m, from.ptr, to.ptr, keep.ptr);
  ABX_TRY(exclusiveScanI32(s, keep.ptr, pos.ptr, m + 1, total64.ptr));
  unsigned long long h_total = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&h_total, total64.ptr, sizeof(h_total), cudaMemcpyDeviceToHost, s));
  ABX_LAUNCH(compactRowsI64Kernel, divUp(m, 256), 256, 0, s, m, from.ptr, to.ptr, keep.ptr, pos.ptr, tmp_from.ptr,
             tmp_to.ptr);
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  *m_out = (int64_t)h_total;
  ABX_CUDA_TRY(cudaMemcpyAsync(from.ptr, tmp_from.ptr, sizeof(unsigned long long) * h_total, cudaMemcpyDeviceToDevice, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(to.ptr, tmp_to.ptr, sizeof(unsigned long long) * h_total, cudaMemcpyDeviceToDevice, s));
  return ABX_OK;
}

// sortAndFilterMergePairs (Helpers.hpp:501-561): arrays hold m rows and have room for 2 m
abx_status sortAndFilterPairs(cudaStream_t s, int64_t m, TempBuffer<unsigned long long> &from,
                              TempBuffer<unsigned long long> &to, int64_t *m_out)
{
  int64_t u = 0;
  ABX_TRY(sortUniqueRows(s, m, from, to, &u));
  *m_out = u;
  if (u == 0)
    return ABX_OK;
  // the rows to re-link are staged behind the table (slots u .. 2u), marked "none" first
  ABX_CUDA_TRY(cudaMemsetAsync(from.ptr + u, 0xff, sizeof(unsigned long long) * u, s));
  ABX_LAUNCH(relinkRowsKernel, divUp(u, 256), 256, 0, s, u, from.ptr, to.ptr);
  ABX_LAUNCH(applyRelinkKernel, divUp(u, 256), 256, 0, s, u, from.ptr, to.ptr);
  return sortUniqueRows(s, u, from, to, m_out);
}
} // namespace

abx_status distDbscan(abx_comm *comm, cudaStream_t s, float const *xyz, int64_t n, float eps, int minpts, int impl,
                      int algo, long long *labels)
{
  if (!(eps > 0))
  {
    setError("SearchException: dbscan requires eps > 0");
    return ABX_ERR_SEARCH;
  }
  if (minpts < 2)
  {
    setError("SearchException: dbscan requires core_min_size >= 2");
    return ABX_ERR_SEARCH;
  }
  if (n < 0 || n >= (int64_t)1 << 30 || (n > 0 && (!xyz || !labels)))
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  int const R = comm->size, rank = comm->rank;
  if (R > 64)
  {
    setError("distributed dbscan: at most 64 ranks");
    return ABX_ERR_ARG;
  }
  // a throw-away "tree" object carries the communicator scratch the exchange helpers use
  abx_dist_tree t;
  t.comm = comm;
  t.R = R;
  t.rank = rank;
  ABX_TRY(takePinnedScratch(&t.h_pin));
  struct Scratch
  {
    abx_dist_tree &t;
    ~Scratch()
    {
      returnPinnedScratch(t.h_pin);
      t.h_pin = nullptr;
    }
  } scratch{t};

  // 1. rank boxes and sizes
  TempBuffer<unsigned> enc;
  TempBuffer<uint32_t> meta, all;
  TempBuffer<float> boxes_dev;
  ABX_TRY(enc.alloc(6, s));
  ABX_TRY(meta.alloc(8, s));
  ABX_TRY(all.alloc((size_t)8 * R, s));
  ABX_TRY(boxes_dev.alloc((size_t)6 * R, s));
  ABX_TRY(sceneBounds(s, ABX_PRIM_POINT3F, xyz, n, enc.ptr));
  ABX_TRY(decodeBounds(s, enc.ptr, (float *)meta.ptr));
  ABX_CUDA_TRY(cudaMemcpyAsync(meta.ptr + 6, &n, sizeof(int64_t), cudaMemcpyHostToDevice, s));
  ABX_TRY(comm->allGather(meta.ptr, all.ptr, 8 * sizeof(uint32_t), s));
  std::vector<uint32_t> h(8 * (size_t)R);
  ABX_CUDA_TRY(cudaMemcpyAsync(h.data(), all.ptr, sizeof(uint32_t) * 8 * R, cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 1
  std::vector<float> boxes(6 * (size_t)R);
  std::vector<long long> rank_off(R + 1, 0);
  for (int r = 0; r < R; ++r)
  {
    memcpy(&boxes[6 * (size_t)r], &h[8 * (size_t)r], 6 * sizeof(float));
    int64_t sz = 0;
    memcpy(&sz, &h[8 * (size_t)r + 6], sizeof(int64_t));
    rank_off[r + 1] = rank_off[r] + sz;
  }
  ABX_CUDA_TRY(cudaMemcpyAsync(boxes_dev.ptr, boxes.data(), sizeof(float) * 6 * R, cudaMemcpyHostToDevice, s));
  t.boxes_dev = boxes_dev.ptr;

  // 2. halo: points within the ghost distance of another rank's box
  float const e32 = eps;
  float const ghost = minpts == 2 ? e32 : nextafterf(2.f * e32, 10.f * e32);
  TempBuffer<float> radius;
  TempBuffer<uint32_t> counts, matrix;
  ABX_TRY(radius.alloc(1, s));
  ABX_TRY(counts.alloc(R, s));
  ABX_TRY(matrix.alloc((size_t)R * R, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(radius.ptr, &ghost, sizeof(float), cudaMemcpyHostToDevice, s));
  ABX_CUDA_TRY(cudaMemsetAsync(counts.ptr, 0, sizeof(uint32_t) * R, s));
  ABX_TRY(routeLaunch(s, false, ABX_PRED_SPHERE3F, xyz, n, radius.ptr, 0, boxes_dev.ptr, R, rank, counts.ptr, nullptr,
                      nullptr, nullptr));
  ABX_TRY(gatherCountMatrix(&t, s, counts.ptr, matrix.ptr));
  ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 2
  ExchangePlan fwd;
  fwd.fromMatrix(t.h_pin, R, rank);
  TempBuffer<uint32_t> ghost_pts;
  TempBuffer<int32_t> ghost_ids;
  ABX_TRY(forwardPredicates(&t, s, ABX_PRED_SPHERE3F, xyz, 3, n, radius.ptr, 0, fwd, ghost_pts, ghost_ids));
  int64_t const G = fwd.n_recv, n_all = n + G;
  if (n_all >= (int64_t)1 << 30)
  {
    setError("distributed dbscan: local + ghost points exceed 2^30");
    return ABX_ERR_ARG;
  }

  // 3. local DBSCAN on local + ghost points
  TempBuffer<float> unified;
  TempBuffer<int32_t> local_labels, core;
  TempBuffer<long long> glob;
  ABX_TRY(unified.alloc(3 * (size_t)std::max<int64_t>(n_all, 1), s));
  ABX_TRY(local_labels.alloc((size_t)std::max<int64_t>(n_all, 1), s));
  ABX_TRY(core.alloc((size_t)std::max<int64_t>(n_all, 1), s));
  ABX_TRY(glob.alloc((size_t)std::max<int64_t>(n_all, 1), s));
  if (n > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(unified.ptr, xyz, sizeof(float) * 3 * n, cudaMemcpyDeviceToDevice, s));
  if (G > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(unified.ptr + 3 * n, ghost_pts.ptr, sizeof(float) * 3 * G, cudaMemcpyDeviceToDevice, s));
  // The local clustering can fail on one rank alone (the DenseBox precision guard looks at the rank's own bounds,
  // an allocation can fail): the ranks agree on the outcome before the next collective, so that nobody is left
  // waiting in it.
  abx_status const local_st =
      n_all > 0 ? dbscan(s, unified.ptr, n_all, eps, minpts, impl, algo, local_labels.ptr, core.ptr) : ABX_OK;
  {
    TempBuffer<uint32_t> st_dev, st_all;
    ABX_TRY(st_dev.alloc(1, s));
    ABX_TRY(st_all.alloc((size_t)R, s));
    uint32_t const word = (uint32_t)local_st;
    ABX_CUDA_TRY(cudaMemcpyAsync(st_dev.ptr, &word, sizeof(word), cudaMemcpyHostToDevice, s));
    ABX_TRY(comm->allGather(st_dev.ptr, st_all.ptr, sizeof(uint32_t), s));
    ABX_CUDA_TRY(cudaMemcpyAsync(t.h_pin, st_all.ptr, sizeof(uint32_t) * R, cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s));
    if (local_st != ABX_OK)
      return local_st; // this rank's own message stands
    for (int r = 0; r < R; ++r)
      if (t.h_pin[r] != (uint32_t)ABX_OK)
      {
        setError("distributed dbscan: rank " + std::to_string(r) + " failed in its local clustering (status " +
                 std::to_string(t.h_pin[r]) + ")");
        return (abx_status)t.h_pin[r];
      }
  }

  // 4. local -> global labels
  TempBuffer<int32_t> seg;
  TempBuffer<long long> rank_off_dev;
  ABX_TRY(seg.alloc(R + 1, s));
  ABX_TRY(rank_off_dev.alloc(R + 1, s));
  std::vector<int32_t> h_seg(R + 1);
  for (int r = 0; r <= R; ++r)
    h_seg[r] = (int32_t)fwd.recv_off[r];
  ABX_CUDA_TRY(cudaMemcpyAsync(seg.ptr, h_seg.data(), sizeof(int32_t) * (R + 1), cudaMemcpyHostToDevice, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(rank_off_dev.ptr, rank_off.data(), sizeof(long long) * (R + 1), cudaMemcpyHostToDevice, s));
  if (n_all > 0)
    ABX_LAUNCH(globalLabelsKernel, divUp(n_all, 256), 256, 0, s, n_all, n, local_labels.ptr, ghost_ids.ptr, seg.ptr, R,
               rank_off_dev.ptr, rank, glob.ptr);
  if (n > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(labels, glob.ptr, sizeof(long long) * n, cudaMemcpyDeviceToDevice, s));
  if (fwd.global == 0)
    return ABX_OK; // no rank has a ghost: nothing to reconcile (every rank takes this branch)

  // 5. ghost labels back to their owners (noise is not sent); ghosts are grouped by owner already
  TempBuffer<int32_t> flags, pos, back_ids;
  TempBuffer<long long> back_labels;
  ABX_TRY(flags.alloc((size_t)G + 1, s));
  ABX_TRY(pos.alloc((size_t)G + 1, s));
  ABX_TRY(back_ids.alloc((size_t)std::max<int64_t>(G, 1), s));
  ABX_TRY(back_labels.alloc((size_t)std::max<int64_t>(G, 1), s));
  ABX_CUDA_TRY(cudaMemsetAsync(flags.ptr, 0, sizeof(int32_t) * ((size_t)G + 1), s));
  if (G > 0)
    ABX_LAUNCH(flagLabelledKernel, divUp(G, 256), 256, 0, s, G, glob.ptr + n, flags.ptr);
  ABX_TRY(exclusiveScanI32(s, flags.ptr, pos.ptr, G + 1));
  if (G > 0)
    ABX_LAUNCH(scatterLabelledKernel, divUp(G, 256), 256, 0, s, G, glob.ptr + n, ghost_ids.ptr, pos.ptr, back_labels.ptr,
               back_ids.ptr);
  ABX_LAUNCH(segmentTotalsKernel, 1, 64, 0, s, pos.ptr, seg.ptr, R, counts.ptr);
  ABX_TRY(gatherCountMatrix(&t, s, counts.ptr, matrix.ptr));
  ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 3
  ExchangePlan back;
  back.fromMatrix(t.h_pin, R, rank);
  int64_t const M = back.n_recv;
  TempBuffer<int32_t> got_ids;
  TempBuffer<long long> got_labels, got_sorted;
  ABX_TRY(got_ids.alloc((size_t)std::max<int64_t>(M, 1), s));
  ABX_TRY(got_labels.alloc((size_t)std::max<int64_t>(M, 1), s));
  ABX_TRY(got_sorted.alloc((size_t)std::max<int64_t>(M, 1), s));
  ExchangeColumn cols[2] = {{back_labels.ptr, got_labels.ptr, sizeof(long long)}, {back_ids.ptr, got_ids.ptr, sizeof(int32_t)}};
  ABX_TRY(comm->allToAllV(cols, 2, back.send_off.data(), back.recv_off.data(), s));

  // 6. merge pairs of this rank's multi-labelled points
  TempBuffer<long long> pairs2;
  TempBuffer<unsigned> n_pairs;
  ABX_TRY(pairs2.alloc((size_t)std::max<int64_t>(4 * M, 2), s)); // at most one own pair + cnt pairs per point
  ABX_TRY(n_pairs.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(n_pairs.ptr, 0, sizeof(unsigned), s));
  if (M > 0)
  {
    TempBuffer<uint32_t> perm;
    ABX_TRY(perm.alloc((size_t)M, s));
    ABX_TRY(sortPairsU32(s, (uint32_t *)got_ids.ptr, perm.ptr, M, true, bitsFor(std::max<int64_t>(n, 2)), /*fixup=*/false));
    ABX_LAUNCH(gatherI64Kernel, divUp(M, 256), 256, 0, s, got_labels.ptr, perm.ptr, M, got_sorted.ptr);
    ABX_LAUNCH(mergePairsKernel, divUp(M, 256), 256, 0, s, M, got_ids.ptr, got_sorted.ptr, core.ptr, labels, pairs2.ptr,
               n_pairs.ptr);
  }
  unsigned h_pairs = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&h_pairs, n_pairs.ptr, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  int64_t m_local = h_pairs;
  TempBuffer<unsigned long long> from, to;
  ABX_TRY(from.alloc((size_t)std::max<int64_t>(2 * m_local, 2), s));
  ABX_TRY(to.alloc((size_t)std::max<int64_t>(2 * m_local, 2), s));
  if (m_local > 0)
    ABX_LAUNCH(splitPairsI64Kernel, divUp(m_local, 256), 256, 0, s, m_local, pairs2.ptr, from.ptr, to.ptr);
  ABX_TRY(sortAndFilterPairs(s, m_local, from, to, &m_local));

  // 7. all-gather-v of the merge pairs (counts, then rows padded to the longest contribution)
  uint32_t const my_count = (uint32_t)m_local;
  ABX_CUDA_TRY(cudaMemcpyAsync(counts.ptr, &my_count, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  ABX_TRY(comm->allGather(counts.ptr, matrix.ptr, sizeof(uint32_t), s));
  ABX_CUDA_TRY(cudaMemcpyAsync(t.h_pin, matrix.ptr, sizeof(uint32_t) * R, cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s)); // blocking point 4
  std::vector<int64_t> pair_counts(R);
  int64_t mx = 0, total_pairs = 0;
  for (int r = 0; r < R; ++r)
  {
    pair_counts[r] = t.h_pin[r];
    mx = std::max(mx, pair_counts[r]);
    total_pairs += pair_counts[r];
  }
  if (total_pairs == 0)
    return ABX_OK;
  TempBuffer<unsigned long long> send_rows, all_rows, gfrom, gto;
  ABX_TRY(send_rows.alloc((size_t)2 * mx, s));
  ABX_TRY(all_rows.alloc((size_t)2 * mx * R, s));
  ABX_CUDA_TRY(cudaMemsetAsync(send_rows.ptr, 0, sizeof(unsigned long long) * 2 * mx, s));
  if (m_local > 0)
  {
    ABX_CUDA_TRY(cudaMemcpyAsync(send_rows.ptr, from.ptr, sizeof(unsigned long long) * m_local, cudaMemcpyDeviceToDevice, s));
    ABX_CUDA_TRY(cudaMemcpyAsync(send_rows.ptr + mx, to.ptr, sizeof(unsigned long long) * m_local, cudaMemcpyDeviceToDevice, s));
  }
  ABX_TRY(comm->allGather(send_rows.ptr, all_rows.ptr, sizeof(unsigned long long) * 2 * mx, s));
  ABX_TRY(gfrom.alloc((size_t)2 * total_pairs, s));
  ABX_TRY(gto.alloc((size_t)2 * total_pairs, s));
  int64_t o = 0;
  for (int r = 0; r < R; ++r)
  {
    if (pair_counts[r] == 0)
      continue;
    ABX_CUDA_TRY(cudaMemcpyAsync(gfrom.ptr + o, all_rows.ptr + (size_t)2 * mx * r, sizeof(unsigned long long) * pair_counts[r],
                                 cudaMemcpyDeviceToDevice, s));
    ABX_CUDA_TRY(cudaMemcpyAsync(gto.ptr + o, all_rows.ptr + (size_t)2 * mx * r + mx,
                                 sizeof(unsigned long long) * pair_counts[r], cudaMemcpyDeviceToDevice, s));
    o += pair_counts[r];
  }
  int64_t m_global = 0;
  ABX_TRY(sortAndFilterPairs(s, total_pairs, gfrom, gto, &m_global));
  // 8. flatten
  if (n > 0 && m_global > 0)
    ABX_LAUNCH(relabelKernel, divUp(n, 256), 256, 0, s, n, labels, m_global, gfrom.ptr, gto.ptr);
  return ABX_OK;
}

} // namespace abx

// ------------------------------------------------------------------------- C ABI ----
extern "C"
{

abx_status abx_comm_from_nccl(void *nccl_comm, abx_comm **out)
{
  if (!nccl_comm || !out)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  NcclApi *api = ncclApi();
  if (!api)
    return ABX_ERR_CUDA;
  auto c = std::make_unique<NcclComm>();
  c->api = api;
  c->comm = (ncclComm_t)nccl_comm;
  ABX_NCCL_TRY(api, api->CommCount(c->comm, &c->size));
  ABX_NCCL_TRY(api, api->CommUserRank(c->comm, &c->rank));
  *out = c.release();
  return ABX_OK;
}

abx_status abx_comm_unique_id(char id_out[ABX_COMM_UNIQUE_ID_BYTES])
{
  static_assert(sizeof(ncclUniqueId) == ABX_COMM_UNIQUE_ID_BYTES, "unique id size");
  NcclApi *api = ncclApi();
  if (!api)
    return ABX_ERR_CUDA;
  ncclUniqueId id;
  ABX_NCCL_TRY(api, api->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return ABX_OK;
}

abx_status abx_comm_init_rank(const char id[ABX_COMM_UNIQUE_ID_BYTES], int32_t n_ranks, int32_t rank, abx_comm **out)
{
  if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks)
  {
    setError("bad communicator arguments");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  NcclApi *api = ncclApi();
  if (!api)
    return ABX_ERR_CUDA;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  auto c = std::make_unique<NcclComm>();
  c->api = api;
  c->owned = true;
  ABX_NCCL_TRY(api, api->CommInitRank(&c->comm, n_ranks, uid, rank));
  c->size = n_ranks;
  c->rank = rank;
  *out = c.release();
  return ABX_OK;
}

abx_status abx_comm_create_local(int32_t n_ranks, abx_comm **out_array)
{
  if (!out_array || n_ranks < 1 || n_ranks > 64)
  {
    setError("bad communicator arguments");
    return ABX_ERR_ARG;
  }
  auto g = std::make_shared<LocalGroup>();
  g->size = n_ranks;
  g->ptr.assign(n_ranks, nullptr);
  g->cols.assign(n_ranks, nullptr);
  g->off.assign(n_ranks, nullptr);
  for (int r = 0; r < n_ranks; ++r)
  {
    LocalComm *c = new LocalComm;
    c->g = g;
    c->rank = r;
    c->size = n_ranks;
    out_array[r] = c;
  }
  return ABX_OK;
}

abx_status abx_comm_destroy(abx_comm *comm)
{
  delete comm;
  return ABX_OK;
}
int32_t abx_comm_rank(const abx_comm *comm) { return comm ? comm->rank : -1; }
int32_t abx_comm_size(const abx_comm *comm) { return comm ? comm->size : 0; }

static abx_status distCreate(abx_comm *comm, cudaStream_t s, int prim_kind, void const *prims_dev, int64_t n,
                             abx_dist_tree **out)
{
  if (comm->size > 64)
  {
    setError("DistributedTree: at most 64 ranks (the top tree is evaluated as a flat list of rank boxes)");
    return ABX_ERR_ARG;
  }
  std::unique_ptr<abx_dist_tree, abx_status (*)(abx_dist_tree *)> t(new abx_dist_tree, abx_dist_destroy);
  t->comm = comm;
  t->R = comm->size;
  t->rank = comm->rank;
  t->kind = prim_kind;
  int const R = t->R;
  // bottom tree (ArborX_DistributedTree.hpp:183-186)
  ABX_TRY(buildTree(s, prim_kind, prims_dev, n, nullptr, &t->bottom, /*want_wide=*/true));
  // all-gather of (rank box, size) (:208-227, :243-245): 6 floats + a 64-bit count per rank
  TempBuffer<uint32_t> meta, all;
  ABX_TRY(meta.alloc(8, s));
  ABX_TRY(all.alloc((size_t)8 * R, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(meta.ptr, t->bottom->bounds_dev, 6 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  int64_t const n_local = n;
  ABX_CUDA_TRY(cudaMemcpyAsync(meta.ptr + 6, &n_local, sizeof(int64_t), cudaMemcpyHostToDevice, s));
  ABX_TRY(comm->allGather(meta.ptr, all.ptr, 8 * sizeof(uint32_t), s));
  std::vector<uint32_t> h(8 * (size_t)R);
  ABX_CUDA_TRY(cudaMemcpyAsync(h.data(), all.ptr, sizeof(uint32_t) * 8 * R, cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  t->boxes.resize(6 * (size_t)R);
  t->sizes.resize(R);
  t->total = 0;
  for (int d = 0; d < 3; ++d)
  {
    t->bounds[d] = FLT_MAX;
    t->bounds[3 + d] = -FLT_MAX;
  }
  for (int r = 0; r < R; ++r)
  {
    memcpy(&t->boxes[6 * (size_t)r], &h[8 * (size_t)r], 6 * sizeof(float));
    memcpy(&t->sizes[r], &h[8 * (size_t)r + 6], sizeof(int64_t));
    t->total += t->sizes[r];
    if (t->sizes[r] > 0)
      for (int d = 0; d < 3; ++d)
      {
        t->bounds[d] = std::min(t->bounds[d], t->boxes[6 * (size_t)r + d]);
        t->bounds[3 + d] = std::max(t->bounds[3 + d], t->boxes[6 * (size_t)r + 3 + d]);
      }
  }
  ABX_TRY(deviceAlloc((void **)&t->boxes_dev, sizeof(float) * 6 * R, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(t->boxes_dev, t->boxes.data(), sizeof(float) * 6 * R, cudaMemcpyHostToDevice, s));
  ABX_TRY(takePinnedScratch(&t->h_pin));
  ABX_TRY(takeSideStream(&t->side));
  for (auto &e : t->ev)
    ABX_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *out = t.release();
  return ABX_OK;
}

abx_status abx_dist_create(abx_comm *comm, void *stream, int prim_kind, const void *prims_dev, int64_t n,
                           abx_dist_tree **out)
{
  if (!comm || !out)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  return distCreate(comm, (cudaStream_t)stream, prim_kind, prims_dev, n, out);
}

abx_status abx_dist_create_host(abx_comm *comm, void *stream, int prim_kind, const void *prims_host, int64_t n,
                                abx_dist_tree **out)
{
  if (!comm || !out || n < 0 || prim_kind < 0 || prim_kind > ABX_PRIM_TRI3F)
  {
    setError("bad argument");
    return ABX_ERR_ARG;
  }
  *out = nullptr;
  ABX_TRY(ensureDevice());
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> dev;
  ABX_TRY(dev.alloc((size_t)primWords(prim_kind) * (size_t)std::max<int64_t>(n, 1), s));
  if (n > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(dev.ptr, prims_host, sizeof(float) * primWords(prim_kind) * (size_t)n,
                                 cudaMemcpyHostToDevice, s));
  return distCreate(comm, s, prim_kind, dev.ptr, n, out);
}

abx_status abx_dist_destroy(abx_dist_tree *t)
{
  if (!t)
    return ABX_OK;
  if (t->bottom)
  {
    deviceFree(t->boxes_dev, t->bottom->stream);
    abx_bvh_destroy(t->bottom);
  }
  if (t->h_pin)
    returnPinnedScratch(t->h_pin);
  for (auto &e : t->ev)
    if (e)
      cudaEventDestroy(e);
  if (t->side)
    returnSideStream(t->side);
  delete t;
  return ABX_OK;
}

int64_t abx_dist_size(const abx_dist_tree *t) { return t ? t->total : 0; }
int abx_dist_empty(const abx_dist_tree *t) { return !t || t->total == 0; }
abx_status abx_dist_bounds(const abx_dist_tree *t, float out6[6])
{
  if (!t || !out6)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  memcpy(out6, t->bounds, sizeof(t->bounds));
  return ABX_OK;
}

abx_status abx_dist_query_spatial_crs(abx_dist_tree *t, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                      abx_alloc_fn alloc, void *user, int32_t **offsets_dev, int32_t **values2_dev,
                                      int64_t *nnz)
{
  if (!t || !offsets_dev || !values2_dev || !nnz)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  void *vals = nullptr;
  abx_status const st = distSpatial(t, (cudaStream_t)stream, pred_kind, preds_dev, q, false, alloc, user, offsets_dev,
                                    &vals, nnz, nullptr, nullptr, nullptr);
  *values2_dev = (int32_t *)vals;
  return st;
}

abx_status abx_dist_query_nearest_crs(abx_dist_tree *t, void *stream, const void *points_dev, int64_t q, int32_t k,
                                      abx_alloc_fn alloc, void *user, int32_t **offsets_dev, int32_t **values2_dev,
                                      float **distances_dev, int64_t *nnz)
{
  if (!t || !offsets_dev || !values2_dev || !nnz)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  void *vals = nullptr;
  abx_status const st = distNearest(t, (cudaStream_t)stream, points_dev, q, k, false, distances_dev != nullptr, alloc,
                                    user, offsets_dev, &vals, distances_dev, nnz, nullptr, nullptr, nullptr);
  *values2_dev = (int32_t *)vals;
  return st;
}

// host variants: predicates up, compact results down
static abx_status copyOut(abx_alloc_fn alloc_host, void *user, int which, void const *dev, size_t bytes, cudaStream_t s,
                          void **host_out)
{
  *host_out = alloc_host(user, which, bytes);
  if (bytes == 0)
    return ABX_OK;
  if (!*host_out)
  {
    setError("output allocator returned NULL");
    return ABX_ERR_ARG;
  }
  ABX_CUDA_TRY(cudaMemcpyAsync(*host_out, dev, bytes, cudaMemcpyDeviceToHost, s));
  return ABX_OK;
}

abx_status abx_dist_query_spatial_crs_host(abx_dist_tree *t, void *stream, int pred_kind, const void *preds_host,
                                           int64_t q, abx_alloc_fn alloc_host, void *user, int32_t **offsets_host,
                                           uint32_t **indices_host, int64_t *nnz, int32_t **remote_pos_host,
                                           int32_t **remote_rank_host, int64_t *n_remote)
{
  if (!t || !alloc_host || !offsets_host || !indices_host || !nnz || !remote_pos_host || !remote_rank_host || !n_remote ||
      q < 0 || (q > 0 && !preds_host))
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  if (pred_kind != ABX_PRED_SPHERE3F && pred_kind != ABX_PRED_BOX3F && pred_kind != ABX_PRED_POINT3F)
  {
    setError("DistributedTree: spatial predicates are intersects(Sphere | Box | Point)");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int const W = predWords(pred_kind);
  TempBuffer<float> preds;
  ABX_TRY(preds.alloc((size_t)W * (size_t)std::max<int64_t>(q, 1), s));
  if (q > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(preds.ptr, preds_host, sizeof(float) * W * (size_t)q, cudaMemcpyHostToDevice, s));
  int32_t *off = nullptr;
  void *vals = nullptr;
  TempBuffer<int32_t> rpos, rrank;
  abx_status st = distSpatial(t, s, pred_kind, preds.ptr, q, true, nullptr, nullptr, &off, &vals, nnz, &rpos, &rrank,
                              n_remote);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 0, off, sizeof(int32_t) * (size_t)(q + 1), s, (void **)offsets_host);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 1, vals, sizeof(uint32_t) * (size_t)*nnz, s, (void **)indices_host);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 3, rpos.ptr, sizeof(int32_t) * (size_t)*n_remote, s, (void **)remote_pos_host);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 4, rrank.ptr, sizeof(int32_t) * (size_t)*n_remote, s, (void **)remote_rank_host);
  if (st == ABX_OK && cudaStreamSynchronize(s) != cudaSuccess)
  {
    setError("result copy failed");
    st = ABX_ERR_CUDA;
  }
  deviceFree(off, s);
  deviceFree(vals, s);
  return st;
}

abx_status abx_dist_query_nearest_crs_host(abx_dist_tree *t, void *stream, const void *points_host, int64_t q, int32_t k,
                                           abx_alloc_fn alloc_host, void *user, int32_t **offsets_host,
                                           uint32_t **indices_host, float **distances_host, int64_t *nnz,
                                           int32_t **remote_pos_host, int32_t **remote_rank_host, int64_t *n_remote)
{
  if (!t || !alloc_host || !offsets_host || !indices_host || !nnz || !remote_pos_host || !remote_rank_host || !n_remote ||
      q < 0 || (q > 0 && !points_host))
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> pts;
  ABX_TRY(pts.alloc(3 * (size_t)std::max<int64_t>(q, 1), s));
  if (q > 0)
    ABX_CUDA_TRY(cudaMemcpyAsync(pts.ptr, points_host, sizeof(float) * 3 * (size_t)q, cudaMemcpyHostToDevice, s));
  int32_t *off = nullptr;
  void *vals = nullptr;
  float *dist = nullptr;
  TempBuffer<int32_t> rpos, rrank;
  abx_status st = distNearest(t, s, pts.ptr, q, k, true, distances_host != nullptr, nullptr, nullptr, &off, &vals,
                              distances_host ? &dist : nullptr, nnz, &rpos, &rrank, n_remote);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 0, off, sizeof(int32_t) * (size_t)(q + 1), s, (void **)offsets_host);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 1, vals, sizeof(uint32_t) * (size_t)*nnz, s, (void **)indices_host);
  if (st == ABX_OK && distances_host)
    st = copyOut(alloc_host, user, 2, dist, sizeof(float) * (size_t)*nnz, s, (void **)distances_host);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 3, rpos.ptr, sizeof(int32_t) * (size_t)*n_remote, s, (void **)remote_pos_host);
  if (st == ABX_OK)
    st = copyOut(alloc_host, user, 4, rrank.ptr, sizeof(int32_t) * (size_t)*n_remote, s, (void **)remote_rank_host);
  if (st == ABX_OK && cudaStreamSynchronize(s) != cudaSuccess)
  {
    setError("result copy failed");
    st = ABX_ERR_CUDA;
  }
  deviceFree(off, s);
  deviceFree(vals, s);
  deviceFree(dist, s);
  return st;
}

abx_status abx_dist_dbscan_points3f(abx_comm *comm, void *stream, const float *xyz_dev, int64_t n, float eps,
                                    int32_t minpts, int implementation, int algorithm, int64_t *labels_dev)
{
  if (!comm)
  {
    setError("null communicator");
    return ABX_ERR_ARG;
  }
  ABX_TRY(ensureDevice());
  static_assert(sizeof(long long) == sizeof(int64_t), "labels are 64-bit");
  return distDbscan(comm, (cudaStream_t)stream, xyz_dev, n, eps, minpts, implementation, algorithm,
                    (long long *)labels_dev);
}

abx_status abx_dist_nearest_pairs(abx_bvh *bvh, void *stream, const void *points_dev, int64_t q, int32_t k, int32_t rank,
                                  int32_t *values2_dev, float *distances_dev, int64_t *missing_out)
{
  ABX_TRY(ensureDevice());
  if (!bvh || !missing_out || q < 0 || rank < 0)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  *missing_out = 0;
  if (q == 0 || k < 1)
    return ABX_OK;
  if (!points_dev || !values2_dev)
  {
    setError("null argument");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<unsigned long long> missing;
  ABX_TRY(missing.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(missing.ptr, 0, sizeof(unsigned long long), s));
  ABX_TRY(localKnnPairs(bvh, s, points_dev, q, k, rank, values2_dev, distances_dev, missing.ptr));
  unsigned long long h_missing = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(&h_missing, missing.ptr, sizeof(h_missing), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  *missing_out = (int64_t)h_missing;
  return ABX_OK;
}

} // extern "C"
