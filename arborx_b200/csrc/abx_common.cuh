// abx_common.cuh -- shared device/host helpers for libabx (sm_100a only).
//
// Float arithmetic that feeds a comparison uses the __f*_rn intrinsics so that it
// is never contracted into FMAs: results must be bit-identical to the reference's
// unfused float chains (geometry/algorithms/ArborX_Distance.hpp:54-70,
// ArborX_Intersects.hpp:84-114).
#pragma once

#include <atomic>
#include <cfloat>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>

#include <cuda_runtime.h>

#include "../../include/abx.h"

namespace abx
{

// ---------------------------------------------------------------- errors ----
void setError(std::string const &msg);
extern std::atomic<int64_t> g_launch_count;

#define ABX_CUDA_TRY(expr)                                                                                            \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                           \
    if (_e != cudaSuccess)                                                                                             \
    {                                                                                                                  \
      ::abx::setError(std::string(#expr) + ": " + cudaGetErrorString(_e));                                            \
      return ABX_ERR_CUDA;                                                                                             \
    }                                                                                                                  \
  } while (0)

#define ABX_TRY(expr)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    abx_status _s = (expr);                                                                                            \
    if (_s != ABX_OK)                                                                                                  \
      return _s;                                                                                                       \
  } while (0)

// per-kernel device timing (abx_profile_enable / abx_profile_report): CUDA events on
// the launching stream around every launch, aggregated by kernel name
extern bool g_profile;
int profileBegin(char const *name, cudaStream_t s); // returns a record handle (thread-safe)
void profileEnd(int handle, cudaStream_t s);

// every kernel launch goes through this so that launches are counted, checked and
// (optionally) timed
#define ABX_LAUNCH_TAGGED(tag, kernel, grid, block, smem, stream, ...)                                                \
  do                                                                                                                   \
  {                                                                                                                    \
    int const _prof = ::abx::g_profile ? ::abx::profileBegin(tag, stream) : -1;                                        \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                                       \
    if (_prof >= 0)                                                                                                    \
      ::abx::profileEnd(_prof, stream);                                                                                \
    ++::abx::g_launch_count;                                                                                           \
    ABX_CUDA_TRY(cudaGetLastError());                                                                                  \
  } while (0)
#define ABX_LAUNCH(kernel, grid, block, smem, stream, ...)                                                            \
  ABX_LAUNCH_TAGGED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

constexpr int kNumSMs = 148; // B200

// Tuning switches: the release library has none.  Built with -DABX_TUNING (make TUNING=1, a separate
// libabx_tuning.so for scripts/), tuneInt reads the named environment variable once per call site;
// otherwise it is the compile-time default and the alternative kernel instantiations are not compiled.
#ifdef ABX_TUNING
int tuneIntEnv(char const *name, int dflt);
#define ABX_TUNE_INT(name, dflt) ([] { static int const v = ::abx::tuneIntEnv(name, dflt); return v; }())
#else
#define ABX_TUNE_INT(name, dflt) (dflt)
#endif

// "has this (kernel attribute) set-up been done on the current device": one bit per device ordinal
struct PerDeviceOnce
{
  std::atomic<unsigned long long> mask{0};
  bool needed() const
  {
    int dev = 0;
    cudaGetDevice(&dev);
    return !(mask.load(std::memory_order_acquire) & (1ull << (dev & 63)));
  }
  void done()
  {
    int dev = 0;
    cudaGetDevice(&dev);
    mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
};

static inline int divUp(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// stream-ordered temporary buffer (cudaMallocAsync pool, release threshold maxed)
abx_status deviceAlloc(void **p, size_t bytes, cudaStream_t s);
void deviceFree(void *p, cudaStream_t s);

template <class T>
struct TempBuffer
{
  T *ptr = nullptr;
  cudaStream_t stream = nullptr;
  TempBuffer() = default;
  TempBuffer(TempBuffer const &) = delete;
  TempBuffer &operator=(TempBuffer const &) = delete;
  abx_status alloc(size_t count, cudaStream_t s)
  {
    release();
    stream = s;
    return deviceAlloc((void **)&ptr, count * sizeof(T), s);
  }
  void release()
  {
    if (ptr)
      deviceFree(ptr, stream);
    ptr = nullptr;
  }
  T *take()
  {
    T *p = ptr;
    ptr = nullptr;
    return p;
  }
  ~TempBuffer() { release(); }
};

// Measured and rejected for the traversal kernels (B200, 10M points, r01): prefetch.global.L2 of
// the pushed (far) child: kNN 11.2 -> 11.5 ms; prefetch.global.L1 of both children as soon as a
// node arrives: kNN 10.6 -> 11.5 ms, radius 4.8 -> 5.6 ms.  The kernels wait on the dependent
// node-load chain, but extra requests cost more in the L1 pipeline than the overlap returns.

// ------------------------------------------------------------- geometry ----
struct Box
{
  float lo[3];
  float hi[3];
};

__host__ __device__ inline Box emptyBox()
{
  // ArborX_Box.hpp:35-44
  Box b;
  for (int d = 0; d < 3; ++d)
  {
    b.lo[d] = FLT_MAX;
    b.hi[d] = -FLT_MAX;
  }
  return b;
}

__device__ __forceinline__ void boxUnion(Box &a, Box const &b)
{
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    a.lo[d] = fminf(a.lo[d], b.lo[d]);
    a.hi[d] = fmaxf(a.hi[d], b.hi[d]);
  }
}

// distance^2 from point c to box [lo,hi], reference operation order:
// closestPoint (ClosestPoint.hpp:49-66) then sum of squares in index order
// (Distance.hpp:54-70).  tmp = projected - point.
__device__ __forceinline__ float pointBoxDist2(float cx, float cy, float cz, float lx, float ly, float lz, float hx,
                                               float hy, float hz)
{
  // min(max(c, lo), hi) equals the reference's branchy clamp for every valid box; for an "empty" box
  // (lo = +FLT_MAX > hi) both give an infinite distance, which is all any caller compares
  float px = fminf(fmaxf(cx, lx), hx);
  float py = fminf(fmaxf(cy, ly), hy);
  float pz = fminf(fmaxf(cz, lz), hz);
  float tx = __fsub_rn(px, cx), ty = __fsub_rn(py, cy), tz = __fsub_rn(pz, cz);
  float d2 = __fmul_rn(tx, tx);
  d2 = __fadd_rn(d2, __fmul_rn(ty, ty));
  d2 = __fadd_rn(d2, __fmul_rn(tz, tz));
  return d2;
}

// Largest float t with sqrtf(t) <= r, so that (sqrtf(d2) <= r) == (d2 <= t) for
// every non-negative float d2 (sqrtf is correctly rounded and monotone).  Lets the
// traversal loops skip the square root without changing a single result.
__host__ __device__ inline float sqrtThreshold(float r)
{
  if (!(r >= 0.f))
    return -1.f; // sqrt(d2) <= r is false for every d2 >= 0 (also NaN radius)
  if (r > FLT_MAX)
    return r; // +inf radius: every non-NaN d2 qualifies
#ifdef __CUDA_ARCH__
  float t = __fmul_rn(r, r);
  if (t > FLT_MAX)
    t = FLT_MAX;
  // walk to the boundary (a couple of ulps at most)
  while (__fsqrt_rn(t) > r)
    t = __uint_as_float(__float_as_uint(t) - 1u);
  for (;;)
  {
    float u = __uint_as_float(__float_as_uint(t) + 1u);
    if (u <= FLT_MAX && __fsqrt_rn(u) <= r)
      t = u;
    else
      break;
  }
  return t;
#else
  return r * r;
#endif
}

// ------------------------------------------------------------ tree node ----
// Node64: one 64-byte record per internal node k (Karras index, root = 0), holding
// BOTH children's boxes so that one aligned 64-byte load decides both branches:
//   f[0] = (L.lo.xyz, bits(left_ref))   f[1] = (L.hi.xyz, bits(right_ref))
//   f[2] = (R.lo.xyz, bits(range_lo))   f[3] = (R.hi.xyz, bits(range_hi))
// ref >= 0: internal node index; ref < 0: leaf, original index = ~ref.
// [range_lo, range_hi] is the node's range of sorted leaf positions; a leaf left
// child sits at position range_lo, a leaf right child at range_hi.
// The reference layout ({left_child, rope, box}, detail/ArborX_Node.hpp:24-44) is
// derivable from it (abx_bvh_export_reference_layout).
struct Node64
{
  float4 f[4];
};
static_assert(sizeof(Node64) == 64, "Node64 must be 64 bytes");

// 4-wide node of the spatial query kernels: one 64-byte record per internal node of
// the binary tree holding up to four children -- the node's grandchildren, or a child itself when that child is a
// leaf or a subtree of <= 4 leaves (then a "leaf run" of sorted positions).  Child boxes are quantised to 8 bits
// per coordinate against the node's own box and are CONSERVATIVE (decoded box contains the exact one: the encoder
// checks it with the decoder's own arithmetic), so a traversal that tests every reported leaf exactly returns the
// same result set with half the dependent node loads.
//   w[0] = (origin.x, origin.y, origin.z, scale.x)   w[1] = (scale.y, scale.z, q[0..3], q[4..7])
//   w[2] = (q[8..11], q[12..15], q[16..19], q[20..23])   q[6k + d] = min_d, q[6k + 3 + d] = max_d of child k
//   w[3] = child refs: >= 0 wide node (Karras index), kWideEmpty none, else ~((first << 2) | (leaves - 1))
struct Wide64
{
  uint4 w[4];
};
static_assert(sizeof(Wide64) == 64, "Wide64 must be 64 bytes");
constexpr int kWideEmpty = (int)0x80000000;
constexpr int kWideRun = 4; // leaves per leaf run (two bits)
__device__ __forceinline__ float wideByte(unsigned word, int j)
{
  // exact float(byte j of word): 0x4B000000 | b is 2^23 + b
  return __fsub_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7540 | j)), 8388608.0f);
}
__device__ __forceinline__ float wideLo(float q, float scale, float origin) { return __fmaf_rd(q, scale, origin); }
__device__ __forceinline__ float wideHi(float q, float scale, float origin) { return __fmaf_ru(q, scale, origin); }

__device__ __forceinline__ int refLeaf(unsigned orig) { return ~(int)orig; }
__device__ __forceinline__ bool refIsLeaf(int ref) { return ref < 0; }
__device__ __forceinline__ unsigned refOrig(int ref) { return (unsigned)(~ref); }

// ------------------------------------------------------------ Morton ----
// spatial/detail/ArborX_MortonCode.hpp:187-197
__host__ __device__ inline unsigned long long expandBits2_64(unsigned long long x)
{
  x &= 0x1fffffllu;
  x = (x | x << 32) & 0x1f00000000ffffllu;
  x = (x | x << 16) & 0x1f0000ff0000ffllu;
  x = (x | x << 8) & 0x100f00f00f00f00fllu;
  x = (x | x << 4) & 0x10c30c30c30c30c3llu;
  x = (x | x << 2) & 0x1249249249249249llu;
  return x;
}
// :63-73
__host__ __device__ inline unsigned expandBits2_32(unsigned x)
{
  x &= 0x000003ffu;
  x = (x ^ (x << 16)) & 0xff0000ffu;
  x = (x ^ (x << 8)) & 0x0300f00fu;
  x = (x ^ (x << 4)) & 0x030c30c3u;
  x = (x ^ (x << 2)) & 0x09249249u;
  return x;
}

// ordered-uint encoding of floats for atomicMin/atomicMax
__host__ __device__ inline unsigned floatToOrdered(float f)
{
#ifdef __CUDA_ARCH__
  unsigned u = __float_as_uint(f);
#else
  unsigned u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float orderedToFloat(unsigned u)
{
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

__device__ __forceinline__ float4 ldcg4(float4 const *p) { return __ldcg(p); }

} // namespace abx

// ------------------------------------------------------- internal C++ API ----
struct abx_bvh
{
  int kind = 0;
  int64_t n = 0;
  int device = 0;
  cudaStream_t stream = nullptr; // stream the tree was built on
  abx::Node64 *nodes = nullptr;  // n-1 records
  float4 *leaf_box = nullptr;    // sorted leaves: points 1 float4 (xyz, bits(orig)); others 2 float4 (lo,orig)(hi,0)
  float4 *leaf_tri = nullptr;    // triangles only: 3 float4 per sorted leaf (a, b, c)
  uint32_t *perm = nullptr;      // sorted position -> original index
  uint64_t *codes = nullptr;     // sorted Morton64 codes
  bool want_wide = false;        // user-facing tree: the first spatial query writes the 4-wide records
  cudaEvent_t wide_ready = nullptr; // recorded behind the conversion kernel
  cudaStream_t wide_stream = nullptr;
  abx::Wide64 *wide = nullptr;   // 4-wide quantised nodes for the spatial kernels
  unsigned *wide_bad = nullptr;  // device counter: records the converter could not make conservative (non-finite
                                 // boxes); non-zero => the kernels keep the Node64 walk
  float *bounds_dev = nullptr;   // 6 floats, root box (scene bounds)
  float bounds_host[6];
  bool bounds_host_valid = false;
  int64_t bytes = 0;
};

namespace abx
{
// sort.cu
// stable sort of (key, value) pairs by the low key_bits bits of the key (the rest must be zero)
// fixup = false forces plain LSD over all key_bits (keys with long runs of equal top bits, e.g. grid cells)
abx_status sortPairsU64(cudaStream_t s, uint64_t *keys, uint32_t *vals, int64_t n, bool iota_vals, int key_bits = 64,
                        bool fixup = true);
abx_status sortPairsU32(cudaStream_t s, uint32_t *keys, uint32_t *vals, int64_t n, bool iota_vals, int key_bits = 32,
                        bool fixup = true);
// double-buffer forms: input in keys[0]/vals[0], output in keys[*cur]/vals[*cur].  approx_top_bits > 0
// orders by the top approx_top_bits of the key_bits only (an ordering hint, e.g. predicate sorting).
abx_status sortPairsU64DB(cudaStream_t s, uint64_t *const keys[2], uint32_t *const vals[2], int *cur, int64_t n,
                          bool iota_vals, int key_bits);
abx_status sortPairsU32DB(cudaStream_t s, uint32_t *const keys[2], uint32_t *const vals[2], int *cur, int64_t n,
                          bool iota_vals, int key_bits, int approx_top_bits);
abx_status exclusiveScanI32(cudaStream_t s, int32_t const *in, int32_t *out, int64_t n_plus_1,
                            unsigned long long *total64 = nullptr);
// build.cu
abx_status sceneBounds(cudaStream_t s, int kind, void const *prims, int64_t n, unsigned *bounds_enc6);
abx_status decodeBounds(cudaStream_t s, unsigned const *bounds_enc6, float *bounds6);
abx_status morton64(cudaStream_t s, int kind, void const *prims, int64_t n, float const *bounds6, uint64_t *codes);
abx_status morton32(cudaStream_t s, int pred_kind, void const *preds, int64_t q, float const *bounds6, uint32_t *codes);
abx_status buildHierarchy(cudaStream_t s, abx_bvh *bvh, void const *prims);
// want_wide: also write the 4-wide records the spatial query kernels walk (user-facing trees; the trees DBSCAN
// builds for its own kernels do not need them)
abx_status buildTree(cudaStream_t s, int kind, void const *prims, int64_t n, uint64_t const *sorted_codes_or_null,
                     abx_bvh **out, bool want_wide = false);
abx_status ensureWide(cudaStream_t s, abx_bvh *t);
abx_status exportReference(cudaStream_t s, abx_bvh *bvh, int32_t *leaf_rope, uint32_t *leaf_index, int32_t *left_child,
                           int32_t *rope, float *boxes6, uint64_t *codes);
// query.cu
abx_status predicatePermutation(cudaStream_t s, abx_bvh *bvh, int pred_kind, void const *preds, int64_t q,
                                TempBuffer<uint32_t> &perm);
// the same order with the points within `near` of another rank's box first (DistributedTree's two-stage kNN);
// *n_near_dev receives their number
abx_status pointPermutationNearFirst(cudaStream_t s, abx_bvh *bvh, float const *pts, int64_t q, float const *boxes6,
                                     int R, int self_rank, float near, TempBuffer<uint32_t> &perm,
                                     unsigned long long *n_near_dev);
abx_status spatialCount(cudaStream_t s, abx_bvh *bvh, int pred_kind, void const *preds, int64_t q,
                        uint32_t const *qperm, int32_t limit, int32_t *counts);
// out_offsets / pair_rank (also spatialCompact): rows start at out_offsets[i] (lengths still from `offsets`) and
// values are (index, pair_rank) pairs when pair_rank >= 0
abx_status spatialFill(cudaStream_t s, abx_bvh *bvh, int pred_kind, void const *preds, int64_t q,
                       uint32_t const *qperm, int32_t const *offsets, uint32_t *indices,
                       int32_t const *out_offsets = nullptr, int pair_rank = -1);
int spatialStageSlots();
abx_status spatialStage(cudaStream_t s, abx_bvh *bvh, int pred_kind, void const *preds, int64_t q,
                        uint32_t const *qperm, int32_t *counts, uint32_t *staging);
abx_status spatialCompact(cudaStream_t s, abx_bvh *bvh, int pred_kind, void const *preds, int64_t q,
                          uint32_t const *qperm, int32_t const *offsets, uint32_t *indices, uint32_t const *staging,
                          int32_t const *out_offsets = nullptr, int pair_rank = -1);
abx_status nearestQuery(cudaStream_t s, abx_bvh *bvh, float const *pts, int64_t q, int32_t k,
                        int32_t const *k_per_query, uint32_t const *qperm, int32_t const *offsets, int64_t total_rows,
                        int32_t *counts, uint32_t *indices, float *distances,
                        unsigned long long *missing = nullptr, int pair_rank = -1, bool pad_pairs = true);
// (index, rank) rows of k slots: the slots behind counts[i] become (-1, -1) / +inf; ids = the rows to visit (or all)
abx_status padShortRows(cudaStream_t s, int64_t rows, int k, int32_t const *counts, int32_t *vals2, float *dist,
                        uint32_t const *ids = nullptr);
abx_status nearestGeomQuery(cudaStream_t s, abx_bvh *t, int pred_kind, float const *preds, int64_t q, int32_t k,
                            uint32_t const *qperm, int64_t total_rows, int32_t *counts, uint32_t *indices,
                            float *distances, unsigned long long *missing);
abx_status sphereCentres(cudaStream_t s, float const *spheres4, int64_t q, float *pts3);
abx_status sphereDistances(cudaStream_t s, int64_t total, int row, float const *spheres4, float *dist);
abx_status compactRows(cudaStream_t s, int64_t q, int32_t const *old_offsets, int32_t const *new_offsets,
                       uint32_t const *old_idx, float const *old_dist, uint32_t *new_idx, float *new_dist);
abx_status routeLaunch(cudaStream_t s, bool fill, int pred_kind, void const *preds, int64_t q, float const *radius,
                       int64_t radius_stride, float const *boxes6, int R, int self_rank, unsigned *counts,
                       unsigned const *base, unsigned *cursors, int32_t *out_qid, uint32_t const *ids = nullptr);
abx_status mergeSorted(cudaStream_t s, int64_t q, int32_t const *local_off, int32_t const *local_idx, int rank,
                       int64_t m, int32_t const *remote_ids, int32_t const *remote_vals2, int32_t *out_off,
                       int32_t *out_vals2);
abx_status mergeCounts(cudaStream_t s, int64_t q, int32_t const *local_off, int64_t m, int32_t const *remote_ids,
                       int32_t *out_off);
abx_status mergeRemoteRows(cudaStream_t s, int64_t m, int32_t const *remote_ids, int32_t const *remote_vals2,
                           int32_t const *local_off, int32_t const *out_off, int32_t *out_vals2);
abx_status knnMerge(cudaStream_t s, int64_t m, int32_t const *ids, int32_t const *cand2, float const *cand_d, int k,
                    int32_t *vals2, float *dists);
abx_status pairWithRank(cudaStream_t s, int32_t const *indices, int64_t n, int rank, int32_t *out2);
abx_status mergeCrs(cudaStream_t s, int64_t q, int32_t const *local_off, int32_t const *local_idx, int rank,
                    int32_t const *remote_off, int32_t const *remote_vals2, int32_t *out_off, int32_t *out_vals2);
abx_status clipK(cudaStream_t s, int32_t const *k_per_query, int k, int n, int64_t q, int32_t *out);
abx_status halfTraversalPairs(cudaStream_t s, abx_bvh *bvh, float r, uint32_t *pairs, int64_t capacity,
                              unsigned long long *count_dev);
// capi.cu: the CRS drivers (also used by the DistributedTree host code, abx_dist.cu)
struct SpatialCrsCall // state of one spatial CRS query between its two halves; must not move in between
{
  abx_bvh *bvh = nullptr;
  cudaStream_t s = nullptr;
  int pred_kind = 0;
  void const *preds = nullptr;
  int64_t q = 0;
  abx_policy policy;
  abx_alloc_fn alloc = nullptr;
  void *user = nullptr;
  int32_t *offsets = nullptr;
  bool trivial = false, staged = false;
  TempBuffer<uint32_t> qperm, staging;
  TempBuffer<int> overflow;
  TempBuffer<unsigned long long> total64;
  unsigned long long h_total = 0;
  int h_overflow = 0;
  // where the two read-backs land.  The defaults are pageable (a device-to-host cudaMemcpyAsync into pageable
  // memory returns only when the copy is done: Begin then blocks like the one-shot call would anyway); a caller
  // that wants to keep enqueueing other work while the traversal runs points them at PINNED words it owns.
  unsigned long long *total_out = &h_total;
  int *overflow_out = &h_overflow;
  SpatialCrsCall() = default;
  SpatialCrsCall(SpatialCrsCall const &) = delete;
};
abx_status spatialCrsBegin(SpatialCrsCall &c, abx_bvh *bvh, cudaStream_t s, int pred_kind, void const *preds, int64_t q,
                           abx_policy const &policy, abx_alloc_fn alloc, void *user);
abx_status spatialCrsEnd(SpatialCrsCall &c, int32_t **offsets_out, uint32_t **indices_out, int64_t *nnz_out,
                         bool sync_if_trivial = false);
// End in two steps for a caller that places the rows itself: Wait blocks for the number of results (offsets are in
// c.offsets); FillInto writes row i at out_offsets[i] of `values` (4-byte indices, or (index, pair_rank) pairs)
abx_status spatialCrsWait(SpatialCrsCall &c, int64_t *nnz_out);
abx_status spatialCrsFillInto(SpatialCrsCall &c, int64_t nnz, int32_t const *out_offsets, void *values, int pair_rank);
abx_status spatialCrs(abx_bvh *bvh, cudaStream_t s, int pred_kind, void const *preds, int64_t q,
                      abx_policy const &policy, abx_alloc_fn alloc, void *user, int32_t **offsets_out,
                      uint32_t **indices_out, int64_t *nnz_out,
                      std::function<abx_status()> const &before_sync = nullptr);
// pred_kind: ABX_PRED_POINT3F (pts = 3 floats each), ABX_PRED_BOX3F, ABX_PRED_SPHERE3F or ABX_PRED_RAY3F
abx_status nearestCrs(abx_bvh *bvh, cudaStream_t s, void const *pts, int64_t q, int32_t k,
                      int32_t const *k_per_query, abx_policy const &policy, abx_alloc_fn alloc, void *user,
                      int32_t **offsets_out, uint32_t **indices_out, float **distances_out, int64_t *nnz_out,
                      int pred_kind = ABX_PRED_POINT3F);
abx_status ensureDevice();
// dbscan.cu
// core_flags (optional): 1 for core points (all points when minpts == 2)
abx_status dbscan(cudaStream_t s, float const *xyz, int64_t n, float eps, int32_t minpts, int impl, int algo,
                  int32_t *labels, int32_t *core_flags = nullptr);
// mst.cu
// dendrogram_parents != nullptr: BoruvkaMode::HDBSCAN -- edges in (chain, weight) order, parents [2 n - 1] and
// heights [n - 1] of the dendrogram over them
abx_status minimumSpanningTree(cudaStream_t s, float const *xyz, int64_t n, int32_t k, int32_t *edges2, float *weights,
                               int *iterations, int32_t *dendrogram_parents = nullptr,
                               float *dendrogram_heights = nullptr);
abx_status dendrogramUnionFind(cudaStream_t s, int32_t const *edges2, float const *weights, int64_t m, int32_t *parents,
                               float *heights);
} // namespace abx
