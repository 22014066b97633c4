// abx_query.cu -- tree traversals: spatial (count / fill), nearest (kNN), half.
//
// Behavioural contract: spatial/detail/ArborX_TreeTraversal.hpp:34-120 (spatial),
// :122-336 (nearest), spatial/detail/ArborX_HalfTraversal.hpp:24-75 (half).  The
// reference walks {left_child, rope} nodes one box test per step; here one
// 64-byte Node64 load tests both children, leaf boxes live in the parent record, and
// a short per-thread stack replaces the ropes.  The spatial kernels queue the leaves
// they have to test and test them with the warp converged (traverseSpatialDeferred);
// the nearest kernel keeps its candidates in an unsorted shared-memory set and never
// dereferences point leaves.  Result sets are identical; order inside a row is
// traversal order, which the reference does not specify either.
#include "abx_traverse.cuh"

// minimum resident blocks per SM of the traversal kernels (register cap; tuning aid)
#ifndef ABX_NEAREST_MINB
#define ABX_NEAREST_MINB 12 // 40 registers: 10.6 ms at 10M / k = 10 (1: 48 regs 11.2 ms, 16: 32 regs 10.8 ms)
#endif
#ifndef ABX_SPATIAL_MINB
#define ABX_SPATIAL_MINB 1
#endif


namespace abx
{

namespace
{

enum
{
  MODE_COUNT = 0,   // counts only (query with a counting callback, CountUpToN)
  MODE_FILL = 1,    // second traversal writing rows at their CRS offsets
  MODE_STAGE = 2,   // count AND keep the first kStage results of every query in a staging buffer
  MODE_COMPACT = 3  // staged results -> CRS rows; re-traverses only the queries that overflowed
};
constexpr int kPredicateSortBits = 24;
constexpr int kWideDefault = 1; // 1: Wide64 walk in the spatial kernels (r02: stage 4.98 -> 3.76 ms at 10M); 0: Node64
constexpr int kSpatialVariantDefault = 1; // r01: immediate 5.85 ms; deferred (4,12) 4.89, (4,16) 4.99, (4,8) 4.93, (2,16) 5.75
// Staging buffer of the single-traversal CRS path: one 128-byte row of kStage slots per query,
// indexed by the ORIGINAL query id.  CRS rows are in original query order while the traversal
// runs in Morton order, so one side of the hand-over is a scatter; here it is the stage
// kernel's stores (each lane appends to its own row; the kernel is latency-bound and uses a
// tenth of the DRAM bandwidth), and the compaction runs in original order with coalesced row
// reads and fully coalesced CRS writes.  (Slot-major staging with the compaction in Morton order
// cost 1.5 ms at 10M queries: 2.7 GB of DRAM traffic for 0.9 GB of payload, read-for-ownership
// of the partially written CRS sectors.)
constexpr int kStage = 32;

// QCAP > 0: deferred leaf tests (traverseSpatialDeferred) with QCAP queue slots per thread
// WIDE: walk the tree's 4-wide quantised records (Wide64, half the dependent node loads; needs QCAP >= 5) unless
// the converter flagged the tree (*wide_bad != 0: non-finite boxes), in which case the Node64 walk runs instead --
// decided on the device, so no host round trip sits between the build and the first query.
template <int PRED, int MODE, int LEAF_F4, bool TRI, int BUCKET, int QCAP, bool WIDE = false>
__global__ void __launch_bounds__(kThreads, ABX_SPATIAL_MINB)
    spatialKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                  float4 const *__restrict__ leaf_tri, int n, float const *__restrict__ preds, int64_t q,
                  unsigned const *__restrict__ qperm, int limit, int32_t *__restrict__ counts,
                  int32_t const *__restrict__ offsets, uint32_t *__restrict__ indices, uint32_t *__restrict__ staging,
                  Wide64 const *__restrict__ wide, unsigned const *__restrict__ wide_bad,
                  int32_t const *__restrict__ out_offsets, int pair_rank)
{
  // out_offsets / pair_rank (fill and compact forms): rows start at out_offsets[qi] instead of offsets[qi] (their
  // lengths still come from `offsets`), and values are written as (index, pair_rank) pairs -- DistributedTree writes
  // the local rows straight into the merged result this way
  __shared__ unsigned squeue[(QCAP > 0 ? QCAP : 1) * kThreads];
  int64_t const t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  bool active = t < q; // deferred form: the whole warp stays for the converged leaf phase
  if (QCAP == 0 && !active)
    return;
  int64_t const qi = active ? (qperm ? (int64_t)qperm[t] : t) : 0;
  int64_t base = 0;
  if (active && (MODE == MODE_FILL || MODE == MODE_COMPACT))
    base = (int64_t)(out_offsets ? out_offsets[qi] : offsets[qi]);
  auto put = [&](int64_t pos, unsigned orig) {
    if (pair_rank >= 0)
      reinterpret_cast<int2 *>(indices)[pos] = make_int2((int)orig, pair_rank);
    else
      indices[pos] = orig;
  };
  if (active && MODE == MODE_COMPACT)
  {
    int const c = offsets[qi + 1] - offsets[qi];
    if (c <= kStage)
    {
      for (int s = 0; s < c; ++s)
        put(base + s, staging[(size_t)qi * kStage + s]);
      active = false;
      if (QCAP == 0)
        return;
    }
    // overflowed the staging slots: fall through to a filling traversal
  }
  if (QCAP > 0 && MODE == MODE_COMPACT && !__any_sync(0xffffffffu, active))
    return;
  Pred<PRED> pred;
  pred.load(preds, qi);
  int count = 0;
  auto emit = [&](unsigned orig, int pos) {
    if (TRI && !triangleLeafTest<PRED>(pred, leaf_tri, pos))
      return false;
    if (MODE == MODE_FILL || MODE == MODE_COMPACT)
      put(base + count, orig);
    if (MODE == MODE_STAGE && count < kStage)
      staging[(size_t)qi * kStage + count] = orig;
    ++count;
    return limit > 0 && count >= limit;
  };
  bool const had_query = active;
  if (WIDE && __ldg(wide_bad) == 0u)
    traverseWideDeferred<LEAF_F4, (QCAP >= 5 ? QCAP : 5)>(wide, leaf_box, pred, active, squeue, emit);
  else if (QCAP > 0)
    traverseSpatialDeferred<LEAF_F4, (BUCKET >= 1 && BUCKET <= 4 ? BUCKET : 4), (QCAP > 0 ? QCAP : 3)>(nodes, leaf_box, pred, active,
                                                                                        squeue, emit);
  else
    traverseSpatial<LEAF_F4, BUCKET>(nodes, leaf_box, pred, emit);
  if (had_query && (MODE == MODE_COUNT || MODE == MODE_STAGE))
    counts[qi] = count;
}

// n == 1: test the predicate against the single leaf (TreeTraversal.hpp:80-90)
template <int PRED, int MODE>
__global__ void __launch_bounds__(kThreads)
    spatialSingleLeafKernel(float4 const *__restrict__ leaf_box, float4 const *__restrict__ leaf_tri, int prim_kind,
                            float const *__restrict__ preds, int64_t q, int32_t *__restrict__ counts,
                            int32_t const *__restrict__ offsets, uint32_t *__restrict__ indices,
                            int32_t const *__restrict__ out_offsets, int pair_rank)
{
  int64_t const qi = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (qi >= q)
    return;
  Pred<PRED> pred;
  pred.load(preds, qi);
  float4 lo = __ldg(leaf_box);
  float4 hi = prim_kind == ABX_PRIM_POINT3F ? lo : __ldg(leaf_box + 1);
  // triangles: the value itself is tested, not its box (see spatialLaunch)
  bool const hit = prim_kind == ABX_PRIM_TRI3F ? triangleLeafTest<PRED>(pred, leaf_tri, 0) : pred.box(lo, hi);
  if (MODE == MODE_COUNT)
    counts[qi] = hit ? 1 : 0;
  else if (hit)
  {
    int64_t const pos = out_offsets ? out_offsets[qi] : offsets[qi];
    if (pair_rank >= 0)
      reinterpret_cast<int2 *>(indices)[pos] = make_int2(0, pair_rank);
    else
      indices[pos] = 0u;
  }
}

// ---- nearest ---------------------------------------------------------------------
// The traversal works on SQUARED distances and takes the (correctly rounded) root
// only when a row is written.  sqrtf is monotone, so the k smallest squared
// distances give exactly the k smallest distances of the reference
// (TreeTraversal.hpp:180-335, Distance.hpp:54-80): reported distances are
// bit-identical; indices can differ from the reference only among candidates whose
// reported distances are equal -- the ties the reference leaves to its heap order.
__device__ __forceinline__ float pointBoxDist2v(float px, float py, float pz, float4 lo, float4 hi)
{
  return pointBoxDist2(px, py, pz, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
}

// Bounded candidate list, ascending, entirely in registers (K is a compile-time
// constant so every index below is static).  A candidate enters iff d2 < d[K-1]
// (strict, like the reference's `distance < radius`).
template <int K>
struct RegList
{
  float d[K];
  unsigned id[K];
  __device__ __forceinline__ void init()
  {
#pragma unroll
    for (int i = 0; i < K; ++i)
    {
      d[i] = __int_as_float(0x7f800000); // +inf
      id[i] = 0xffffffffu;
    }
  }
  __device__ __forceinline__ float radius() const { return d[K - 1]; }
  __device__ __forceinline__ void insert(float dist, unsigned idx)
  {
    // replace the worst entry, then one bubble pass toward the front; strict <
    // keeps earlier arrivals ahead of later ones at equal distance
    d[K - 1] = dist;
    id[K - 1] = idx;
#pragma unroll
    for (int i = K - 1; i >= 1; --i)
    {
      bool const sw = d[i] < d[i - 1];
      float const lo = sw ? d[i] : d[i - 1], hi = sw ? d[i - 1] : d[i];
      unsigned const ilo = sw ? id[i] : id[i - 1], ihi = sw ? id[i - 1] : id[i];
      d[i - 1] = lo;
      d[i] = hi;
      id[i - 1] = ilo;
      id[i] = ihi;
    }
  }
};

// max-heap of (squared distance, index) in global scratch for large or per-query k
// (reference: NearestBufferProvider.hpp:24-72, misc/ArborX_Heap.hpp)
struct GlobalHeap
{
  float2 *h; // x = squared distance, y = bits(index)
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

// K > 0: register list of exactly K candidates (the first min(k, found) are
// reported; K >= k).  K == 0: global heap with run-time k.
constexpr int kNearestBucket = 1; // 1 = leaves only

// Candidate set of the K > 0 path: K (distance, index) slots per thread in shared memory,
// UNSORTED, plus the position and value of the largest distance in registers.  The traversal
// only ever needs "the k-th smallest so far" and "replace the worst": filling a slot costs two
// stores, a replacement two stores and a K-slot rescan.  A sorted register list pays a K-step
// compare-and-shift for every candidate, and ncu showed that code taking 40 % of the kernel's
// issue slots with 3 of 32 lanes active.  The row is sorted once, at the end, with all lanes
// converged.  Which of several equal largest distances is replaced is arbitrary: like the
// reference's heap this only permutes candidates of equal distance.
// Measured and rejected in round 2 (profiles/r02_validate_wide.log, r02_knn_chunk_experiment.log; 10M / k = 10,
// this kernel 10.6 ms): the 4-wide quantised nodes of the spatial kernels 13.9 ms (decoding four boxes and ranking
// four children costs more than the halved chain saves); a chunked form where a warp owns 64 / 96 / 128 sorted
// queries and a finished lane takes the next one (rows ranked at the end, converged) 14.2 / 15.3 / 15.8 ms.  ncu
// says why: the L1 data pipe is at 85 % of its peak (l1tex__data_pipe_lsu_wavefronts), so what counts is the
// number of distinct cache lines a warp's loads touch -- lanes that walk neighbouring queries in lock step share
// them, lanes that drift apart do not.  Relieving that pipe at the price of instructions does not pay either
// (issue slots are at 73 %): candidate distances in registers instead of shared memory (no re-read of the K slots
// after a replacement, 48 registers) 12.2 ms, the same plus the first 8 stack entries in shared memory 12.8 ms
// (profiles/r02_knn_dreg_sstack_experiment.log).
template <int K, int LEAF_F4, bool TRI>
__global__ void __launch_bounds__(kThreads, (K > 0 && K <= 16) ? ABX_NEAREST_MINB : 1)
    nearestKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box,
                  float4 const *__restrict__ leaf_tri, int n, int prim_kind, float const *__restrict__ pts, int64_t q,
                  unsigned const *__restrict__ qperm, int k_uniform, int row_stride,
                  int32_t const *__restrict__ k_per_query, int32_t const *__restrict__ offsets,
                  int32_t *__restrict__ counts, uint32_t *__restrict__ indices, float *__restrict__ distances,
                  float2 *__restrict__ scratch, unsigned long long *__restrict__ missing, int pair_rank)
{
  int64_t const t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= q)
    return;
  int64_t const qi = qperm ? (int64_t)qperm[t] : t;
  int const k = k_per_query ? k_per_query[qi] : k_uniform;
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
    float const d2 = TRI ? pointTriangleDist2(px, py, pz, __ldg(leaf_tri), __ldg(leaf_tri + 1), __ldg(leaf_tri + 2))
                         : pointBoxDist2v(px, py, pz, lo, hi);
    if (pair_rank >= 0)
    {
      reinterpret_cast<int2 *>(indices)[base] = make_int2(0, pair_rank);
      if (missing && row_stride > 1) // DistributedTree rows have k slots: the rest is padding (padShortRowsKernel)
        atomicAdd(missing, (unsigned long long)(row_stride - 1));
    }
    else
      indices[base] = 0u;
    if (distances)
      distances[base] = __fsqrt_rn(d2);
    if (counts)
      counts[qi] = 1;
    return;
  }

  constexpr bool USE_REGS = K > 0;
  constexpr int KS = USE_REGS ? K : 1;
  __shared__ float set_d[KS * kThreads];
  __shared__ unsigned set_i[KS * kThreads];
  float *const my_d = set_d + threadIdx.x; // slot j at my_d[j * kThreads]: conflict-free across the warp
  unsigned *const my_i = set_i + threadIdx.x;
  int worst = 0; // slot holding the largest distance once the set is full
  GlobalHeap heap;
  heap.h = nullptr;
  heap.size = 0;
  if (!USE_REGS)
    heap.h = scratch + base;
  float radius2 = __int_as_float(0x7f800000);
  int found = 0;

  auto offer = [&](float d2, unsigned idx, int pos) {
    // leaf whose (box) squared distance is < radius2
    if (TRI)
    {
      d2 = pointTriangleDist2(px, py, pz, __ldg(leaf_tri + 3 * (size_t)pos), __ldg(leaf_tri + 3 * (size_t)pos + 1),
                              __ldg(leaf_tri + 3 * (size_t)pos + 2));
      if (!(d2 < radius2))
        return;
    }
    if (USE_REGS)
    {
      // d2 < radius2 here; radius2 stays +inf until K candidates are known
      int const slot = found < KS ? found : worst;
      my_d[slot * kThreads] = d2;
      my_i[slot * kThreads] = idx;
      if (found < KS)
        ++found;
      if (found == KS)
      {
        float m = my_d[0];
        int p = 0;
#pragma unroll
        for (int j = 1; j < KS; ++j)
        {
          float const v = my_d[j * kThreads];
          if (v > m)
          {
            m = v;
            p = j;
          }
        }
        radius2 = m;
        worst = p;
      }
    }
    else
    {
      if (heap.size < k)
        heap.push(d2, idx);
      else
        heap.replaceTop(d2, idx);
      found = heap.size;
      if (found == k)
        radius2 = heap.top();
    }
  };

  // stack of (squared box distance, node) for the farther child
  unsigned long long stack[kStackSize];
  int sp = 0;
  int node = 0;
  while (true)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
    float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
    int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
    int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
    float const dl = pointBoxDist2v(px, py, pz, a0, a1);
    float const dr = pointBoxDist2v(px, py, pz, a2, a3);
    int const l_hi = refIsLeaf(lref) ? rl : lref;
    int const r_lo = refIsLeaf(rref) ? rr : rref;
    // leaves and small subtrees (<= kBucket contiguous sorted leaves) are consumed on
    // the spot, nearer one first; radius2 may shrink between the two
    // (kNN prunes better by descending: sub-boxes reject most of a bucket once the list is
    // full, so only subtrees of <= kNearestBucket leaves are scanned; measured on B200 at
    // 10M / k = 10: bucket 8 = 13.3 ms, bucket 2 = 14.5 ms, leaves only = 11.4 ms)
    bool const l_small = kNearestBucket == 1 ? refIsLeaf(lref) : (l_hi - rl < kNearestBucket);
    bool const r_small = kNearestBucket == 1 ? refIsLeaf(rref) : (rr - r_lo < kNearestBucket);
    auto consume = [&](bool is_leaf, float d, int ref, int lo, int hi) {
      // Triangle leaves are offered with their own distance whatever their box says, as the reference does (its
      // leaves have no box, TreeTraversal.hpp:232-262): the computed closest point can fall an ulp outside the
      // triangle's box, so the box distance is not a safe pre-test for the last bit of the k-th distance.
      if (!(TRI && is_leaf) && !(d < radius2))
        return;
      if (is_leaf)
      {
        offer(d, refOrig(ref), lo);
        return;
      }
      for (int j = lo; j <= hi; ++j)
      {
        float d2;
        unsigned orig;
        if (LEAF_F4 == 1)
        {
          float4 const p = __ldg(leaf_box + j);
          float tx = __fsub_rn(p.x, px), ty = __fsub_rn(p.y, py), tz = __fsub_rn(p.z, pz);
          d2 = __fmul_rn(tx, tx);
          d2 = __fadd_rn(d2, __fmul_rn(ty, ty));
          d2 = __fadd_rn(d2, __fmul_rn(tz, tz));
          orig = __float_as_uint(p.w);
        }
        else
        {
          float4 const l = __ldg(leaf_box + 2 * (size_t)j), h = __ldg(leaf_box + 2 * (size_t)j + 1);
          d2 = pointBoxDist2v(px, py, pz, l, h);
          orig = __float_as_uint(l.w);
        }
        if (d2 < radius2)
          offer(d2, orig, j);
      }
    };
    // one consume site per slot (first / second): lanes whose only candidate is the left child
    // and lanes whose only candidate is the right child run the same insertion code together
    // instead of four serialised inlined copies
    bool const swap = l_small && r_small && dr < dl;
    bool const first_is_left = l_small && !swap;
    if (l_small || r_small)
      consume(first_is_left ? refIsLeaf(lref) : refIsLeaf(rref), first_is_left ? dl : dr, first_is_left ? lref : rref,
              first_is_left ? rl : r_lo, first_is_left ? l_hi : rr);
    if (l_small && r_small)
      consume(swap ? refIsLeaf(lref) : refIsLeaf(rref), swap ? dl : dr, swap ? lref : rref, swap ? rl : r_lo,
              swap ? l_hi : rr);
    bool const go_l = !l_small && dl < radius2;
    bool const go_r = !r_small && dr < radius2;
    if (go_l || go_r)
    {
      // nearer child first; left on ties (TreeTraversal.hpp:310-313)
      bool const left_first = go_l && (dl <= dr || !go_r);
      if (go_l && go_r)
      {
        float const fd = left_first ? dr : dl;
        int const fn = left_first ? rref : lref;
        stack[sp++] = ((unsigned long long)__float_as_uint(fd) << 32) | (unsigned)fn;
      }
      node = left_first ? lref : rref;
      continue;
    }
    // pop until a node that can still contain a closer leaf
    bool popped = false;
    while (sp > 0)
    {
      unsigned long long const e = stack[--sp];
      if (__uint_as_float((unsigned)(e >> 32)) < radius2)
      {
        node = (int)(unsigned)e;
        popped = true;
        break;
      }
    }
    if (!popped)
      break;
  }

  if (USE_REGS)
  {
    // sort the row: insertion into a register list, in slot order (all lanes are here together)
    RegList<KS> list;
    list.init();
#pragma unroll 1
    for (int j = 0; j < found; ++j)
      list.insert(my_d[j * kThreads], my_i[j * kThreads]);
    int const m = min(min(found, k), USE_REGS ? K : 1);
#pragma unroll
    for (int i = 0; i < (USE_REGS ? K : 1); ++i)
      if (i < m)
      {
        // pair_rank >= 0: DistributedTree's (index, rank) values, written in place of the index
        if (pair_rank >= 0)
          reinterpret_cast<int2 *>(indices)[base + i] = make_int2((int)list.id[i], pair_rank);
        else
          indices[base + i] = list.id[i];
        if (distances)
          distances[base + i] = __fsqrt_rn(list.d[i]);
      }
    found = m;
  }
  else
  {
    heap.sortAscending();
    for (int i = 0; i < found; ++i)
    {
      float2 e = heap.h[i];
      if (pair_rank >= 0)
        reinterpret_cast<int2 *>(indices)[base + i] = make_int2((int)__float_as_uint(e.y), pair_rank);
      else
        indices[base + i] = __float_as_uint(e.y);
      if (distances)
        distances[base + i] = __fsqrt_rn(e.x);
    }
  }
  if (counts)
    counts[qi] = found;
  if (missing)
  {
    int const expected = offsets ? (offsets[qi + 1] - (int)base) : row_stride;
    if (found < expected)
      atomicAdd(missing, (unsigned long long)(expected - found));
  }
}

// DistributedTree rows ((index, rank) pairs, k slots per query): slots behind the counts[i] entries found are
// padded with (-1, -1) / +inf so that candidates from other ranks can be merged in place.  Kept out of the
// traversal kernel: short rows are rare (fewer than k reachable leaves) and the walk should not carry the code.
__global__ void padShortRowsKernel(int64_t q, int k, int32_t const *__restrict__ counts, int2 *__restrict__ vals2,
                                   float *__restrict__ dist, uint32_t const *__restrict__ ids)
{
  int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= q)
    return;
  int64_t const i = ids ? (int64_t)ids[t] : t;
  for (int j = counts[i]; j < k; ++j)
  {
    vals2[i * k + j] = make_int2(-1, -1);
    if (dist)
      dist[i * k + j] = __int_as_float(0x7f800000);
  }
}

// ---- nearest(Box, k) / nearest(Ray, k) ---------------------------------------------------------------
// Nearest<Geometry> with a query geometry other than a point (spatial/detail/ArborX_Predicates.hpp:58-80).  The
// walk is the general (global-heap) form of nearestKernel with the predicate geometry's distance:
//   box   distance(Box, Box) (geometry/algorithms/ArborX_Distance.hpp:166-197), kept squared until a row is
//         written; a point leaf is the degenerate box, which gives the same per-axis deltas as
//         distance(Point, Box) (:72-80 through ReverseDispatch)
//   ray   distance(Ray, Box) = max(tmin, 0) if the ray hits the box, else +inf (geometry/ArborX_Ray.hpp:433-444):
//         a length along the ray, not squared; leaves the ray misses are never accepted, so rows can be short
// nearest(Sphere, k) needs no kernel: distance(Sphere, X) = max(distance(centre, X) - r, 0) (:83-108, :199-209)
// orders like the centre's distance, so the host runs the point form and rewrites the distances.
template <int QK>
struct GeomQuery;
template <>
struct GeomQuery<ABX_PRED_BOX3F>
{
  float lo[3], hi[3];
  __device__ __forceinline__ void load(float const *__restrict__ p, int64_t i)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      lo[d] = p[6 * i + d];
      hi[d] = p[6 * i + 3 + d];
    }
  }
  __device__ __forceinline__ float dist(float4 blo, float4 bhi) const
  {
    float const b_lo[3] = {blo.x, blo.y, blo.z}, b_hi[3] = {bhi.x, bhi.y, bhi.z};
    float d2 = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      float delta = 0.f;
      if (lo[d] > b_hi[d])
        delta = __fsub_rn(lo[d], b_hi[d]);
      else if (b_lo[d] > hi[d])
        delta = __fsub_rn(b_lo[d], hi[d]);
      d2 = __fadd_rn(d2, __fmul_rn(delta, delta));
    }
    return d2;
  }
  static __device__ __forceinline__ float report(float d) { return __fsqrt_rn(d); }
};
template <>
struct GeomQuery<ABX_PRED_RAY3F>
{
  Pred<ABX_PRED_RAY3F> ray;
  __device__ __forceinline__ void load(float const *__restrict__ p, int64_t i) { ray.load(p, i); }
  __device__ __forceinline__ float dist(float4 blo, float4 bhi) const { return ray.distance(blo, bhi); }
  static __device__ __forceinline__ float report(float d) { return d; }
};

template <int QK, int LEAF_F4>
__global__ void __launch_bounds__(kThreads)
    nearestGeomKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n,
                      float const *__restrict__ preds, int64_t q, unsigned const *__restrict__ qperm, int k,
                      int row_stride, int32_t *__restrict__ counts, uint32_t *__restrict__ indices,
                      float *__restrict__ distances, float2 *__restrict__ scratch,
                      unsigned long long *__restrict__ missing)
{
  int64_t const t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t >= q)
    return;
  int64_t const qi = qperm ? (int64_t)qperm[t] : t;
  int64_t const base = qi * (int64_t)row_stride;
  GeomQuery<QK> query;
  query.load(preds, qi);
  float const inf = __int_as_float(0x7f800000);
  int found = 0;
  if (n == 1)
  {
    // TreeTraversal.hpp:168-178: the single value is reported unconditionally
    float4 const lo = __ldg(leaf_box);
    float4 const hi = LEAF_F4 == 1 ? lo : __ldg(leaf_box + 1);
    indices[base] = 0u;
    if (distances)
      distances[base] = GeomQuery<QK>::report(query.dist(lo, hi));
    found = 1;
  }
  else
  {
    GlobalHeap heap;
    heap.h = scratch + base;
    heap.size = 0;
    float radius = inf;
    auto offer = [&](float d, unsigned idx) {
      if (heap.size < k)
        heap.push(d, idx);
      else
        heap.replaceTop(d, idx);
      if (heap.size == k)
        radius = heap.top();
    };
    unsigned long long stack[kStackSize];
    int sp = 0;
    int node = 0;
    while (true)
    {
      float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
      float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
      int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
      float const dl = query.dist(a0, a1), dr = query.dist(a2, a3);
      bool const l_leaf = refIsLeaf(lref), r_leaf = refIsLeaf(rref);
      // leaves are consumed on the spot, nearer one first
      bool const swap = l_leaf && r_leaf && dr < dl;
      if (l_leaf && !swap && dl < radius)
        offer(dl, refOrig(lref));
      if (r_leaf && dr < radius)
        offer(dr, refOrig(rref));
      if (swap && dl < radius)
        offer(dl, refOrig(lref));
      bool const go_l = !l_leaf && dl < radius;
      bool const go_r = !r_leaf && dr < radius;
      if (go_l || go_r)
      {
        bool const left_first = go_l && (dl <= dr || !go_r);
        if (go_l && go_r)
          stack[sp++] = ((unsigned long long)__float_as_uint(left_first ? dr : dl) << 32) |
                        (unsigned)(left_first ? rref : lref);
        node = left_first ? lref : rref;
        continue;
      }
      bool popped = false;
      while (sp > 0)
      {
        unsigned long long const e = stack[--sp];
        if (__uint_as_float((unsigned)(e >> 32)) < radius)
        {
          node = (int)(unsigned)e;
          popped = true;
          break;
        }
      }
      if (!popped)
        break;
    }
    heap.sortAscending();
    found = heap.size;
    for (int i = 0; i < found; ++i)
    {
      float2 const e = heap.h[i];
      indices[base + i] = __float_as_uint(e.y);
      if (distances)
        distances[base + i] = GeomQuery<QK>::report(e.x);
    }
  }
  if (counts)
    counts[qi] = found;
  if (missing && found < row_stride)
    atomicAdd(missing, (unsigned long long)(row_stride - found));
}

// distances[i] = max(distances[i] - r(query of i), 0): distance(Sphere, X) from distance(centre, X); rows of
// `row` entries per query
__global__ void sphereDistanceKernel(int64_t total, int row, float const *__restrict__ spheres4, float *__restrict__ dist)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total)
    dist[i] = fmaxf(__fsub_rn(dist[i], spheres4[4 * (i / row) + 3]), 0.f);
}
__global__ void sphereCentresKernel(int64_t q, float const *__restrict__ spheres4, float *__restrict__ pts3)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q)
  {
    pts3[3 * i] = spheres4[4 * i];
    pts3[3 * i + 1] = spheres4[4 * i + 1];
    pts3[3 * i + 2] = spheres4[4 * i + 2];
  }
}

// rows [old_offsets[i], +count_i) -> [new_offsets[i], +count_i), count_i = new_offsets[i+1] - new_offsets[i]
__global__ void compactRowsKernel(int64_t q, int32_t const *__restrict__ old_offsets,
                                  int32_t const *__restrict__ new_offsets, uint32_t const *__restrict__ old_idx,
                                  float const *__restrict__ old_dist, uint32_t *__restrict__ new_idx,
                                  float *__restrict__ new_dist)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q)
    return;
  int const src = old_offsets[i], dst = new_offsets[i], c = new_offsets[i + 1] - dst;
  for (int j = 0; j < c; ++j)
  {
    new_idx[dst + j] = old_idx[src + j];
    if (new_dist)
      new_dist[dst + j] = old_dist[src + j];
  }
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
  traverseHalf(nodes, leaf_box, i, pred, [&](unsigned orig, int) {
    unsigned long long slot = atomicAdd(count, 1ull);
    if (pairs && slot < capacity)
    {
      pairs[2 * slot] = me;
      pairs[2 * slot + 1] = orig;
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
  case ABX_PRED_RAY3F:                                                                                                 \
  {                                                                                                                    \
    constexpr int P = ABX_PRED_RAY3F;                                                                                  \
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
                         uint32_t *indices, uint32_t *staging = nullptr, int32_t const *out_offsets = nullptr,
                         int pair_rank = -1)
{
  if (q <= 0)
    return ABX_OK;
  int const grid = divUp(q, kThreads);
  int const n = (int)t->n;
  char const *tag = MODE == MODE_COUNT   ? "spatialKernel<count>"
                    : MODE == MODE_FILL  ? "spatialKernel<fill>"
                    : MODE == MODE_STAGE ? "spatialKernel<stage>"
                                         : "spatialKernel<compact>";
  if (t->kind == ABX_PRIM_TRI3F && pred_kind != ABX_PRED_SPHERE3F && pred_kind != ABX_PRED_RAY3F)
  {
    setError("only intersects(Sphere) and intersects(Ray) are defined for triangle primitives");
    return ABX_ERR_ARG;
  }
  if (t->kind == ABX_PRIM_POINT3F && pred_kind == ABX_PRED_RAY3F)
  {
    setError("intersects(Ray) is defined for box and triangle primitives");
    return ABX_ERR_ARG;
  }
  if (n == 0)
  {
    if (MODE == MODE_COUNT || MODE == MODE_STAGE)
      ABX_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * q, s));
    return ABX_OK;
  }
  if (n == 1)
  {
    // one leaf: the count and fill forms are all that is needed (no staging)
    constexpr int M1 = (MODE == MODE_COUNT || MODE == MODE_STAGE) ? MODE_COUNT : MODE_FILL;
    ABX_DISPATCH_PRED(pred_kind, ABX_LAUNCH((spatialSingleLeafKernel<P, M1>), grid, kThreads, 0, s, t->leaf_box,
                                            t->leaf_tri, t->kind, (float const *)preds, q, counts, offsets, indices,
                                            out_offsets, pair_rank));
    return ABX_OK;
  }
  // tuning aid: ABX_SPATIAL_VARIANT picks (leaf-run size, deferred-queue slots); 0 slots = immediate leaf tests
  int const variant = ABX_TUNE_INT("ABX_SPATIAL_VARIANT", kSpatialVariantDefault);
  // 4-wide records, written by the first spatial query of a user-facing tree
  bool wide = ABX_TUNE_INT("ABX_WIDE", kWideDefault) != 0 && t->want_wide;
  if (wide)
  {
    ABX_TRY(ensureWide(s, t));
    wide = t->wide != nullptr;
  }
#define ABX_SPATIAL_W(LF4, TRIFLAG)                                                                                   \
  ABX_DISPATCH_PRED(pred_kind,                                                                                         \
                    ABX_LAUNCH_TAGGED(tag, (spatialKernel<P, MODE, LF4, TRIFLAG, 4, 16, true>), grid, kThreads, 0, s,  \
                                      t->nodes, t->leaf_box, t->leaf_tri, n, (float const *)preds, q, qperm, limit,    \
                                      counts, offsets, indices, staging, t->wide, t->wide_bad, out_offsets,      \
                                      pair_rank))
#define ABX_SPATIAL_B(LF4, TRIFLAG, B, QC)                                                                            \
  ABX_DISPATCH_PRED(pred_kind, ABX_LAUNCH_TAGGED(tag, (spatialKernel<P, MODE, LF4, TRIFLAG, B, QC>), grid, kThreads,  \
                                                 0, s, t->nodes, t->leaf_box, t->leaf_tri, n, (float const *)preds,   \
                                                 q, qperm, limit, counts, offsets, indices, staging,                   \
                                                 (Wide64 const *)nullptr, (unsigned const *)nullptr, out_offsets,     \
                                                 pair_rank))
#ifdef ABX_TUNING
#define ABX_SPATIAL_VARIANTS(LF4, TRIFLAG)                                                                            \
  switch (variant)                                                                                                     \
  {                                                                                                                    \
  case 1: ABX_SPATIAL_B(LF4, TRIFLAG, 4, 12); break;                                                                   \
  case 2: ABX_SPATIAL_B(LF4, TRIFLAG, 4, 16); break;                                                                   \
  case 3: ABX_SPATIAL_B(LF4, TRIFLAG, 2, 16); break;                                                                   \
  case 4: ABX_SPATIAL_B(LF4, TRIFLAG, 4, 8); break;                                                                    \
  default: ABX_SPATIAL_B(LF4, TRIFLAG, 4, 0); break;                                                                   \
  }
#else
#define ABX_SPATIAL_VARIANTS(LF4, TRIFLAG)                                                                            \
  (void)variant;                                                                                                       \
  ABX_SPATIAL_B(LF4, TRIFLAG, 4, 12)
#endif
#define ABX_SPATIAL(LF4, TRIFLAG)                                                                                     \
  do                                                                                                                   \
  {                                                                                                                    \
    if (wide)                                                                                                          \
    {                                                                                                                  \
      ABX_SPATIAL_W(LF4, TRIFLAG);                                                                                     \
      break;                                                                                                           \
    }                                                                                                                  \
    ABX_SPATIAL_VARIANTS(LF4, TRIFLAG);                                                                                \
  } while (0)
  if (t->kind == ABX_PRIM_TRI3F)
  {
    // Triangles: the exact leaf test is not bounded by the leaf's box (the ray - triangle test accepts hits within
    // its tolerances outside the triangle, ArborX_Ray.hpp:340-355: 4 % of the rays of the 20M-triangle icosphere
    // gain a neighbouring triangle that way; the sphere - triangle distance can differ from the box distance in
    // the last place), so the reference's rule is kept to the letter: a leaf is tested whenever its parent is
    // visited.  That rules out leaf runs, the deferred queue and the quantised wide records for this combination.
    ABX_SPATIAL_B(2, true, 0, 0);
  }
  else if (t->kind == ABX_PRIM_BOX3F)
  {
    ABX_SPATIAL(2, false);
  }
  else
  {
    ABX_SPATIAL(1, false);
  }
#undef ABX_SPATIAL
#undef ABX_SPATIAL_VARIANTS
#undef ABX_SPATIAL_B
#undef ABX_SPATIAL_W
  return ABX_OK;
}

} // namespace

// Morton32 permutation of the predicates (CrsGraphWrapperImpl.hpp:407-419)
abx_status predicatePermutation(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q,
                                TempBuffer<uint32_t> &perm)
{
  // The permutation only has to make neighbouring threads traverse neighbouring subtrees; the
  // results do not depend on it.  Morton32 codes use 30 bits; ordering by the top
  // kPredicateSortBits of them (ABX_QUERY_SORT_BITS overrides) costs one digit pass less than the
  // reference's full sort and leaves cells far smaller than a warp's worth of queries unordered.
  int const sort_bits = ABX_TUNE_INT("ABX_QUERY_SORT_BITS", kPredicateSortBits);
  TempBuffer<uint32_t> codes, codes_alt, perm_alt;
  ABX_TRY(codes.alloc(q, s));
  ABX_TRY(codes_alt.alloc(q, s));
  ABX_TRY(perm.alloc(q, s));
  ABX_TRY(perm_alt.alloc(q, s));
  ABX_TRY(morton32(s, pred_kind, preds, q, t->bounds_dev, codes.ptr));
  uint32_t *kb[2] = {codes.ptr, codes_alt.ptr};
  uint32_t *vb[2] = {perm.ptr, perm_alt.ptr};
  int cur = 0;
  ABX_TRY(sortPairsU32DB(s, kb, vb, &cur, q, true, 30, sort_bits >= 30 ? 0 : sort_bits));
  if (cur != 0)
    std::swap(perm.ptr, perm_alt.ptr);
  return ABX_OK;
}

// DistributedTree's two-stage kNN: the Morton-ordered permutation is split (stably) into the points within `near`
// of another rank's box and the rest, so both groups keep the order that makes neighbouring threads walk
// neighbouring subtrees.
__global__ void __launch_bounds__(256)
    nearFlagsKernel(float const *__restrict__ pts, uint32_t const *__restrict__ perm, int64_t q,
                    float const *__restrict__ boxes6, int R, int self_rank, float near2, int32_t *__restrict__ flags)
{
  __shared__ float sbox[64 * 6];
  for (int i = threadIdx.x; i < R * 6; i += blockDim.x)
    sbox[i] = boxes6[i];
  __syncthreads();
  int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= q)
    return;
  int64_t const i = perm ? (int64_t)perm[t] : t;
  float const c[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  bool is_near = false;
  for (int rk = 0; rk < R && !is_near; ++rk)
  {
    if (rk == self_rank)
      continue;
    float const *b = sbox + 6 * rk;
    if (b[0] > b[3] || b[1] > b[4] || b[2] > b[5])
      continue;
    float d2 = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      float const p = fminf(fmaxf(c[d], b[d]), b[3 + d]) - c[d];
      d2 += p * p;
    }
    is_near = !(d2 > near2); // NaN coordinates count as near
  }
  flags[t] = is_near ? 1 : 0;
}
__global__ void splitByFlagKernel(uint32_t const *__restrict__ perm, int64_t q, int32_t const *__restrict__ flags,
                                  int32_t const *__restrict__ pos /* exclusive scan of flags, q + 1 */,
                                  uint32_t *__restrict__ out)
{
  int64_t const t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= q)
    return;
  uint32_t const v = perm ? perm[t] : (uint32_t)t;
  int32_t const before = pos[t];
  int64_t const dst = flags[t] ? (int64_t)before : (int64_t)pos[q] + (t - before);
  out[dst] = v;
}

abx_status pointPermutationNearFirst(cudaStream_t s, abx_bvh *t, float const *pts, int64_t q, float const *boxes6, int R,
                                     int self_rank, float near, TempBuffer<uint32_t> &perm,
                                     unsigned long long *n_near_dev)
{
  TempBuffer<uint32_t> sorted;
  if (t->n > 1)
    ABX_TRY(predicatePermutation(s, t, ABX_PRED_POINT3F, pts, q, sorted));
  TempBuffer<int32_t> flags, pos;
  ABX_TRY(flags.alloc((size_t)q + 1, s));
  ABX_TRY(pos.alloc((size_t)q + 1, s));
  ABX_TRY(perm.alloc((size_t)q, s));
  ABX_LAUNCH(nearFlagsKernel, divUp(q, 256), 256, 0, s, pts, sorted.ptr, q, boxes6, R, self_rank, near * near, flags.ptr);
  ABX_TRY(exclusiveScanI32(s, flags.ptr, pos.ptr, q + 1, n_near_dev));
  ABX_LAUNCH(splitByFlagKernel, divUp(q, 256), 256, 0, s, sorted.ptr, q, flags.ptr, pos.ptr, perm.ptr);
  return ABX_OK;
}

abx_status spatialCount(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q, uint32_t const *qperm,
                        int32_t limit, int32_t *counts)
{
  return spatialLaunch<MODE_COUNT>(s, t, pred_kind, preds, q, qperm, limit, counts, nullptr, nullptr);
}

abx_status spatialFill(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q, uint32_t const *qperm,
                       int32_t const *offsets, uint32_t *indices, int32_t const *out_offsets, int pair_rank)
{
  return spatialLaunch<MODE_FILL>(s, t, pred_kind, preds, q, qperm, 0, nullptr, offsets, indices, nullptr, out_offsets,
                                  pair_rank);
}

// single-traversal CRS: stage (count + keep first kStage results) ... scan ... compact
int spatialStageSlots() { return kStage; }
abx_status spatialStage(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q, uint32_t const *qperm,
                        int32_t *counts, uint32_t *staging)
{
  return spatialLaunch<MODE_STAGE>(s, t, pred_kind, preds, q, qperm, 0, counts, nullptr, nullptr, staging);
}
abx_status spatialCompact(cudaStream_t s, abx_bvh *t, int pred_kind, void const *preds, int64_t q,
                          uint32_t const *qperm, int32_t const *offsets, uint32_t *indices, uint32_t const *staging,
                          int32_t const *out_offsets, int pair_rank)
{
  return spatialLaunch<MODE_COMPACT>(s, t, pred_kind, preds, q, qperm, 0, nullptr, offsets, indices,
                                     const_cast<uint32_t *>(staging), out_offsets, pair_rank);
}

// uniform k: offsets == nullptr and rows start at i * min(k, n); per-query k:
// offsets = CRS offsets of min(k_i, n).  total_rows = size of indices.
abx_status nearestQuery(cudaStream_t s, abx_bvh *t, float const *pts, int64_t q, int32_t k, int32_t const *k_per_query,
                        uint32_t const *qperm, int32_t const *offsets, int64_t total_rows, int32_t *counts,
                        uint32_t *indices, float *distances, unsigned long long *missing, int pair_rank, bool pad_pairs)
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
  // DistributedTree rows (pair_rank >= 0) always have k slots: short rows are padded
  int const row_stride = pair_rank >= 0 ? std::max(0, k) : std::max(0, std::min(k, n));
  TempBuffer<int32_t> counts_tmp;
  if (pair_rank >= 0 && !counts)
  {
    ABX_TRY(counts_tmp.alloc((size_t)q, s));
    counts = counts_tmp.ptr;
  }
#define ABX_NEAREST(KCAP, SCRATCH)                                                                                    \
  do                                                                                                                   \
  {                                                                                                                    \
    if (tri)                                                                                                           \
      ABX_LAUNCH_TAGGED("nearestKernel<" #KCAP ",tri>", (nearestKernel<KCAP, 2, true>), grid, kThreads, 0, s,          \
                        t->nodes, t->leaf_box, t->leaf_tri, n, t->kind, pts, q, qperm, k, row_stride, k_per_query,     \
                        offsets, counts, indices, distances, SCRATCH, missing, pair_rank);                                                 \
    else if (t->kind == ABX_PRIM_BOX3F)                                                                                \
      ABX_LAUNCH_TAGGED("nearestKernel<" #KCAP ",box>", (nearestKernel<KCAP, 2, false>), grid, kThreads, 0, s,         \
                        t->nodes, t->leaf_box, t->leaf_tri, n, t->kind, pts, q, qperm, k, row_stride, k_per_query,     \
                        offsets, counts, indices, distances, SCRATCH, missing, pair_rank);                                                 \
    else                                                                                                               \
      ABX_LAUNCH_TAGGED("nearestKernel<" #KCAP ">", (nearestKernel<KCAP, 1, false>), grid, kThreads, 0, s, t->nodes,   \
                        t->leaf_box, t->leaf_tri, n, t->kind, pts, q, qperm, k, row_stride, k_per_query, offsets,      \
                        counts, indices, distances, SCRATCH, missing, pair_rank);                                                          \
  } while (0)
  if (kmax <= 16)
  {
    // exact-k register lists
    switch (std::max(kmax, 1))
    {
    case 1: ABX_NEAREST(1, nullptr); break;
    case 2: ABX_NEAREST(2, nullptr); break;
    case 3: ABX_NEAREST(3, nullptr); break;
    case 4: ABX_NEAREST(4, nullptr); break;
    case 5: ABX_NEAREST(5, nullptr); break;
    case 6: ABX_NEAREST(6, nullptr); break;
    case 7: ABX_NEAREST(7, nullptr); break;
    case 8: ABX_NEAREST(8, nullptr); break;
    case 9: ABX_NEAREST(9, nullptr); break;
    case 10: ABX_NEAREST(10, nullptr); break;
    case 11: ABX_NEAREST(11, nullptr); break;
    case 12: ABX_NEAREST(12, nullptr); break;
    case 13: ABX_NEAREST(13, nullptr); break;
    case 14: ABX_NEAREST(14, nullptr); break;
    case 15: ABX_NEAREST(15, nullptr); break;
    default: ABX_NEAREST(16, nullptr); break;
    }
  }
  else if (kmax <= 24)
    ABX_NEAREST(24, nullptr); // keeps the 24 nearest, reports the first k
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
  if (pair_rank >= 0 && row_stride > 0 && pad_pairs)
    ABX_LAUNCH(padShortRowsKernel, divUp(q, 256), 256, 0, s, q, row_stride, counts, (int2 *)indices, distances,
               (uint32_t const *)nullptr);
  return ABX_OK;
}

abx_status padShortRows(cudaStream_t s, int64_t rows, int k, int32_t const *counts, int32_t *vals2, float *dist,
                        uint32_t const *ids)
{
  if (rows > 0 && k > 0)
    ABX_LAUNCH(padShortRowsKernel, divUp(rows, 256), 256, 0, s, rows, k, counts, (int2 *)vals2, dist, ids);
  return ABX_OK;
}

// nearest(Box | Ray, k): rows of min(k, n) slots, counts per row (rows can be short: leaves a ray misses)
abx_status nearestGeomQuery(cudaStream_t s, abx_bvh *t, int pred_kind, float const *preds, int64_t q, int32_t k,
                            uint32_t const *qperm, int64_t total_rows, int32_t *counts, uint32_t *indices,
                            float *distances, unsigned long long *missing)
{
  if (q <= 0)
    return ABX_OK;
  int const n = (int)t->n;
  if (t->kind == ABX_PRIM_TRI3F)
  {
    setError("nearest(Box | Ray, k) is defined for point and box primitives");
    return ABX_ERR_ARG;
  }
  if (n == 0)
  {
    if (counts)
      ABX_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * q, s));
    return ABX_OK;
  }
  int const grid = divUp(q, kThreads);
  int const row_stride = std::max(0, std::min(k, n));
  TempBuffer<float2> scratch;
  ABX_TRY(scratch.alloc((size_t)std::max<int64_t>(total_rows, 1), s));
  bool const boxes = t->kind == ABX_PRIM_BOX3F;
#define ABX_GEOM(QK)                                                                                                  \
  do                                                                                                                   \
  {                                                                                                                    \
    if (boxes)                                                                                                         \
      ABX_LAUNCH((nearestGeomKernel<QK, 2>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n, preds, q, qperm, k,       \
                 row_stride, counts, indices, distances, scratch.ptr, missing);                                        \
    else                                                                                                               \
      ABX_LAUNCH((nearestGeomKernel<QK, 1>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n, preds, q, qperm, k,       \
                 row_stride, counts, indices, distances, scratch.ptr, missing);                                        \
  } while (0)
  if (pred_kind == ABX_PRED_BOX3F)
    ABX_GEOM(ABX_PRED_BOX3F);
  else if (pred_kind == ABX_PRED_RAY3F)
    ABX_GEOM(ABX_PRED_RAY3F);
  else
  {
    setError("nearestGeomQuery: box or ray predicates");
    return ABX_ERR_ARG;
  }
#undef ABX_GEOM
  return ABX_OK;
}

abx_status sphereCentres(cudaStream_t s, float const *spheres4, int64_t q, float *pts3)
{
  if (q > 0)
    ABX_LAUNCH(sphereCentresKernel, divUp(q, 256), 256, 0, s, q, spheres4, pts3);
  return ABX_OK;
}
abx_status sphereDistances(cudaStream_t s, int64_t total, int row, float const *spheres4, float *dist)
{
  if (total > 0 && row > 0)
    ABX_LAUNCH(sphereDistanceKernel, divUp(total, 256), 256, 0, s, total, row, spheres4, dist);
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

// DistributedTree: merge the CRS rows of local results (indices) and of remote results
// ((index, rank) pairs) into one CRS of (index, rank) pairs (countResults + sort by query id,
// distributed/detail/ArborX_DistributedTreeUtils.hpp:229-263, without a sort: both inputs are
// already grouped by query)
__global__ void mergeCountsKernel(int64_t q, int32_t const *__restrict__ local_off,
                                  int32_t const *__restrict__ remote_off, int32_t *__restrict__ counts)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q)
    counts[i] = (local_off[i + 1] - local_off[i]) + (remote_off[i + 1] - remote_off[i]);
}
__global__ void mergeRowsKernel(int64_t q, int32_t const *__restrict__ local_off, int32_t const *__restrict__ local_idx,
                                int rank, int32_t const *__restrict__ remote_off,
                                int2 const *__restrict__ remote_vals, int32_t const *__restrict__ out_off,
                                int2 *__restrict__ out_vals)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q)
    return;
  int dst = out_off[i];
  for (int j = local_off[i]; j < local_off[i + 1]; ++j)
    out_vals[dst++] = make_int2(local_idx[j], rank);
  for (int j = remote_off[i]; j < remote_off[i + 1]; ++j)
    out_vals[dst++] = remote_vals[j];
}

// DistributedTree routing: which OTHER ranks' boxes may satisfy each predicate.  The test is
// conservative for spheres (a rank that receives a predicate it has nothing for returns
// nothing), exact comparisons for boxes / points.  Two passes: per-destination counts, then the
// query ids grouped by destination (the send buffer order of the all-to-all-v).
template <int PRED, bool FILL>
__global__ void __launch_bounds__(256)
    routeKernel(float const *__restrict__ preds, int64_t q, float const *__restrict__ radius, int64_t radius_stride,
                float const *__restrict__ boxes6, int R, int self_rank, unsigned *__restrict__ counts /*[R]*/,
                unsigned const *__restrict__ base /*[R]*/, unsigned *__restrict__ cursors /*[R]*/,
                int32_t *__restrict__ out_qid, uint32_t const *__restrict__ ids /* q predicates to visit, or null */)
{
  __shared__ float sbox[64 * 6];
  __shared__ unsigned scount[64];
  for (int i = threadIdx.x; i < R * 6; i += blockDim.x)
    sbox[i] = boxes6[i];
  if (threadIdx.x < 64)
    scount[threadIdx.x] = 0;
  __syncthreads();
  int64_t const slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < q)
  {
    int64_t const i = ids ? (int64_t)ids[slot] : slot;
    float c[3], r2 = 0.f, hi3[3];
    if (PRED == ABX_PRED_SPHERE3F)
    {
      // spheres either packed (centre, radius) or points + a separate (strided) radius array
      int const st = radius ? 3 : 4;
      c[0] = preds[st * i], c[1] = preds[st * i + 1], c[2] = preds[st * i + 2];
      float const r = radius ? radius[i * radius_stride] : preds[4 * i + 3];
      r2 = r * r * 1.0001f + 1e-30f;
    }
    else if (PRED == ABX_PRED_BOX3F)
    {
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        c[d] = preds[6 * i + d];
        hi3[d] = preds[6 * i + 3 + d];
      }
    }
    else
      c[0] = preds[3 * i], c[1] = preds[3 * i + 1], c[2] = preds[3 * i + 2];
    for (int rk = 0; rk < R; ++rk)
    {
      if (rk == self_rank)
        continue;
      float const *b = sbox + 6 * rk;
      if (b[0] > b[3] || b[1] > b[4] || b[2] > b[5])
        continue; // rank without primitives
      bool hit;
      if (PRED == ABX_PRED_BOX3F)
        hit = !(c[0] > b[3] || hi3[0] < b[0] || c[1] > b[4] || hi3[1] < b[1] || c[2] > b[5] || hi3[2] < b[2]);
      else
      {
        float d2 = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          float const p = fminf(fmaxf(c[d], b[d]), b[3 + d]) - c[d];
          d2 += p * p;
        }
        hit = PRED == ABX_PRED_SPHERE3F ? (d2 <= r2 || !(r2 < __int_as_float(0x7f800000))) : (d2 == 0.f);
      }
      if (hit)
      {
        if (FILL)
          out_qid[base[rk] + atomicAdd(&cursors[rk], 1u)] = (int32_t)i;
        else
          atomicAdd(&scount[rk], 1u);
      }
    }
  }
  if (!FILL)
  {
    __syncthreads();
    if (threadIdx.x < R && scount[threadIdx.x])
      atomicAdd(&counts[threadIdx.x], scount[threadIdx.x]);
  }
}

abx_status routeLaunch(cudaStream_t s, bool fill, int pred_kind, void const *preds, int64_t q, float const *radius,
                       int64_t radius_stride, float const *boxes6, int R, int self_rank, unsigned *counts,
                       unsigned const *base, unsigned *cursors, int32_t *out_qid, uint32_t const *ids)
{
  if (R > 64)
  {
    setError("route: at most 64 ranks");
    return ABX_ERR_ARG;
  }
  if (q <= 0)
    return ABX_OK;
  int const grid = divUp(q, 256);
#define ABX_ROUTE(P)                                                                                                  \
  do                                                                                                                   \
  {                                                                                                                    \
    if (fill)                                                                                                          \
      ABX_LAUNCH_TAGGED("routeKernel<fill>", (routeKernel<P, true>), grid, 256, 0, s, (float const *)preds, q, radius, \
                        radius_stride, boxes6, R, self_rank, counts, base, cursors, out_qid, ids);                     \
    else                                                                                                               \
      ABX_LAUNCH_TAGGED("routeKernel<count>", (routeKernel<P, false>), grid, 256, 0, s, (float const *)preds, q,       \
                        radius, radius_stride, boxes6, R, self_rank, counts, base, cursors, out_qid, ids);             \
  } while (0)
  switch (pred_kind)
  {
  case ABX_PRED_SPHERE3F: ABX_ROUTE(ABX_PRED_SPHERE3F); break;
  case ABX_PRED_BOX3F: ABX_ROUTE(ABX_PRED_BOX3F); break;
  case ABX_PRED_POINT3F: ABX_ROUTE(ABX_PRED_POINT3F); break;
  default: setError("unknown predicate kind"); return ABX_ERR_ARG;
  }
#undef ABX_ROUTE
  return ABX_OK;
}

// values2[i] = (indices[i], rank)
__global__ void pairWithRankKernel(int32_t const *__restrict__ indices, int64_t n, int rank, int2 *__restrict__ out)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = make_int2(indices[i], rank);
}
// candidates of one query are contiguous (ids ascending): the thread at the start of a segment merges
// the segment into the query's row by insertion (k and the segment are tens of entries; a few percent
// of the queries have any)
__global__ void knnMergeKernel(int64_t m, int const *__restrict__ ids, int2 const *__restrict__ cand,
                               float const *__restrict__ cand_d, int k, int2 *vals, float *dists)
{
  int64_t const c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m)
    return;
  int const qid = ids[c];
  if (c > 0 && ids[c - 1] == qid)
    return;
  int2 *row_v = vals + (int64_t)qid * k;
  float *row_d = dists + (int64_t)qid * k;
  for (int64_t e = c; e < m && ids[e] == qid; ++e)
  {
    float const d = cand_d[e];
    if (!(d < row_d[k - 1]))
      continue;
    // strict <: a local entry stays ahead of a remote one at the same distance
    int pos = k - 1;
    while (pos > 0 && d < row_d[pos - 1])
    {
      row_d[pos] = row_d[pos - 1];
      row_v[pos] = row_v[pos - 1];
      --pos;
    }
    row_d[pos] = d;
    row_v[pos] = cand[e];
  }
}

abx_status knnMerge(cudaStream_t s, int64_t m, int32_t const *ids, int32_t const *cand2, float const *cand_d, int k,
                    int32_t *vals2, float *dists)
{
  if (m > 0 && k > 0)
    ABX_LAUNCH(knnMergeKernel, divUp(m, 128), 128, 0, s, m, ids, (int2 const *)cand2, cand_d, k,
               (int2 *)vals2, dists);
  return ABX_OK;
}

abx_status pairWithRank(cudaStream_t s, int32_t const *indices, int64_t n, int rank, int32_t *out2)
{
  if (n > 0)
    ABX_LAUNCH(pairWithRankKernel, divUp(n, 256), 256, 0, s, indices, n, rank, (int2 *)out2);
  return ABX_OK;
}

abx_status mergeCrs(cudaStream_t s, int64_t q, int32_t const *local_off, int32_t const *local_idx, int rank,
                    int32_t const *remote_off, int32_t const *remote_vals2, int32_t *out_off, int32_t *out_vals2)
{
  if (q < 0)
    return ABX_OK;
  if (q > 0)
    ABX_LAUNCH(mergeCountsKernel, divUp(q, 256), 256, 0, s, q, local_off, remote_off, out_off);
  ABX_TRY(exclusiveScanI32(s, out_off, out_off, q + 1));
  if (q > 0)
    ABX_LAUNCH(mergeRowsKernel, divUp(q, 256), 256, 0, s, q, local_off, local_idx, rank, remote_off,
               (int2 const *)remote_vals2, out_off, (int2 *)out_vals2);
  return ABX_OK;
}

// ---- merge with the remote results given as (query id ascending, value) records -------------------
__global__ void mergeLocalCountsKernel(int64_t q, int32_t const *__restrict__ local_off, int32_t *__restrict__ counts)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q)
    counts[i] = local_off[i + 1] - local_off[i];
}
// the thread at the start of a query's segment adds the segment length (one writer per query)
__global__ void mergeRemoteCountsKernel(int64_t m, int const *__restrict__ ids, int32_t *__restrict__ counts)
{
  int64_t const c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m)
    return;
  int const qid = ids[c];
  if (c > 0 && ids[c - 1] == qid)
    return;
  int len = 1;
  while (c + len < m && ids[c + len] == qid)
    ++len;
  counts[qid] += len;
}
// local rows -> (index, rank) pairs at their new offsets.  A warp takes 32 consecutive rows and walks
// their concatenated elements 32 at a time (coalesced in, nearly coalesced out: rows that gained no
// remote entries keep their relative positions); the row of an element is found by a 5-step search
// over the 32 row starts held one per lane.
__global__ void __launch_bounds__(256)
    mergeLocalRowsKernel(int64_t q, int32_t const *__restrict__ local_off, int32_t const *__restrict__ local_idx,
                         int rank, int32_t const *__restrict__ out_off, int2 *__restrict__ out_vals)
{
  int const lane = threadIdx.x & 31;
  int64_t const r0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  if (r0 >= q)
    return;
  int64_t const r = min(r0 + lane, q); // lanes past the last row hold its end (empty rows)
  int const lo = local_off[r];
  int const shift = (r < q ? out_off[r] : 0) - lo;
  int const begin = __shfl_sync(0xffffffffu, lo, 0);
  int const end = local_off[min(r0 + 32, q)];
  for (int j = begin + lane; j - lane < end; j += 32)
  {
    // largest lane l with lo_l <= j
    int l = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
    {
      int const probe = __shfl_sync(0xffffffffu, lo, min(l + step, 31));
      if (l + step < 32 && probe <= j)
        l += step;
    }
    int const sh = __shfl_sync(0xffffffffu, shift, l);
    if (j < end)
      out_vals[j + sh] = make_int2(local_idx[j], rank);
  }
}
__global__ void mergeRemoteRowsKernel(int64_t m, int const *__restrict__ ids, int2 const *__restrict__ vals,
                                      int32_t const *__restrict__ local_off, int32_t const *__restrict__ out_off,
                                      int2 *__restrict__ out_vals)
{
  int64_t const c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m)
    return;
  int const qid = ids[c];
  if (c > 0 && ids[c - 1] == qid)
    return;
  int dst = out_off[qid] + (local_off[qid + 1] - local_off[qid]);
  for (int64_t e = c; e < m && ids[e] == qid; ++e)
    out_vals[dst++] = vals[e];
}

// out_off[q + 1] = offsets of the merged rows (local row lengths + records per query id)
abx_status mergeCounts(cudaStream_t s, int64_t q, int32_t const *local_off, int64_t m, int32_t const *remote_ids,
                       int32_t *out_off)
{
  if (q < 0)
    return ABX_OK;
  if (q > 0)
    ABX_LAUNCH(mergeLocalCountsKernel, divUp(q, 256), 256, 0, s, q, local_off, out_off);
  if (m > 0)
    ABX_LAUNCH(mergeRemoteCountsKernel, divUp(m, 256), 256, 0, s, m, remote_ids, out_off);
  return exclusiveScanI32(s, out_off, out_off, q + 1);
}

// remote (index, rank) records, query id ascending, into the rows of a merged CRS behind their local parts
abx_status mergeRemoteRows(cudaStream_t s, int64_t m, int32_t const *remote_ids, int32_t const *remote_vals2,
                           int32_t const *local_off, int32_t const *out_off, int32_t *out_vals2)
{
  if (m > 0)
    ABX_LAUNCH(mergeRemoteRowsKernel, divUp(m, 256), 256, 0, s, m, remote_ids, (int2 const *)remote_vals2, local_off,
               out_off, (int2 *)out_vals2);
  return ABX_OK;
}

abx_status mergeSorted(cudaStream_t s, int64_t q, int32_t const *local_off, int32_t const *local_idx, int rank,
                       int64_t m, int32_t const *remote_ids, int32_t const *remote_vals2, int32_t *out_off,
                       int32_t *out_vals2)
{
  if (q < 0)
    return ABX_OK;
  ABX_TRY(mergeCounts(s, q, local_off, m, remote_ids, out_off));
  if (q > 0)
    ABX_LAUNCH(mergeLocalRowsKernel, divUp(divUp(q, 32) * 32, 256), 256, 0, s, q, local_off, local_idx, rank, out_off,
               (int2 *)out_vals2);
  if (m > 0)
    ABX_LAUNCH(mergeRemoteRowsKernel, divUp(m, 256), 256, 0, s, m, remote_ids, (int2 const *)remote_vals2, local_off,
               out_off, (int2 *)out_vals2);
  return ABX_OK;
}

abx_status compactRows(cudaStream_t s, int64_t q, int32_t const *old_offsets, int32_t const *new_offsets,
                       uint32_t const *old_idx, float const *old_dist, uint32_t *new_idx, float *new_dist)
{
  if (q <= 0)
    return ABX_OK;
  ABX_LAUNCH(compactRowsKernel, divUp(q, 256), 256, 0, s, q, old_offsets, new_offsets, old_idx, old_dist, new_idx,
             new_dist);
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
