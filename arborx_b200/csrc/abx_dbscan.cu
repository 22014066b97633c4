// abx_dbscan.cu -- ArborX::dbscan: FDBSCAN and FDBSCAN-DenseBox.
//
// Behavioural contract: cluster/ArborX_DBSCAN.hpp:219-522,
// cluster/detail/ArborX_FDBSCAN.hpp:31-110, ArborX_FDBSCANDenseBox.hpp:32-307,
// ArborX_UnionFind.hpp:74-181 (ECL-CC), ArborX_CartesianGrid.hpp:27-142.
#include "abx_traverse.cuh"

namespace abx
{

namespace
{

// ---- lock-free union-find (UnionFind.hpp:113-181) ---------------------------------
// Stale reads are benign (every value ever stored in labels[x] is an ancestor of x
// in the final forest); only the root link needs an atomic CAS.
__device__ __forceinline__ int ufLoad(int const *labels, int i) { return __ldcg(labels + i); }

__device__ __forceinline__ int ufRepresentative(int *labels, int i)
{
  int curr = ufLoad(labels, i);
  if (curr != i)
  {
    int const parent = curr;
    int next, prev = i;
    while (curr > (next = ufLoad(labels, curr)))
    {
      __stcg(labels + prev, next); // path halving (UnionFind.hpp:113-128)
      prev = curr;
      curr = next;
    }
    // and point the queried element at the root it found: the same elements (a dense cell's first
    // point, a thread's own point) are asked over and over, and the one-load same-set test of the
    // main kernels only succeeds on direct children of the root.  i is not a root here (roots are
    // only ever changed by the CAS in ufMerge), and any ancestor is a valid parent.
    if (curr != parent)
      __stcg(labels + i, curr);
  }
  return curr;
}

__device__ __forceinline__ void ufMergeInto(int *labels, int i, int j) { __stcg(labels + i, ufRepresentative(labels, j)); }

__device__ __forceinline__ void ufMerge(int *labels, int i, int j)
{
  int vstat = ufRepresentative(labels, i);
  int ostat = ufRepresentative(labels, j);
  while (vstat != ostat)
  {
    if (vstat < ostat)
      ostat = atomicCAS(labels + ostat, ostat, vstat);
    else
      vstat = atomicCAS(labels + vstat, vstat, ostat);
  }
}

__global__ void iotaLabelsKernel(int *labels, int n)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    labels[i] = i;
}

// ---- FDBSCAN: core-point counting (CountUpToN, FDBSCAN.hpp:31-46) ----------------
// one thread per sorted leaf: queries are the tree's own points, already in
// Morton order, so neighbouring threads walk neighbouring subtrees
__global__ void __launch_bounds__(kThreads)
    countCoreKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n, float eps,
                    int minpts, int *__restrict__ num_neigh)
{
  int const t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= n)
    return;
  float4 const p = __ldg(leaf_box + t);
  Pred<ABX_PRED_SPHERE3F> pred;
  pred.cx = p.x, pred.cy = p.y, pred.cz = p.z, pred.r = eps;
  pred.t = sqrtThreshold(eps);
  int count = 0;
  traverseSpatial<1>(nodes, leaf_box, pred, [&](unsigned, int) { return ++count >= minpts; });
  num_neigh[__float_as_uint(p.w)] = count;
}

constexpr int kDbscanQueue = 16; // queued leaf runs per thread (deferred leaf tests)

// ---- FDBSCAN: half traversal + FDBSCANCallback (FDBSCAN.hpp:49-110) ---------------
template <bool SPECIAL /*minpts == 2*/, bool STAR>
__global__ void __launch_bounds__(kThreads)
    fdbscanMainKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n, float eps,
                      int minpts, int const *__restrict__ num_neigh, int *labels)
{
  // deferred leaf tests (abx_traverse.cuh): inside a cluster a point has hundreds of neighbours, and
  // testing them with the warp converged matters more than anywhere else; all lanes stay
  __shared__ unsigned squeue[kDbscanQueue * kThreads];
  int const t0 = blockIdx.x * kThreads + threadIdx.x;
  bool active = t0 < n;
  int const t = active ? t0 : 0;
  float4 const p = __ldg(leaf_box + t);
  Pred<ABX_PRED_SPHERE3F> pred;
  pred.cx = p.x, pred.cy = p.y, pred.cz = p.z, pred.r = eps;
  pred.t = sqrtThreshold(eps);
  int const i = (int)__float_as_uint(p.w);
  bool const i_core = SPECIAL ? true : (num_neigh[i] >= minpts);
  if (STAR && !i_core)
    active = false; // border points do not take part in DBSCAN* (callback would return at once)
  // Inside a cluster almost every pair is already in one set.  rep_i is a (possibly stale) root of
  // i's set; sets only ever merge, so labels[j] == rep_i proves j is in i's set with one load and
  // no chase; anything else goes through the full merge and refreshes rep_i.
  int rep_i = i_core ? ufRepresentative(labels, i) : -1;
  auto mergeCore = [&](int j) {
    if (ufLoad(labels, j) == rep_i)
      return;
    ufMerge(labels, i, j);
    rep_i = ufRepresentative(labels, i);
  };
  auto pair = [&](unsigned orig_j, int) {
    int const j = (int)orig_j;
    bool const j_core = SPECIAL ? true : (num_neigh[j] >= minpts);
    if (STAR)
    {
      if (j_core)
        mergeCore(j);
      return false;
    }
    if (!i_core)
    {
      if (j_core)
        ufMergeInto(labels, i, j); // never merge(): a border point must not bridge clusters
    }
    else
    {
      if (j_core)
        mergeCore(j);
      else
        ufMergeInto(labels, j, i);
    }
    return false;
  };
  traverseSpatialDeferred<1, 4, kDbscanQueue>(nodes, leaf_box, pred, active, squeue, pair, t);
}

// ---- finalize_labels (:489-506) and mark_noise (:512-518) -------------------------
__global__ void finalizeLabelsKernel(int *labels, int *cluster_sizes, int n)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  int vstat = -1;
  if (i < n)
  {
    int next;
    vstat = ufLoad(labels, i);
    int const old = vstat;
    while (vstat > (next = ufLoad(labels, vstat)))
      vstat = next;
    if (vstat != old)
      __stcg(labels + i, vstat);
  }
  // one atomic per distinct cluster per warp: neighbouring (Morton-ordered) points mostly share
  // their cluster, and a few giant clusters would otherwise serialise 10M atomics on a few words
  unsigned const peers = __match_any_sync(0xffffffffu, vstat);
  if (vstat >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1)
    atomicAdd(cluster_sizes + vstat, __popc(peers));
}

__global__ void markNoiseKernel(int *labels, int const *__restrict__ cluster_sizes, int const *__restrict__ num_neigh,
                                int minpts, int n)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  bool const special = (minpts == 2);
  int const l = labels[i];
  if (cluster_sizes[l] == 1 && (special || !(num_neigh[i] >= minpts)))
    labels[i] = -1;
}

// ---- FDBSCAN-DenseBox --------------------------------------------------------------
struct Grid
{
  float lo[3];
  float h;
  unsigned long long n[3];
};

// CartesianGrid::cellIndex (CartesianGrid.hpp:51-63)
__device__ __forceinline__ unsigned long long cellIndex(Grid const &g, float x, float y, float z)
{
  float const p[3] = {x, y, z};
  unsigned long long s = 0;
#pragma unroll
  for (int d = 2; d >= 0; --d)
  {
    int const i = (int)floorf(__fdiv_rn(__fsub_rn(p[d], g.lo[d]), g.h));
    s = s * g.n[d] + (unsigned long long)(long long)i;
  }
  return s;
}

__global__ void cellIndicesKernel(float const *__restrict__ xyz, int n, Grid g, unsigned long long *__restrict__ cells)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    cells[i] = cellIndex(g, xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]);
}

// flags[i] = 1 when sorted position i starts a new cell (computeOffsetsInOrderedView,
// misc/ArborX_Utils.hpp:25-51); flags has n+1 slots, the last one is scratch
__global__ void cellStartFlagsKernel(unsigned long long const *__restrict__ cells, int n, int *__restrict__ flags)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    flags[i] = (i == 0 || cells[i] != cells[i - 1]) ? 1 : 0;
}

// cell_offsets[rank[i]] = i for every cell start, cell_offsets[num_cells] = n
__global__ void cellOffsetsKernel(int const *__restrict__ flags, int const *__restrict__ rank, int n,
                                  int *__restrict__ cell_offsets)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    if (flags[i])
      cell_offsets[rank[i]] = i;
  }
  else if (i == n)
    cell_offsets[rank[n]] = n;
}

// per cell: dense size (>= minpts ? size : 0) and sparse size, scanned separately so
// that dense cells come first and cells keep their sorted order inside each class
// (reorderDenseAndSparseCells, FDBSCANDenseBox.hpp:215-283; the reference's order of
// cells inside a class is arbitrary)
__global__ void cellClassSizesKernel(int const *__restrict__ cell_offsets, int num_cells, int minpts,
                                     int *__restrict__ dense_size, int *__restrict__ sparse_size,
                                     int *__restrict__ dense_flag)
{
  int const c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= num_cells)
    return;
  int const sz = cell_offsets[c + 1] - cell_offsets[c];
  bool const dense = sz >= minpts;
  dense_size[c] = dense ? sz : 0;
  sparse_size[c] = dense ? 0 : sz;
  dense_flag[c] = dense ? 1 : 0;
}

// scatter points to their reordered position; records for each dense cell its
// start in the reordered arrays (dense_cell_offsets) and its grid cell
__global__ void reorderCellsKernel(int const *__restrict__ flags_rank /*cell id per sorted position*/,
                                   int const *__restrict__ cell_offsets, int const *__restrict__ dense_off,
                                   int const *__restrict__ sparse_off, int const *__restrict__ dense_flag,
                                   int const *__restrict__ dense_rank, int num_points_dense,
                                   unsigned const *__restrict__ perm_in, unsigned long long const *__restrict__ cells_in,
                                   int n, unsigned *__restrict__ perm_out, int *__restrict__ dense_cell_offsets,
                                   unsigned long long *__restrict__ dense_cell_ids)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int const c = flags_rank[i];
  int const within = i - cell_offsets[c];
  bool const dense = dense_flag[c] != 0;
  int const dst = dense ? dense_off[c] + within : num_points_dense + sparse_off[c] + within;
  perm_out[dst] = perm_in[i];
  if (dense && within == 0)
  {
    dense_cell_offsets[dense_rank[c]] = dst;
    dense_cell_ids[dense_rank[c]] = cells_in[i];
  }
}

// cell id of every sorted position = (inclusive scan of start flags) - 1
__global__ void cellIdKernel(int const *__restrict__ flags, int const *__restrict__ rank, int n, int *__restrict__ id)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    id[i] = rank[i] + flags[i] - 1;
}

// unionFindWithinEachDenseCell (FDBSCANDenseBox.hpp:286-307): points of a dense cell
// are mutually within eps, link neighbours in the reordered order
__global__ void denseCellUnionKernel(int const *__restrict__ dense_cell_offsets, int num_dense,
                                     unsigned const *__restrict__ perm, int num_points_dense, int *labels)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1 || i >= num_points_dense)
    return;
  // i and i-1 are in the same dense cell unless i starts a cell: binary search
  int lo = 0, hi = num_dense;
  while (hi - lo > 1)
  {
    int const mid = (lo + hi) / 2;
    if (dense_cell_offsets[mid] <= i)
      lo = mid;
    else
      hi = mid;
  }
  if (dense_cell_offsets[lo] != i)
    ufMerge(labels, (int)perm[i], (int)perm[i - 1]);
}

// MixedBoxPrimitives (ArborX_DBSCAN.hpp:73-175): dense-cell boxes (CartesianGrid::cellBox,
// CartesianGrid.hpp:66-82) followed by degenerate boxes of the sparse points
__global__ void mixedPrimitivesKernel(float const *__restrict__ xyz, unsigned const *__restrict__ perm, Grid g,
                                      unsigned long long const *__restrict__ dense_cell_ids, int num_dense,
                                      int num_points_dense, int n_prims, float *__restrict__ boxes6)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_prims)
    return;
  float lo[3], hi[3];
  if (i < num_dense)
  {
    unsigned long long cell = dense_cell_ids[i];
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      unsigned long long const k = cell % g.n[d];
      cell /= g.n[d];
      // max = min + (i+1)*h ; min += i*h   (size_t -> float conversions as in the reference)
      hi[d] = __fadd_rn(g.lo[d], __fmul_rn((float)(k + 1), g.h));
      lo[d] = __fadd_rn(g.lo[d], __fmul_rn((float)k, g.h));
    }
  }
  else
  {
    unsigned const o = perm[num_points_dense + (i - num_dense)];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      lo[d] = hi[d] = xyz[3 * (size_t)o + d];
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    boxes6[6 * (size_t)i + d] = lo[d];
    boxes6[6 * (size_t)i + 3 + d] = hi[d];
  }
}

__global__ void markDenseCoreKernel(unsigned const *__restrict__ perm, int num_points_dense, int *num_neigh)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < num_points_dense)
    num_neigh[perm[i]] = INT_MAX; // ArborX_DBSCAN.hpp:438-441
}

// the points of the dense cells gathered in reordered (cell by cell) order: (x, y, z, bits(original index)).
// Every loop over "the points of dense cell k" below reads a contiguous run of this array instead of
// chasing perm -> xyz.
__global__ void gatherDensePointsKernel(float const *__restrict__ xyz, unsigned const *__restrict__ perm,
                                        int num_points_dense, float4 *__restrict__ dense_pts)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_points_dense)
    return;
  unsigned const o = perm[i];
  dense_pts[i] = make_float4(xyz[3 * (size_t)o], xyz[3 * (size_t)o + 1], xyz[3 * (size_t)o + 2], __uint_as_float(o));
}

// distance(query_point, point_j) <= eps with the reference's operand order (tmp = point_j - query)
__device__ __forceinline__ bool withinEps4(float4 pj, float px, float py, float pz, float t)
{
  float tx = __fsub_rn(pj.x, px), ty = __fsub_rn(pj.y, py), tz = __fsub_rn(pj.z, pz);
  float d2 = __fmul_rn(tx, tx);
  d2 = __fadd_rn(d2, __fmul_rn(ty, ty));
  d2 = __fadd_rn(d2, __fmul_rn(tz, tz));
  return d2 <= t;
}

// CountUpToN_DenseBox (FDBSCANDenseBox.hpp:32-96) for the points of sparse cells
__global__ void __launch_bounds__(kThreads)
    denseCountKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n_prims,
                     float const *__restrict__ xyz, unsigned const *__restrict__ perm,
                     float4 const *__restrict__ dense_pts, int const *__restrict__ dense_cell_offsets, int num_dense,
                     int num_points_dense, int n, float eps, int minpts, int *__restrict__ num_neigh)
{
  int const s = blockIdx.x * kThreads + threadIdx.x;
  if (s >= n - num_points_dense)
    return;
  int const i = (int)perm[num_points_dense + s];
  float const px = xyz[3 * (size_t)i], py = xyz[3 * (size_t)i + 1], pz = xyz[3 * (size_t)i + 2];
  Pred<ABX_PRED_SPHERE3F> pred;
  pred.cx = px, pred.cy = py, pred.cz = pz, pred.r = eps;
  pred.t = sqrtThreshold(eps);
  int count = 0;
  if (n_prims >= 2)
    traverseSpatial<2>(nodes, leaf_box, pred, [&](unsigned prim, int) {
      int const k = (int)prim;
      if (k < num_dense)
      {
        int const ce = dense_cell_offsets[k + 1];
        for (int jj = dense_cell_offsets[k]; jj < ce; ++jj)
          if (withinEps4(__ldg(dense_pts + jj), px, py, pz, pred.t))
            if (++count >= minpts)
              return true;
        return false;
      }
      return ++count >= minpts;
    });
  else
    count = 1; // a single primitive: the point itself
  num_neigh[i] = count;
}

// ---- FDBSCANDenseBoxCallback (FDBSCANDenseBox.hpp:98-205), B200 shape ----------------------------------------
// The reference runs one full traversal per POINT; on clustered data nearly every point sits in a dense cell,
// and all points of a cell find the same neighbouring cells and repeat the same "is any point of that cell
// within eps of me" loops (ncu on GanTao 10M: 82 ms, 6 GB/s of DRAM, 68 warps stalled on dependent label loads
// per issue).  What the callback computes for dense cells is a relation between CELLS: dense cells A and B end up
// in one cluster iff some a in A, b in B are within eps (all points of a dense cell are core and mutually
// within eps).  So:
//   denseCellPairsKernel   one WARP per dense cell A: a warp-uniform traversal with A's box finds the dense cells
//                          B > A whose box is within eps, skips the ones already in A's set (one label load) and
//                          tests the rest cooperatively -- points of A within eps of B's box against the points of
//                          B (and the mirror image), 32 pairs per step, stopping at the first hit.
//   sparseMainKernel       one thread per point of a SPARSE cell (a small minority): the reference callback, plus
//                          border points attaching themselves to the first core neighbour they find (the reference
//                          lets every core neighbour overwrite the border point's label: same set of outcomes).
// Core points end up with the same partition as the reference (labels = smallest index of the component).
__device__ __forceinline__ float boxBoxDist2(float const *A, float4 lo, float4 hi)
{
  float d2 = 0.f;
  float const blo[3] = {lo.x, lo.y, lo.z}, bhi[3] = {hi.x, hi.y, hi.z};
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    float const gap = fmaxf(fmaxf(__fsub_rn(blo[d], A[3 + d]), __fsub_rn(A[d], bhi[d])), 0.f);
    d2 = __fadd_rn(d2, __fmul_rn(gap, gap));
  }
  return d2;
}

__global__ void __launch_bounds__(128)
    denseCellPairsKernel(Node64 const *__restrict__ nodes, float const *__restrict__ boxes6,
                         float4 const *__restrict__ dense_pts, int const *__restrict__ dense_cell_offsets, int num_dense,
                         float eps, int *labels)
{
  unsigned const full = 0xffffffffu;
  int const lane = threadIdx.x & 31;
  int const c = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= num_dense)
    return; // the whole warp
  float const t = sqrtThreshold(eps);
  // culling threshold for box-to-box distances, inflated: a box pair is only dropped when no point pair can pass
  // the exact tests below
  float const t_cull = __fadd_rn(__fmul_rn(t, 1.0001f), 1e-30f);
  float A[6];
#pragma unroll
  for (int d = 0; d < 6; ++d)
    A[d] = __ldg(boxes6 + 6 * (size_t)c + d);
  int const a0 = dense_cell_offsets[c], a1 = dense_cell_offsets[c + 1];
  int const first_c = (int)__float_as_uint(__ldg(dense_pts + a0).w);
  int rep_c = 0;
  if (lane == 0)
    rep_c = ufRepresentative(labels, first_c);
  rep_c = __shfl_sync(full, rep_c, 0);

  // one candidate cell k (the same value in every lane)
  auto process = [&](int k) {
    int const b0 = dense_cell_offsets[k], b1 = dense_cell_offsets[k + 1];
    int const first_k = (int)__float_as_uint(__ldg(dense_pts + b0).w);
    if (ufLoad(labels, first_k) == rep_c)
      return;
    int rc = 0, rk = 0;
    if (lane == 0)
    {
      rc = ufRepresentative(labels, first_c);
      rk = ufRepresentative(labels, first_k);
    }
    rc = __shfl_sync(full, rc, 0);
    rk = __shfl_sync(full, rk, 0);
    rep_c = rc;
    if (rc == rk)
      return;
    float B[6];
#pragma unroll
    for (int d = 0; d < 6; ++d)
      B[d] = __ldg(boxes6 + 6 * (size_t)k + d);
    bool found = false;
    // side = 0: points of A that see B's box (what a's own traversal would have reported) against every point
    // of B; side = 1: the mirror image.  A pair within eps is found by at least one side.
    for (int side = 0; side < 2 && !found; ++side)
    {
      int const p0 = side == 0 ? a0 : b0, p1 = side == 0 ? a1 : b1; // the "query" cell
      int const o0 = side == 0 ? b0 : a0, o1 = side == 0 ? b1 : a1; // the other cell
      float const *obox = side == 0 ? B : A;
      for (int base = p0; base < p1 && !found; base += 32)
      {
        int const pi = base + lane;
        float4 const pp = pi < p1 ? __ldg(dense_pts + pi) : make_float4(0.f, 0.f, 0.f, 0.f);
        bool const sees = pi < p1 && pointBoxDist2(pp.x, pp.y, pp.z, obox[0], obox[1], obox[2], obox[3], obox[4],
                                                   obox[5]) <= t;
        unsigned m = __ballot_sync(full, sees);
        while (m && !found)
        {
          int const src = __ffs(m) - 1;
          m &= m - 1;
          float const qx = __shfl_sync(full, pp.x, src), qy = __shfl_sync(full, pp.y, src),
                      qz = __shfl_sync(full, pp.z, src);
          for (int ob = o0; ob < o1; ob += 32)
          {
            int const oi = ob + lane;
            bool const hit = oi < o1 && withinEps4(__ldg(dense_pts + oi), qx, qy, qz, t);
            if (__any_sync(full, hit))
            {
              found = true;
              break;
            }
          }
        }
      }
    }
    if (found)
    {
      if (lane == 0)
      {
        ufMerge(labels, first_c, first_k);
        rc = ufRepresentative(labels, first_c);
      }
      rep_c = __shfl_sync(full, rc, 0);
    }
  };

  // warp-uniform traversal of the mixed tree with A's box grown by eps
  int stack[kStackSize];
  int sp = 0;
  int node = 0;
  while (true)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
    float4 const n0 = __ldg(f), n1 = __ldg(f + 1), n2 = __ldg(f + 2), n3 = __ldg(f + 3);
    int const lref = __float_as_int(n0.w), rref = __float_as_int(n1.w);
    bool hit_l = boxBoxDist2(A, n0, n1) <= t_cull;
    bool hit_r = boxBoxDist2(A, n2, n3) <= t_cull;
    if (hit_l && refIsLeaf(lref))
    {
      int const k = (int)refOrig(lref);
      if (k < num_dense && k > c)
        process(k);
      hit_l = false;
    }
    if (hit_r && refIsLeaf(rref))
    {
      int const k = (int)refOrig(rref);
      if (k < num_dense && k > c)
        process(k);
      hit_r = false;
    }
    if (hit_l)
    {
      if (hit_r)
        stack[sp++] = rref;
      node = lref;
    }
    else if (hit_r)
      node = rref;
    else
    {
      if (sp == 0)
        break;
      node = stack[--sp];
    }
  }
}

template <bool SPECIAL, bool STAR>
__global__ void __launch_bounds__(kThreads)
    sparseMainKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n_prims,
                     float const *__restrict__ xyz, unsigned const *__restrict__ perm,
                     float4 const *__restrict__ dense_pts, int const *__restrict__ dense_cell_offsets, int num_dense,
                     int num_points_dense, int n, float eps, int minpts, int const *__restrict__ num_neigh, int *labels)
{
  __shared__ unsigned squeue[kDbscanQueue * kThreads];
  if (n_prims < 2)
    return;
  int const s0 = blockIdx.x * kThreads + threadIdx.x;
  bool active = s0 < n - num_points_dense;
  int const i = (int)perm[num_points_dense + (active ? s0 : 0)];
  bool const i_core = SPECIAL ? true : (num_neigh[i] >= minpts);
  if (STAR && !i_core)
    active = false; // DBSCAN*: border points stay unlabelled
  float const px = xyz[3 * (size_t)i], py = xyz[3 * (size_t)i + 1], pz = xyz[3 * (size_t)i + 2];
  Pred<ABX_PRED_SPHERE3F> pred;
  pred.cx = px, pred.cy = py, pred.cz = pz, pred.r = eps;
  pred.t = sqrtThreshold(eps);
  // rep_i: a (possibly stale) root of i's set -- see fdbscanMainKernel
  int rep_i = i_core ? ufRepresentative(labels, i) : -1;
  traverseSpatialDeferred<2, 4, kDbscanQueue>(nodes, leaf_box, pred, active, squeue, [&](unsigned prim, int) {
    int const k = (int)prim;
    if (k < num_dense)
    {
      int const cs = dense_cell_offsets[k], ce = dense_cell_offsets[k + 1];
      if (i_core)
      {
        int const first = (int)__float_as_uint(__ldg(dense_pts + cs).w);
        if (ufLoad(labels, first) == rep_i)
          return false;
        if (ufRepresentative(labels, i) == ufRepresentative(labels, first))
          return false;
        for (int jj = cs; jj < ce; ++jj)
        {
          float4 const pj = __ldg(dense_pts + jj);
          if (withinEps4(pj, px, py, pz, pred.t))
          {
            ufMerge(labels, i, (int)__float_as_uint(pj.w));
            break;
          }
        }
        rep_i = ufRepresentative(labels, i);
        return false;
      }
      // border point: it joins the cluster of the first core point found within eps (every point of a dense
      // cell is core) and stops
      for (int jj = cs; jj < ce; ++jj)
      {
        float4 const pj = __ldg(dense_pts + jj);
        if (withinEps4(pj, px, py, pz, pred.t))
        {
          ufMergeInto(labels, i, (int)__float_as_uint(pj.w));
          return true;
        }
      }
      return false;
    }
    int const j = (int)perm[num_points_dense + (k - num_dense)];
    if (j == i)
      return false;
    bool const j_core = SPECIAL ? true : (num_neigh[j] >= minpts);
    if (i_core)
    {
      // core-core pairs are merged by the larger index; border neighbours attach themselves
      if (j_core && i > j && ufLoad(labels, j) != rep_i)
      {
        ufMerge(labels, i, j);
        rep_i = ufRepresentative(labels, i);
      }
      return false;
    }
    if (j_core)
    {
      ufMergeInto(labels, i, j);
      return true;
    }
    return false;
  });
}

abx_status readInt(cudaStream_t s, int const *dev, int &out)
{
  ABX_CUDA_TRY(cudaMemcpyAsync(&out, dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  return ABX_OK;
}

// core_flags[i] = 1 when point i is a core point (every point counts as core for minpts == 2, where the
// reference never counts neighbours: ArborX_DBSCAN.hpp:273-283)
__global__ void coreFlagsKernel(int n, int const *__restrict__ num_neigh, int minpts, int *__restrict__ core_flags)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    core_flags[i] = (!num_neigh || num_neigh[i] >= minpts) ? 1 : 0;
}

abx_status finalize(cudaStream_t s, int n, int minpts, int const *num_neigh, int *labels, int *core_flags)
{
  if (core_flags)
    ABX_LAUNCH(coreFlagsKernel, divUp(n, 256), 256, 0, s, n, num_neigh, minpts, core_flags);
  TempBuffer<int> cluster_sizes;
  ABX_TRY(cluster_sizes.alloc(n, s));
  ABX_CUDA_TRY(cudaMemsetAsync(cluster_sizes.ptr, 0, sizeof(int) * (size_t)n, s));
  ABX_LAUNCH(finalizeLabelsKernel, divUp(n, 256), 256, 0, s, labels, cluster_sizes.ptr, n);
  ABX_LAUNCH(markNoiseKernel, divUp(n, 256), 256, 0, s, labels, cluster_sizes.ptr, num_neigh, minpts, n);
  return ABX_OK;
}

struct TreeGuard
{
  abx_bvh *t = nullptr;
  ~TreeGuard()
  {
    if (t)
      abx_bvh_destroy(t);
  }
};

abx_status fdbscan(cudaStream_t s, float const *xyz, int n, float eps, int minpts, int algo, int *labels,
                   int *core_flags)
{
  TreeGuard tree;
  ABX_TRY(buildTree(s, ABX_PRIM_POINT3F, xyz, n, nullptr, &tree.t));
  abx_bvh *t = tree.t;
  ABX_LAUNCH(iotaLabelsKernel, divUp(n, 256), 256, 0, s, labels, n);
  bool const special = (minpts == 2);
  bool const star = (algo == ABX_DBSCAN_DBSCAN_STAR);
  TempBuffer<int> num_neigh;
  int const grid = divUp(n, kThreads);
  if (!special)
  {
    ABX_TRY(num_neigh.alloc(n, s));
    if (n >= 2)
      ABX_LAUNCH(countCoreKernel, grid, kThreads, 0, s, t->nodes, t->leaf_box, n, eps, minpts, num_neigh.ptr);
    else
      ABX_CUDA_TRY(cudaMemsetAsync(num_neigh.ptr, 0, sizeof(int) * (size_t)n, s)); // 1 point: count 1 < minpts
  }
  if (n >= 2)
  {
    if (special)
      ABX_LAUNCH((fdbscanMainKernel<true, false>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n, eps, minpts,
                 (int const *)nullptr, labels);
    else if (star)
      ABX_LAUNCH((fdbscanMainKernel<false, true>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n, eps, minpts,
                 (int const *)num_neigh.ptr, labels);
    else
      ABX_LAUNCH((fdbscanMainKernel<false, false>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n, eps, minpts,
                 (int const *)num_neigh.ptr, labels);
  }
  return finalize(s, n, minpts, num_neigh.ptr, labels, core_flags);
}

abx_status denseBox(cudaStream_t s, float const *xyz, int n, float eps, int minpts, int algo, int *labels,
                    int *core_flags)
{
  bool const special = (minpts == 2);
  bool const star = (algo == ABX_DBSCAN_DBSCAN_STAR);
  // scene bounds -> grid (ArborX_DBSCAN.hpp:333-342); the grid lives on the host,
  // like the reference's CartesianGrid
  TempBuffer<unsigned> enc;
  TempBuffer<float> bounds_dev;
  ABX_TRY(enc.alloc(6, s));
  ABX_TRY(bounds_dev.alloc(6, s));
  ABX_TRY(sceneBounds(s, ABX_PRIM_POINT3F, xyz, n, enc.ptr));
  ABX_TRY(decodeBounds(s, enc.ptr, bounds_dev.ptr));
  float b[6];
  ABX_CUDA_TRY(cudaMemcpyAsync(b, bounds_dev.ptr, sizeof(b), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  Grid g;
  g.h = eps / sqrtf(3.f); // ArborX_DBSCAN.hpp:341
  for (int d = 0; d < 3; ++d)
  {
    g.lo[d] = b[d];
    float const delta = b[3 + d] - b[d];
    g.n[d] = delta != 0 ? (unsigned long long)std::ceil(delta / g.h) : 1ull; // CartesianGrid.hpp:92-108
  }
  {
    // overflow guard (:110-119) and loss-of-precision guard (:121-136)
    unsigned long long m = ~0ull;
    for (int d = 1; d < 3; ++d)
    {
      m /= g.n[d - 1];
      if (!(g.n[d] < m))
      {
        setError("DenseBox grid cell index would overflow");
        return ABX_ERR_SEARCH;
      }
    }
    float const tol = 5 * FLT_EPSILON;
    for (int d = 0; d < 3; ++d)
      if (std::fabs(g.h / g.lo[d]) < tol)
      {
        setError("ArborX exception: FDBSCAN-DenseBox algorithm will experience loss of precision, undetectably "
                 "producing wrong results. Please switch to using FDBSCAN.");
        return ABX_ERR_PRECISION;
      }
  }
  int const grid_n = divUp(n, 256);
  TempBuffer<unsigned long long> cells;
  TempBuffer<unsigned> perm, perm2;
  ABX_TRY(cells.alloc(n, s));
  ABX_TRY(perm.alloc(n, s));
  ABX_TRY(perm2.alloc(n, s));
  ABX_LAUNCH(cellIndicesKernel, grid_n, 256, 0, s, xyz, n, g, cells.ptr);
  // cell ids are < nx*ny*nz: only the digits that can be non-zero are sorted.  Cells hold long
  // runs of equal keys, so the plain LSD form is used
  int cell_bits = 1;
  {
    long double const total = (long double)g.n[0] * (long double)g.n[1] * (long double)g.n[2];
    while (cell_bits < 64 && (long double)((unsigned long long)1 << cell_bits) < total)
      ++cell_bits;
  }
  ABX_TRY(sortPairsU64(s, (uint64_t *)cells.ptr, perm.ptr, n, true, cell_bits, /*fixup=*/false));

  // distinct cells
  TempBuffer<int> flags, rank, cell_id, cell_offsets;
  ABX_TRY(flags.alloc((size_t)n + 1, s));
  ABX_TRY(rank.alloc((size_t)n + 1, s));
  ABX_TRY(cell_id.alloc(n, s));
  ABX_LAUNCH(cellStartFlagsKernel, grid_n, 256, 0, s, cells.ptr, n, flags.ptr);
  ABX_TRY(exclusiveScanI32(s, flags.ptr, rank.ptr, (int64_t)n + 1));
  int num_cells = 0;
  ABX_TRY(readInt(s, rank.ptr + n, num_cells));
  ABX_TRY(cell_offsets.alloc((size_t)num_cells + 1, s));
  ABX_LAUNCH(cellOffsetsKernel, divUp((int64_t)n + 1, 256), 256, 0, s, flags.ptr, rank.ptr, n, cell_offsets.ptr);
  ABX_LAUNCH(cellIdKernel, grid_n, 256, 0, s, flags.ptr, rank.ptr, n, cell_id.ptr);

  // dense / sparse classes
  TempBuffer<int> dense_size, sparse_size, dense_flag, dense_off, sparse_off, dense_rank;
  size_t const nc1 = (size_t)num_cells + 1;
  ABX_TRY(dense_size.alloc(nc1, s));
  ABX_TRY(sparse_size.alloc(nc1, s));
  ABX_TRY(dense_flag.alloc(nc1, s));
  ABX_TRY(dense_off.alloc(nc1, s));
  ABX_TRY(sparse_off.alloc(nc1, s));
  ABX_TRY(dense_rank.alloc(nc1, s));
  ABX_LAUNCH(cellClassSizesKernel, divUp(num_cells, 256), 256, 0, s, cell_offsets.ptr, num_cells, minpts,
             dense_size.ptr, sparse_size.ptr, dense_flag.ptr);
  ABX_TRY(exclusiveScanI32(s, dense_size.ptr, dense_off.ptr, (int64_t)nc1));
  ABX_TRY(exclusiveScanI32(s, sparse_size.ptr, sparse_off.ptr, (int64_t)nc1));
  ABX_TRY(exclusiveScanI32(s, dense_flag.ptr, dense_rank.ptr, (int64_t)nc1));
  int num_points_dense = 0, num_dense = 0;
  ABX_TRY(readInt(s, dense_off.ptr + num_cells, num_points_dense));
  ABX_TRY(readInt(s, dense_rank.ptr + num_cells, num_dense));

  TempBuffer<int> dense_cell_offsets;
  TempBuffer<unsigned long long> dense_cell_ids;
  ABX_TRY(dense_cell_offsets.alloc((size_t)num_dense + 1, s));
  ABX_TRY(dense_cell_ids.alloc((size_t)std::max(num_dense, 1), s));
  ABX_LAUNCH(reorderCellsKernel, grid_n, 256, 0, s, cell_id.ptr, cell_offsets.ptr, dense_off.ptr, sparse_off.ptr,
             dense_flag.ptr, dense_rank.ptr, num_points_dense, perm.ptr, cells.ptr, n, perm2.ptr,
             dense_cell_offsets.ptr, dense_cell_ids.ptr);
  ABX_CUDA_TRY(cudaMemcpyAsync(dense_cell_offsets.ptr + num_dense, &num_points_dense, sizeof(int),
                               cudaMemcpyHostToDevice, s));
  // (the host int outlives the copy: every path below synchronises or the copy is
  // from pageable memory, which cudaMemcpyAsync stages before returning)

  ABX_LAUNCH(iotaLabelsKernel, grid_n, 256, 0, s, labels, n);
  if (num_points_dense > 1)
    ABX_LAUNCH(denseCellUnionKernel, divUp(num_points_dense, 256), 256, 0, s, dense_cell_offsets.ptr, num_dense,
               perm2.ptr, num_points_dense, labels);

  // BVH over the mixed primitives
  int const n_sparse = n - num_points_dense;
  int const n_prims = num_dense + n_sparse;
  TempBuffer<float> boxes;
  ABX_TRY(boxes.alloc(6 * (size_t)n_prims, s));
  ABX_LAUNCH(mixedPrimitivesKernel, divUp(n_prims, 256), 256, 0, s, xyz, perm2.ptr, g, dense_cell_ids.ptr, num_dense,
             num_points_dense, n_prims, boxes.ptr);
  TreeGuard tree;
  ABX_TRY(buildTree(s, ABX_PRIM_BOX3F, boxes.ptr, n_prims, nullptr, &tree.t));
  abx_bvh *t = tree.t;

  TempBuffer<float4> dense_pts;
  ABX_TRY(dense_pts.alloc((size_t)std::max(num_points_dense, 1), s));
  if (num_points_dense > 0)
    ABX_LAUNCH(gatherDensePointsKernel, divUp(num_points_dense, 256), 256, 0, s, xyz, perm2.ptr, num_points_dense,
               dense_pts.ptr);
  TempBuffer<int> num_neigh;
  if (!special)
  {
    ABX_TRY(num_neigh.alloc(n, s));
    if (num_points_dense > 0)
      ABX_LAUNCH(markDenseCoreKernel, divUp(num_points_dense, 256), 256, 0, s, perm2.ptr, num_points_dense,
                 num_neigh.ptr);
    if (n_sparse > 0)
      ABX_LAUNCH(denseCountKernel, divUp(n_sparse, kThreads), kThreads, 0, s, t->nodes, t->leaf_box, n_prims, xyz,
                 perm2.ptr, dense_pts.ptr, dense_cell_offsets.ptr, num_dense, num_points_dense, n, eps, minpts,
                 num_neigh.ptr);
  }
  // dense cells among themselves: one warp per cell
  if (num_dense > 1 && n_prims >= 2)
    ABX_LAUNCH(denseCellPairsKernel, divUp(num_dense, 4), 128, 0, s, t->nodes, boxes.ptr, dense_pts.ptr,
               dense_cell_offsets.ptr, num_dense, eps, labels);
  // points of sparse cells: one thread per point
  if (n_sparse > 0)
  {
    int const grid = divUp(n_sparse, kThreads);
    if (special)
      ABX_LAUNCH((sparseMainKernel<true, false>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n_prims, xyz, perm2.ptr,
                 dense_pts.ptr, dense_cell_offsets.ptr, num_dense, num_points_dense, n, eps, minpts,
                 (int const *)nullptr, labels);
    else if (star)
      ABX_LAUNCH((sparseMainKernel<false, true>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n_prims, xyz, perm2.ptr,
                 dense_pts.ptr, dense_cell_offsets.ptr, num_dense, num_points_dense, n, eps, minpts,
                 (int const *)num_neigh.ptr, labels);
    else
      ABX_LAUNCH((sparseMainKernel<false, false>), grid, kThreads, 0, s, t->nodes, t->leaf_box, n_prims, xyz,
                 perm2.ptr, dense_pts.ptr, dense_cell_offsets.ptr, num_dense, num_points_dense, n, eps, minpts,
                 (int const *)num_neigh.ptr, labels);
  }
  return finalize(s, n, minpts, num_neigh.ptr, labels, core_flags);
}

} // namespace

abx_status dbscan(cudaStream_t s, float const *xyz, int64_t n, float eps, int32_t minpts, int impl, int algo,
                  int32_t *labels, int32_t *core_flags)
{
  // ArborX_DBSCAN.hpp:240-241
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
  if (n < 0 || n >= (int64_t)1 << 30)
  {
    setError("number of points must be in [0, 2^30)");
    return ABX_ERR_ARG;
  }
  if (algo != ABX_DBSCAN_DBSCAN && algo != ABX_DBSCAN_DBSCAN_STAR)
  {
    setError("unknown DBSCAN algorithm");
    return ABX_ERR_ARG;
  }
  if (n == 0)
    return ABX_OK;
  if (impl == ABX_DBSCAN_FDBSCAN)
    return fdbscan(s, xyz, (int)n, eps, minpts, algo, labels, core_flags);
  if (impl == ABX_DBSCAN_FDBSCAN_DENSEBOX)
    return denseBox(s, xyz, (int)n, eps, minpts, algo, labels, core_flags);
  setError("unknown DBSCAN implementation");
  return ABX_ERR_ARG;
}

} // namespace abx
