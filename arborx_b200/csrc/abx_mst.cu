// abx_mst.cu -- Euclidean / mutual-reachability minimum spanning tree (Boruvka over the BVH) and the dendrogram of
// its edges: the HDBSCAN pipeline.
//
// Behavioural contract:
//   cluster/ArborX_MinimumSpanningTree.hpp:46-258   MinimumSpanningTree(space, points, k) in MST mode: BVH over the
//       points, core distances (k > 1: distance to the k-th nearest point, the point itself included), Boruvka
//       rounds until one component is left, edges back in the caller's indices
//   cluster/detail/ArborX_BoruvkaHelpers.hpp        DirectedEdge order (:37-105: weight, then the unordered pair of
//       LEAF POSITIONS, then the direction bit), component-aware nearest neighbour with `<=` pruning (:160-312),
//       retrieveEdges (:336-372), UpdateComponentsAndEdges (:381-470), finalizeEdges (:472-488), the Morton-neighbour
//       bound on every component's shortest edge (:735-778)
//   cluster/detail/ArborX_MutualReachabilityDistance.hpp:27-77   max(core_i, core_j, d)
//   spatial/detail/ArborX_TreeNodeLabeling.hpp:27-93  parents of nodes; an internal node carries a component label
//       iff all leaves below it do
//   cluster/ArborX_Dendrogram.hpp:47-76 + detail/ArborX_DendrogramHelpers.hpp:31-80   UNION_FIND dendrogram: edges
//       sorted by weight on the device, the union-find loop itself on the host (in the reference as well: it is a
//       sequential algorithm)
//   cluster/ArborX_HDBSCAN.hpp:29-53                hdbscan(space, points, core_min_size, implementation): BORUVKA =
//       MinimumSpanningTree in HDBSCAN mode, which grows the dendrogram with the rounds and never leaves the device
//       (MinimumSpanningTree.hpp:176-230,266-297, BoruvkaHelpers.hpp:439-447,490-733); UNION_FIND = MST(k) + the
//       dendrogram above
//
// Under the strict edge order above the minimum spanning tree is unique, so the edge SET is comparable bit for bit
// with the reference's (the leaf positions are: the trees are bit-identical); the order of the
// edges in the output array is the order components happen to append them and carries no meaning (the reference's
// tests sort before comparing).
//
// Shape here: the tree is the library's Node64 tree (both children's boxes in one 64-byte record), one thread per
// sorted leaf position (threads of a warp are Morton neighbours and walk the same subtrees), ordered descent with
// a (distance, node) stack like the kNN kernel, component labels of internal nodes in a side array indexed like
// the records.  Weights are compared as the reference compares them (square roots taken, not squared distances:
// two different squared distances can round to the same weight, and then the pair decides).
#include "abx_traverse.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

namespace abx
{
namespace
{

constexpr int kIndeterminate = -1; // TreeNodeLabeling.hpp:55
constexpr int kUntouched = -2;     // :56
constexpr unsigned kInfBits = 0x7f800000u;

// BoruvkaHelpers.hpp:76-99: | 0 | 31 bits smaller position | 31 bits larger position | direction |
__device__ __forceinline__ unsigned long long edgeKey(int source, int target)
{
  unsigned long long const lo = (unsigned)min(source, target), hi = (unsigned)max(source, target);
  return (lo << 32) | (hi << 1) | (source < target ? 0ull : 1ull);
}
__device__ __forceinline__ int keySource(unsigned long long key)
{
  int const lo = (int)((key >> 32) & 0x7fffffffull), hi = (int)((key >> 1) & 0x7fffffffull);
  return (key & 1ull) ? hi : lo;
}
__device__ __forceinline__ int keyTarget(unsigned long long key)
{
  int const lo = (int)((key >> 32) & 0x7fffffffull), hi = (int)((key >> 1) & 0x7fffffffull);
  return (key & 1ull) ? lo : hi;
}

// TreeNodeLabeling.hpp:27-42
__global__ void findParentsKernel(Node64 const *__restrict__ nodes, int n, int *__restrict__ par_int,
                                  int *__restrict__ par_leaf)
{
  int const k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n - 1)
    return;
  float4 const *f = reinterpret_cast<float4 const *>(nodes + k);
  int const lref = __float_as_int(__ldg(f).w), rref = __float_as_int(__ldg(f + 1).w);
  int const rl = __float_as_int(__ldg(f + 2).w), rr = __float_as_int(__ldg(f + 3).w);
  if (refIsLeaf(lref))
    par_leaf[rl] = k;
  else
    par_int[lref] = k;
  if (refIsLeaf(rref))
    par_leaf[rr] = k;
  else
    par_int[rref] = k;
  if (k == 0)
    par_int[0] = -1;
}

__global__ void fillIntKernel(int *__restrict__ a, int64_t n, int v)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    a[i] = v;
}
__global__ void iotaKernel(int *__restrict__ a, int64_t n)
{
  int64_t const i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    a[i] = (int)i;
}

// TreeNodeLabeling.hpp:44-93: every leaf walks up; the first arrival at a node leaves its label and stops, the
// second compares and continues, so a node is settled after both of its children
__global__ void reduceLabelsKernel(int n, int const *__restrict__ par_int, int const *__restrict__ par_leaf,
                                   int const *__restrict__ lab_leaf, int *lab_int)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int label = lab_leaf[i];
  int node = par_leaf[i];
  while (true)
  {
    int const old = atomicCAS(&lab_int[node], kUntouched, label);
    if (old == kUntouched)
      break;
    if (old != label)
    {
      label = kIndeterminate;
      lab_int[node] = kIndeterminate;
    }
    if (node == 0)
      break;
    node = par_int[node];
  }
}

// one component per round: weight of its shortest outgoing edge (float bits: weights are non-negative, so the
// unsigned order is the float order), then the smallest pair key among the edges of that weight
__global__ void resetRoundKernel(int n, unsigned *__restrict__ comp_w, unsigned long long *__restrict__ comp_key,
                                 unsigned *__restrict__ radii)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  comp_w[i] = kInfBits;
  comp_key[i] = ~0ull;
  radii[i] = kInfBits;
}

__device__ __forceinline__ float pointDist(float4 a, float4 b)
{
  // Distance.hpp:54-70
  float const tx = __fsub_rn(b.x, a.x), ty = __fsub_rn(b.y, a.y), tz = __fsub_rn(b.z, a.z);
  float d2 = __fmul_rn(tx, tx);
  d2 = __fadd_rn(d2, __fmul_rn(ty, ty));
  d2 = __fadd_rn(d2, __fmul_rn(tz, tz));
  return __fsqrt_rn(d2);
}

// BoruvkaHelpers.hpp:735-778: Morton neighbours in different components bound both components' shortest edges
template <bool MUTUAL>
__global__ void neighbourBoundsKernel(int n, float4 const *__restrict__ leaf_box, int const *__restrict__ lab_leaf,
                                      float const *__restrict__ core_pos, unsigned *__restrict__ radii)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;
  int const li = lab_leaf[i], lj = lab_leaf[i + 1];
  if (li == lj)
    return;
  float r = pointDist(__ldg(leaf_box + i), __ldg(leaf_box + i + 1));
  if (MUTUAL)
    r = fmaxf(fmaxf(core_pos[i], core_pos[i + 1]), r);
  atomicMin(&radii[li], __float_as_uint(r));
  atomicMin(&radii[lj], __float_as_uint(r));
}

// BoruvkaHelpers.hpp:160-312: the nearest point of another component, for the point at sorted position i
template <bool MUTUAL>
__global__ void __launch_bounds__(kThreads)
    componentNearestKernel(Node64 const *__restrict__ nodes, float4 const *__restrict__ leaf_box, int n,
                           int const *__restrict__ lab_leaf, int const *__restrict__ lab_int,
                           float const *__restrict__ core_pos, unsigned const *__restrict__ radii,
                           unsigned *__restrict__ comp_w, unsigned *__restrict__ best_w,
                           unsigned long long *__restrict__ best_key)
{
  int const i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n)
    return;
  float4 const me = __ldg(leaf_box + i);
  int const component = lab_leaf[i];
  float const core_i = MUTUAL ? core_pos[i] : 0.f;
  float radius = __uint_as_float(radii[component]);
  float bw = __int_as_float(kInfBits);
  unsigned long long bk = ~0ull;

  auto offer = [&](int pos, float d) {
    float const w = MUTUAL ? fmaxf(fmaxf(core_i, core_pos[pos]), d) : d;
    unsigned long long const key = edgeKey(i, pos);
    if (w < bw || (w == bw && key < bk))
    {
      bw = w;
      bk = key;
      radius = w;
    }
  };

  unsigned long long stack[kStackSize];
  int sp = 0;
  int node = 0;
  while (true)
  {
    float4 const *f = reinterpret_cast<float4 const *>(nodes + node);
    float4 const a0 = __ldg(f), a1 = __ldg(f + 1), a2 = __ldg(f + 2), a3 = __ldg(f + 3);
    int const lref = __float_as_int(a0.w), rref = __float_as_int(a1.w);
    int const rl = __float_as_int(a2.w), rr = __float_as_int(a3.w);
    float const dl = __fsqrt_rn(pointBoxDist2(me.x, me.y, me.z, a0.x, a0.y, a0.z, a1.x, a1.y, a1.z));
    float const dr = __fsqrt_rn(pointBoxDist2(me.x, me.y, me.z, a2.x, a2.y, a2.z, a3.x, a3.y, a3.z));
    bool const l_leaf = refIsLeaf(lref), r_leaf = refIsLeaf(rref);
    bool go_l = false, go_r = false;
    // `<=`, not `<`: equidistant candidates must all be seen, the pair key decides among them (:206-211)
    if (dl <= radius && (l_leaf ? lab_leaf[rl] : lab_int[lref]) != component)
    {
      if (l_leaf)
        offer(rl, dl);
      else
        go_l = true;
    }
    if (dr <= radius && (r_leaf ? lab_leaf[rr] : lab_int[rref]) != component)
    {
      if (r_leaf)
        offer(rr, dr);
      else
        go_r = true;
    }
    if (go_l || go_r)
    {
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
    bool popped = false;
    while (sp > 0)
    {
      unsigned long long const e = stack[--sp];
      if (__uint_as_float((unsigned)(e >> 32)) <= radius)
      {
        node = (int)(unsigned)e;
        popped = true;
        break;
      }
    }
    if (!popped)
      break;
  }
  best_w[i] = __float_as_uint(bw);
  best_key[i] = bk;
  // :300-310: the comparison before the atomic keeps large components from hammering one address
  if (bw < __int_as_float(kInfBits) && __float_as_uint(bw) <= comp_w[component])
    atomicMin(&comp_w[component], __float_as_uint(bw));
}

// retrieveEdges (:336-372): among a component's candidates of the winning weight, the smallest pair key
__global__ void componentEdgeKernel(int n, int const *__restrict__ lab_leaf, unsigned const *__restrict__ comp_w,
                                    unsigned const *__restrict__ best_w,
                                    unsigned long long const *__restrict__ best_key,
                                    unsigned long long *__restrict__ comp_key)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int const c = lab_leaf[i];
  unsigned const w = best_w[i];
  if (w != kInfBits && w == comp_w[c])
    atomicMin(&comp_key[c], best_key[i]);
}

// UpdateComponentsAndEdges::computeNextComponent (:391-405): a pair of components that chose each other is
// resolved to the smaller label
__device__ __forceinline__ int nextComponent(int c, int const *__restrict__ lab_leaf,
                                             unsigned long long const *__restrict__ comp_key)
{
  // a component without an outgoing edge (possible only with non-finite coordinates) stays where it is: the
  // round then merges nothing and the host reports it
  unsigned long long const kc = comp_key[c];
  if (kc == ~0ull)
    return c;
  int const next = lab_leaf[keyTarget(kc)];
  unsigned long long const kn = comp_key[next];
  if (kn == ~0ull)
    return next;
  int const next_next = lab_leaf[keyTarget(kn)];
  return next_next != c ? next : min(c, next);
}

// UnidirectionalEdgesTag (:422-437)
__global__ void appendEdgesKernel(int n, int const *__restrict__ lab_leaf,
                                  unsigned long long const *__restrict__ comp_key,
                                  unsigned const *__restrict__ comp_w, int *__restrict__ num_edges,
                                  int2 *__restrict__ edge_pos, float *__restrict__ weights,
                                  int *__restrict__ edges_mapping /* HDBSCAN mode: component -> its edge, or null */)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || lab_leaf[i] != i)
    return;
  if (nextComponent(i, lab_leaf, comp_key) == i)
    return;
  int const e = atomicAdd(num_edges, 1);
  unsigned long long const key = comp_key[i];
  edge_pos[e] = make_int2(keySource(key), keyTarget(key));
  weights[e] = __uint_as_float(comp_w[i]);
  if (edges_mapping)
    edges_mapping[i] = e;
}

// LabelsTag (:415-420) into a second array (the reference updates in place; the chains it follows end at the same
// component either way)
__global__ void updateLabelsKernel(int n, int const *__restrict__ lab_leaf,
                                   unsigned long long const *__restrict__ comp_key, int *__restrict__ lab_next)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int prev = lab_leaf[i], next;
  while ((next = nextComponent(prev, lab_leaf, comp_key)) != prev)
    prev = next;
  lab_next[i] = next;
}

// finalizeEdges (:472-488)
__global__ void finalizeEdgesKernel(int64_t m, int2 const *__restrict__ edge_pos, uint32_t const *__restrict__ perm,
                                    int2 *__restrict__ edges)
{
  int64_t const e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m)
    return;
  int2 const p = edge_pos[e];
  edges[e] = make_int2((int)perm[p.x], (int)perm[p.y]);
}

// core distance of the point at sorted position pos = last entry of its kNN row (MaxDistance over the k nearest)
__global__ void coreDistanceKernel(int n, uint32_t const *__restrict__ perm, float const *__restrict__ rows,
                                   int32_t const *__restrict__ counts, int stride, float *__restrict__ core_pos)
{
  int const pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n)
    return;
  uint32_t const o = perm[pos];
  int const c = counts[o];
  core_pos[pos] = c > 0 ? rows[(size_t)o * stride + (c - 1)] : 0.f;
}

__global__ void gatherEdgesKernel(int64_t m, uint32_t const *__restrict__ order, int2 const *__restrict__ edges,
                                  int2 *__restrict__ sorted_edges)
{
  int64_t const e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m)
    sorted_edges[e] = edges[order[e]];
}
__global__ void weightBitsKernel(int64_t m, float const *__restrict__ w, unsigned *__restrict__ bits, int *bad)
{
  int64_t const e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m)
    return;
  unsigned const b = __float_as_uint(w[e]);
  // finite, non-negative weights order like their bit patterns (BoruvkaHelpers.hpp:567-573 asserts the same)
  if (b >= kInfBits && b != 0x80000000u)
    *bad = 1;
  bits[e] = b == 0x80000000u ? 0u : b;
}

// ---- BoruvkaMode::HDBSCAN: the dendrogram grows with the rounds (MinimumSpanningTree.hpp:176-230,266-297) -----------
constexpr int kRootChain = -2;   // BoruvkaHelpers.hpp:34
constexpr int kFollowChain = -3; // :35

// WeightedEdge order (WeightedEdge.hpp:30-50): weight, smaller vertex, larger vertex -- vertices are leaf positions here
__device__ __forceinline__ bool edgeLess(float wa, int2 a, float wb, int2 b)
{
  if (wa != wb)
    return wa < wb;
  int const amin = min(a.x, a.y), bmin = min(b.x, b.y);
  if (amin != bmin)
    return amin < bmin;
  return max(a.x, a.y) < max(b.x, b.y);
}

// BidirectionalEdgesTag (BoruvkaHelpers.hpp:439-447): of two components that chose each other only the larger label
// appended the edge; the smaller one takes its index from the partner
__global__ void sharedEdgeMappingKernel(int n, int const *__restrict__ lab_leaf,
                                        unsigned long long const *__restrict__ comp_key, int *edges_mapping)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || lab_leaf[i] != i)
    return;
  if (comp_key[i] == ~0ull || nextComponent(i, lab_leaf, comp_key) != i)
    return;
  edges_mapping[i] = edges_mapping[lab_leaf[keyTarget(comp_key[i])]];
}

// assignVertexParents (:527-548), first round: every vertex is a component, its parent is the edge it picked
__global__ void vertexParentsKernel(int n, int const *__restrict__ lab_leaf,
                                    unsigned long long const *__restrict__ comp_key,
                                    int const *__restrict__ edges_mapping, uint32_t const *__restrict__ perm,
                                    int *__restrict__ parents)
{
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || comp_key[e] == ~0ull)
    return;
  int const i = lab_leaf[keySource(comp_key[e])];
  parents[(int)perm[i] + (n - 1)] = edges_mapping[i];
}

// updateSidedParents (:490-525): the edges of the previous round hang below the edge their (merged) component picks
// in this round -- on its source or its target side -- or, when they are heavier than it, follow its chain upwards
__global__ void sidedParentsKernel(int first, int last, int const *__restrict__ lab_leaf,
                                   int2 const *__restrict__ edge_pos, float const *__restrict__ weights,
                                   int const *__restrict__ edges_mapping, int *__restrict__ sided_parents)
{
  int const e = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= last)
    return;
  int2 const edge = edge_pos[e];
  int const component = lab_leaf[edge.x];
  int const alpha = edges_mapping[component];
  if (alpha < 0)
  {
    // the component picked nothing this round (only with non-finite coordinates; the host reports the round)
    sided_parents[e] = kRootChain;
    return;
  }
  int2 const alpha_edge = edge_pos[alpha];
  if (edgeLess(weights[e], edge, weights[alpha], alpha_edge))
    sided_parents[e] = 2 * alpha + (lab_leaf[alpha_edge.x] == component ? 1 : 0);
  else
    sided_parents[e] = kFollowChain - alpha;
}

// computeParentsAndReorderEdges (:550-733), step 1: (sided parent, weight) in one 64-bit key
__global__ void chainKeysKernel(int m, int const *__restrict__ sided_parents, int2 const *__restrict__ edge_pos,
                                float const *__restrict__ weights, unsigned long long *__restrict__ keys)
{
  int const e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m)
    return;
  long long key = sided_parents[e];
  if (key <= kFollowChain)
  {
    int2 const edge = edge_pos[e];
    float const w = weights[e];
    int next = kFollowChain - (int)key;
    while (true)
    {
      key = sided_parents[next];
      if (key <= kFollowChain)
        next = kFollowChain - (int)key;
      else if (key >= 0)
      {
        next = (int)(key / 2);
        if (edgeLess(w, edge, weights[next], edge_pos[next]))
          break;
      }
      else if (key == kRootChain)
        break;
    }
  }
  if (key == kRootChain)
    key = 0x7fffffff;
  keys[e] = ((unsigned long long)key << 32) | (unsigned long long)__float_as_uint(weights[e]);
}

// step 2 (:625-660): within a chain the smallest edge goes first even among equal weights, so that no edge ends up
// with three children
__global__ void fixSameWeightOrderKernel(int m, unsigned long long const *__restrict__ keys, uint32_t *permute,
                                         int2 const *__restrict__ edge_pos, float const *__restrict__ weights)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m - 1)
    return;
  unsigned long long const key = keys[i];
  if (i != 0 && (keys[i - 1] >> 32) == (key >> 32))
    return;
  int best = i;
  for (int k = i + 1; k < m && keys[k] == key; ++k)
  {
    uint32_t const a = permute[k], b = permute[best];
    if (edgeLess(weights[a], edge_pos[a], weights[b], edge_pos[b]))
      best = k;
  }
  if (best != i)
  {
    uint32_t const tmp = permute[i];
    permute[i] = permute[best];
    permute[best] = tmp;
  }
}

__global__ void inversePermutationKernel(int m, uint32_t const *__restrict__ permute, int *__restrict__ rev)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
    rev[permute[i]] = i;
}

// steps 3-5 (:662-712): parents of the vertices and of the edges in the new edge order, the edges themselves
__global__ void dendrogramParentsKernel(int n, unsigned long long const *__restrict__ keys,
                                        uint32_t const *__restrict__ permute, int const *__restrict__ rev,
                                        int2 const *__restrict__ edge_pos, float const *__restrict__ weights,
                                        uint32_t const *__restrict__ perm, int *__restrict__ parents,
                                        int2 *__restrict__ edges_out, float *__restrict__ weights_out,
                                        float *__restrict__ heights)
{
  int const m = n - 1;
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    parents[m + i] = rev[parents[m + i]];
  if (i >= m)
    return;
  if (i == m - 1)
    parents[i] = -1;
  else if ((keys[i] >> 32) == (keys[i + 1] >> 32))
    parents[i] = i + 1;
  else
    parents[i] = rev[(int)((keys[i] >> 32) / 2)];
  uint32_t const old = permute[i];
  int2 const p = edge_pos[old];
  float const w = weights[old];
  edges_out[i] = make_int2((int)perm[p.x], (int)perm[p.y]); // finalizeEdges (:472-488)
  weights_out[i] = w;
  heights[i] = w; // MinimumSpanningTree.hpp:288-295
}

template <bool MUTUAL>
abx_status boruvka(cudaStream_t s, abx_bvh *t, float const *core_pos, int2 *edges, float *weights_out, int *iterations,
                   int *dendrogram_parents, float *dendrogram_heights)
{
  int const n = (int)t->n;
  int const grid = divUp(n, 256);
  // HDBSCAN mode: the rounds also record which edge every component picked; the edges leave in (chain, weight) order
  bool const hd = dendrogram_parents != nullptr;
  TempBuffer<int> edges_mapping, sided_parents;
  TempBuffer<float> round_weights;
  float *weights = weights_out;
  if (hd)
  {
    ABX_TRY(edges_mapping.alloc((size_t)n, s));
    ABX_CUDA_TRY(cudaMemsetAsync(edges_mapping.ptr, 0xff, sizeof(int) * (size_t)n, s)); // -1: no edge picked yet
    ABX_TRY(sided_parents.alloc((size_t)n, s));
    ABX_TRY(round_weights.alloc((size_t)n, s));
    weights = round_weights.ptr; // append order here, reordered into weights_out at the end
  }
  int edges_start = 0, edges_end = 0;
  TempBuffer<int> par_int, par_leaf, lab_a, lab_b, lab_int, num_edges;
  TempBuffer<unsigned> comp_w, radii, best_w;
  TempBuffer<unsigned long long> comp_key, best_key;
  TempBuffer<int2> edge_pos;
  ABX_TRY(par_int.alloc((size_t)n, s));
  ABX_TRY(par_leaf.alloc((size_t)n, s));
  ABX_TRY(lab_a.alloc((size_t)n, s));
  ABX_TRY(lab_b.alloc((size_t)n, s));
  ABX_TRY(lab_int.alloc((size_t)n, s));
  ABX_TRY(num_edges.alloc(1, s));
  ABX_TRY(comp_w.alloc((size_t)n, s));
  ABX_TRY(radii.alloc((size_t)n, s));
  ABX_TRY(best_w.alloc((size_t)n, s));
  ABX_TRY(comp_key.alloc((size_t)n, s));
  ABX_TRY(best_key.alloc((size_t)n, s));
  ABX_TRY(edge_pos.alloc((size_t)n, s));
  ABX_LAUNCH(findParentsKernel, grid, 256, 0, s, t->nodes, n, par_int.ptr, par_leaf.ptr);
  ABX_LAUNCH(iotaKernel, grid, 256, 0, s, lab_a.ptr, (int64_t)n);
  ABX_CUDA_TRY(cudaMemsetAsync(num_edges.ptr, 0, sizeof(int), s));
  int *lab = lab_a.ptr, *lab_next = lab_b.ptr;
  int components = n, rounds = 0;
  while (components > 1)
  {
    ++rounds;
    ABX_LAUNCH(fillIntKernel, grid, 256, 0, s, lab_int.ptr, (int64_t)n - 1, kUntouched);
    ABX_LAUNCH(reduceLabelsKernel, grid, 256, 0, s, n, par_int.ptr, par_leaf.ptr, lab, lab_int.ptr);
    ABX_LAUNCH(resetRoundKernel, grid, 256, 0, s, n, comp_w.ptr, comp_key.ptr, radii.ptr);
    ABX_LAUNCH_TAGGED("neighbourBoundsKernel", (neighbourBoundsKernel<MUTUAL>), grid, 256, 0, s, n, t->leaf_box, lab,
                      core_pos, radii.ptr);
    ABX_LAUNCH_TAGGED("componentNearestKernel", (componentNearestKernel<MUTUAL>), divUp(n, kThreads), kThreads, 0, s,
                      t->nodes, t->leaf_box, n, lab, lab_int.ptr, core_pos, radii.ptr, comp_w.ptr, best_w.ptr,
                      best_key.ptr);
    ABX_LAUNCH(componentEdgeKernel, grid, 256, 0, s, n, lab, comp_w.ptr, best_w.ptr, best_key.ptr, comp_key.ptr);
    ABX_LAUNCH(appendEdgesKernel, grid, 256, 0, s, n, lab, comp_key.ptr, comp_w.ptr, num_edges.ptr, edge_pos.ptr,
               weights, hd ? edges_mapping.ptr : (int *)nullptr);
    if (hd)
    {
      ABX_LAUNCH(sharedEdgeMappingKernel, grid, 256, 0, s, n, lab, comp_key.ptr, edges_mapping.ptr);
      if (rounds == 1)
        ABX_LAUNCH(vertexParentsKernel, grid, 256, 0, s, n, lab, comp_key.ptr, edges_mapping.ptr, t->perm,
                   dendrogram_parents);
      else if (edges_end > edges_start)
        ABX_LAUNCH(sidedParentsKernel, divUp(edges_end - edges_start, 256), 256, 0, s, edges_start, edges_end, lab,
                   edge_pos.ptr, weights, edges_mapping.ptr, sided_parents.ptr);
    }
    ABX_LAUNCH(updateLabelsKernel, grid, 256, 0, s, n, lab, comp_key.ptr, lab_next);
    int h_edges = 0;
    ABX_CUDA_TRY(cudaMemcpyAsync(&h_edges, num_edges.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
    ABX_CUDA_TRY(cudaStreamSynchronize(s)); // the reference's per-round fence (MinimumSpanningTree.hpp:196-198)
    std::swap(lab, lab_next);
    int const left = n - h_edges;
    if (left >= components)
    {
      setError("MinimumSpanningTree: a Boruvka round merged nothing (non-finite coordinates?)");
      return ABX_ERR_ARG;
    }
    components = left;
    edges_start = edges_end;
    edges_end = h_edges;
  }
  if (iterations)
    *iterations = rounds;
  if (!hd)
  {
    ABX_LAUNCH(finalizeEdgesKernel, divUp(n - 1, 256), 256, 0, s, (int64_t)n - 1, edge_pos.ptr, t->perm, edges);
    return ABX_OK;
  }
  // the edges of the last round form the root chain (MinimumSpanningTree.hpp:271-275)
  int const m = n - 1;
  if (edges_end > edges_start)
    ABX_LAUNCH(fillIntKernel, divUp(edges_end - edges_start, 256), 256, 0, s, sided_parents.ptr + edges_start,
               (int64_t)(edges_end - edges_start), kRootChain);
  TempBuffer<unsigned long long> keys;
  TempBuffer<uint32_t> permute;
  TempBuffer<int> rev;
  ABX_TRY(keys.alloc((size_t)m, s));
  ABX_TRY(permute.alloc((size_t)m, s));
  ABX_TRY(rev.alloc((size_t)m, s));
  ABX_LAUNCH(chainKeysKernel, divUp(m, 256), 256, 0, s, m, sided_parents.ptr, edge_pos.ptr, weights, keys.ptr);
  ABX_TRY(sortPairsU64(s, (uint64_t *)keys.ptr, permute.ptr, m, true, 63));
  ABX_LAUNCH(fixSameWeightOrderKernel, divUp(m, 256), 256, 0, s, m, keys.ptr, permute.ptr, edge_pos.ptr, weights);
  ABX_LAUNCH(inversePermutationKernel, divUp(m, 256), 256, 0, s, m, permute.ptr, rev.ptr);
  ABX_LAUNCH(dendrogramParentsKernel, grid, 256, 0, s, n, keys.ptr, permute.ptr, rev.ptr, edge_pos.ptr, weights, t->perm,
             dendrogram_parents, edges, weights_out, dendrogram_heights);
  return ABX_OK;
}

} // namespace

// edges2: (n - 1) x (source, target), weights: n - 1, both on the device
abx_status minimumSpanningTree(cudaStream_t s, float const *xyz, int64_t n, int32_t k, int32_t *edges2, float *weights,
                               int *iterations, int32_t *dendrogram_parents, float *dendrogram_heights)
{
  if (iterations)
    *iterations = 0;
  if (n == 1 && dendrogram_parents)
  {
    int const root = -1;
    ABX_CUDA_TRY(cudaMemcpyAsync(dendrogram_parents, &root, sizeof(int), cudaMemcpyHostToDevice, s));
  }
  if (n < 2)
    return ABX_OK;
  if (n >= (int64_t)1 << 30)
  {
    setError("MinimumSpanningTree: n must be < 2^30");
    return ABX_ERR_ARG;
  }
  abx_bvh *tree = nullptr;
  ABX_TRY(buildTree(s, ABX_PRIM_POINT3F, xyz, n, nullptr, &tree));
  struct Guard
  {
    abx_bvh *t;
    ~Guard() { abx_bvh_destroy(t); }
  } guard{tree};
  if (k <= 1)
    return boruvka<false>(s, tree, nullptr, (int2 *)edges2, weights, iterations, dendrogram_parents, dendrogram_heights);
  // core distances: nearest(point, k) of every point in the tree's own order (MinimumSpanningTree.hpp:70-81)
  int const stride = (int)std::min<int64_t>(k, n);
  TempBuffer<float> core_pos;
  ABX_TRY(core_pos.alloc((size_t)n, s));
  {
    TempBuffer<uint32_t> idx;
    TempBuffer<float> dist;
    TempBuffer<int32_t> counts;
    ABX_TRY(idx.alloc((size_t)n * stride, s));
    ABX_TRY(dist.alloc((size_t)n * stride, s));
    ABX_TRY(counts.alloc((size_t)n, s));
    ABX_TRY(nearestQuery(s, tree, xyz, n, k, nullptr, tree->perm, nullptr, n * stride, counts.ptr, idx.ptr, dist.ptr));
    ABX_LAUNCH(coreDistanceKernel, divUp(n, 256), 256, 0, s, (int)n, tree->perm, dist.ptr, counts.ptr, stride,
               core_pos.ptr);
  }
  return boruvka<true>(s, tree, core_pos.ptr, (int2 *)edges2, weights, iterations, dendrogram_parents,
                       dendrogram_heights);
}

// parents: 2 m + 1 entries (edges in ascending weight order first, then the m + 1 vertices); heights: m
abx_status dendrogramUnionFind(cudaStream_t s, int32_t const *edges2, float const *weights, int64_t m, int32_t *parents,
                               float *heights)
{
  if (m < 0 || m >= (int64_t)1 << 30)
  {
    setError("Dendrogram: bad edge count");
    return ABX_ERR_ARG;
  }
  if (m == 0)
  {
    int const root = -1; // a single vertex
    ABX_CUDA_TRY(cudaMemcpyAsync(parents, &root, sizeof(int), cudaMemcpyHostToDevice, s));
    return ABX_OK;
  }
  // Dendrogram.hpp:66-68: sortByKey(weights, edges), stable
  TempBuffer<unsigned> bits;
  TempBuffer<uint32_t> order;
  TempBuffer<int2> sorted;
  TempBuffer<int> bad;
  ABX_TRY(bits.alloc((size_t)m, s));
  ABX_TRY(order.alloc((size_t)m, s));
  ABX_TRY(sorted.alloc((size_t)m, s));
  ABX_TRY(bad.alloc(1, s));
  ABX_CUDA_TRY(cudaMemsetAsync(bad.ptr, 0, sizeof(int), s));
  ABX_LAUNCH(weightBitsKernel, divUp(m, 256), 256, 0, s, m, weights, bits.ptr, bad.ptr);
  ABX_TRY(sortPairsU32(s, bits.ptr, order.ptr, m, true, 32));
  ABX_LAUNCH(gatherEdgesKernel, divUp(m, 256), 256, 0, s, m, order.ptr, (int2 const *)edges2, sorted.ptr);
  ABX_CUDA_TRY(cudaMemcpyAsync(heights, bits.ptr, sizeof(float) * (size_t)m, cudaMemcpyDeviceToDevice, s));
  std::vector<int2> h_edges((size_t)m);
  int h_bad = 0;
  ABX_CUDA_TRY(cudaMemcpyAsync(h_edges.data(), sorted.ptr, sizeof(int2) * (size_t)m, cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  if (h_bad)
  {
    setError("Dendrogram: edge weights must be finite and non-negative");
    return ABX_ERR_ARG;
  }
  // DendrogramHelpers.hpp:31-80 (dendrogramUnionFindHost): sequential by construction, on the host in the reference
  int64_t const nv = m + 1;
  std::vector<int> labels((size_t)nv), set_edges((size_t)nv, -1), h_parents((size_t)(2 * m + 1), -1);
  std::iota(labels.begin(), labels.end(), 0);
  auto find = [&](int i) {
    // UnionFind::representative with path halving (UnionFind.hpp:113-128)
    int curr = labels[i];
    if (curr != i)
    {
      int next, prev = i;
      while (curr > (next = labels[curr]))
      {
        labels[prev] = next;
        prev = curr;
        curr = next;
      }
    }
    return curr;
  };
  for (int64_t e = 0; e < m; ++e)
  {
    int const a = h_edges[e].x, b = h_edges[e].y;
    if (a < 0 || b < 0 || a >= nv || b >= nv)
    {
      setError("Dendrogram: vertex index out of range");
      return ABX_ERR_ARG;
    }
    int const i = find(a), j = find(b);
    for (int v : {i, j})
    {
      int const child = set_edges[v];
      if (child != -1)
        h_parents[child] = (int)e;
      else
        h_parents[m + v] = (int)e;
    }
    int const lo = std::min(i, j), hi = std::max(i, j);
    labels[hi] = lo; // UnionFind::merge of two representatives (:139-181)
    set_edges[lo] = (int)e;
  }
  h_parents[m - 1] = -1; // root
  ABX_CUDA_TRY(cudaMemcpyAsync(parents, h_parents.data(), sizeof(int) * (size_t)(2 * m + 1), cudaMemcpyHostToDevice, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s)); // h_parents goes out of scope
  return ABX_OK;
}

} // namespace abx

using namespace abx;

extern "C" {

abx_status abx_mst_points3f(void *stream, const float *xyz_dev, int64_t n, int32_t k, int32_t *edges2_dev,
                            float *weights_dev, int32_t *iterations)
{
  ABX_TRY(ensureDevice());
  if (n < 0 || k < 1 || (n > 0 && !xyz_dev) || (n > 1 && (!edges2_dev || !weights_dev)))
  {
    setError("MinimumSpanningTree: bad argument (n >= 0, k >= 1, non-null arrays)");
    return ABX_ERR_ARG;
  }
  int it = 0;
  abx_status const st = minimumSpanningTree((cudaStream_t)stream, xyz_dev, n, k, edges2_dev, weights_dev, &it);
  if (iterations)
    *iterations = it;
  return st;
}

abx_status abx_mst_points3f_host(void *stream, const float *xyz_host, int64_t n, int32_t k, int32_t *edges2_host,
                                 float *weights_host, int32_t *iterations)
{
  ABX_TRY(ensureDevice());
  if (n < 0 || k < 1 || (n > 0 && !xyz_host) || (n > 1 && (!edges2_host || !weights_host)))
  {
    setError("MinimumSpanningTree: bad argument (n >= 0, k >= 1, non-null arrays)");
    return ABX_ERR_ARG;
  }
  if (iterations)
    *iterations = 0;
  if (n < 2)
    return ABX_OK;
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> xyz, w;
  TempBuffer<int32_t> e;
  ABX_TRY(xyz.alloc(3 * (size_t)n, s));
  ABX_TRY(w.alloc((size_t)n - 1, s));
  ABX_TRY(e.alloc(2 * ((size_t)n - 1), s));
  ABX_CUDA_TRY(cudaMemcpyAsync(xyz.ptr, xyz_host, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  int it = 0;
  ABX_TRY(minimumSpanningTree(s, xyz.ptr, n, k, e.ptr, w.ptr, &it));
  ABX_CUDA_TRY(cudaMemcpyAsync(edges2_host, e.ptr, sizeof(int32_t) * 2 * ((size_t)n - 1), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaMemcpyAsync(weights_host, w.ptr, sizeof(float) * ((size_t)n - 1), cudaMemcpyDeviceToHost, s));
  ABX_CUDA_TRY(cudaStreamSynchronize(s));
  if (iterations)
    *iterations = it;
  return ABX_OK;
}

abx_status abx_dendrogram_union_find(void *stream, const int32_t *edges2_dev, const float *weights_dev,
                                     int64_t num_edges, int32_t *parents_dev, float *parent_heights_dev)
{
  ABX_TRY(ensureDevice());
  if (num_edges < 0 || !parents_dev || (num_edges > 0 && (!edges2_dev || !weights_dev || !parent_heights_dev)))
  {
    setError("Dendrogram: bad argument");
    return ABX_ERR_ARG;
  }
  return dendrogramUnionFind((cudaStream_t)stream, edges2_dev, weights_dev, num_edges, parents_dev, parent_heights_dev);
}

abx_status abx_mst_hdbscan_points3f(void *stream, const float *xyz_dev, int64_t n, int32_t k, int32_t *edges2_dev,
                                    float *weights_dev, int32_t *parents_dev, float *parent_heights_dev,
                                    int32_t *iterations)
{
  ABX_TRY(ensureDevice());
  if (n < 1 || k < 1 || !xyz_dev || !parents_dev || (n > 1 && (!edges2_dev || !weights_dev || !parent_heights_dev)))
  {
    setError("MinimumSpanningTree (HDBSCAN mode): bad argument (n >= 1, k >= 1, non-null arrays)");
    return ABX_ERR_ARG;
  }
  int it = 0;
  abx_status const st = minimumSpanningTree((cudaStream_t)stream, xyz_dev, n, k, edges2_dev, weights_dev, &it,
                                            parents_dev, parent_heights_dev);
  if (iterations)
    *iterations = it;
  return st;
}

abx_status abx_hdbscan_points3f(void *stream, const float *xyz_dev, int64_t n, int32_t core_min_size,
                                int dendrogram_impl, int32_t *parents_dev, float *parent_heights_dev)
{
  ABX_TRY(ensureDevice());
  if (n < 1 || core_min_size < 1 || !xyz_dev || !parents_dev || (n > 1 && !parent_heights_dev) ||
      (dendrogram_impl != ABX_DENDROGRAM_BORUVKA && dendrogram_impl != ABX_DENDROGRAM_UNION_FIND))
  {
    setError("hdbscan: bad argument (n >= 1, core_min_size >= 1, non-null arrays, a known dendrogram implementation)");
    return ABX_ERR_ARG;
  }
  cudaStream_t s = (cudaStream_t)stream;
  TempBuffer<float> w;
  TempBuffer<int32_t> e;
  ABX_TRY(w.alloc((size_t)std::max<int64_t>(n - 1, 1), s));
  ABX_TRY(e.alloc(2 * (size_t)std::max<int64_t>(n - 1, 1), s));
  if (dendrogram_impl == ABX_DENDROGRAM_BORUVKA) // HDBSCAN.hpp:41-47: the hybrid, all on the device
    return minimumSpanningTree(s, xyz_dev, n, core_min_size, e.ptr, w.ptr, nullptr, parents_dev, parent_heights_dev);
  ABX_TRY(minimumSpanningTree(s, xyz_dev, n, core_min_size, e.ptr, w.ptr, nullptr));
  return dendrogramUnionFind(s, e.ptr, w.ptr, n - 1, parents_dev, parent_heights_dev);
}

} // extern "C"
