/*
 * abx.h -- C ABI of the B200-native geometric-search engine (libabx.so).
 *
 * This is the drop-in boundary for ArborX's hot path: BVH construction,
 * spatial / nearest queries with CRS output, DBSCAN.  The reference has no FFI
 * (its boundary is a C++ template API), so every entry point below cites the
 * reference interface it replaces (paths relative to the reference's src/).
 * A header-only C++ facade with the reference's spellings sits on top of this
 * ABI (include/ArborX_B200.hpp); INTEGRATION.md shows the bindings.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross the boundary;
 *  - `stream` is a cudaStream_t passed as void* (the reference's execution
 *    space instance, e.g. spatial/ArborX_LinearBVH.hpp:68-73): every kernel of a
 *    call is enqueued on it;
 *  - `*_dev` pointers are device memory, `*_host` pointers host memory;
 *  - all functions return an abx_status; abx_last_error() gives the message of
 *    the last failure on the calling thread.  ABX_ERR_SEARCH corresponds to the
 *    reference throwing ArborX::SearchException (misc/ArborX_Exception.hpp:19-38);
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with ABX_ERR_CUDA.
 */
#ifndef ABX_H
#define ABX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABX_VERSION 100
#if defined(__GNUC__)
#define ABX_API __attribute__((visibility("default")))
#else
#define ABX_API
#endif

typedef int abx_status;
enum
{
  ABX_OK = 0,
  ABX_ERR_SEARCH = 1,    /* precondition failure == ArborX::SearchException */
  ABX_ERR_CUDA = 2,      /* CUDA runtime error / no device */
  ABX_ERR_PRECISION = 3, /* FDBSCAN-DenseBox loss-of-precision guard (cluster/detail/ArborX_CartesianGrid.hpp:121-136) */
  ABX_ERR_ARG = 4        /* malformed argument (null pointer, unknown enum) */
};

/* Primitive (indexable) layouts: geometry/ArborX_Point.hpp:24-38, ArborX_Box.hpp:32-67,
 * ArborX_Triangle.hpp:21-26 -- tightly packed float arrays. */
enum
{
  ABX_PRIM_POINT3F = 0, /* 3 floats: x y z */
  ABX_PRIM_BOX3F = 1,   /* 6 floats: min xyz, max xyz */
  ABX_PRIM_TRI3F = 2    /* 9 floats: a, b, c */
};
/* Predicate geometries: spatial/detail/ArborX_Predicates.hpp:83-101,130-147 */
enum
{
  ABX_PRED_SPHERE3F = 0, /* intersects(Sphere): 4 floats centre xyz, radius (geometry/ArborX_Sphere.hpp:25-47) */
  ABX_PRED_BOX3F = 1,    /* intersects(Box):    6 floats */
  ABX_PRED_POINT3F = 2,  /* intersects(Point) / nearest(Point, k): 3 floats */
  ABX_PRED_RAY3F = 3     /* intersects(Experimental::Ray): 6 floats origin xyz, direction xyz (any non-zero
                            vector; normalised in double like the Ray constructor, geometry/ArborX_Ray.hpp:47-55);
                            box and triangle primitives */
};

/* Experimental::TraversalPolicy (spatial/detail/ArborX_TraversalPolicy.hpp:19-48) */
typedef struct abx_policy
{
  int32_t buffer_size;     /* 0: two-pass; +b: soft preallocation; -b: hard (overflow -> ABX_ERR_SEARCH) */
  int32_t sort_predicates; /* non-zero: Morton32-sort predicates (default in the reference) */
} abx_policy;

/* Output allocation.  The reference re-allocates the caller's Views
 * (spatial/detail/ArborX_CrsGraphWrapperImpl.hpp:257,286,337,349); across a C ABI
 * the caller supplies the allocator instead.  `which`: 0 offsets, 1 indices / values,
 * 2 distances, 3 remote positions, 4 remote ranks (DistributedTree host results).  Must return device memory of at least `bytes` (may return NULL
 * for bytes == 0).  Passing a NULL allocator makes the library allocate with
 * cudaMallocAsync on `stream`; release those with abx_free(). */
typedef void *(*abx_alloc_fn)(void *user, int which, size_t bytes);

typedef struct abx_bvh abx_bvh;

ABX_API const char *abx_last_error(void);
ABX_API int abx_version(void);
/* number of kernel launches issued by this library on the calling process so far
 * (bench.py reports the delta over the timed region as gpu_launches) */
ABX_API int64_t abx_launch_count(void);
ABX_API abx_status abx_free(void *stream, void *ptr_dev);
/* The library recycles its device buffers through per-stream free lists; abx_trim()
 * synchronises the device and returns all cached blocks to the driver (bytes released). */
ABX_API int64_t abx_trim(void);
/* Per-kernel device timing (the analogue of the reference's Kokkos-Tools regions,
 * SURVEY.md section 5): CUDA events on the launching stream around every launch.
 * enable(1) clears and starts recording, enable(0) stops.  report() synchronises
 * and writes "name\tlaunches\ttotal_ms\tmax_ms\n" lines (max_ms: the longest single launch), most expensive first; it
 * returns the bytes needed including the NUL. */
ABX_API abx_status abx_profile_enable(int on);
ABX_API int64_t abx_profile_report(char *buf, int64_t capacity);

/* ---- BoundingVolumeHierarchy (spatial/ArborX_LinearBVH.hpp:50-142,171-256) ---- */
/* ctor BVH(space, values): builds over n primitives; leaf value = original index */
ABX_API abx_status abx_bvh_build(void *stream, int prim_kind, const void *prims_dev, int64_t n, abx_bvh **out);
/* same with primitives in host memory: H2D copy is part of the call */
ABX_API abx_status abx_bvh_build_host(void *stream, int prim_kind, const void *prims_host, int64_t n, abx_bvh **out);
/* triangles as vertex-index triples over a shared vertex array (the Triangles AccessTraits of
 * benchmarks/triangulated_surface_distance/triangulated_surface_distance.cpp:34-58): vertices_dev 3 floats each,
 * triangles_dev 3 int32 each; the tree is the one abx_bvh_build(ABX_PRIM_TRI3F) gives for the expanded corners */
ABX_API abx_status abx_bvh_build_indexed_triangles(void *stream, const float *vertices_dev, int64_t n_vertices,
                                                   const int32_t *triangles_dev, int64_t n, abx_bvh **out);
ABX_API abx_status abx_bvh_destroy(abx_bvh *bvh);
ABX_API int64_t abx_bvh_size(const abx_bvh *bvh);               /* size()  :75-76 */
ABX_API int abx_bvh_empty(const abx_bvh *bvh);                  /* empty() :78-79 */
ABX_API abx_status abx_bvh_bounds(abx_bvh *bvh, float out6[6]); /* bounds() :81-82; blocks like the ctor's root-box copy
                                                           (spatial/detail/ArborX_TreeConstruction.hpp:108-113) */
/* bytes of device memory held by the tree */
ABX_API int64_t abx_bvh_memory_bytes(const abx_bvh *bvh);

/* Structural parity hook (tests): writes the tree in the reference's node layout
 * (spatial/detail/ArborX_Node.hpp:24-44, ArborX_HappyTreeFriends.hpp:26-85).  All
 * pointers are device memory and may be NULL: leaf_rope[n], leaf_index[n],
 * left_child[n-1], rope[n-1], boxes6[6(n-1)], sorted_codes[n]. */
ABX_API abx_status abx_bvh_export_reference_layout(abx_bvh *bvh, void *stream, int32_t *leaf_rope, uint32_t *leaf_index,
                                           int32_t *left_child, int32_t *rope, float *boxes6, uint64_t *sorted_codes);

/* ---- query(space, predicates, indices, offsets, policy)  (ArborX_LinearBVH.hpp:90-110,
 *      ArborX_CrsGraphWrapper.hpp:22-35, detail/ArborX_CrsGraphWrapperImpl.hpp:148-446) ---- */
/* Spatial predicates -> CRS.  offsets has q+1 ints, row i = ORIGINAL query i. */
ABX_API abx_status abx_query_spatial_crs(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                 const abx_policy *policy, abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                 uint32_t **indices_dev, int64_t *nnz);
/* query(space, predicates, callback) with a counting callback: counts_dev[i] =
 * number of matches of query i; limit > 0 stops each query at `limit` matches
 * (CountUpToN early exit, cluster/detail/ArborX_FDBSCAN.hpp:31-46;
 * CallbackTreeTraversalControl, detail/ArborX_Callbacks.hpp:24-28). */
ABX_API abx_status abx_query_spatial_count(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                   int sort_predicates, int32_t limit, int32_t *counts_dev);
/* nearest(Point, k) -> CRS, rows ascending by distance (detail/ArborX_TreeTraversal.hpp:180-335).
 * k_per_query_dev may be NULL (uniform k).  distances_dev may be NULL. */
ABX_API abx_status abx_query_nearest_crs(abx_bvh *bvh, void *stream, const void *points_dev, int64_t q, int32_t k,
                                 const int32_t *k_per_query_dev, const abx_policy *policy, abx_alloc_fn alloc,
                                 void *user, int32_t **offsets_dev, uint32_t **indices_dev, float **distances_dev,
                                 int64_t *nnz);
/* nearest(Geometry, k) for the other predicate geometries of detail/ArborX_Predicates.hpp:58-80 over point and box
 * primitives: ABX_PRED_BOX3F (distance(Box, Box), geometry/algorithms/ArborX_Distance.hpp:166-197),
 * ABX_PRED_SPHERE3F (max(distance(centre, X) - r, 0), :83-108,199-209), ABX_PRED_RAY3F (distance(Ray, Box) = the
 * length along the ray to where it enters the box, geometry/ArborX_Ray.hpp:433-444: the reference's way of ray
 * casting with nearest queries; boxes the ray misses are at infinite distance and never reported, so rows can be
 * shorter than k), ABX_PRED_POINT3F (same as abx_query_nearest_crs).  Uniform k. */
ABX_API abx_status abx_query_nearest_geom_crs(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_dev,
                                              int64_t q, int32_t k, const abx_policy *policy, abx_alloc_fn alloc,
                                              void *user, int32_t **offsets_dev, uint32_t **indices_dev,
                                              float **distances_dev, int64_t *nnz);
/* Host-buffer variants (end-to-end path): predicates in host memory, results
 * copied into host arrays obtained from `alloc_host` (which: as above). */
ABX_API abx_status abx_query_spatial_crs_host(abx_bvh *bvh, void *stream, int pred_kind, const void *preds_host, int64_t q,
                                      const abx_policy *policy, abx_alloc_fn alloc_host, void *user,
                                      int32_t **offsets_host, uint32_t **indices_host, int64_t *nnz);
ABX_API abx_status abx_query_nearest_crs_host(abx_bvh *bvh, void *stream, const void *points_host, int64_t q, int32_t k,
                                      const abx_policy *policy, abx_alloc_fn alloc_host, void *user,
                                      int32_t **offsets_host, uint32_t **indices_host, float **distances_host,
                                      int64_t *nnz);

/* Experimental::HalfTraversal over point leaves with WithinRadiusGetter
 * (detail/ArborX_HalfTraversal.hpp:24-75, cluster/ArborX_DBSCAN.hpp:55-70): every
 * unordered pair of leaves within r exactly once, as (original index, original
 * index).  pairs_dev may be NULL (count only); at most `capacity` pairs are stored. */
ABX_API abx_status abx_half_traversal_pairs(abx_bvh *bvh, void *stream, float r, uint32_t *pairs_dev, int64_t capacity,
                                    int64_t *count);

/* Experimental::findHalfNeighborList / findFullNeighborList (spatial/detail/ArborX_NeighborList.hpp:47-192,
 * ArborX_ExpandHalfToFull.hpp:24-72): CRS neighbour lists of a point cloud within `radius`.  Half: every unordered
 * pair once, in the row of the point the half traversal reports second; full: every pair in both rows.  The order
 * inside a row is unspecified (the reference fills rows with atomics). */
ABX_API abx_status abx_find_half_neighbor_list(void *stream, const float *xyz_dev, int64_t n, float radius,
                                               abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                               uint32_t **indices_dev, int64_t *nnz);
ABX_API abx_status abx_find_full_neighbor_list(void *stream, const float *xyz_dev, int64_t n, float radius,
                                               abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                               uint32_t **indices_dev, int64_t *nnz);

/* ---- ArborX::BruteForce (spatial/ArborX_BruteForce.hpp:42-160, detail/ArborX_BruteForceImpl.hpp:40-233): the
 * interface of the BVH answered by testing every predicate against every primitive (tiles of primitives staged in
 * shared memory).  Point and box primitives; rows of a spatial query in ascending primitive order. ---- */
typedef struct abx_brute abx_brute;
ABX_API abx_status abx_brute_create(void *stream, int prim_kind, const void *prims_dev, int64_t n, abx_brute **out);
ABX_API abx_status abx_brute_destroy(abx_brute *brute);
ABX_API int64_t abx_brute_size(const abx_brute *brute);
ABX_API abx_status abx_brute_bounds(abx_brute *brute, float out6[6]);
ABX_API abx_status abx_brute_query_spatial_crs(abx_brute *brute, void *stream, int pred_kind, const void *preds_dev,
                                               int64_t q, abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                               uint32_t **indices_dev, int64_t *nnz);
ABX_API abx_status abx_brute_query_nearest_crs(abx_brute *brute, void *stream, const void *points_dev, int64_t q,
                                               int32_t k, abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                               uint32_t **indices_dev, float **distances_dev, int64_t *nnz);

/* ---- ArborX::dbscan(space, points, eps, minpts, labels, params) (cluster/ArborX_DBSCAN.hpp:180-223) ---- */
enum
{
  ABX_DBSCAN_FDBSCAN = 0,
  ABX_DBSCAN_FDBSCAN_DENSEBOX = 1
};
enum
{
  ABX_DBSCAN_DBSCAN = 0,
  ABX_DBSCAN_DBSCAN_STAR = 1
};
/* labels_dev[n]: cluster id (smallest original index among the cluster's core
 * points) or -1 for noise.  eps <= 0 or minpts < 2 -> ABX_ERR_SEARCH (:240-241). */
ABX_API abx_status abx_dbscan(void *stream, const float *xyz_dev, int64_t n, float eps, int32_t minpts, int implementation,
                      int algorithm, int32_t *labels_dev);
ABX_API abx_status abx_dbscan_host(void *stream, const float *xyz_host, int64_t n, float eps, int32_t minpts,
                           int implementation, int algorithm, int32_t *labels_host);

/* ---- ArborX::Experimental::MinimumSpanningTree / Dendrogram / hdbscan ----
 * cluster/ArborX_MinimumSpanningTree.hpp:46-101: Boruvka over the BVH.  k = 1: Euclidean distances; k > 1: mutual
 * reachability max(core_i, core_j, d) with core_i = distance to the k-th nearest point, i itself included (:70-88,
 * detail/ArborX_MutualReachabilityDistance.hpp:27-77).  edges2_dev: (n - 1) x (source, target) in the caller's
 * indices, weights_dev: n - 1.  Equal weights are ordered by the pair of leaf positions like the reference's
 * DirectedEdge (detail/ArborX_BoruvkaHelpers.hpp:37-105), which makes the tree unique: the edge SET equals the
 * reference's, the order of the edges in the array is unspecified (as there).  n < 2: nothing is written.
 * iterations (optional): number of Boruvka rounds.  Blocks once per round (:196-198). */
ABX_API abx_status abx_mst_points3f(void *stream, const float *xyz_dev, int64_t n, int32_t k, int32_t *edges2_dev,
                                    float *weights_dev, int32_t *iterations);
ABX_API abx_status abx_mst_points3f_host(void *stream, const float *xyz_host, int64_t n, int32_t k,
                                         int32_t *edges2_host, float *weights_host, int32_t *iterations);
/* cluster/ArborX_Dendrogram.hpp:47-76 (DendrogramImplementation::UNION_FIND): the edges are sorted by weight on the
 * device; the union-find pass over them is sequential and runs on the host, as it does in the reference
 * (detail/ArborX_DendrogramHelpers.hpp:31-80).  parents_dev: 2 * num_edges + 1 entries -- the edges in ascending
 * weight order first, then the num_edges + 1 vertices; the root's parent is -1.  parent_heights_dev: num_edges
 * (the sorted weights). */
ABX_API abx_status abx_dendrogram_union_find(void *stream, const int32_t *edges2_dev, const float *weights_dev,
                                             int64_t num_edges, int32_t *parents_dev, float *parent_heights_dev);
/* MinimumSpanningTree<MemorySpace, BoruvkaMode::HDBSCAN> (cluster/ArborX_MinimumSpanningTree.hpp:31-297): the same
 * rounds also record which edge every component picked, so the dendrogram is complete when the tree is
 * (detail/ArborX_BoruvkaHelpers.hpp:439-447,490-733) -- no host pass.  The edges come out in the hybrid algorithm's
 * (chain, weight) order; parents_dev [2 n - 1] (edges first, then the n vertices; root -> -1) and
 * parent_heights_dev [n - 1] index that order, like the reference's dendrogram_parents / dendrogram_parent_heights. */
ABX_API abx_status abx_mst_hdbscan_points3f(void *stream, const float *xyz_dev, int64_t n, int32_t k,
                                            int32_t *edges2_dev, float *weights_dev, int32_t *parents_dev,
                                            float *parent_heights_dev, int32_t *iterations);
/* cluster/ArborX_HDBSCAN.hpp:29-53: hdbscan(space, primitives, core_min_size, dendrogram_impl) */
enum
{
  ABX_DENDROGRAM_BORUVKA = 0,   /* the reference's default: the hybrid above */
  ABX_DENDROGRAM_UNION_FIND = 1 /* MST(core_min_size), then abx_dendrogram_union_find */
};
ABX_API abx_status abx_hdbscan_points3f(void *stream, const float *xyz_dev, int64_t n, int32_t core_min_size,
                                        int dendrogram_impl, int32_t *parents_dev, float *parent_heights_dev);

/* ---- device view for user callbacks (include/ArborX_B200_Callbacks.cuh) ----
 * The reference instantiates user callbacks inside its traversal templates
 * (spatial/detail/ArborX_Callbacks.hpp:79-150, ArborX_TreeTraversal.hpp:97-119,180-335).  Here the tree's
 * device arrays are exported and the same traversal cores the library kernels use are instantiated in
 * the user's .cu through the header.  The view is valid until abx_bvh_destroy. */
typedef struct abx_device_view
{
  const void *nodes;    /* Node64[n - 1], NULL when n < 2 */
  const void *leaf_box; /* sorted leaves: float4 per point, 2 x float4 per box / triangle */
  const void *leaf_tri; /* 3 x float4 per triangle, else NULL */
  int64_t n;
  int32_t prim_kind;
} abx_device_view;
ABX_API abx_status abx_bvh_device_view(const abx_bvh *bvh, abx_device_view *view);

/* ---- ArborX::DistributedTree (distributed/ArborX_DistributedTree.hpp:33-252) ----
 * One process (or host thread) per GPU.  Primitives are sharded by the caller: every rank builds a
 * bottom tree over its own primitives, the rank boxes and sizes are all-gathered (:208-245) and form the
 * replicated top tree.  Queries are COLLECTIVE over the communicator (:120-121): every rank calls with
 * its own predicates and gets, per predicate, the matching values of ALL ranks as (index, rank) pairs
 * (index = position in the owner rank's primitives; the reference returns user values or {index, rank}
 * from a callback, examples/distributed_tree/distributed_knn.cpp:62-104).
 *
 * Exchange (detail/ArborX_DistributedTreeUtils.hpp:52-263, ArborX_Distributor.hpp): the local tree
 * answers every predicate directly; only predicates that also touch OTHER ranks' boxes are forwarded
 * (one grouped NCCL send/recv each way; counts travel as one all-gather of the R x R count matrix),
 * queried there and merged back per query.  Two blocking points per query call.
 *
 * `comm`: an abx_comm made from the caller's ncclComm_t (abx_comm_from_nccl), bootstrapped from a
 * unique id (abx_comm_unique_id + abx_comm_init_rank; libnccl.so.2 is resolved at run time, so the
 * library loads without NCCL), or an in-process group of host threads sharing one GPU
 * (abx_comm_create_local: the multi-rank protocol on a single-GPU test box). */
typedef struct abx_comm abx_comm;
typedef struct abx_dist_tree abx_dist_tree;
#define ABX_COMM_UNIQUE_ID_BYTES 128
ABX_API abx_status abx_comm_from_nccl(void *nccl_comm /* ncclComm_t, not owned */, abx_comm **out);
ABX_API abx_status abx_comm_unique_id(char id_out[ABX_COMM_UNIQUE_ID_BYTES]);
ABX_API abx_status abx_comm_init_rank(const char id[ABX_COMM_UNIQUE_ID_BYTES], int32_t n_ranks, int32_t rank,
                                      abx_comm **out);
/* n_ranks communicators of one in-process group; rank r's calls must come from its own host thread */
ABX_API abx_status abx_comm_create_local(int32_t n_ranks, abx_comm **out_array);
ABX_API abx_status abx_comm_destroy(abx_comm *comm);
ABX_API int32_t abx_comm_rank(const abx_comm *comm);
ABX_API int32_t abx_comm_size(const abx_comm *comm);

/* DistributedTree(comm, space, values) :129-151.  Collective. */
ABX_API abx_status abx_dist_create(abx_comm *comm, void *stream, int prim_kind, const void *prims_dev, int64_t n,
                                   abx_dist_tree **out);
ABX_API abx_status abx_dist_destroy(abx_dist_tree *tree);
ABX_API int64_t abx_dist_size(const abx_dist_tree *tree);  /* global number of primitives :114-116 */
ABX_API int abx_dist_empty(const abx_dist_tree *tree);
ABX_API abx_status abx_dist_bounds(const abx_dist_tree *tree, float out6[6]); /* union of the rank boxes :122-127 */
/* query(space, intersects(...) predicates, values, offsets) :84-102; detail/ArborX_DistributedTreeSpatial.hpp:31-60.
 * Collective.  pred_kind: SPHERE3F, BOX3F or POINT3F.  offsets (which = 0): q + 1 ints; values2 (which = 1):
 * nnz (index, rank) pairs, row i = results of predicate i, local results first. */
ABX_API abx_status abx_dist_query_spatial_crs(abx_dist_tree *tree, void *stream, int pred_kind, const void *preds_dev,
                                              int64_t q, abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                              int32_t **values2_dev, int64_t *nnz);
/* query(space, nearest(point, k) predicates, ...) detail/ArborX_DistributedTreeNearest.hpp:41-261.  Collective.
 * Rows ascending by distance, min(k, global size) entries (fewer when leaves are at infinite distance);
 * distances_dev may be NULL. */
ABX_API abx_status abx_dist_query_nearest_crs(abx_dist_tree *tree, void *stream, const void *points_dev, int64_t q,
                                              int32_t k, abx_alloc_fn alloc, void *user, int32_t **offsets_dev,
                                              int32_t **values2_dev, float **distances_dev, int64_t *nnz);
/* ArborX::Experimental::dbscan(comm, space, primitives, eps, core_min_size, labels, params)
 * (cluster/ArborX_DistributedDBSCAN.hpp:29-190).  Collective.  The points are sharded as the caller sharded them;
 * labels_dev[n] (64-bit): global id (offset of the owner rank + index there) of the cluster's representative point,
 * -1 for noise; equal labels on different ranks mean the same cluster.  eps <= 0 or minpts < 2 -> ABX_ERR_SEARCH. */
ABX_API abx_status abx_dist_dbscan_points3f(abx_comm *comm, void *stream, const float *xyz_dev, int64_t n, float eps,
                                            int32_t minpts, int implementation, int algorithm, int64_t *labels_dev);
/* Host-buffer variants (end-to-end path): primitives / predicates in host memory, results in host arrays
 * from `alloc_host`.  Results come back in the COMPACT form: indices (which = 1, one uint32 per result) all
 * belong to the calling rank except the n_remote entries listed in remote_pos (which = 3, ascending positions
 * into indices) whose owner ranks are remote_rank (which = 4) -- a few percent of the results on spatially
 * compact shards, so the D2H volume is that of a single-tree query. */
ABX_API abx_status abx_dist_create_host(abx_comm *comm, void *stream, int prim_kind, const void *prims_host, int64_t n,
                                        abx_dist_tree **out);
ABX_API abx_status abx_dist_query_spatial_crs_host(abx_dist_tree *tree, void *stream, int pred_kind,
                                                   const void *preds_host, int64_t q, abx_alloc_fn alloc_host,
                                                   void *user, int32_t **offsets_host, uint32_t **indices_host,
                                                   int64_t *nnz, int32_t **remote_pos_host,
                                                   int32_t **remote_rank_host, int64_t *n_remote);
ABX_API abx_status abx_dist_query_nearest_crs_host(abx_dist_tree *tree, void *stream, const void *points_host,
                                                   int64_t q, int32_t k, abx_alloc_fn alloc_host, void *user,
                                                   int32_t **offsets_host, uint32_t **indices_host,
                                                   float **distances_host, int64_t *nnz, int32_t **remote_pos_host,
                                                   int32_t **remote_rank_host, int64_t *n_remote);

/* ---- DistributedTree building block (distributed/detail/ArborX_DistributedTreeUtils.hpp:229-263) ----
 * Merges, per query, the CRS rows of the local tree's results (indices) with the CRS rows of the
 * results that came back from other ranks ((index, rank) pairs, grouped by query) into one CRS
 * of (index, rank) pairs: out_offsets[q+1], out_values2[2 * (nnz_local + nnz_remote)].  These are the
 * kernels behind abx_dist_query_*; exported so that tests can check each one against a restatement. */
ABX_API abx_status abx_dist_merge_crs(void *stream, int64_t q, const int32_t *local_offsets_dev,
                                      const int32_t *local_indices_dev, int32_t rank,
                                      const int32_t *remote_offsets_dev, const int32_t *remote_values2_dev,
                                      int32_t *out_offsets_dev, int32_t *out_values2_dev);

/* Same merge with the remote results as they arrive from the exchange: n_remote records ordered by query id
 * (remote_query_ids_dev ascending), one (index, rank) pair each; no remote CRS offsets needed. */
ABX_API abx_status abx_dist_merge_sorted(void *stream, int64_t q, const int32_t *local_offsets_dev,
                                         const int32_t *local_indices_dev, int32_t rank, int64_t n_remote,
                                         const int32_t *remote_query_ids_dev, const int32_t *remote_values2_dev,
                                         int32_t *out_offsets_dev, int32_t *out_values2_dev);

/* Routing of forwarded predicates (the top-tree query of DistributedTreeSpatial.hpp:52-55 for R <= 64
 * ranks, evaluated directly against the R rank boxes; conservative for spheres).  Pass 1 counts the
 * predicates that must be forwarded to every OTHER rank; pass 2 writes their query ids grouped by
 * destination (base = exclusive scan of the counts), i.e. the send order of the all-to-all-v. */
ABX_API abx_status abx_dist_route_count(void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                        const float *radius_dev, int64_t radius_stride,
                                        const float *rank_boxes6_dev, int32_t n_ranks, int32_t self_rank,
                                        uint32_t *counts_dev);
ABX_API abx_status abx_dist_route_fill(void *stream, int pred_kind, const void *preds_dev, int64_t q,
                                       const float *radius_dev, int64_t radius_stride,
                                       const float *rank_boxes6_dev, int32_t n_ranks, int32_t self_rank,
                                       const uint32_t *base_dev, uint32_t *cursors_dev, int32_t *query_ids_dev);
/* radius_dev != NULL (ABX_PRED_SPHERE3F only): preds_dev holds points (3 floats) and the radius of
 * predicate i is radius_dev[i * radius_stride] -- the k-th distances of a kNN result, phase II of
 * distributed/detail/ArborX_DistributedTreeNearest.hpp:131-176, without building the spheres. */
/* values2[i] = (indices[i], rank): local results in the (index, rank) form DistributedTree returns */
ABX_API abx_status abx_dist_pair_with_rank(void *stream, const int32_t *indices_dev, int64_t n, int32_t rank,
                                           int32_t *values2_dev);
/* DistributedTree kNN, local phase (distributed/detail/ArborX_DistributedTreeNearest.hpp:131-176): the k
 * nearest of every point in rows of exactly k slots, written directly in the (index, rank) form:
 * values2_dev[2 * (i * k + j)] = index, [.. + 1] = rank; distances_dev[i * k + j].  Rows with fewer than k
 * reachable leaves are padded with (-1, -1) / +inf; *missing_out = number of padded slots. */
ABX_API abx_status abx_dist_nearest_pairs(abx_bvh *bvh, void *stream, const void *points_dev, int64_t q, int32_t k,
                                          int32_t rank, int32_t *values2_dev, float *distances_dev,
                                          int64_t *missing_out);
/* DistributedTree kNN, final ranking (same file :178-233): candidates received from other ranks
 * (query_ids_dev ascending, one (index, rank) pair and one distance each) are merged into the rows of
 * their queries (k entries each, ascending): the k smallest survive, local entries first among equals. */
ABX_API abx_status abx_dist_knn_merge(void *stream, int64_t n_candidates, const int32_t *query_ids_dev,
                                      const int32_t *cand_values2_dev, const float *cand_distances_dev, int32_t k,
                                      int32_t *values2_dev, float *distances_dev);

/* ---- stage-level entry points (tests localise mismatches with these) ---- */
/* TreeConstruction::calculateBoundingBoxOfTheScene (detail/ArborX_TreeConstruction.hpp:27-39) */
ABX_API abx_status abx_scene_bounds(void *stream, int prim_kind, const void *prims_dev, int64_t n, float *bounds6_dev);
/* projectOntoSpaceFillingCurve with Morton64 (detail/ArborX_SpaceFillingCurves.hpp:45-57,67-84) */
ABX_API abx_status abx_morton64(void *stream, int prim_kind, const void *prims_dev, int64_t n, const float *bounds6_dev,
                        uint64_t *codes_dev);
/* Morton32 of predicate centroids (ArborX_LinearBVH.hpp:287-298) */
ABX_API abx_status abx_morton32(void *stream, int pred_kind, const void *preds_dev, int64_t q, const float *bounds6_dev,
                        uint32_t *codes_dev);
/* sortObjects (misc/ArborX_SortUtils.hpp:28-43): sorts keys in place (stable) and
 * writes the permutation */
ABX_API abx_status abx_sort_u64(void *stream, uint64_t *keys_dev, uint32_t *perm_dev, int64_t n);
ABX_API abx_status abx_sort_u32(void *stream, uint32_t *keys_dev, uint32_t *perm_dev, int64_t n);
/* generateHierarchy on caller-sorted codes, primitives taken in the given order
 * (test/tstDetailsTreeConstruction.cpp:152-175) */
ABX_API abx_status abx_bvh_build_from_sorted_codes(void *stream, int prim_kind, const void *prims_dev,
                                           const uint64_t *sorted_codes_dev, int64_t n, abx_bvh **out);
/* exclusive scan used by the CRS path (kokkos_ext/ArborX_KokkosExtStdAlgorithms.hpp) */
ABX_API abx_status abx_exclusive_scan_i32(void *stream, const int32_t *in_dev, int32_t *out_dev, int64_t n_plus_1);

#ifdef __cplusplus
}
#endif
#endif /* ABX_H */
