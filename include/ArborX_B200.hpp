// ArborX_B200.hpp -- header-only C++17 facade over the C ABI of libabx.so that keeps
// the reference's spellings for the hot path, so that a translation unit using
//
//   ArborX::BoundingVolumeHierarchy bvh(space, points);
//   ArborX::query(bvh, space, predicates, indices, offsets);      // or bvh.query(...)
//   ArborX::dbscan(space, points, eps, minpts, labels, params);
//
// only has to swap `#include <ArborX.hpp>` for this header and Kokkos::View for
// ArborX::DeviceView (a thin owning device array; the library re-allocates output
// views like the reference does, spatial/detail/ArborX_CrsGraphWrapperImpl.hpp:257,286).
//
// What maps to what (paths relative to the reference's src/):
//   Point/Box/Sphere/Triangle            geometry/ArborX_{Point,Box,Sphere,Triangle}.hpp
//   intersects / nearest / Intersects<G> spatial/detail/ArborX_Predicates.hpp:58-147
//   Experimental::TraversalPolicy        spatial/detail/ArborX_TraversalPolicy.hpp:19-48
//   BoundingVolumeHierarchy              spatial/ArborX_LinearBVH.hpp:50-142
//   query (free function)                spatial/ArborX_CrsGraphWrapper.hpp:22-35
//   dbscan, DBSCAN::Parameters           cluster/ArborX_DBSCAN.hpp:180-223
//   Experimental::MinimumSpanningTree,   cluster/ArborX_MinimumSpanningTree.hpp:31-101,
//     Dendrogram, hdbscan                cluster/ArborX_Dendrogram.hpp:24-76, cluster/ArborX_HDBSCAN.hpp:29-53
//   SearchException                      misc/ArborX_Exception.hpp:19-38
//
// Not covered by a C ABI: arbitrary device callbacks (functors cannot cross it); see
// INTEGRATION.md ("callbacks").  Compile with any C++17 host compiler and link
// -labx plus the CUDA runtime (cudart) for the DeviceView helpers.
#ifndef ARBORX_B200_HPP
#define ARBORX_B200_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <cuda_runtime_api.h>

#include "abx.h"

namespace ArborX
{

// misc/ArborX_Exception.hpp:19-38
class SearchException : public std::logic_error
{
public:
  using std::logic_error::logic_error;
};

namespace Details
{
inline void check(abx_status st)
{
  if (st == ABX_OK)
    return;
  std::string msg = abx_last_error();
  if (st == ABX_ERR_SEARCH)
    throw SearchException(msg);
  if (st == ABX_ERR_ARG)
    throw std::invalid_argument(msg);
  throw std::runtime_error(msg); // ABX_ERR_PRECISION (CartesianGrid.hpp:132-135) and CUDA failures
}
inline void cudaCheck(cudaError_t e)
{
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e));
}
} // namespace Details

// ---- geometry (same memory layout as the reference's value types) -------------------
template <int DIM = 3, class Coordinate = float>
struct Point
{
  static_assert(DIM == 3 && std::is_same_v<Coordinate, float>, "the B200 hot path is 3-D float");
  Coordinate _coords[DIM];
  Coordinate &operator[](int d) { return _coords[d]; }
  Coordinate const &operator[](int d) const { return _coords[d]; }
};
template <int DIM = 3, class Coordinate = float>
struct Box
{
  Point<DIM, Coordinate> _min_corner{{3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}};
  Point<DIM, Coordinate> _max_corner{{-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f}};
  auto &minCorner() { return _min_corner; }
  auto const &minCorner() const { return _min_corner; }
  auto &maxCorner() { return _max_corner; }
  auto const &maxCorner() const { return _max_corner; }
};
template <int DIM = 3, class Coordinate = float>
struct Sphere
{
  Point<DIM, Coordinate> _centroid;
  Coordinate _radius;
  auto const &centroid() const { return _centroid; }
  auto radius() const { return _radius; }
};
template <int DIM = 3, class Coordinate = float>
struct Triangle
{
  Point<DIM, Coordinate> a, b, c;
};
static_assert(sizeof(Point<>) == 12 && sizeof(Box<>) == 24 && sizeof(Sphere<>) == 16 && sizeof(Triangle<>) == 36);

// ---- predicates: Predicates.hpp:58-147 -----------------------------------------------
template <class Geometry>
struct Intersects
{
  Geometry _geometry;
};
template <class Geometry>
struct Nearest
{
  Geometry _geometry;
  int _k = 0;
};
template <class Geometry>
Intersects<Geometry> intersects(Geometry const &g)
{
  return {g};
}
template <class Geometry>
Nearest<Geometry> nearest(Geometry const &g, int k = 1)
{
  return {g, k};
}

namespace Experimental
{
// geometry/ArborX_Ray.hpp:33-68: origin + direction (normalised by the library, in double, like the constructor)
struct Ray
{
  float _origin[3];
  float _direction[3];
};
static_assert(sizeof(Ray) == 24);
// TraversalPolicy.hpp:19-48
struct TraversalPolicy
{
  int _buffer_size = 0;
  bool _sort_predicates = true;
  TraversalPolicy &setBufferSize(int b)
  {
    _buffer_size = b;
    return *this;
  }
  TraversalPolicy &setPredicateSorting(bool s)
  {
    _sort_predicates = s;
    return *this;
  }
};
} // namespace Experimental

// ---- execution space instance = CUDA stream --------------------------------------------
class Cuda
{
  cudaStream_t _stream = nullptr;

public:
  Cuda() = default;
  explicit Cuda(cudaStream_t s)
      : _stream(s)
  {}
  cudaStream_t cuda_stream() const { return _stream; }
  void fence() const { Details::cudaCheck(cudaStreamSynchronize(_stream)); }
};

// ---- DeviceView<T>: minimal stand-in for Kokkos::View<T*, CudaSpace> -------------------
template <class T>
class DeviceView
{
  T *_data = nullptr;
  std::size_t _size = 0;

public:
  using value_type = T;
  DeviceView() = default;
  explicit DeviceView(std::size_t n) { realloc(n); }
  DeviceView(DeviceView const &) = delete;
  DeviceView &operator=(DeviceView const &) = delete;
  DeviceView(DeviceView &&o) noexcept { swap(o); }
  DeviceView &operator=(DeviceView &&o) noexcept
  {
    swap(o);
    return *this;
  }
  ~DeviceView()
  {
    if (_data)
      cudaFree(_data);
  }
  void swap(DeviceView &o)
  {
    std::swap(_data, o._data);
    std::swap(_size, o._size);
  }
  void realloc(std::size_t n)
  {
    if (_data)
      cudaFree(_data);
    _data = nullptr;
    _size = n;
    if (n)
      Details::cudaCheck(cudaMalloc((void **)&_data, n * sizeof(T)));
  }
  T *data() const { return _data; }
  std::size_t size() const { return _size; }
  std::size_t extent(int) const { return _size; }
  // host <-> device helpers (Kokkos::deep_copy)
  void assign(std::vector<T> const &h, Cuda const &space = {})
  {
    realloc(h.size());
    if (!h.empty())
      Details::cudaCheck(cudaMemcpyAsync(_data, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice,
                                         space.cuda_stream()));
    space.fence();
  }
  std::vector<T> to_host(Cuda const &space = {}) const
  {
    std::vector<T> h(_size);
    if (_size)
      Details::cudaCheck(cudaMemcpyAsync(h.data(), _data, _size * sizeof(T), cudaMemcpyDeviceToHost,
                                         space.cuda_stream()));
    space.fence();
    return h;
  }
};

namespace Details
{
template <class G>
struct PrimKind;
template <>
struct PrimKind<Point<>>
{
  static constexpr int value = ABX_PRIM_POINT3F;
};
template <>
struct PrimKind<Box<>>
{
  static constexpr int value = ABX_PRIM_BOX3F;
};
template <>
struct PrimKind<Triangle<>>
{
  static constexpr int value = ABX_PRIM_TRI3F;
};
template <class G>
struct PredKind;
template <>
struct PredKind<Sphere<>>
{
  static constexpr int value = ABX_PRED_SPHERE3F;
};
template <>
struct PredKind<Box<>>
{
  static constexpr int value = ABX_PRED_BOX3F;
};
template <>
struct PredKind<Point<>>
{
  static constexpr int value = ABX_PRED_POINT3F;
};
template <>
struct PredKind<Experimental::Ray>
{
  static constexpr int value = ABX_PRED_RAY3F;
};

// abx_alloc_fn that resizes the caller's views
template <class Indices, class Offsets>
struct ViewAllocator
{
  Indices *indices;
  Offsets *offsets;
  DeviceView<float> *distances;
  static void *call(void *user, int which, std::size_t bytes)
  {
    auto *self = static_cast<ViewAllocator *>(user);
    if (which == 0)
    {
      self->offsets->realloc(bytes / sizeof(int));
      return self->offsets->data();
    }
    if (which == 1)
    {
      self->indices->realloc(bytes / sizeof(typename Indices::value_type));
      return self->indices->data();
    }
    self->distances->realloc(bytes / sizeof(float));
    return self->distances->data();
  }
};
} // namespace Details

// ---- BoundingVolumeHierarchy: ArborX_LinearBVH.hpp:50-142 -------------------------------
// Value = int (index of the primitive): the bvh_driver / attach_indices configuration.
class BoundingVolumeHierarchy
{
  abx_bvh *_h = nullptr;
  Box<> _bounds;

public:
  using size_type = std::int64_t;
  BoundingVolumeHierarchy() = default;
  template <class Geometry>
  BoundingVolumeHierarchy(Cuda const &space, DeviceView<Geometry> const &values)
  {
    Details::check(abx_bvh_build(space.cuda_stream(), Details::PrimKind<Geometry>::value, values.data(),
                                 (std::int64_t)values.size(), &_h));
    float b[6];
    Details::check(abx_bvh_bounds(_h, b)); // the reference's ctor also blocks for the root box
    for (int d = 0; d < 3; ++d)
    {
      _bounds._min_corner[d] = b[d];
      _bounds._max_corner[d] = b[3 + d];
    }
  }
  BoundingVolumeHierarchy(BoundingVolumeHierarchy const &) = delete;
  BoundingVolumeHierarchy &operator=(BoundingVolumeHierarchy const &) = delete;
  BoundingVolumeHierarchy(BoundingVolumeHierarchy &&o) noexcept
      : _h(o._h)
      , _bounds(o._bounds)
  {
    o._h = nullptr;
  }
  ~BoundingVolumeHierarchy()
  {
    if (_h)
      abx_bvh_destroy(_h);
  }
  size_type size() const { return _h ? abx_bvh_size(_h) : 0; }
  bool empty() const { return size() == 0; }
  Box<> bounds() const { return _bounds; }
  abx_bvh *handle() const { return _h; }

  // query(space, predicates, indices, offsets, policy): spatial predicates
  template <class Geometry>
  void query(Cuda const &space, DeviceView<Intersects<Geometry>> const &predicates, DeviceView<int> &indices,
             DeviceView<int> &offsets, Experimental::TraversalPolicy const &policy = {}) const
  {
    static_assert(sizeof(Intersects<Geometry>) == sizeof(Geometry));
    abx_policy p{policy._buffer_size, policy._sort_predicates ? 1 : 0};
    Details::ViewAllocator<DeviceView<int>, DeviceView<int>> alloc{&indices, &offsets, nullptr};
    std::int32_t *off = nullptr;
    std::uint32_t *idx = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_query_spatial_crs(_h, space.cuda_stream(), Details::PredKind<Geometry>::value,
                                         predicates.data(), (std::int64_t)predicates.size(), &p,
                                         &Details::ViewAllocator<DeviceView<int>, DeviceView<int>>::call, &alloc, &off,
                                         &idx, &nnz));
  }
  // nearest(Box | Sphere, k) / nearest(Experimental::Ray, k) with the same k for every predicate (Predicates.hpp:58-80)
  template <class Geometry>
  void query(Cuda const &space, DeviceView<Geometry> const &geometries, int k, DeviceView<int> &indices,
             DeviceView<int> &offsets, DeviceView<float> *distances = nullptr,
             Experimental::TraversalPolicy const &policy = {}) const
  {
    abx_policy p{policy._buffer_size, policy._sort_predicates ? 1 : 0};
    Details::ViewAllocator<DeviceView<int>, DeviceView<int>> alloc{&indices, &offsets, distances};
    std::int32_t *off = nullptr;
    std::uint32_t *idx = nullptr;
    float *dist = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_query_nearest_geom_crs(_h, space.cuda_stream(), Details::PredKind<Geometry>::value,
                                              geometries.data(), (std::int64_t)geometries.size(), k, &p,
                                              &Details::ViewAllocator<DeviceView<int>, DeviceView<int>>::call, &alloc,
                                              &off, &idx, distances ? &dist : nullptr, &nnz));
  }
  // nearest(Point, k) with the same k for every predicate (Experimental::make_nearest)
  void query(Cuda const &space, DeviceView<Point<>> const &points, int k, DeviceView<int> &indices,
             DeviceView<int> &offsets, DeviceView<float> *distances = nullptr,
             Experimental::TraversalPolicy const &policy = {}) const
  {
    abx_policy p{policy._buffer_size, policy._sort_predicates ? 1 : 0};
    Details::ViewAllocator<DeviceView<int>, DeviceView<int>> alloc{&indices, &offsets, distances};
    std::int32_t *off = nullptr;
    std::uint32_t *idx = nullptr;
    float *dist = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_query_nearest_crs(_h, space.cuda_stream(), points.data(), (std::int64_t)points.size(), k, nullptr,
                                         &p, &Details::ViewAllocator<DeviceView<int>, DeviceView<int>>::call, &alloc,
                                         &off, &idx, distances ? &dist : nullptr, &nnz));
  }
};
using BVH = BoundingVolumeHierarchy;

// ArborX_CrsGraphWrapper.hpp:22-35
template <class... Args>
void query(BoundingVolumeHierarchy const &tree, Cuda const &space, Args &&...args)
{
  tree.query(space, std::forward<Args>(args)...);
}

// ---- BruteForce: spatial/ArborX_BruteForce.hpp:42-160 ------------------------------------------
class BruteForce
{
  abx_brute *_h = nullptr;

public:
  BruteForce() = default;
  template <class Geometry>
  BruteForce(Cuda const &space, DeviceView<Geometry> const &values)
  {
    Details::check(abx_brute_create(space.cuda_stream(), Details::PrimKind<Geometry>::value, values.data(),
                                    (std::int64_t)values.size(), &_h));
  }
  BruteForce(BruteForce const &) = delete;
  BruteForce &operator=(BruteForce const &) = delete;
  ~BruteForce()
  {
    if (_h)
      abx_brute_destroy(_h);
  }
  std::int64_t size() const { return _h ? abx_brute_size(_h) : 0; }
  bool empty() const { return size() == 0; }
  Box<> bounds() const
  {
    Box<> bx;
    float b[6];
    Details::check(abx_brute_bounds(_h, b));
    for (int d = 0; d < 3; ++d)
    {
      bx._min_corner[d] = b[d];
      bx._max_corner[d] = b[3 + d];
    }
    return bx;
  }
  template <class Geometry>
  void query(Cuda const &space, DeviceView<Intersects<Geometry>> const &predicates, DeviceView<int> &indices,
             DeviceView<int> &offsets) const
  {
    Details::ViewAllocator<DeviceView<int>, DeviceView<int>> alloc{&indices, &offsets, nullptr};
    std::int32_t *off = nullptr;
    std::uint32_t *idx = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_brute_query_spatial_crs(_h, space.cuda_stream(), Details::PredKind<Geometry>::value,
                                               predicates.data(), (std::int64_t)predicates.size(),
                                               &Details::ViewAllocator<DeviceView<int>, DeviceView<int>>::call, &alloc,
                                               &off, &idx, &nnz));
  }
  void query(Cuda const &space, DeviceView<Point<>> const &points, int k, DeviceView<int> &indices,
             DeviceView<int> &offsets, DeviceView<float> *distances = nullptr) const
  {
    Details::ViewAllocator<DeviceView<int>, DeviceView<int>> alloc{&indices, &offsets, distances};
    std::int32_t *off = nullptr;
    std::uint32_t *idx = nullptr;
    float *dist = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_brute_query_nearest_crs(_h, space.cuda_stream(), points.data(), (std::int64_t)points.size(), k,
                                               &Details::ViewAllocator<DeviceView<int>, DeviceView<int>>::call, &alloc,
                                               &off, &idx, distances ? &dist : nullptr, &nnz));
  }
};

// ---- DistributedTree: distributed/ArborX_DistributedTree.hpp:33-252 ----------------------------------
// One process (or host thread) per GPU.  The reference takes an MPI_Comm; here the communicator is the caller's
// ncclComm_t (passed as void *: this header does not need nccl.h).  Values returned by queries are
// PairIndexRank{index, rank} like the reference's default (examples/distributed_tree/distributed_knn.cpp:62-104).
struct PairIndexRank
{
  int index;
  int rank;
};
class DistributedTree
{
  abx_comm *_comm = nullptr;
  abx_dist_tree *_h = nullptr;

public:
  template <class Geometry>
  DistributedTree(void *nccl_comm, Cuda const &space, DeviceView<Geometry> const &values)
  {
    Details::check(abx_comm_from_nccl(nccl_comm, &_comm));
    Details::check(abx_dist_create(_comm, space.cuda_stream(), Details::PrimKind<Geometry>::value, values.data(),
                                   (std::int64_t)values.size(), &_h));
  }
  DistributedTree(DistributedTree const &) = delete;
  DistributedTree &operator=(DistributedTree const &) = delete;
  ~DistributedTree()
  {
    if (_h)
      abx_dist_destroy(_h);
    if (_comm)
      abx_comm_destroy(_comm);
  }
  std::int64_t size() const { return _h ? abx_dist_size(_h) : 0; }
  bool empty() const { return size() == 0; }
  Box<> bounds() const
  {
    Box<> bx;
    float b[6];
    Details::check(abx_dist_bounds(_h, b));
    for (int d = 0; d < 3; ++d)
    {
      bx._min_corner[d] = b[d];
      bx._max_corner[d] = b[3 + d];
    }
    return bx;
  }
  // collective: query(space, intersects(...) predicates, values, offsets) :84-102
  template <class Geometry>
  void query(Cuda const &space, DeviceView<Intersects<Geometry>> const &predicates, DeviceView<PairIndexRank> &values,
             DeviceView<int> &offsets) const
  {
    Details::ViewAllocator<DeviceView<PairIndexRank>, DeviceView<int>> alloc{&values, &offsets, nullptr};
    std::int32_t *off = nullptr, *vals = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_dist_query_spatial_crs(_h, space.cuda_stream(), Details::PredKind<Geometry>::value,
                                              predicates.data(), (std::int64_t)predicates.size(),
                                              &Details::ViewAllocator<DeviceView<PairIndexRank>, DeviceView<int>>::call,
                                              &alloc, &off, &vals, &nnz));
  }
  // collective: nearest(Point, k)
  void query(Cuda const &space, DeviceView<Point<>> const &points, int k, DeviceView<PairIndexRank> &values,
             DeviceView<int> &offsets, DeviceView<float> *distances = nullptr) const
  {
    Details::ViewAllocator<DeviceView<PairIndexRank>, DeviceView<int>> alloc{&values, &offsets, distances};
    std::int32_t *off = nullptr, *vals = nullptr;
    float *dist = nullptr;
    std::int64_t nnz = 0;
    Details::check(abx_dist_query_nearest_crs(_h, space.cuda_stream(), points.data(), (std::int64_t)points.size(), k,
                                              &Details::ViewAllocator<DeviceView<PairIndexRank>, DeviceView<int>>::call,
                                              &alloc, &off, &vals, distances ? &dist : nullptr, &nnz));
  }
};

// ---- dbscan: cluster/ArborX_DBSCAN.hpp:180-223 --------------------------------------------
namespace DBSCAN
{
enum class Implementation
{
  FDBSCAN,
  FDBSCAN_DenseBox
};
enum class Algorithm
{
  DBSCAN,
  DBSCAN_STAR
};
struct Parameters
{
  bool _verbose = false;
  Implementation _implementation = Implementation::FDBSCAN_DenseBox;
  Algorithm _algorithm = Algorithm::DBSCAN;
  Parameters &setVerbosity(bool v)
  {
    _verbose = v;
    return *this;
  }
  Parameters &setImplementation(Implementation i)
  {
    _implementation = i;
    return *this;
  }
  Parameters &setAlgorithm(Algorithm a)
  {
    _algorithm = a;
    return *this;
  }
};
} // namespace DBSCAN

inline void dbscan(Cuda const &space, DeviceView<Point<>> const &primitives, double eps, int core_min_size,
                   DeviceView<int> &labels, DBSCAN::Parameters const &parameters = {})
{
  labels.realloc(primitives.size());
  Details::check(abx_dbscan(space.cuda_stream(), reinterpret_cast<float const *>(primitives.data()),
                            (std::int64_t)primitives.size(), static_cast<float>(eps), core_min_size,
                            parameters._implementation == DBSCAN::Implementation::FDBSCAN ? ABX_DBSCAN_FDBSCAN
                                                                                          : ABX_DBSCAN_FDBSCAN_DENSEBOX,
                            parameters._algorithm == DBSCAN::Algorithm::DBSCAN ? ABX_DBSCAN_DBSCAN
                                                                               : ABX_DBSCAN_DBSCAN_STAR,
                            labels.data()));
}

// ---- MinimumSpanningTree / Dendrogram / hdbscan ------------------------------------------------
// cluster/ArborX_MinimumSpanningTree.hpp:31-101, cluster/ArborX_Dendrogram.hpp:24-76, cluster/ArborX_HDBSCAN.hpp:29-53
namespace Experimental
{
// cluster/detail/ArborX_WeightedEdge.hpp:20-50; the library keeps (source, target) and the weights in two arrays
struct UnweightedEdge
{
  int source;
  int target;
};

struct MinimumSpanningTree
{
  DeviceView<UnweightedEdge> edges;
  DeviceView<float> weights;
  int iterations = 0;
  MinimumSpanningTree(Cuda const &space, DeviceView<Point<>> const &points, int k = 1)
  {
    std::size_t const n = points.size();
    edges.realloc(n > 0 ? n - 1 : 0);
    weights.realloc(n > 0 ? n - 1 : 0);
    std::int32_t it = 0;
    Details::check(abx_mst_points3f(space.cuda_stream(), reinterpret_cast<float const *>(points.data()),
                                    (std::int64_t)n, k, reinterpret_cast<std::int32_t *>(edges.data()),
                                    weights.data(), &it));
    iterations = it;
  }
};

enum class DendrogramImplementation
{
  BORUVKA,
  UNION_FIND
};

struct Dendrogram
{
  DeviceView<int> _parents;
  DeviceView<float> _parent_heights;
  Dendrogram() = default;
  Dendrogram(Cuda const &space, DeviceView<UnweightedEdge> const &edges, DeviceView<float> const &weights)
  {
    std::size_t const m = edges.size();
    _parents.realloc(2 * m + 1);
    _parent_heights.realloc(m);
    Details::check(abx_dendrogram_union_find(space.cuda_stream(), reinterpret_cast<std::int32_t const *>(edges.data()),
                                             weights.data(), (std::int64_t)m, _parents.data(),
                                             _parent_heights.data()));
  }
};

inline Dendrogram hdbscan(Cuda const &space, DeviceView<Point<>> const &primitives, int core_min_size,
                          DendrogramImplementation impl = DendrogramImplementation::BORUVKA)
{
  Dendrogram d;
  std::size_t const n = primitives.size();
  d._parents.realloc(n > 0 ? 2 * n - 1 : 0);
  d._parent_heights.realloc(n > 0 ? n - 1 : 0);
  Details::check(abx_hdbscan_points3f(space.cuda_stream(), reinterpret_cast<float const *>(primitives.data()),
                                      (std::int64_t)n, core_min_size,
                                      impl == DendrogramImplementation::BORUVKA ? ABX_DENDROGRAM_BORUVKA
                                                                                : ABX_DENDROGRAM_UNION_FIND,
                                      d._parents.data(), d._parent_heights.data()));
  return d;
}
} // namespace Experimental

} // namespace ArborX

#endif
