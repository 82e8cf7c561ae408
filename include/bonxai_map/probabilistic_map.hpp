// bonxai_b200 drop-in for bonxai_map/include/bonxai_map/probabilistic_map.hpp (+ src/probabilistic_map.cpp).
//
// Bonxai::ProbabilisticMap with the reference's public names (probabilistic_map.hpp:27-139): Options in int32
// log-odds, insertPointCloud, addHitPoint/addMissPoint, isOccupied/isUnknown/isFree, getOccupiedVoxels /
// getFreeVoxels, grid(), logods()/prob(), RayIterator/ComputeRay. The map lives in B200 HBM; every member
// forwards to the C ABI of include/bonxai_b200.h and the result is bit-identical to the reference's
// (same cell words, same set of cells) after every call.
//
// Header-only: unlike the reference there is no probabilistic_map.cpp to compile, link libbonxai_b200.so.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <type_traits>
#include <vector>

#include "bonxai/bonxai.hpp"

#if __has_include(<eigen3/Eigen/Geometry>)
#include <eigen3/Eigen/Geometry>
#define BONXAI_B200_HAS_EIGEN 1
#elif __has_include(<Eigen/Geometry>)
#include <Eigen/Geometry>
#define BONXAI_B200_HAS_EIGEN 1
#endif

namespace Bonxai {

#ifndef BONXAI_B200_HAS_EIGEN
// minimal stand-in used only when Eigen is not installed: callers that pass Eigen::Vector3d get Eigen's type
struct Vector3dLite {
  double v[3];
  Vector3dLite() : v{0, 0, 0} {}
  Vector3dLite(double x, double y, double z) : v{x, y, z} {}
  double& x() { return v[0]; }
  double& y() { return v[1]; }
  double& z() { return v[2]; }
  const double& x() const { return v[0]; }
  const double& y() const { return v[1]; }
  const double& z() const { return v[2]; }
};
#endif

// Cells of the ray key_origin -> key_end in visiting order; the end cell is NOT visited and an empty ray
// results when both keys are equal (probabilistic_map.hpp:162-203). The reference advances an integer error
// term per axis; the same cells follow from the closed form used by the device kernels:
//   cell k = origin + sign(d) * floor((2*k*|d| + m) / (2*m)),  m = max|d|,  k = 0 .. m-1.
template <class Functor>
inline void RayIterator(const CoordT& key_origin, const CoordT& key_end, const Functor& func) {
  const int64_t d[3] = {(int64_t)key_end.x - key_origin.x, (int64_t)key_end.y - key_origin.y, (int64_t)key_end.z - key_origin.z};
  int64_t a[3], s[3], m = 0;
  for (int i = 0; i < 3; ++i) {
    a[i] = d[i] < 0 ? -d[i] : d[i];
    s[i] = d[i] < 0 ? -1 : 1;
    if (a[i] > m) m = a[i];
  }
  for (int64_t k = 0; k < m; ++k) {
    const CoordT c = {(int32_t)(key_origin.x + s[0] * ((2 * k * a[0] + m) / (2 * m))),
                      (int32_t)(key_origin.y + s[1] * ((2 * k * a[1] + m) / (2 * m))),
                      (int32_t)(key_origin.z + s[2] * ((2 * k * a[2] + m) / (2 * m)))};
    if (!func(c)) return;
  }
}

inline void ComputeRay(const CoordT& key_origin, const CoordT& key_end, std::vector<CoordT>& ray) {
  ray.clear();
  RayIterator(key_origin, key_end, [&ray](const CoordT& c) {
    ray.push_back(c);
    return true;
  });
}

class ProbabilisticMap {
 public:
#ifdef BONXAI_B200_HAS_EIGEN
  using Vector3D = Eigen::Vector3d;
#else
  using Vector3D = Vector3dLite;
#endif

  // log-odds as a fixed-point integer with 6 decimals (probabilistic_map.hpp:34-36): float argument,
  // double arithmetic, truncation toward zero.
  [[nodiscard]] static int32_t logods(float prob) { return int32_t(1e6 * std::log(prob / (1.0 - prob))); }
  // probabilistic_map.hpp:39-42
  [[nodiscard]] static float prob(int32_t logods_fixed) {
    const float l = float(logods_fixed) * 1e-6;
    return (1.0 - 1.0 / (1.0 + std::exp(l)));
  }

  // the 32-bit cell word: bits 0-3 update_id, bits 4-31 probability_log (probabilistic_map.hpp:44-53)
  struct CellT {
    int32_t update_id : 4;
    int32_t probability_log : 28;
    CellT() : update_id(0), probability_log(0) {}  // UnknownProbability = logods(0.5f) = 0
  };
  static_assert(sizeof(CellT) == 4, "CellT must be one 32-bit word");

  // same defaults as OctoMap (probabilistic_map.hpp:56-64)
  struct Options {
    int32_t prob_miss_log = logods(0.4f);
    int32_t prob_hit_log = logods(0.7f);
    int32_t clamp_min_log = logods(0.12f);
    int32_t clamp_max_log = logods(0.97f);
    int32_t occupancy_threshold_log = logods(0.5f);
  };

  static inline const int32_t UnknownProbability = 0;

  explicit ProbabilisticMap(double resolution) : _resolution(resolution) {
    detail::check(bnx_map_create(resolution, &_map));
    bnx_grid_t* g = nullptr;
    detail::check(bnx_map_grid(_map, &g));
    _grid = VoxelGrid<CellT>(g, resolution, 2, 3);
    int32_t o[5];
    detail::check(bnx_map_get_options(_map, o));
    _options = fromArray(o);
  }
  ProbabilisticMap(const ProbabilisticMap&) = delete;
  ProbabilisticMap& operator=(const ProbabilisticMap&) = delete;
  ~ProbabilisticMap() {
    if (_map) bnx_map_destroy(_map);
  }

  [[nodiscard]] VoxelGrid<CellT>& grid() { return _grid; }
  [[nodiscard]] const VoxelGrid<CellT>& grid() const { return _grid; }
  [[nodiscard]] const Options& options() const { return _options; }
  void setOptions(const Options& options) {
    _options = options;
    const int32_t o[5] = {options.prob_miss_log, options.prob_hit_log, options.clamp_min_log, options.clamp_max_log,
                          options.occupancy_threshold_log};
    detail::check(bnx_map_set_options(_map, o));
  }
  bnx_map_t* handle() const { return _map; }

  // probabilistic_map.hpp:141-160. PointT: anything ConvertPoint understands. Point types whose x,y,z are
  // three consecutive floats or doubles (pcl::PointXYZ, Eigen::Vector3f/3d, Point3D, struct{float x,y,z}) are
  // read in place; anything else is converted to double triplets first. Use Bonxai::PinnedAllocator for the
  // vector to make the host->device copy a direct DMA.
  template <typename PointT, typename Allocator>
  void insertPointCloud(const std::vector<PointT, Allocator>& points, const PointT& origin, double max_range) {
    const Point3D o = ConvertPoint<Point3D>(origin);
    const int64_t n = (int64_t)points.size();
    const char* base = nullptr;
    int scalar = 0;  // 4 float, 8 double, 0 no in-place layout
    if (n > 0) inPlaceLayout(points[0], base, scalar);
    if (n == 0 || scalar == 8) {
      const double od[3] = {o.x, o.y, o.z};
      detail::check(bnx_map_insert_f64(_map, base, n ? (int64_t)sizeof(PointT) : 24, n, od, max_range, BNX_HOST));
    } else if (scalar == 4) {
      const float of[3] = {(float)o.x, (float)o.y, (float)o.z};  // exact: the origin is a PointT of floats
      detail::check(bnx_map_insert_f32(_map, base, (int64_t)sizeof(PointT), n, of, max_range, BNX_HOST));
    } else {
      insertConverted(points, o, max_range);
    }
  }

  // probabilistic_map.cpp:30-54: the endpoint is updated now; its ray is cast by the next insertPointCloud
  void addHitPoint(const Vector3D& point) {
    const double p[3] = {point.x(), point.y(), point.z()};
    detail::check(bnx_map_add_hit(_map, p));
  }
  void addMissPoint(const Vector3D& point) {
    const double p[3] = {point.x(), point.y(), point.z()};
    detail::check(bnx_map_add_miss(_map, p));
  }

  // probabilistic_map.cpp:56-75
  [[nodiscard]] bool isOccupied(const CoordT& coord) const { return query(coord, BNX_OCCUPIED); }
  [[nodiscard]] bool isUnknown(const CoordT& coord) const { return query(coord, BNX_UNKNOWN); }
  [[nodiscard]] bool isFree(const CoordT& coord) const { return query(coord, BNX_FREE); }

  // probabilistic_map.cpp:108-126
  void getOccupiedVoxels(std::vector<CoordT>& coords) { voxels(BNX_OCCUPIED, coords); }
  void getFreeVoxels(std::vector<CoordT>& coords) { voxels(BNX_FREE, coords); }

  // probabilistic_map.hpp:115-124: appends coord * resolution (voxel corner) per occupied voxel
  template <typename PointT>
  void getOccupiedVoxels(std::vector<PointT>& points) {
    int64_t n = 0;
    detail::check(bnx_map_get_voxel_points(_map, BNX_OCCUPIED, nullptr, 0, &n, BNX_HOST));
    std::vector<double> xyz((size_t)n * 3);
    if (n) detail::check(bnx_map_get_voxel_points(_map, BNX_OCCUPIED, xyz.data(), n, &n, BNX_HOST));
    points.reserve(points.size() + (size_t)n);
    for (int64_t i = 0; i < n; ++i) points.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }

 private:
  static Options fromArray(const int32_t o[5]) {
    Options r;
    r.prob_miss_log = o[0];
    r.prob_hit_log = o[1];
    r.clamp_min_log = o[2];
    r.clamp_max_log = o[3];
    r.occupancy_threshold_log = o[4];
    return r;
  }
  bool query(const CoordT& c, int kind) const {
    uint8_t out = 0;
    detail::check(bnx_map_query(_map, &c.x, 1, kind, &out, BNX_HOST));
    return out != 0;
  }
  void voxels(int kind, std::vector<CoordT>& coords) {
    int64_t n = 0;
    detail::check(bnx_map_get_voxels(_map, kind, nullptr, 0, &n, BNX_HOST));
    coords.resize((size_t)n);
    if (n) detail::check(bnx_map_get_voxels(_map, kind, &coords[0].x, n, &n, BNX_HOST));
    coords.resize((size_t)n);
  }

  // point types without an in-place xyz layout: converted to double triplets first (ConvertPoint, like the reference)
  template <typename PointT, typename Allocator>
#if defined(__GNUC__)
  __attribute__((noinline))
#endif
  void insertConverted(const std::vector<PointT, Allocator>& points, const Point3D& o, double max_range) {
    std::vector<double> xyz;
    xyz.reserve(points.size() * 3);
    for (const auto& pt : points) {
      const Point3D p = ConvertPoint<Point3D>(pt);
      xyz.push_back(p.x);
      xyz.push_back(p.y);
      xyz.push_back(p.z);
    }
    const double od[3] = {o.x, o.y, o.z};
    detail::check(bnx_map_insert_f64(_map, xyz.data(), 24, (int64_t)points.size(), od, max_range, BNX_HOST));
  }

  // x,y,z as three consecutive floats/doubles inside PointT?
  template <typename PointT>
  static void inPlaceLayout(const PointT& p, const char*& base, int& scalar) {
    if constexpr (detail::has_xyz_fields<PointT>::value) {
      using S = std::remove_cv_t<decltype(PointT::x)>;
      if constexpr (std::is_same_v<S, float> || std::is_same_v<S, double>) {
        if (&p.y == &p.x + 1 && &p.z == &p.x + 2 && sizeof(PointT) % sizeof(S) == 0) {
          base = reinterpret_cast<const char*>(&p.x);
          scalar = (int)sizeof(S);
        }
      }
    } else if constexpr (detail::has_xyz_methods<PointT>::value) {
      if constexpr (std::is_lvalue_reference_v<decltype(p.x())>) {
        using S = std::remove_cv_t<std::remove_reference_t<decltype(p.x())>>;
        if constexpr (std::is_same_v<S, float> || std::is_same_v<S, double>) {
          if (&p.y() == &p.x() + 1 && &p.z() == &p.x() + 2 && sizeof(PointT) % sizeof(S) == 0) {
            base = reinterpret_cast<const char*>(&p.x());
            scalar = (int)sizeof(S);
          }
        }
      }
    }
  }

  double _resolution;
  bnx_map_t* _map = nullptr;
  VoxelGrid<CellT> _grid{nullptr, 1.0, 2, 3};
  Options _options;
};

// std::vector allocator over pinned host memory (bnx_host_alloc): point clouds kept in such a vector are copied
// to the GPU by DMA without a staging pass.
template <class T>
struct PinnedAllocator {
  using value_type = T;
  PinnedAllocator() = default;
  template <class U>
  PinnedAllocator(const PinnedAllocator<U>&) {}
  T* allocate(size_t n) {
    void* p = nullptr;
    detail::check(bnx_host_alloc(&p, n * sizeof(T)));
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { bnx_host_free(p); }
  template <class U>
  bool operator==(const PinnedAllocator<U>&) const { return true; }
  template <class U>
  bool operator!=(const PinnedAllocator<U>&) const { return false; }
};

}  // namespace Bonxai
