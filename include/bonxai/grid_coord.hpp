// bonxai_b200 drop-in for bonxai_core/include/bonxai/grid_coord.hpp — host-side value types only.
//
// Same public names as the reference (Bonxai::CoordT :62-78, Point3D :32-60, ConvertPoint :134-162,
// PosToCoord :80-85, CoordToPos :87-91, std::hash<CoordT> :209-218) so that caller code compiles unchanged.
// Everything here is plain host arithmetic; the voxel storage lives on the GPU behind include/bonxai_b200.h.
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace Bonxai {

struct CoordT {
  int32_t x, y, z;

  int32_t& operator[](size_t i) {
    if (i > 2) throw std::runtime_error("out of bound index");
    return i == 0 ? x : (i == 1 ? y : z);
  }
  bool operator==(const CoordT& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const CoordT& o) const { return !(*this == o); }
  CoordT operator+(const CoordT& o) const { return {x + o.x, y + o.y, z + o.z}; }
  CoordT operator-(const CoordT& o) const { return {x - o.x, y - o.y, z - o.z}; }
  CoordT& operator+=(const CoordT& o) { return *this = *this + o; }
  CoordT& operator-=(const CoordT& o) { return *this = *this - o; }
};
static_assert(sizeof(CoordT) == 12, "CoordT must be three packed int32 (the C ABI passes arrays of them)");

namespace detail {
template <class T, class = void>
struct has_xyz_methods : std::false_type {};
template <class T>
struct has_xyz_methods<T, std::void_t<decltype(std::declval<const T&>().x()), decltype(std::declval<const T&>().y()),
                                      decltype(std::declval<const T&>().z())>> : std::true_type {};
template <class T, class = void>
struct has_xyz_fields : std::false_type {};
template <class T>
struct has_xyz_fields<T, std::void_t<decltype(T::x), decltype(T::y), decltype(T::z)>> : std::true_type {};
template <class T>
struct is_triplet : std::false_type {};
template <class T, class A>
struct is_triplet<std::vector<T, A>> : std::true_type {};
template <class T>
struct is_triplet<std::array<T, 3>> : std::true_type {};
}  // namespace detail

template <typename PointOut, typename PointIn>
PointOut ConvertPoint(const PointIn& v);

struct Point3D {
  double x, y, z;

  Point3D() = default;
  Point3D(const Point3D&) = default;
  Point3D(Point3D&&) = default;
  Point3D& operator=(const Point3D&) = default;
  Point3D& operator=(Point3D&&) = default;
  Point3D(double px, double py, double pz) : x(px), y(py), z(pz) {}

  template <typename T>
  Point3D(const T& v) {
    *this = ConvertPoint<Point3D>(v);
  }
  template <typename T>
  Point3D& operator=(const T& v) {
    *this = ConvertPoint<Point3D>(v);
    return *this;
  }
  double& operator[](size_t i) {
    if (i > 2) throw std::runtime_error("out of bound index");
    return i == 0 ? x : (i == 1 ? y : z);
  }
};

// Any {x,y,z} representation to any other: types with x()/y()/z() (Eigen), with public fields (pcl::PointXYZ,
// Point3D) or indexable triplets (std::array<T,3>, std::vector<T>).
template <typename PointOut, typename PointIn>
inline PointOut ConvertPoint(const PointIn& v) {
  constexpr bool same = std::is_same_v<PointIn, PointOut>;
  static_assert(same || detail::has_xyz_methods<PointIn>::value || detail::has_xyz_fields<PointIn>::value ||
                    detail::is_triplet<PointIn>::value,
                "Can't convert from the specified type");
  static_assert(same || detail::has_xyz_methods<PointOut>::value || detail::has_xyz_fields<PointOut>::value ||
                    detail::is_triplet<PointOut>::value,
                "Can't convert to the specified type");
  if constexpr (same) {
    return v;
  } else if constexpr (detail::has_xyz_methods<PointIn>::value) {
    return {v.x(), v.y(), v.z()};
  } else if constexpr (detail::has_xyz_fields<PointIn>::value) {
    return {v.x, v.y, v.z};
  } else {
    return {v[0], v[1], v[2]};
  }
}

// floor(p * inv_resolution) per axis, then the C cast to int32 — one rounded fp64 multiply each, exactly the
// arithmetic the device kernels reproduce with __dmul_rn / __double2int_rd.
inline CoordT PosToCoord(const Point3D& p, double inv_resolution) {
  return {static_cast<int32_t>(std::floor(p.x * inv_resolution)), static_cast<int32_t>(std::floor(p.y * inv_resolution)),
          static_cast<int32_t>(std::floor(p.z * inv_resolution))};
}

// voxel corner, not centre
inline Point3D CoordToPos(const CoordT& c, double resolution) {
  return {static_cast<double>(c.x) * resolution, static_cast<double>(c.y) * resolution, static_cast<double>(c.z) * resolution};
}

}  // namespace Bonxai

namespace std {
// For callers that keep CoordT in unordered containers. (The device root table does NOT use this hash: for
// root keys its low bits are dead) Same value as the reference's, which callers may rely on.
template <>
struct hash<Bonxai::CoordT> {
  size_t operator()(const Bonxai::CoordT& p) const {
    const int64_t h = static_cast<int64_t>(p.x) * 73856093 ^ static_cast<int64_t>(p.y) * 19349669 ^ static_cast<int64_t>(p.z) * 83492791;
    return static_cast<size_t>(h & ((1 << 20) - 1));
  }
};
}  // namespace std
