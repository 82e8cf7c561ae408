// TEST INFRASTRUCTURE ONLY — puts the UNMODIFIED reference behind oracle/oracle_api.h.
//
// This TU is compiled together with /root/reference/bonxai_map/src/probabilistic_map.cpp from where
// the reference lies (see oracle/Makefile); no reference source is copied into this repository and
// the resulting binary goes to oracle/_ref/ (git-ignored). Every function below only *calls* the
// reference's public API; there is no algorithm in this file.
#include <chrono>
#include <cstring>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "bonxai/bonxai.hpp"
#include "bonxai/serialization.hpp"
#include "bonxai_map/probabilistic_map.hpp"
#include "oracle_api.h"

namespace {

using Grid32 = Bonxai::VoxelGrid<uint32_t>;
using Map = Bonxai::ProbabilisticMap;

struct P3 {
  float x, y, z;
};
struct P4 {
  float x, y, z, pad;
};

struct MapBox {
  explicit MapBox(double res) : map(res) {}
  Map map;
  bool track = false;
  int64_t counters[4] = {0, -1, -1, -1};
  double last_seconds = 0.0;
};

inline uint32_t cellWord(const Map::CellT& c) {
  uint32_t w;
  static_assert(sizeof(Map::CellT) == 4, "CellT must be one 32-bit word");
  std::memcpy(&w, &c, 4);
  return w;
}

struct CoordHash {
  size_t operator()(const Bonxai::CoordT& c) const {
    uint64_t h = (uint64_t)(uint32_t)c.x * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)c.y + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= ((uint64_t)(uint32_t)c.z + 0x165667B1ull) * 0xD6E8FEB86659FD93ull;
    return (size_t)(h ^ (h >> 29));
  }
};
using Snapshot = std::unordered_map<Bonxai::CoordT, uint32_t, CoordHash>;

void snapshot(Map& map, Snapshot& s) {
  s.clear();
  s.reserve(map.grid().activeCellsCount() * 2);
  map.grid().forEachCell(
      [&](Map::CellT& c, const Bonxai::CoordT& p) { s.emplace(p, cellWord(c)); });
}

template <class PointT>
void timedInsert(MapBox* b, const std::vector<PointT>& pts, const PointT& origin, double max_range) {
  Snapshot before;
  if (b->track) {
    snapshot(b->map, before);
  }
  const auto t0 = std::chrono::steady_clock::now();
  b->map.insertPointCloud(pts, origin, max_range);
  const auto t1 = std::chrono::steady_clock::now();
  b->last_seconds = std::chrono::duration<double>(t1 - t0).count();
  b->counters[0] = (int64_t)pts.size();
  b->counters[1] = -1;
  b->counters[2] = -1;
  b->counters[3] = -1;
  if (b->track) {
    int64_t changed = 0;
    b->map.grid().forEachCell([&](Map::CellT& c, const Bonxai::CoordT& p) {
      auto it = before.find(p);
      if (it == before.end() || it->second != cellWord(c)) {
        ++changed;
      }
    });
    b->counters[3] = changed;
  }
}

inline Bonxai::CoordT C(const int32_t* p) {
  return {p[0], p[1], p[2]};
}

}  // namespace

extern "C" {

const char* orc_kind(void) {
  return "reference";
}

int32_t orc_logods(float prob) {
  return Map::logods(prob);
}
float orc_prob(int32_t v) {
  return Map::prob(v);
}

void orc_pos_to_coord(double resolution, const double* xyz, int64_t n, int32_t* out) {
  Grid32 g(resolution);
  for (int64_t i = 0; i < n; ++i) {
    const auto c = g.posToCoord(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    out[3 * i] = c.x;
    out[3 * i + 1] = c.y;
    out[3 * i + 2] = c.z;
  }
}

void orc_coord_to_pos(double resolution, const int32_t* xyz, int64_t n, double* out) {
  Grid32 g(resolution);
  for (int64_t i = 0; i < n; ++i) {
    const auto p = g.coordToPos(C(xyz + 3 * i));
    out[3 * i] = p.x;
    out[3 * i + 1] = p.y;
    out[3 * i + 2] = p.z;
  }
}

int64_t orc_compute_ray(const int32_t a[3], const int32_t b[3], int32_t* out_xyz, int64_t cap) {
  std::vector<Bonxai::CoordT> ray;
  Bonxai::ComputeRay(C(a), C(b), ray);
  for (int64_t i = 0; i < (int64_t)ray.size() && i < cap; ++i) {
    out_xyz[3 * i] = ray[i].x;
    out_xyz[3 * i + 1] = ray[i].y;
    out_xyz[3 * i + 2] = ray[i].z;
  }
  return (int64_t)ray.size();
}

// ---------------------------------------------------------------- VoxelGrid<uint32_t>
void* orc_grid_create(double voxel_size, int inner_bits, int leaf_bits) {
  try {
    return new Grid32(voxel_size, (uint8_t)inner_bits, (uint8_t)leaf_bits);
  } catch (const std::exception&) {
    return nullptr;
  }
}
void orc_grid_destroy(void* g) {
  delete static_cast<Grid32*>(g);
}

void orc_grid_set_values(void* g, const int32_t* xyz, const uint32_t* vals, int64_t n,
                         uint8_t* was_on) {
  auto acc = static_cast<Grid32*>(g)->createAccessor();
  for (int64_t i = 0; i < n; ++i) {
    const bool w = acc.setValue(C(xyz + 3 * i), vals[i]);
    if (was_on) {
      was_on[i] = w;
    }
  }
}

void orc_grid_get_values(void* g, const int32_t* xyz, int64_t n, uint32_t* out, uint8_t* found) {
  auto acc = static_cast<const Grid32*>(g)->createConstAccessor();
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t* v = acc.value(C(xyz + 3 * i));
    found[i] = v != nullptr;
    if (v) {
      out[i] = *v;
    }
  }
}

void orc_grid_get_or_create(void* g, const int32_t* xyz, int64_t n, uint32_t* out) {
  auto acc = static_cast<Grid32*>(g)->createAccessor();
  for (int64_t i = 0; i < n; ++i) {
    out[i] = *acc.value(C(xyz + 3 * i), true);
  }
}

void orc_grid_set_on(void* g, const int32_t* xyz, int64_t n, uint32_t default_value,
                     uint8_t* was_on) {
  auto acc = static_cast<Grid32*>(g)->createAccessor();
  for (int64_t i = 0; i < n; ++i) {
    const bool w = acc.setCellOn(C(xyz + 3 * i), default_value);
    if (was_on) {
      was_on[i] = w;
    }
  }
}

void orc_grid_set_off(void* g, const int32_t* xyz, int64_t n, uint8_t* was_on) {
  auto acc = static_cast<Grid32*>(g)->createAccessor();
  for (int64_t i = 0; i < n; ++i) {
    const bool w = acc.setCellOff(C(xyz + 3 * i));
    if (was_on) {
      was_on[i] = w;
    }
  }
}

void orc_grid_is_on(void* g, const int32_t* xyz, int64_t n, uint8_t* out) {
  auto acc = static_cast<const Grid32*>(g)->createConstAccessor();
  for (int64_t i = 0; i < n; ++i) {
    out[i] = acc.isCellOn(C(xyz + 3 * i));
  }
}

int64_t orc_grid_active_count(void* g) {
  return (int64_t) static_cast<Grid32*>(g)->activeCellsCount();
}

int64_t orc_grid_dump(void* g, int32_t* xyz, uint32_t* vals, int64_t cap) {
  int64_t n = 0;
  static_cast<Grid32*>(g)->forEachCell([&](uint32_t& v, const Bonxai::CoordT& p) {
    if (n < cap) {
      xyz[3 * n] = p.x;
      xyz[3 * n + 1] = p.y;
      xyz[3 * n + 2] = p.z;
      vals[n] = v;
    }
    ++n;
  });
  return n;
}

void orc_grid_clear(void* g, int opt) {
  static_cast<Grid32*>(g)->clear(opt == 0 ? Bonxai::CLEAR_MEMORY : Bonxai::SET_ALL_CELLS_OFF);
}

void orc_grid_release_unused(void* g) {
  static_cast<Grid32*>(g)->releaseUnusedMemory();
}

int64_t orc_grid_serialize(void* g, uint8_t* out, int64_t cap) {
  std::ostringstream ss(std::ios::out | std::ios::binary);
  Bonxai::Serialize(ss, *static_cast<Grid32*>(g));
  const std::string s = ss.str();
  if ((int64_t)s.size() <= cap && out) {
    std::memcpy(out, s.data(), s.size());
  }
  return (int64_t)s.size();
}

void* orc_grid_deserialize(const uint8_t* data, int64_t len) {
  try {
    std::istringstream ss(std::string(reinterpret_cast<const char*>(data), (size_t)len),
                          std::ios::in | std::ios::binary);
    char header[256];
    ss.getline(header, 256);
    const auto info = Bonxai::GetHeaderInfo(header);
    return new Grid32(Bonxai::Deserialize<uint32_t>(ss, info));
  } catch (const std::exception&) {
    return nullptr;
  }
}

// ---------------------------------------------------------------- ProbabilisticMap
void* orc_map_create(double resolution) {
  return new MapBox(resolution);
}
void orc_map_destroy(void* m) {
  delete static_cast<MapBox*>(m);
}

void orc_map_set_options(void* m, const int32_t o[5]) {
  Map::Options opt;
  opt.prob_miss_log = o[0];
  opt.prob_hit_log = o[1];
  opt.clamp_min_log = o[2];
  opt.clamp_max_log = o[3];
  opt.occupancy_threshold_log = o[4];
  static_cast<MapBox*>(m)->map.setOptions(opt);
}

void orc_map_get_options(void* m, int32_t o[5]) {
  const auto& opt = static_cast<MapBox*>(m)->map.options();
  o[0] = opt.prob_miss_log;
  o[1] = opt.prob_hit_log;
  o[2] = opt.clamp_min_log;
  o[3] = opt.clamp_max_log;
  o[4] = opt.occupancy_threshold_log;
}

void orc_map_insert_f32(void* m, const void* pts, int64_t stride_bytes, int64_t n,
                        const float origin[3], double max_range) {
  auto* b = static_cast<MapBox*>(m);
  if (stride_bytes == 16) {
    std::vector<P4> v((size_t)n);
    if (n) {
      std::memcpy(v.data(), pts, (size_t)n * 16);
    }
    timedInsert(b, v, P4{origin[0], origin[1], origin[2], 0.f}, max_range);
  } else {
    std::vector<P3> v((size_t)n);
    if (n) {
      std::memcpy(v.data(), pts, (size_t)n * 12);
    }
    timedInsert(b, v, P3{origin[0], origin[1], origin[2]}, max_range);
  }
}

void orc_map_insert_f64(void* m, const double* pts, int64_t n, const double origin[3],
                        double max_range) {
  auto* b = static_cast<MapBox*>(m);
  std::vector<Eigen::Vector3d> v;
  v.reserve((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    v.emplace_back(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  }
  timedInsert(b, v, Eigen::Vector3d(origin[0], origin[1], origin[2]), max_range);
}

void orc_map_add_hit(void* m, const double p[3]) {
  static_cast<MapBox*>(m)->map.addHitPoint(Eigen::Vector3d(p[0], p[1], p[2]));
}
void orc_map_add_miss(void* m, const double p[3]) {
  static_cast<MapBox*>(m)->map.addMissPoint(Eigen::Vector3d(p[0], p[1], p[2]));
}

void orc_map_query(void* m, const int32_t* xyz, int64_t n, int kind, uint8_t* out) {
  const Map& map = static_cast<MapBox*>(m)->map;
  for (int64_t i = 0; i < n; ++i) {
    const auto c = C(xyz + 3 * i);
    out[i] = kind == 0 ? map.isOccupied(c) : kind == 1 ? map.isUnknown(c) : map.isFree(c);
  }
}

int64_t orc_map_get_voxels(void* m, int kind, int32_t* xyz, int64_t cap) {
  std::vector<Bonxai::CoordT> coords;
  if (kind == 0) {
    static_cast<MapBox*>(m)->map.getOccupiedVoxels(coords);
  } else {
    static_cast<MapBox*>(m)->map.getFreeVoxels(coords);
  }
  for (int64_t i = 0; i < (int64_t)coords.size() && i < cap; ++i) {
    xyz[3 * i] = coords[i].x;
    xyz[3 * i + 1] = coords[i].y;
    xyz[3 * i + 2] = coords[i].z;
  }
  return (int64_t)coords.size();
}

int64_t orc_map_active_count(void* m) {
  return (int64_t) static_cast<MapBox*>(m)->map.grid().activeCellsCount();
}

int64_t orc_map_dump(void* m, int32_t* xyz, uint32_t* words, int64_t cap) {
  int64_t n = 0;
  static_cast<MapBox*>(m)->map.grid().forEachCell([&](Map::CellT& c, const Bonxai::CoordT& p) {
    if (n < cap) {
      xyz[3 * n] = p.x;
      xyz[3 * n + 1] = p.y;
      xyz[3 * n + 2] = p.z;
      words[n] = cellWord(c);
    }
    ++n;
  });
  return n;
}

void orc_map_counters(void* m, int64_t out[4]) {
  std::memcpy(out, static_cast<MapBox*>(m)->counters, sizeof(int64_t) * 4);
}
void orc_map_track_updates(void* m, int enable) {
  static_cast<MapBox*>(m)->track = enable != 0;
}
double orc_map_last_insert_seconds(void* m) {
  return static_cast<MapBox*>(m)->last_seconds;
}

}  // extern "C"
