// One caller program, two builds:
//   * against the REFERENCE headers (+ oracle/shim Eigen stand-in)  -> tests/golden/dropin_expected.txt
//     (tests/golden/make_dropin_expected.sh, run where /root/reference exists)
//   * against include/ (this repo's drop-in headers) + libbonxai_b200.so on the GPU box
// The two outputs must be identical: that is the "drop-in" claim for the C++ API. Only API that exists in both
// is used; every printed quantity is order independent (forEachCell order is unspecified in both).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <sstream>

#include "bonxai/bonxai.hpp"
#include "bonxai/serialization.hpp"
#include "bonxai_map/probabilistic_map.hpp"

struct PointXYZ {  // pcl::PointXYZ layout
  float x, y, z, pad;
};

static uint64_t lcg(uint64_t& s) {
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return s >> 33;
}
static double uni(uint64_t& s, double lo, double hi) { return lo + (hi - lo) * (double)(lcg(s) % 1000003) / 1000003.0; }

template <class GridT>
static void printDigest(const char* name, GridT& grid) {
  // order-independent digest of (coord, value bytes)
  uint64_t sum = 0, xr = 0, n = 0;
  grid.forEachCell([&](auto& v, const Bonxai::CoordT& c) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](const void* p, size_t len) {
      const unsigned char* b = static_cast<const unsigned char*>(p);
      for (size_t i = 0; i < len; ++i) h = (h ^ b[i]) * 1099511628211ull;
    };
    mix(&c.x, 4);
    mix(&c.y, 4);
    mix(&c.z, 4);
    mix(&v, sizeof(v));
    sum += h;
    xr ^= h * 0x9E3779B97F4A7C15ull;
    ++n;
  });
  std::printf("%s: cells=%llu sum=%016llx xor=%016llx\n", name, (unsigned long long)n, (unsigned long long)sum, (unsigned long long)xr);
}

int main() {
  // ---------------------------------------------------------------- VoxelGrid (examples/tutorial.cpp scenario)
  {
    const double RES = 0.1;
    Bonxai::VoxelGrid<long> grid(RES);
    Bonxai::BinaryVoxelGrid binaryGrid(RES);
    auto accessor = grid.createAccessor();
    auto binaryAccessor = binaryGrid.createAccessor();
    long count = 0;
    for (double x = -1.0; x < 1.0; x += RES)
      for (double y = -1.0; y < 1.0; y += RES)
        for (double z = -1.0; z < 1.0; z += RES) {
          const Bonxai::CoordT coord = grid.posToCoord(x, y, z);
          accessor.setValue(coord, count++);
          binaryAccessor.setCellOn(coord);
        }
    std::printf("cells grid/binary: %zu/%zu\n", grid.activeCellsCount(), binaryGrid.activeCellsCount());
    auto* origin_ptr = accessor.value(grid.posToCoord(0, 0, 0));
    *origin_ptr = 500;
    std::printf("value at origin: %ld\n", *origin_ptr);
    auto* far_voxel = accessor.value(grid.posToCoord(10, 10, 10), true);
    (*far_voxel)++;
    std::printf("value at (10,10,10): %ld\n", *far_voxel);
    std::printf("setValue on existing returns %d, on new returns %d\n", (int)accessor.setValue(grid.posToCoord(0, 0, 0), 501),
                (int)accessor.setValue(grid.posToCoord(20, 0, 0), 7));
    auto mutableVisitor = [&grid, &accessor](auto& value, const Bonxai::CoordT& coord) {
      Bonxai::Point3D pos = grid.coordToPos(coord);
      if (pos.z < -0.1) {
        accessor.setCellOff(coord);
      } else {
        value = 1;
      }
    };
    grid.forEachCell(mutableVisitor);
    std::printf("cells after visitor: %zu\n", grid.activeCellsCount());
    printDigest("grid<long>", grid);
    auto constAccessor = grid.createConstAccessor();
    const auto* cell = constAccessor.value(grid.posToCoord(0, 0, 0));
    std::printf("const value at origin: %s\n", cell ? std::to_string(*cell).c_str() : "nullptr");
    cell = constAccessor.value(grid.posToCoord(0, 0, -0.2));
    std::printf("const value at (0,0,-0.2): %s\n", cell ? std::to_string(*cell).c_str() : "nullptr");
    std::printf("isCellOn: %d %d\n", (int)constAccessor.isCellOn(grid.posToCoord(0.5, 0.5, 0.5)), (int)constAccessor.isCellOn(grid.posToCoord(5, 5, 5)));
    const Bonxai::CoordT k = {-37, 100, 7};
    const auto rk = grid.getRootKey(k), ik = grid.getInnerKey(k);
    std::printf("keys: root %d %d %d inner %d %d %d idx %u %u\n", rk.x, rk.y, rk.z, ik.x, ik.y, ik.z, grid.getInnerIndex(k), grid.getLeafIndex(k));
    grid.clear(Bonxai::SET_ALL_CELLS_OFF);
    std::printf("after SET_ALL_CELLS_OFF: %zu\n", grid.activeCellsCount());
    auto* again = accessor.value(grid.posToCoord(0, 0, 0), true);
    std::printf("re-created value: %ld\n", *again);
    grid.releaseUnusedMemory();
    grid.clear(Bonxai::CLEAR_MEMORY);
    std::printf("after CLEAR_MEMORY: %zu\n", grid.activeCellsCount());
    bool threw = false;
    try {
      Bonxai::VoxelGrid<int> bad(0.1, 0, 3);
    } catch (const std::runtime_error&) {
      threw = true;
    }
    std::printf("bits<1 throws: %d\n", (int)threw);
  }
  // ---------------------------------------------------------------- ProbabilisticMap
  {
    using Map = Bonxai::ProbabilisticMap;
    std::printf("logods: %d %d %d %d %d prob(847297)=%.7f\n", Map::logods(0.4f), Map::logods(0.7f), Map::logods(0.12f), Map::logods(0.97f),
                Map::logods(0.5f), (double)Map::prob(847297));
    Map map(0.05);
    uint64_t seed = 42;
    for (int scan = 0; scan < 4; ++scan) {
      std::vector<PointXYZ> cloud;
      const PointXYZ origin = {(float)(0.3 * scan), 0.1f, 0.2f, 0.f};
      for (int i = 0; i < 3000; ++i) {
        const double az = uni(seed, -3.14159, 3.14159), el = uni(seed, -0.5, 0.5), r = uni(seed, 0.3, 4.5);
        cloud.push_back({(float)(origin.x + r * std::cos(el) * std::cos(az)), (float)(origin.y + r * std::cos(el) * std::sin(az)),
                         (float)(origin.z + r * std::sin(el)), 0.f});
      }
      map.insertPointCloud(cloud, origin, 3.0);
      std::printf("scan %d: active=%zu\n", scan, map.grid().activeCellsCount());
    }
    printDigest("map after float scans", map.grid());
    std::vector<Map::Vector3D> cloud_d;
    for (int i = 0; i < 2000; ++i) cloud_d.emplace_back(uni(seed, -2, 2), uni(seed, -2, 2), uni(seed, -0.5, 1.0));
    map.insertPointCloud(cloud_d, Map::Vector3D(0.0, 0.0, 0.3), 1.5);
    printDigest("map after double scan", map.grid());
    map.addHitPoint(Map::Vector3D(1.0, 1.0, 1.0));
    map.addMissPoint(Map::Vector3D(-1.0, 0.5, 0.25));
    std::vector<Bonxai::Point3D> cloud_p = {{0.5, 0.5, 0.5}, {1.0, 1.0, 1.0}, {-0.7, 0.2, 0.1}};
    map.insertPointCloud(cloud_p, Bonxai::Point3D(0.0, 0.0, 0.0), 10.0);
    printDigest("map after queued points", map.grid());
    std::vector<Bonxai::CoordT> occ, fre;
    map.getOccupiedVoxels(occ);
    map.getFreeVoxels(fre);
    std::printf("occupied=%zu free=%zu\n", occ.size(), fre.size());
    const auto c_hit = map.grid().posToCoord(1.0, 1.0, 1.0);
    std::printf("query hit voxel: occ %d unk %d free %d; far voxel: occ %d unk %d free %d\n", (int)map.isOccupied(c_hit), (int)map.isUnknown(c_hit),
                (int)map.isFree(c_hit), (int)map.isOccupied({9999, 0, 0}), (int)map.isUnknown({9999, 0, 0}), (int)map.isFree({9999, 0, 0}));
    std::vector<Bonxai::Point3D> occ_pts;
    map.getOccupiedVoxels(occ_pts);
    double sx = 0, sy = 0, sz = 0;
    for (const auto& p : occ_pts) {
      sx += p.x;
      sy += p.y;
      sz += p.z;
    }
    std::printf("occupied points=%zu centroid*n=%.6f %.6f %.6f\n", occ_pts.size(), sx, sy, sz);
    Map::Options opt = map.options();
    opt.prob_hit_log = 1500000;
    opt.occupancy_threshold_log = 300000;
    map.setOptions(opt);
    map.insertPointCloud(cloud_p, Bonxai::Point3D(0.0, 0.0, 0.0), 10.0);
    map.getOccupiedVoxels(occ);
    std::printf("after setOptions: occupied=%zu hit=%d\n", occ.size(), map.options().prob_hit_log);
    std::vector<Bonxai::CoordT> ray;
    Bonxai::ComputeRay({0, 0, 0}, {12, 6, 80}, ray);
    std::printf("ray cells=%zu last=%d %d %d\n", ray.size(), ray.back().x, ray.back().y, ray.back().z);
    Bonxai::ComputeRay({5, -3, 2}, {-40, 17, -9}, ray);
    long acc = 0;
    for (const auto& c : ray) acc = acc * 31 + c.x * 7 + c.y * 3 + c.z;
    std::printf("ray2 cells=%zu hash=%ld\n", ray.size(), acc);
  }
  // ---------------------------------------------------------------- Serialize / Deserialize (examples/test_serialization.cpp)
  {
    Bonxai::VoxelGrid<int> grid(0.1);
    auto accessor = grid.createAccessor();
    int count = 0;
    for (double x = -0.5; x < 0.5; x += 0.1)
      for (double y = -0.5; y < 0.5; y += 0.1)
        for (double z = -0.5; z < 0.5; z += 0.1) accessor.setValue(grid.posToCoord(x, y, z), count++);
    accessor.setCellOff(grid.posToCoord(0.2, 0.2, 0.2));
    std::ostringstream ofile(std::ios::binary);
    Bonxai::Serialize(ofile, grid);
    const std::string msg = ofile.str();
    std::istringstream ifile(msg, std::ios::binary);
    char header[256];
    ifile.getline(header, 256);
    Bonxai::HeaderInfo info = Bonxai::GetHeaderInfo(header);
    std::printf("stream: %zu bytes, header '%s' -> type %s bits %d/%d res %.3f\n", msg.size(), header, info.type_name.c_str(), info.inner_bits,
                info.leaf_bits, info.resolution);
    auto new_grid = Bonxai::Deserialize<int>(ifile, info);
    std::printf("cells %zu -> %zu\n", grid.activeCellsCount(), new_grid.activeCellsCount());
    printDigest("original", grid);
    printDigest("deserialized", new_grid);
    bool threw = false;
    try {
      std::istringstream again(msg, std::ios::binary);
      again.getline(header, 256);
      auto wrong = Bonxai::Deserialize<float>(again, info);
    } catch (const std::runtime_error&) {
      threw = true;
    }
    std::printf("type mismatch throws: %d\n", (int)threw);
  }
  return 0;
}
