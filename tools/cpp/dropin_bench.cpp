// e2e through the drop-in C++ API: what a Bonxai user who swaps the include path gets.
//   dropin_bench <scans.bin> <resolution> <max_range> <warmup> <mode>
// scans.bin: int64 n_scans, int64 n_points, then per scan {float origin[3]; float xyz_pad[n_points][4]}.
// mode: "vector"   std::vector<PointXYZ> (pageable) + insertPointCloud per scan (synchronous, as in bonxai_server.cpp:176-182)
//       "pinned"   std::vector<PointXYZ, Bonxai::PinnedAllocator<PointXYZ>>: same call, the H2D copy is a direct DMA
//       "publish"  pinned + after EVERY insert the publisher's post-step (bonxai_server.cpp:186,217-251) through the fused
//                  call (occupied voxels -> coord*res -> z window -> float xyz on the host)
//       "publish_ref" pinned + after every insert getOccupiedVoxels(std::vector<Point3D>&) + the caller's own z filter /
//                  float conversion loop, i.e. the reference node's code unchanged
// Prints one JSON object. The clock brackets the calls only (steady_clock), first <warmup> scans untimed.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bonxai/bonxai.hpp"
#include "bonxai_map/probabilistic_map.hpp"

struct PointXYZ {  // pcl::PointXYZ layout
  float x, y, z, pad;
};

template <class Vec>
static int run(const std::vector<std::vector<PointXYZ>>& scans, const std::vector<PointXYZ>& origins, double res, double max_range, int warmup,
               const std::string& mode) {
  using Clock = std::chrono::steady_clock;
  Bonxai::ProbabilisticMap map(res);
  std::vector<Vec> clouds(scans.size());
  for (size_t i = 0; i < scans.size(); ++i) clouds[i].assign(scans[i].begin(), scans[i].end());
  std::vector<float, Bonxai::PinnedAllocator<float>> published;  // pinned: the device-to-host copy of the cloud is a direct DMA
  std::vector<Bonxai::Point3D> voxels;
  std::vector<PointXYZ> pcl_cloud;
  double secs = 0.0;
  size_t last_published = 0;
  const double zmin = -100.0, zmax = 100.0;  // occupancy_min_z / occupancy_max_z defaults of the node
  for (size_t i = 0; i < clouds.size(); ++i) {
    const auto t0 = Clock::now();
    map.insertPointCloud(clouds[i], origins[i], max_range);
    if (mode == "publish") {
      // one call per scan in the steady state: the buffer keeps 25 % of head-room over the last count, and only a scan
      // that outgrows it pays a second pass (BNX_ERR_CAPACITY reports the count without writing)
      int64_t n = 0;
      int st = bnx_map_publish_occupied_f32(map.handle(), zmin, zmax, published.data(), 4, (int64_t)(published.size() / 4), &n, BNX_HOST);
      if (st == BNX_ERR_CAPACITY) {
        published.resize((size_t)(n + n / 4 + 1024) * 4);
        st = bnx_map_publish_occupied_f32(map.handle(), zmin, zmax, published.data(), 4, (int64_t)(published.size() / 4), &n, BNX_HOST);
      }
      if (st != BNX_OK) return 2;
      last_published = (size_t)n;
    } else if (mode == "publish_ref") {
      voxels.clear();
      map.getOccupiedVoxels(voxels);
      pcl_cloud.clear();
      for (const auto& v : voxels)
        if (v.z >= zmin && v.z <= zmax) pcl_cloud.push_back(PointXYZ{(float)v.x, (float)v.y, (float)v.z, 1.0f});
      last_published = pcl_cloud.size();
    }
    const auto t1 = Clock::now();
    if ((int)i >= warmup) secs += std::chrono::duration<double>(t1 - t0).count();
  }
  const size_t timed = clouds.size() - (size_t)warmup;
  std::printf("{\"mode\": \"%s\", \"scans_timed\": %zu, \"points_per_scan\": %zu, \"us_per_scan\": %.3f, \"points_per_s\": %.1f, "
              "\"active_cells\": %zu, \"published_points_last\": %zu}\n",
              mode.c_str(), timed, scans[0].size(), 1e6 * secs / (double)timed, (double)timed * (double)scans[0].size() / secs,
              map.grid().activeCellsCount(), last_published);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s scans.bin resolution max_range warmup vector|pinned|publish|publish_ref\n", argv[0]);
    return 1;
  }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 1;
  int64_t hdr[2];
  if (std::fread(hdr, 8, 2, f) != 2) return 1;
  std::vector<std::vector<PointXYZ>> scans((size_t)hdr[0]);
  std::vector<PointXYZ> origins((size_t)hdr[0]);
  for (auto i = 0; i < hdr[0]; ++i) {
    float o[3];
    if (std::fread(o, 4, 3, f) != 3) return 1;
    origins[(size_t)i] = PointXYZ{o[0], o[1], o[2], 0.f};
    scans[(size_t)i].resize((size_t)hdr[1]);
    if (std::fread(scans[(size_t)i].data(), sizeof(PointXYZ), (size_t)hdr[1], f) != (size_t)hdr[1]) return 1;
  }
  std::fclose(f);
  const double res = std::atof(argv[2]), max_range = std::atof(argv[3]);
  const int warmup = std::atoi(argv[4]);
  const std::string mode = argv[5];
  try {
    if (mode == "vector") return run<std::vector<PointXYZ>>(scans, origins, res, max_range, warmup, mode);
    return run<std::vector<PointXYZ, Bonxai::PinnedAllocator<PointXYZ>>>(scans, origins, res, max_range, warmup, mode);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "dropin_bench: %s\n", e.what());
    return 3;
  }
}
