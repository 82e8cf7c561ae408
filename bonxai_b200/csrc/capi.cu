// extern "C" surface declared in include/bonxai_b200.h. Handles are heap objects of the C++ classes;
// nothing but plain pointers and sizes crosses this file.
#include <cstring>
#include <new>

#include "map.hpp"

using namespace bnx;

struct bnx_grid {
  Grid* g;
  bool owned;  // false for the grid borrowed from a map
  Map* map;    // that map: its pipelined scans are completed before the grid is touched
};
struct bnx_map {
  Map m;
  bnx_grid grid_handle;
};

#define BNX_HANDLE(h)                       \
  do {                                      \
    if (!(h)) {                             \
      set_error("null handle");             \
      return BNX_ERR_INVALID;               \
    }                                       \
  } while (0)

namespace {
// run the call on the device the handle lives on, whatever the caller's current device is
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
    } else {
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace

extern "C" {

int bnx_version(void) { return 100; }
int64_t bnx_launch_count(void) { return (int64_t)launch_count(); }
const char* bnx_last_error(void) { return get_error(); }

int bnx_device_count(int* count) {
  BNX_REQUIRE(count != nullptr, "null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
    cudaGetLastError();
    *count = 0;
    return BNX_OK;
  }
  BNX_CUDA(e);
  *count = n;
  return BNX_OK;
}

int bnx_host_alloc(void** ptr, size_t bytes) {
  BNX_REQUIRE(ptr != nullptr, "null output");
  BNX_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
  return BNX_OK;
}
int bnx_host_free(void* ptr) {
  if (ptr) BNX_CUDA(cudaFreeHost(ptr));
  return BNX_OK;
}

// ---------------------------------------------------------------------------------- VoxelGrid
int bnx_grid_create(double voxel_size, int inner_bits, int leaf_bits, int cell_bytes, bnx_grid_t** out) {
  BNX_REQUIRE(out != nullptr, "null output");
  *out = nullptr;
  Grid* g = new (std::nothrow) Grid();
  if (!g) return BNX_ERR_NOMEM;
  const int s = g->init(voxel_size, inner_bits, leaf_bits, cell_bytes);
  if (s != BNX_OK) {
    delete g;
    return s;
  }
  *out = new bnx_grid{g, true, nullptr};
  return BNX_OK;
}

int bnx_grid_destroy(bnx_grid_t* h) {
  if (!h) return BNX_OK;
  if (!h->owned) {
    set_error("this grid handle is owned by its map");
    return BNX_ERR_INVALID;
  }
  {
    DeviceGuard dg(h->g->device);
    delete h->g;
  }
  delete h;
  return BNX_OK;
}

int bnx_grid_set_stream(bnx_grid_t* h, void* stream) {
  BNX_HANDLE(h);
  h->g->set_stream(static_cast<cudaStream_t>(stream));
  return BNX_OK;
}
int bnx_grid_sync(bnx_grid_t* h) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->g->device);
  return h->g->sync();
}
int bnx_grid_info(const bnx_grid_t* h, double* voxel_size, int* inner_bits, int* leaf_bits, int* cell_bytes) {
  BNX_HANDLE(h);
  if (voxel_size) *voxel_size = h->g->resolution;
  if (inner_bits) *inner_bits = h->g->inner_bits;
  if (leaf_bits) *leaf_bits = h->g->leaf_bits;
  if (cell_bytes) *cell_bytes = h->g->cell_bytes;
  return BNX_OK;
}

#define GRID_CALL(h, expr)                       \
  BNX_HANDLE(h);                                 \
  DeviceGuard dg((h)->g->device);                \
  if ((h)->map) BNX_TRY((h)->map->drain());      \
  return (h)->g->expr

int bnx_grid_pos_to_coord(const bnx_grid_t* h, const double* xyz, int64_t n, int32_t* out, int where) {
  GRID_CALL(h, pos_to_coord(xyz, n, out, where));
}
int bnx_grid_coord_to_pos(const bnx_grid_t* h, const int32_t* xyz, int64_t n, double* out, int where) {
  GRID_CALL(h, coord_to_pos(xyz, n, out, where));
}
int bnx_grid_set_values(bnx_grid_t* h, const int32_t* xyz, const void* values, int64_t n, uint8_t* was_on, int where) {
  GRID_CALL(h, set_values(xyz, values, n, was_on, where));
}
int bnx_grid_get_values(bnx_grid_t* h, const int32_t* xyz, int64_t n, void* values, uint8_t* found, int where) {
  GRID_CALL(h, get_values(xyz, n, values, found, where));
}
int bnx_grid_get_or_create(bnx_grid_t* h, const int32_t* xyz, int64_t n, void* values, int where) {
  GRID_CALL(h, get_or_create(xyz, n, values, where));
}
int bnx_grid_update_values(bnx_grid_t* h, const int32_t* xyz, const void* values, int64_t n, int where) {
  GRID_CALL(h, update_values(xyz, values, n, where));
}
int bnx_grid_set_on(bnx_grid_t* h, const int32_t* xyz, int64_t n, const void* default_value, uint8_t* was_on, int where) {
  GRID_CALL(h, set_on(xyz, n, default_value, was_on, where));
}
int bnx_grid_set_off(bnx_grid_t* h, const int32_t* xyz, int64_t n, uint8_t* was_on, int where) {
  GRID_CALL(h, set_off(xyz, n, was_on, where));
}
int bnx_grid_is_on(bnx_grid_t* h, const int32_t* xyz, int64_t n, uint8_t* out, int where) {
  GRID_CALL(h, is_on(xyz, n, out, where));
}
int bnx_grid_active_count(bnx_grid_t* h, int64_t* count) { GRID_CALL(h, active_count(count)); }
int bnx_grid_digest(bnx_grid_t* h, uint64_t out[3]) { GRID_CALL(h, digest(reinterpret_cast<u64*>(out))); }
int bnx_grid_dump(bnx_grid_t* h, int32_t* xyz, void* values, int64_t cap, int64_t* count, int where) {
  GRID_CALL(h, dump(xyz, nullptr, values, cap, count, where, -1, 0));
}
int bnx_grid_clear(bnx_grid_t* h, int option) { GRID_CALL(h, clear(option)); }
int bnx_grid_release_unused(bnx_grid_t* h) { GRID_CALL(h, release_unused()); }
int bnx_grid_mem_usage(bnx_grid_t* h, int64_t* bytes) { GRID_CALL(h, mem_usage(bytes)); }
int bnx_grid_stats(bnx_grid_t* h, int64_t out[8]) { GRID_CALL(h, stats(out)); }
int bnx_grid_serialize(bnx_grid_t* h, const char* type_name, uint8_t* buffer, int64_t cap, int64_t* size) {
  GRID_CALL(h, serialize(type_name, buffer, cap, size));
}
int bnx_grid_deserialize(const uint8_t* data, int64_t len, int cell_bytes, const char* expect_type_name, bnx_grid_t** out) {
  BNX_REQUIRE(out != nullptr, "null output");
  *out = nullptr;
  Grid* g = nullptr;
  BNX_TRY(Grid::deserialize(data, len, cell_bytes, expect_type_name, &g));
  *out = new bnx_grid{g, true, nullptr};
  return BNX_OK;
}

// ---------------------------------------------------------------------------------- ProbabilisticMap
int bnx_map_create(double resolution, bnx_map_t** out) {
  BNX_REQUIRE(out != nullptr, "null output");
  *out = nullptr;
  bnx_map* h = new (std::nothrow) bnx_map();
  if (!h) return BNX_ERR_NOMEM;
  const int s = h->m.init(resolution);
  if (s != BNX_OK) {
    delete h;
    return s;
  }
  h->grid_handle.g = &h->m.grid;
  h->grid_handle.owned = false;
  h->grid_handle.map = &h->m;
  *out = h;
  return BNX_OK;
}

int bnx_map_destroy(bnx_map_t* h) {
  if (!h) return BNX_OK;
  DeviceGuard dg(h->m.grid.device);
  delete h;
  return BNX_OK;
}

int bnx_map_set_stream(bnx_map_t* h, void* stream) {
  BNX_HANDLE(h);
  h->m.grid.set_stream(static_cast<cudaStream_t>(stream));
  return BNX_OK;
}
int bnx_map_sync(bnx_map_t* h) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  BNX_TRY(h->m.drain());
  return h->m.grid.sync();
}
int bnx_map_grid(bnx_map_t* h, bnx_grid_t** grid) {
  BNX_HANDLE(h);
  BNX_REQUIRE(grid != nullptr, "null output");
  *grid = &h->grid_handle;
  return BNX_OK;
}
int bnx_map_set_options(bnx_map_t* h, const int32_t options[5]) {
  BNX_HANDLE(h);
  BNX_REQUIRE(options != nullptr, "null options");
  std::memcpy(h->m.options, options, sizeof(int32_t) * 5);
  return BNX_OK;
}
int bnx_map_get_options(const bnx_map_t* h, int32_t options[5]) {
  BNX_HANDLE(h);
  BNX_REQUIRE(options != nullptr, "null options");
  std::memcpy(options, h->m.options, sizeof(int32_t) * 5);
  return BNX_OK;
}

int bnx_map_insert_f32(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, const float origin[3], double max_range,
                       int where) {
  BNX_HANDLE(h);
  BNX_REQUIRE(origin != nullptr, "null origin");
  DeviceGuard dg(h->m.grid.device);
  const double o[3] = {(double)origin[0], (double)origin[1], (double)origin[2]};  // ConvertPoint<Vector3D>(origin)
  return h->m.insert(points, stride_bytes, n, false, o, max_range, where);
}
int bnx_map_insert_f64(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, const double origin[3], double max_range,
                       int where) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.insert(points, stride_bytes, n, true, origin, max_range, where);
}
int bnx_map_insert_async_f32(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, const float origin[3], double max_range,
                             int where) {
  BNX_HANDLE(h);
  BNX_REQUIRE(origin != nullptr, "null origin");
  DeviceGuard dg(h->m.grid.device);
  const double o[3] = {(double)origin[0], (double)origin[1], (double)origin[2]};
  return h->m.insert_async(points, stride_bytes, n, false, o, max_range, where);
}
int bnx_map_insert_async_f64(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, const double origin[3], double max_range,
                             int where) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.insert_async(points, stride_bytes, n, true, origin, max_range, where);
}
int bnx_map_insert_transformed_f32(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, const float sensor_to_world[16],
                                   const float origin[3], double max_range, int where, int async) {
  BNX_HANDLE(h);
  BNX_REQUIRE(origin != nullptr && sensor_to_world != nullptr, "null origin / transform");
  DeviceGuard dg(h->m.grid.device);
  const double o[3] = {(double)origin[0], (double)origin[1], (double)origin[2]};
  if (!async) BNX_TRY(h->m.drain());
  h->m.set_next_transform(sensor_to_world);
  const int st = async ? h->m.insert_async(points, stride_bytes, n, false, o, max_range, where) : h->m.insert(points, stride_bytes, n, false, o, max_range, where);
  h->m.clear_next_transform();  // a call that failed before it consumed the transform must not leave it to the next insert
  return st;
}
int bnx_map_totals(bnx_map_t* h, int64_t out[4]) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  BNX_TRY(h->m.drain());
  std::memcpy(out, h->m.totals, sizeof(int64_t) * 4);
  return BNX_OK;
}
int bnx_map_add_hit(bnx_map_t* h, const double point[3]) {
  BNX_HANDLE(h);
  BNX_REQUIRE(point != nullptr, "null point");
  DeviceGuard dg(h->m.grid.device);
  return h->m.add_point(point, false);
}
int bnx_map_add_miss(bnx_map_t* h, const double point[3]) {
  BNX_HANDLE(h);
  BNX_REQUIRE(point != nullptr, "null point");
  DeviceGuard dg(h->m.grid.device);
  return h->m.add_point(point, true);
}
int bnx_map_query(bnx_map_t* h, const int32_t* xyz, int64_t n, int kind, uint8_t* out, int where) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.query(xyz, n, kind, out, where);
}
int bnx_map_get_voxels(bnx_map_t* h, int kind, int32_t* xyz, int64_t cap, int64_t* count, int where) {
  BNX_HANDLE(h);
  BNX_TRY(h->m.drain());
  BNX_REQUIRE(kind == BNX_OCCUPIED || kind == BNX_FREE, "get_voxels: kind must be BNX_OCCUPIED or BNX_FREE");
  DeviceGuard dg(h->m.grid.device);
  return h->m.grid.dump(xyz, nullptr, nullptr, cap, count, where, kind, h->m.options[4]);
}
int bnx_map_get_voxel_points(bnx_map_t* h, int kind, double* xyz, int64_t cap, int64_t* count, int where) {
  BNX_HANDLE(h);
  BNX_TRY(h->m.drain());
  BNX_REQUIRE(kind == BNX_OCCUPIED || kind == BNX_FREE, "get_voxel_points: kind must be BNX_OCCUPIED or BNX_FREE");
  DeviceGuard dg(h->m.grid.device);
  return h->m.grid.dump(nullptr, xyz, nullptr, cap, count, where, kind, h->m.options[4]);
}
int bnx_map_shard_config(bnx_map_t* h, int rank, int world) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_config(rank, world);
}
int bnx_map_shard_begin(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, int is_f64, uint32_t index_base,
                        const double origin[3], double max_range, void* send_records, int64_t cap_records, int where) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_begin(points, stride_bytes, n, is_f64 != 0, index_base, origin, max_range, send_records, cap_records, where);
}
int bnx_map_shard_resolve_mark(bnx_map_t* h, const void* recv_records, void* send_leaves, int64_t cap_leaves) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_resolve_mark(recv_records, send_leaves, cap_leaves);
}
int bnx_map_shard_merge(bnx_map_t* h, const void* recv_leaves, void* flags) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_merge(recv_leaves, flags);
}
int bnx_map_shard_finish(bnx_map_t* h, const void* flags_reduced, int* retry) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_finish(flags_reduced, retry);
}
int bnx_map_publish_occupied_f32(bnx_map_t* h, double z_min, double z_max, float* points, int64_t stride_floats, int64_t cap, int64_t* count,
                                 int where) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  BNX_TRY(h->m.drain());
  return h->m.grid.dump_points_f32(points, stride_floats, 1, z_min, z_max, cap, count, where, h->m.options[4]);
}
int bnx_nccl_unique_id(const char* nccl_library_path, void* out128) {
  BNX_REQUIRE(out128 != nullptr, "null output");
  return Map::nccl_unique_id(nccl_library_path, out128);
}
int bnx_map_shard_comm_init(bnx_map_t* h, const char* nccl_library_path, const void* unique_id128, int rank, int world) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_comm_init(nccl_library_path, unique_id128, rank, world);
}
int bnx_map_shard_insert(bnx_map_t* h, const void* points, int64_t stride_bytes, int64_t n, int is_f64, uint32_t index_base, int64_t n_max,
                         const double origin[3], double max_range, int where, int async) {
  BNX_HANDLE(h);
  BNX_REQUIRE(origin != nullptr, "null origin");
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_insert(points, stride_bytes, n, is_f64 != 0, index_base, n_max, origin, max_range, where, async != 0);
}
int bnx_map_shard_p2p_alloc(bnx_map_t* h, int64_t cap_records, int64_t cap_leaves, void* ipc_handle64, void** device_ptr) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.p2p_alloc(cap_records, cap_leaves, ipc_handle64, device_ptr);
}
int bnx_map_shard_p2p_attach(bnx_map_t* h, const void* ipc_handles, void* const* device_ptrs) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.p2p_attach(ipc_handles, device_ptrs);
}
int bnx_map_shard_host_init(bnx_map_t* h, int rank, int world, bnx_allgather_fn allgather, void* ctx) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.shard_host_init(rank, world, reinterpret_cast<Map::AllGatherFn>(allgather), ctx);
}
int bnx_map_shard_stats(bnx_map_t* h, int64_t out[8]) {
  BNX_HANDLE(h);
  BNX_REQUIRE(out != nullptr, "null output");
  std::memcpy(out, h->m.shard_stats, sizeof(int64_t) * 8);
  return BNX_OK;
}
int bnx_map_shard_set_fleet(bnx_map_t* h, const double* origins) {
  BNX_HANDLE(h);
  return h->m.set_fleet(origins);
}
int bnx_map_shard_exchange(const bnx_map_t* h, int* kind) {
  BNX_HANDLE(h);
  BNX_REQUIRE(kind != nullptr, "null output");
  *kind = h->m.exchange_kind();
  return BNX_OK;
}
int bnx_map_counters(bnx_map_t* h, int64_t out[8]) {
  BNX_HANDLE(h);
  BNX_TRY(h->m.drain());
  std::memcpy(out, h->m.counters, sizeof(int64_t) * 8);
  return BNX_OK;
}
int bnx_map_update_count(const bnx_map_t* h, int* value) {
  BNX_HANDLE(h);
  BNX_REQUIRE(value != nullptr, "null output");
  *value = (int)h->m.update_count;
  return BNX_OK;
}
int bnx_map_set_marking(bnx_map_t* h, int mode) {
  BNX_HANDLE(h);
  DeviceGuard dg(h->m.grid.device);
  return h->m.set_marking(mode);
}
int bnx_map_set_profiling(bnx_map_t* h, int enable) {
  BNX_HANDLE(h);
  h->m.profiling = enable != 0;
  return BNX_OK;
}
int bnx_map_phase_times(bnx_map_t* h, double out_us[8]) {
  BNX_HANDLE(h);
  std::memcpy(out_us, h->m.phase_us, sizeof(double) * 8);
  return BNX_OK;
}

}  // extern "C"
