// Host-side owner of a device-resident sparse voxel grid (the VoxelGrid<DataT> of bonxai.hpp:114-333).
#pragma once

#include <vector>

#include "arena.hpp"
#include "grid_device.cuh"

namespace bnx {

// grow-only device scratch buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  int reserve(size_t need) {
    if (need <= bytes) return BNX_OK;
    release();
    size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      bytes = 0;
      set_error(std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
      return BNX_ERR_NOMEM;
    }
    bytes = want;
    return BNX_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

class Grid {
 public:
  Grid() = default;
  ~Grid();
  Grid(const Grid&) = delete;
  Grid& operator=(const Grid&) = delete;

  int init(double voxel_size, int inner_bits, int leaf_bits, int cell_bytes);

  // ---- batched accessor operations (xyz / values / flags: host or device pointers per `where`)
  int set_values(const i32* xyz, const void* values, i64 n, u8* was_on, int where);
  int get_values(const i32* xyz, i64 n, void* values, u8* found, int where);
  int get_or_create(const i32* xyz, i64 n, void* values, int where);
  int update_values(const i32* xyz, const void* values, i64 n, int where);
  int set_on(const i32* xyz, i64 n, const void* default_value, u8* was_on, int where);
  int set_off(const i32* xyz, i64 n, u8* was_on, int where);
  int is_on(const i32* xyz, i64 n, u8* out, int where);
  int pos_to_coord(const double* xyz, i64 n, i32* out, int where) const;
  int coord_to_pos(const i32* xyz, i64 n, double* out, int where) const;

  // ---- whole-grid operations
  int active_count(i64* count);
  // order-independent digest {sum, xor, count} of all (coord, value) pairs (see k_digest)
  int digest(u64 out[3]);
  // pred: -1 all ON cells; BNX_OCCUPIED/BNX_FREE: CellT probability_log >/< thr (map grids only).
  // xyz_out (int32 triplets) or pos_out (double triplets, coord*resolution); values optional.
  int dump(i32* xyz, double* pos, void* values, i64 cap, i64* count, int where, int pred, i32 thr);
  // occupied voxels of a map grid as float points coord*resolution with an optional z window (publisher post-step)
  int dump_points_f32(float* out, i64 stride_floats, int zfilter, double zmin, double zmax, i64 cap, i64* count, int where, i32 thr);
  int clear(int option);
  int release_unused();
  int mem_usage(i64* bytes);
  int stats(i64 out[8]);
  int serialize(const char* type_name, u8* buffer, i64 cap, i64* size);
  static int deserialize(const u8* data, i64 len, int cell_bytes, const char* expect_type, Grid** out);

  // ---- plumbing shared with the map
  GridDev dev() const { return dev_; }
  cudaStream_t stream() const { return stream_; }
  // any stream handle, including 0 (the legacy default stream); BNX_OWN_STREAM selects the grid's private stream
  void set_stream(cudaStream_t s) { stream_ = (s == reinterpret_cast<cudaStream_t>(~(uintptr_t)0)) ? own_stream_ : s; }
  int sync();
  // read the device counters (synchronises the stream)
  int read_counters(GridCounters* out);
  // after a kernel reported ERR_* bits: clamp counters, grow the failing pools, clear the bits
  int recover(const GridCounters& seen);
  // proactive growth at quiet points (root table load factor, pool head-room)
  // extra_leaves / extra_inner: head-room the caller knows it needs before the next quiet point (e.g. what the last
  // window of a pipelined run allocated), on top of half a growth step
  int maintain(const GridCounters& seen, u64 extra_leaves = 0, u64 extra_inner = 0);
  // zero_stream: where the new memory is zero-filled (default: the grid's stream). Mapping more memory behind the pools
  // does not wait for running kernels (VMM), so a caller may grow AHEAD of need on a side stream while scans run and
  // make its work stream wait for the fill.
  int ensure_leaf_capacity(u64 leaves, cudaStream_t zero_stream = nullptr);
  int ensure_inner_capacity(u64 inner, cudaStream_t zero_stream = nullptr);
  // pools grow in fixed steps — max(BNX_GROW_MB (256), live / 8) — not by doubling: mapped memory stays within ~1.15x of
  // the live nodes of a large map
  u64 leaf_step(u64 live_leaves) const;
  u64 inner_step(u64 live_inner) const;
  int grow_root_table(u64 min_slots);

  double resolution = 0.0, inv_resolution = 0.0;
  int inner_bits = 2, leaf_bits = 3, cell_bytes = 4;
  int device = 0;

 private:
  template <bool CREATE>
  int locate(const i32* d_xyz, i64 n, u32* d_loc);
  int stage_in(const void* src, size_t bytes, int where, DevBuf& buf, const void** dptr);
  int dedupe(const i32* d_xyz, const u32* d_loc, i64 n);
  int free_list_reserve();

  GridDev dev_ = {};
  Arena leaf_arena_, inner_arena_;
  int4* root_ = nullptr;
  u64 root_slots_ = 0;
  GridCounters* d_ctr_ = nullptr;
  GridCounters* h_ctr_ = nullptr;  // pinned
  u64* d_count_ = nullptr;         // generic 64-bit device counter(s)
  u64* h_count_ = nullptr;         // pinned mirror
  u32* free_list_ = nullptr;
  u64 free_list_cap_ = 0;
  cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
  // scratch of the batched operations
  DevBuf b_xyz_, b_val_, b_flag_, b_loc_, b_slot_, b_keys_, b_first_, b_last_, b_out_;
};

}  // namespace bnx
