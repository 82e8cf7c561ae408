/*
 * bonxai_b200 — C ABI of the B200-native (sm_100a) implementation of Bonxai's occupancy-mapping hot path.
 *
 * The reference (facontidavide/Bonxai) has no FFI boundary: callers include its headers and call C++
 * members directly. This header IS the boundary this build creates: each entry point
 * names the reference member it replaces (paths relative to the reference tree). The drop-in C++
 * headers include/bonxai/bonxai.hpp and include/bonxai_map/probabilistic_map.hpp are thin inline
 * wrappers over these functions; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types. Coordinates are int32 xyz triplets
 *     (Bonxai::CoordT, bonxai_core/include/bonxai/grid_coord.hpp:62-66), densely packed.
 *   - every function returns a bnx_status (0 = ok); bnx_last_error() gives the thread-local message.
 *   - `where` says whether the caller's buffers are host (BNX_HOST) or device (BNX_DEVICE) memory.
 *     Host calls are synchronous on return. Device calls are enqueued on the handle's stream
 *     (bnx_*_set_stream) and complete in stream order; scalar outputs (counts) make the call wait.
 *   - one host thread per handle at a time (the reference's single-writer contract, README.md:128-133).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns BNX_ERR_CUDA.
 */
#ifndef BONXAI_B200_H
#define BONXAI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define BNX_API
#else
#define BNX_API __attribute__((visibility("default")))
#endif

typedef struct bnx_grid bnx_grid_t; /* Bonxai::VoxelGrid<DataT>, DataT = opaque cell_bytes blob   */
typedef struct bnx_map bnx_map_t;   /* Bonxai::ProbabilisticMap                                   */

typedef enum {
  BNX_OK = 0,
  BNX_ERR_INVALID = 1,   /* bad argument (reference: std::runtime_error, e.g. bonxai.hpp:399-401)   */
  BNX_ERR_CUDA = 2,      /* CUDA runtime/driver failure or no device                                */
  BNX_ERR_NOMEM = 3,     /* device memory exhausted while growing node pools                        */
  BNX_ERR_CAPACITY = 4,  /* caller's output buffer too small (count is still reported)              */
  BNX_ERR_UNSUPPORTED = 5
} bnx_status;

enum { BNX_HOST = 0, BNX_DEVICE = 1 };
/* Bonxai::ClearOption, bonxai.hpp:107-112 */
enum { BNX_CLEAR_MEMORY = 0, BNX_SET_ALL_CELLS_OFF = 1 };
/* cell predicates of ProbabilisticMap (probabilistic_map.cpp:56-75,108-126) */
enum { BNX_OCCUPIED = 0, BNX_UNKNOWN = 1, BNX_FREE = 2 };

BNX_API int bnx_version(void);
BNX_API const char* bnx_last_error(void);
/* number of CUDA kernels this library has launched in this process (instrumentation for bench.py) */
BNX_API int64_t bnx_launch_count(void);
/* number of usable CUDA devices; 0 with BNX_OK when the driver is present but no device is */
BNX_API int bnx_device_count(int* count);
/* pinned host memory for zero-staging H2D/D2H of point clouds and dumps */
BNX_API int bnx_host_alloc(void** ptr, size_t bytes);
BNX_API int bnx_host_free(void* ptr);

/* ------------------------------------------------------------------------------------------------
 * VoxelGrid<DataT>                                     bonxai_core/include/bonxai/bonxai.hpp:114-333
 * ---------------------------------------------------------------------------------------------- */
/* VoxelGrid(voxel_size, inner_bits=2, leaf_bits=3), bonxai.hpp:138,389-402. BNX_ERR_INVALID when
 * a bit count is < 1 (the reference throws). cell_bytes = sizeof(DataT): 1, 2, 4, 8 or 16.
 * The grid lives on the calling thread's current CUDA device. */
BNX_API int bnx_grid_create(double voxel_size, int inner_bits, int leaf_bits, int cell_bytes, bnx_grid_t** out);
BNX_API int bnx_grid_destroy(bnx_grid_t* g);
/* cudaStream_t to enqueue on. Any stream handle is taken as is — including 0, the legacy default stream
 * (what torch.cuda.current_stream().cuda_stream returns by default); BNX_OWN_STREAM goes back to the handle's
 * private non-blocking stream, which is what a new grid/map uses. */
#define BNX_OWN_STREAM ((void*)~(uintptr_t)0)
BNX_API int bnx_grid_set_stream(bnx_grid_t* g, void* cuda_stream);
BNX_API int bnx_grid_sync(bnx_grid_t* g);
/* innetBits()/leafBits()/voxelSize(), bonxai.hpp:146-154 */
BNX_API int bnx_grid_info(const bnx_grid_t* g, double* voxel_size, int* inner_bits, int* leaf_bits, int* cell_bytes);

/* posToCoord, bonxai.hpp:404-410 (fp64 multiply by 1/voxel_size, floor, cast) — xyz[n][3] doubles */
BNX_API int bnx_grid_pos_to_coord(const bnx_grid_t* g, const double* xyz, int64_t n, int32_t* out, int where);
/* coordToPos, bonxai.hpp:412-417 */
BNX_API int bnx_grid_coord_to_pos(const bnx_grid_t* g, const int32_t* xyz, int64_t n, double* out, int where);

/* Accessor::setValue over a batch, bonxai.hpp:449-466. Sequential semantics are kept: when a
 * coordinate repeats inside the batch the LAST value wins and was_on is false only for the first
 * occurrence of a previously-off cell. was_on (n bytes) may be NULL. */
BNX_API int bnx_grid_set_values(bnx_grid_t* g, const int32_t* xyz, const void* values, int64_t n,
                                uint8_t* was_on, int where);
/* ConstAccessor::value, bonxai.hpp:496-516: found[i] = 0 (and values[i] untouched) for nullptr */
BNX_API int bnx_grid_get_values(bnx_grid_t* g, const int32_t* xyz, int64_t n, void* values, uint8_t* found,
                                int where);
/* Accessor::value(coord, create_if_missing=true), bonxai.hpp:469-494: missing cells are created
 * ON with DataT{} (all-zero bytes); returns the cell values */
BNX_API int bnx_grid_get_or_create(bnx_grid_t* g, const int32_t* xyz, int64_t n, void* values, int where);
/* write through the pointer value() returned: stores values[i] only where the cell is ON */
BNX_API int bnx_grid_update_values(bnx_grid_t* g, const int32_t* xyz, const void* values, int64_t n, int where);
/* Accessor::setCellOn(coord, default_value), bonxai.hpp:537-554 */
BNX_API int bnx_grid_set_on(bnx_grid_t* g, const int32_t* xyz, int64_t n, const void* default_value,
                            uint8_t* was_on, int where);
/* Accessor::setCellOff, bonxai.hpp:557-569 (the value is kept) */
BNX_API int bnx_grid_set_off(bnx_grid_t* g, const int32_t* xyz, int64_t n, uint8_t* was_on, int where);
/* ConstAccessor::isCellOn, bonxai.hpp:518-534 */
BNX_API int bnx_grid_is_on(bnx_grid_t* g, const int32_t* xyz, int64_t n, uint8_t* out, int where);

/* activeCellsCount, bonxai.hpp:689-701 */
BNX_API int bnx_grid_active_count(bnx_grid_t* g, int64_t* count);
/* Order-independent digest of what forEachCell (bonxai.hpp:704-743) would visit: out = {sum, xor, count} over all ON
 * cells of mix64(hash3(x, y, z) + FNV1a64(value bytes) * 0x9E3779B97F4A7C15) (arithmetic mod 2^64; the functions are
 * spelled out in tests/digest.py). Two grids hold the same (coord, value) set iff their digests agree (up to a 2^-64
 * collision); the digests of disjoint shards combine by + / ^ / +. Lets full-size maps be compared with the oracle or
 * across GPUs without copying them to the host. */
BNX_API int bnx_grid_digest(bnx_grid_t* g, uint64_t out[3]);
/* forEachCell, bonxai.hpp:704-743: all ON cells as (coord, value) pairs in unspecified order (the
 * reference's order is unordered_map order). *count always receives the number of ON cells; when
 * cap < *count nothing is written and BNX_ERR_CAPACITY is returned. xyz/values may be NULL to count. */
BNX_API int bnx_grid_dump(bnx_grid_t* g, int32_t* xyz, void* values, int64_t cap, int64_t* count, int where);
/* clear(ClearOption), bonxai.hpp:678-687 */
BNX_API int bnx_grid_clear(bnx_grid_t* g, int clear_option);
/* releaseUnusedMemory, bonxai.hpp:367-387: leaves with every cell OFF go back to the pool, roots
 * whose leaves are all gone leave the table */
BNX_API int bnx_grid_release_unused(bnx_grid_t* g);
/* memUsage, bonxai.hpp:649-676: bytes of device memory in use by live nodes (layout specific,
 * like the reference's figure) */
BNX_API int bnx_grid_mem_usage(bnx_grid_t* g, int64_t* bytes);
/* node statistics: {roots, inner nodes, leaves in use, leaves free-listed, root table slots,
 * leaf pool capacity, mapped bytes, reserved} */
BNX_API int bnx_grid_stats(bnx_grid_t* g, int64_t out[8]);
/* Serialize / Deserialize, bonxai_core/include/bonxai/serialization.hpp:77-116,153-199.
 * type_name is the demangled DataT the header line carries (e.g. "unsigned int", "float").
 * serialize: *size receives the byte count; buffer may be NULL to query (host memory only). */
BNX_API int bnx_grid_serialize(bnx_grid_t* g, const char* type_name, uint8_t* buffer, int64_t cap, int64_t* size);
BNX_API int bnx_grid_deserialize(const uint8_t* data, int64_t len, int cell_bytes, const char* expect_type_name,
                                 bnx_grid_t** out);

/* ------------------------------------------------------------------------------------------------
 * ProbabilisticMap                  bonxai_map/include/bonxai_map/probabilistic_map.hpp:27-139
 * ---------------------------------------------------------------------------------------------- */
/* ProbabilisticMap(resolution), probabilistic_map.cpp:14-16: VoxelGrid<CellT> with default bits 2/3.
 * The cell word is the reference's CellT image: (probability_log << 4) | update_id (hpp:44-53). */
BNX_API int bnx_map_create(double resolution, bnx_map_t** out);
BNX_API int bnx_map_destroy(bnx_map_t* m);
BNX_API int bnx_map_set_stream(bnx_map_t* m, void* cuda_stream);
BNX_API int bnx_map_sync(bnx_map_t* m);
/* grid(), probabilistic_map.cpp:10-12,18-20: borrowed handle, owned by the map */
BNX_API int bnx_map_grid(bnx_map_t* m, bnx_grid_t** grid);
/* Options {prob_miss_log, prob_hit_log, clamp_min_log, clamp_max_log, occupancy_threshold_log},
 * probabilistic_map.hpp:56-64; setOptions/options, probabilistic_map.cpp:22-28 */
BNX_API int bnx_map_set_options(bnx_map_t* m, const int32_t options[5]);
BNX_API int bnx_map_get_options(const bnx_map_t* m, int32_t options[5]);

/* insertPointCloud<PointT>, probabilistic_map.hpp:141-160 (+ addHit/MissPoint, updateFreeCells,
 * RayIterator: probabilistic_map.cpp:30-54,77-106, hpp:162-203).
 *   f32: PointT = {float x,y,z[,...]} every stride_bytes (12 = packed xyz, 16 = pcl::PointXYZ, any
 *        multiple of 4 >= 12); f64: PointT = 3 doubles every stride_bytes (>= 24, multiple of 8).
 *   origin is a PointT of the same scalar type, max_range as in the reference (may be +inf).
 * Bit-exact: after the call a forEachCell dump equals the reference's after the same call. */
BNX_API int bnx_map_insert_f32(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n,
                               const float origin[3], double max_range, int where);
BNX_API int bnx_map_insert_f64(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n,
                               const double origin[3], double max_range, int where);
/* Pipelined insertPointCloud: same result as bnx_map_insert_*, but the call only ENQUEUES the scan and returns; no
 * host synchronisation per scan. The scan is split in two: its FRONT half (host-to-device copy of a host buffer,
 * classification, endpoint dedupe — nothing that touches the map) runs on an internal stream, up to 32 scans ahead of
 * the map updates; its BACK half (resolve, mark, apply) runs on the map's stream, one scan after the other. So the
 * copy and the classification of later scans overlap the map updates of earlier ones.
 *   - the input buffer (device memory, or host memory — pinned for an asynchronous copy) must stay valid and
 *     unchanged until bnx_map_sync() or any other call on the map returns; those complete the queue first;
 *   - a BNX_DEVICE buffer is read on the internal stream: its contents must be COMPLETE when the call is made (the
 *     synchronous bnx_map_insert_* reads on the map's stream instead, ordered after earlier work on that stream).
 * If a queued scan runs out of pool space the device freezes the pipeline at that scan (later scans skip
 * themselves), and the next synchronising call grows the pools and replays from there: the map is always exactly
 * what the synchronous calls would have produced. */
BNX_API int bnx_map_insert_async_f32(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n,
                                     const float origin[3], double max_range, int where);
BNX_API int bnx_map_insert_async_f64(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n,
                                     const double origin[3], double max_range, int where);
/* insertPointCloud with the pre-step of the reference's ROS caller fused into the classify kernel
 * (bonxai_ros/src/bonxai_server.cpp:148-171): points with a non-finite coordinate leave the cloud, the others are
 * transformed by the row-major 4x4 float matrix sensor_to_world exactly like pcl::transformPointCloud's SSE kernel
 * (x*c0 + (y*c1 + (z*c2 + c3)), one rounding per operation), then inserted. `points` are in the SENSOR frame,
 * origin (= the matrix translation in the ROS node) in the world frame. async != 0: pipelined like
 * bnx_map_insert_async_f32. Counter N of the scan excludes the dropped points. */
BNX_API int bnx_map_insert_transformed_f32(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n,
                                           const float sensor_to_world[16], const float origin[3], double max_range,
                                           int where, int async);
/* cumulative {N, E, V, U} over every scan inserted so far (completes the queue) */
BNX_API int bnx_map_totals(bnx_map_t* m, int64_t out[4]);
/* addHitPoint / addMissPoint, probabilistic_map.cpp:30-54: the endpoint cell is updated now, its
 * ray is cast by the next insertPointCloud from that call's origin (updateFreeCells is private in
 * the reference, hpp:136) */
BNX_API int bnx_map_add_hit(bnx_map_t* m, const double point[3]);
BNX_API int bnx_map_add_miss(bnx_map_t* m, const double point[3]);
/* isOccupied / isUnknown / isFree, probabilistic_map.cpp:56-75 — kind = BNX_OCCUPIED|UNKNOWN|FREE */
BNX_API int bnx_map_query(bnx_map_t* m, const int32_t* xyz, int64_t n, int kind, uint8_t* out, int where);
/* getOccupiedVoxels / getFreeVoxels(std::vector<CoordT>&), probabilistic_map.cpp:108-126.
 * Same cap/count protocol as bnx_grid_dump. */
BNX_API int bnx_map_get_voxels(bnx_map_t* m, int kind, int32_t* xyz, int64_t cap, int64_t* count, int where);
/* getOccupiedVoxels<PointT>, probabilistic_map.hpp:115-124: coord * resolution (voxel corner), doubles */
BNX_API int bnx_map_get_voxel_points(bnx_map_t* m, int kind, double* xyz, int64_t cap, int64_t* count, int where);
/* The publisher post-step of the reference's ROS node fused into the compaction (bonxai_ros/src/bonxai_server.cpp:
 * 217-251): every occupied voxel as the point coord*resolution (fp64 product, voxel corner), kept when
 * z_min <= z <= z_max, written as float x,y,z every stride_floats floats (3 = packed, 4 = pcl::PointXYZ with
 * w = 1.0f). Same cap/count protocol as bnx_grid_dump; order unspecified. */
BNX_API int bnx_map_publish_occupied_f32(bnx_map_t* m, double z_min, double z_max, float* points, int64_t stride_floats,
                                         int64_t cap, int64_t* count, int where);
/* counters of the last insert: {N points, E endpoint voxels updated (= rays cast), V = sum of ray
 * cells + N, U cells whose word changed, leaves touched, retries after pool growth, 0, 0} */
BNX_API int bnx_map_counters(bnx_map_t* m, int64_t out[8]);
/* _update_count, probabilistic_map.hpp:129 (cycles 1,2,3) */
BNX_API int bnx_map_update_count(const bnx_map_t* m, int* value);
/* Where insertPointCloud keeps the per-scan marks of updateFreeCells (probabilistic_map.cpp:77-106: which cells some ray
 * crossed, which are hit endpoints): mode 1 = inside the leaves (serves every scan); mode 2 = EXPERIMENTAL: a dense
 * window of 8^3 blocks around the origin, addressed by arithmetic, whenever max_range is finite and the window fits
 * (BNX_DENSE_MAX_MB, default 2048 MB; needs the default inner/leaf bits) — bit-identical results, but measured slower on
 * the B200 (profiles/r2_notes.md); mode 0 (default) = mode 1 unless the environment says BNX_DENSE=1. The call completes
 * queued scans first. */
BNX_API int bnx_map_set_marking(bnx_map_t* m, int mode);
/* device time of the phases of the last insert in microseconds (CUDA events; enabled by
 * bnx_map_set_profiling(m,1)): {h2d, classify, resolve, mark, apply, total, 0, 0} */
BNX_API int bnx_map_set_profiling(bnx_map_t* m, int enable);
BNX_API int bnx_map_phase_times(bnx_map_t* m, double out_us[8]);

/* ------------------------------------------------------------------------------------------------
 * One map sharded over several GPUs (one process per GPU): roots are owned by hash(root key) mod world,
 * every rank holds 1/world of a scan's points. insertPointCloud becomes four stages with two exchanges the
 * CALLER performs on the staged device buffers (NCCL all-to-all, see bonxai_b200/sharded.py and DESIGN.md §7):
 *
 *   shard_begin         classify + local dedupe; endpoint records bucketed by owner -> send_records
 *        exchange 1:    all-to-all of [world][cap_records] x 16-B records (slot 0 of a block = count)
 *   shard_resolve_mark  owner: lowest global index per voxel, stale test, rays; ray cells of foreign roots are
 *                       staged as {leaf origin, 512-bit mask} records -> send_leaves
 *        exchange 2:    all-to-all of [world][cap_leaves] x 80-B records
 *   shard_merge         owner: OR the received masks into its leaves; writes this rank's error flags
 *        all-reduce(MAX) of the 4 x u32 flags
 *   shard_finish        apply (skipped everywhere if any rank ran short); *retry != 0 -> repeat from
 *                       shard_resolve_mark with the same recv_records (bits 8.. say which exchange buffer to grow)
 *
 * All buffers are device memory on the map's stream. The union of the shards equals the unsharded map bit for
 * bit. max_range may be +inf and voxel coordinates arbitrary (the dedupe tables then hold full keys instead of the
 * packed 63-bit ones); addHitPoint/addMissPoint queues are not supported on a sharded map.
 * ---------------------------------------------------------------------------------------------- */
BNX_API int bnx_map_shard_config(bnx_map_t* m, int rank, int world);
BNX_API int bnx_map_shard_begin(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n, int is_f64,
                                uint32_t index_base, const double origin[3], double max_range, void* send_records,
                                int64_t cap_records, int where);
BNX_API int bnx_map_shard_resolve_mark(bnx_map_t* m, const void* recv_records, void* send_leaves, int64_t cap_leaves);
BNX_API int bnx_map_shard_merge(bnx_map_t* m, const void* recv_leaves, void* flags);
BNX_API int bnx_map_shard_finish(bnx_map_t* m, const void* flags_reduced, int* retry);

/* The same protocol driven by the library itself: the two all-to-alls (grouped ncclSend/ncclRecv) and the flag
 * all-reduce are issued on the map's stream through NCCL, which is resolved with dlopen at run time
 * (nccl_library_path may be NULL: "libnccl.so.2"; a copy already loaded by the process, e.g. torch's, is shared).
 *   rank 0: bnx_nccl_unique_id -> 128 bytes, broadcast them by any means -> every rank: bnx_map_shard_comm_init.
 *   every rank, in lock step: bnx_map_shard_insert(points of its slice, index_base = global index of its first
 *   point, n_max = the largest slice over all ranks, ...). async != 0: pipelined, nothing synchronises; a scan that
 *   runs short on ANY rank freezes ALL ranks at that scan (the flags are all-reduced) and the next bnx_map_sync /
 *   query on every rank grows and replays it — so synchronising calls must be made by all ranks together. */
BNX_API int bnx_nccl_unique_id(const char* nccl_library_path, void* out128);
BNX_API int bnx_map_shard_comm_init(bnx_map_t* m, const char* nccl_library_path, const void* unique_id128, int rank, int world);
BNX_API int bnx_map_shard_insert(bnx_map_t* m, const void* points, int64_t stride_bytes, int64_t n, int is_f64,
                                 uint32_t index_base, int64_t n_max, const double origin[3], double max_range, int where,
                                 int async);


/* Peer-memory exchange (NVLink P2P stores instead of collectives). Every rank owns a MAILBOX — one device allocation
 * [arrival flags | endpoint-record inbox [world][cap_records] | leaf-record inbox [world][cap_leaves]] — that all
 * peers map (CUDA IPC between processes, the raw pointer inside one process). The producing kernels of
 * shard_begin / shard_resolve_mark / shard_merge store their records straight into block [rank] of the OWNER's
 * inbox and their last thread block stamps an arrival flag there; the consuming kernels (and the apply pass, for the
 * error flags) spin on the flags of their own mailbox. A scan then needs no collective launch and no staging copy.
 *   bnx_map_shard_p2p_alloc   (re)creates the mailbox of this rank; returns its 64-byte cudaIpcMemHandle_t and/or
 *                             its device pointer (either may be NULL). world <= 16.
 *   bnx_map_shard_p2p_attach  ipc_handles: [world][64] bytes gathered from all ranks, or device_ptrs: [world] raw
 *                             pointers when all shards live in this process. Afterwards the staged calls take NULL
 *                             for their buffer arguments (send_records, recv_records, send_leaves, recv_leaves,
 *                             flags, flags_reduced) and use the mailboxes.
 * bnx_map_shard_comm_init + bnx_map_shard_insert do all of this themselves (handles all-gathered through NCCL,
 * mailboxes re-created collectively when an exchange overflowed) unless BNX_SHARD_EXCHANGE=nccl is set; NCCL then
 * only bootstraps. bnx_map_shard_exchange: 0 caller-run, 1 NCCL collectives, 2 peer memory. A peer that never
 * arrives turns into an error after 20 s (BNX_PEER_TIMEOUT_MS), not a hang. */
BNX_API int bnx_map_shard_p2p_alloc(bnx_map_t* m, int64_t cap_records, int64_t cap_leaves, void* ipc_handle64, void** device_ptr);
BNX_API int bnx_map_shard_p2p_attach(bnx_map_t* m, const void* ipc_handles, void* const* device_ptrs);
BNX_API int bnx_map_shard_exchange(const bnx_map_t* m, int* kind);
/* Fleet step: several sensors feeding ONE sharded map. After bnx_map_shard_set_fleet(m, origins) — origins =
 * [world][3] doubles, the same on every rank — the NEXT bnx_map_shard_insert / bnx_map_shard_begin of every rank takes
 * the WHOLE scan of sensor `rank` (index_base = rank * n_max, its origin argument is ignored in favour of origins[rank])
 * and the step gives exactly the map that world consecutive insertPointCloud calls — sensor 0, 1, ... world-1, each
 * with its own update id (probabilistic_map.cpp:103-105) — would give, with all scans processed concurrently: endpoint
 * records go to the owners as usual, rays are cast from their own sensor's origin, and every touched leaf carries the
 * sensor whose update id stamps it. Exactness needs sensors whose reach does not overlap (pairwise distance > 2 *
 * max_range + 40 voxels: no cell is then visited by two scans, so their order cannot matter); otherwise the insert
 * returns BNX_ERR_UNSUPPORTED and the scans have to be inserted one after the other. One call arms one step. */
BNX_API int bnx_map_shard_set_fleet(bnx_map_t* m, const double* origins);
/* Bootstrap without NCCL (instead of bnx_nccl_unique_id + bnx_map_shard_comm_init): the caller supplies the collective
 * that hands the 64-byte mailbox handles around — allgather(ctx, send, recv, bytes) must gather `bytes` bytes of HOST
 * memory from every rank into recv[world][bytes] and return 0 (any transport: gloo, MPI, a socket). Peer memory is then
 * the only exchange. Ranks may share one GPU (CUDA IPC works between processes on the same device), which is how the
 * multi-process protocol is tested on a single-GPU box. */
typedef int (*bnx_allgather_fn)(void* ctx, const void* send, void* recv, int64_t bytes_per_rank);
BNX_API int bnx_map_shard_host_init(bnx_map_t* m, int rank, int world, bnx_allgather_fn allgather, void* ctx);
/* What the sharded pipeline of this rank really did so far: {resolve_mark attempts, frozen-pipeline replays, mailbox
 * (re)creations, collective drains, synchronous retries, leaf-inbox capacity (records per sender), fullest leaf-inbox
 * block seen (max over ranks), scans completed}. A healthy run has attempts == scans, replays == retries == 0 and ONE
 * mailbox creation. */
BNX_API int bnx_map_shard_stats(bnx_map_t* m, int64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* BONXAI_B200_H */
