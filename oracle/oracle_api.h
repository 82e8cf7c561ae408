/* TEST INFRASTRUCTURE ONLY — not part of the product.
 *
 * One C API, two CPU implementations behind it:
 *   oracle/bonxai_oracle.c      -> oracle/libbonxai_oracle.so    ("port": plain-C restatement)
 *   oracle/ref_wrap.cpp         -> oracle/_ref/libbonxai_ref.so  ("reference": the UNMODIFIED sources
 *                                  under /root/reference compiled behind this API)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * either library. The product (bonxai_b200/, include/) never does.
 *
 * Cells are opaque 4-byte words (`VoxelGrid<uint32_t>`); for the occupancy map the word is the
 * reference's CellT bitfield image: (probability_log << 4) | (update_id & 0xF)
 * (bonxai_map/include/bonxai_map/probabilistic_map.hpp:44-53).
 */
#ifndef BONXAI_ORACLE_API_H
#define BONXAI_ORACLE_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* "reference" or "port" */
const char* orc_kind(void);

/* ---- scalar helpers ------------------------------------------------------------------------ */
/* probabilistic_map.hpp:34-36 / :39-42 */
int32_t orc_logods(float prob);
float orc_prob(int32_t logods_fixed);
/* bonxai.hpp:404-410 — xyz[n][3] doubles -> out[n][3] int32 */
void orc_pos_to_coord(double resolution, const double* xyz, int64_t n, int32_t* out);
/* bonxai.hpp:412-417 */
void orc_coord_to_pos(double resolution, const int32_t* xyz, int64_t n, double* out);
/* probabilistic_map.hpp:13-19,162-203 — returns the number of ray cells; writes min(count, cap) */
int64_t orc_compute_ray(const int32_t a[3], const int32_t b[3], int32_t* out_xyz, int64_t cap);

/* ---- VoxelGrid<uint32_t> ------------------------------------------------------------------- */
/* bonxai.hpp:138,389-402; returns NULL when the reference constructor throws (bits < 1) */
void* orc_grid_create(double voxel_size, int inner_bits, int leaf_bits);
void orc_grid_destroy(void* g);
/* sequential Accessor::setValue over the batch (bonxai.hpp:449-466); was_on may be NULL */
void orc_grid_set_values(void* g, const int32_t* xyz, const uint32_t* vals, int64_t n, uint8_t* was_on);
/* ConstAccessor::value (bonxai.hpp:496-516): found[i]=0 and out[i] untouched when missing */
void orc_grid_get_values(void* g, const int32_t* xyz, int64_t n, uint32_t* out, uint8_t* found);
/* Accessor::value(coord, true) (bonxai.hpp:469-494): creates with DataT{}=0, returns the value */
void orc_grid_get_or_create(void* g, const int32_t* xyz, int64_t n, uint32_t* out);
/* Accessor::setCellOn(coord, default) (bonxai.hpp:537-554) */
void orc_grid_set_on(void* g, const int32_t* xyz, int64_t n, uint32_t default_value, uint8_t* was_on);
/* Accessor::setCellOff (bonxai.hpp:557-569) */
void orc_grid_set_off(void* g, const int32_t* xyz, int64_t n, uint8_t* was_on);
/* ConstAccessor::isCellOn (bonxai.hpp:518-534) */
void orc_grid_is_on(void* g, const int32_t* xyz, int64_t n, uint8_t* out);
/* activeCellsCount (bonxai.hpp:689-701) */
int64_t orc_grid_active_count(void* g);
/* forEachCell (bonxai.hpp:704-743): UNORDERED (coord, value) pairs; returns the total count and
 * writes min(count, cap) entries */
int64_t orc_grid_dump(void* g, int32_t* xyz, uint32_t* vals, int64_t cap);
/* clear (bonxai.hpp:678-687): opt 0 = CLEAR_MEMORY, 1 = SET_ALL_CELLS_OFF */
void orc_grid_clear(void* g, int opt);
/* releaseUnusedMemory (bonxai.hpp:367-387) */
void orc_grid_release_unused(void* g);
/* Serialize (serialization.hpp:77-116) of VoxelGrid<uint32_t>; returns bytes needed, writes <= cap.
 * Root/inner iteration order is implementation-defined (unordered_map), compare after parsing. */
int64_t orc_grid_serialize(void* g, uint8_t* out, int64_t cap);
/* Deserialize (serialization.hpp:153-199); returns a new grid or NULL on a bad header */
void* orc_grid_deserialize(const uint8_t* data, int64_t len);

/* ---- ProbabilisticMap ----------------------------------------------------------------------- */
void* orc_map_create(double resolution); /* probabilistic_map.cpp:14-16 (default bits 2/3) */
void orc_map_destroy(void* m);
/* opts = {prob_miss_log, prob_hit_log, clamp_min_log, clamp_max_log, occupancy_threshold_log}
 * (probabilistic_map.hpp:56-64) */
void orc_map_set_options(void* m, const int32_t opts[5]);
void orc_map_get_options(void* m, int32_t opts[5]);
/* insertPointCloud (probabilistic_map.hpp:141-160). f32: PointT = {float x,y,z[,pad]} with
 * stride_bytes 12 or 16 (pcl::PointXYZ layout); f64: PointT = Eigen::Vector3d, stride 24. */
void orc_map_insert_f32(void* m, const void* pts, int64_t stride_bytes, int64_t n,
                        const float origin[3], double max_range);
void orc_map_insert_f64(void* m, const double* pts, int64_t n, const double origin[3],
                        double max_range);
/* addHitPoint / addMissPoint (probabilistic_map.cpp:30-54): queued until the next insert */
void orc_map_add_hit(void* m, const double p[3]);
void orc_map_add_miss(void* m, const double p[3]);
/* kind 0 isOccupied, 1 isUnknown, 2 isFree (probabilistic_map.cpp:56-75) */
void orc_map_query(void* m, const int32_t* xyz, int64_t n, int kind, uint8_t* out);
/* kind 0 getOccupiedVoxels, 2 getFreeVoxels (probabilistic_map.cpp:108-126); unordered */
int64_t orc_map_get_voxels(void* m, int kind, int32_t* xyz, int64_t cap);
int64_t orc_map_active_count(void* m);
/* forEachCell over grid(): (coord, CellT word) pairs, unordered */
int64_t orc_map_dump(void* m, int32_t* xyz, uint32_t* words, int64_t cap);
/* counters of the LAST insert: {N points, E rays cast, V = sum(ray cells)+N, U cells changed}.
 * The port counts them natively; the reference build cannot see inside the unmodified code and
 * returns -1 for E and V, and U only if orc_map_track_updates(m,1) was set (dump diff, slow). */
void orc_map_counters(void* m, int64_t out[4]);
void orc_map_track_updates(void* m, int enable);
/* seconds spent inside the last insertPointCloud call only (steady_clock, the protocol of
 * bonxai_map/benchmark/benchmark_kitti.cpp:146-152) */
double orc_map_last_insert_seconds(void* m);

/* test helper (port library only): order-independent digest {sum, xor, count} of n (coord, 4-byte word) pairs, the
 * function bnx_grid_digest computes on the device (include/bonxai_b200.h) */
void orc_digest_pairs(const int32_t* xyz, const uint32_t* words, int64_t n, uint64_t out[3]);

#ifdef __cplusplus
}
#endif
#endif
