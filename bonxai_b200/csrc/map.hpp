// ProbabilisticMap on the device (bonxai_map/include/bonxai_map/probabilistic_map.hpp:27-203,
// bonxai_map/src/probabilistic_map.cpp:30-126).
#pragma once

#include "grid.hpp"

namespace bnx {

// sharded map, peer-memory exchange: a rank's mailbox is one device allocation that every peer maps
// (CUDA IPC across processes, the raw pointer inside one process). The producing kernels store their records straight
// into block [rank] of the OWNER's inbox over NVLink and the kernel's last block stamps an arrival flag there; the
// consuming kernel spins on its own flags. No collective launch, no staging copy.
constexpr int MAX_PEERS = 16;
constexpr size_t MBOX_HEADER = 4096;      // flag area in front of the two inboxes
constexpr u32 MBOX_FLAG1 = 0;             // u32[2][MAX_PEERS]  exchange 1 (one set per inbox parity): stamp of the scan whose records have arrived
constexpr u32 MBOX_FLAG2 = 64;            // u32[MAX_PEERS]  exchange 2
constexpr u32 MBOX_FLAGS4 = 128;          // uint4[2][MAX_PEERS] {pool error bits, overflow bits, 0, stamp}, slot = stamp & 1
struct PeerBoxes {
  int4* rec[MAX_PEERS];   // owner o: block of this rank in o's endpoint-record inbox ([0] = {count})
  int4* leaf[MAX_PEERS];  // owner o: block of this rank in o's leaf-record inbox
  u32* flag[MAX_PEERS];   // owner o: its flag area (NULL: the caller moves the staged buffers, no flags)
};

// everything a scan's kernels need besides the grid, passed by value
struct ScanParams {
  double ox, oy, oz;  // origin promoted to fp64 (ConvertPoint, grid_coord.hpp:134-162)
  double max_range, max_range_sqr, inv_res;
  i32 Ox, Oy, Oz;              // origin voxel (probabilistic_map.cpp:91)
  i32 miss, hit, cmin, cmax;   // Options, probabilistic_map.hpp:56-64
  u32 c;                       // _update_count in {1,2,3}, probabilistic_map.hpp:129
  u32 seq;                     // scan serial number: first-touch stamp of leaves
  u32 n;                       // points in this scan
  u32 hash_mask;               // endpoint dedupe table slots - 1
  u32 tile_cap;                // entries of tile_first
  u32 touched_cap;             // entries of the touched-leaf list
  u32 max_chunks;              // per-ray chunk limit that keeps the packed (rays, chunks) counter exact
  u32 packed;                  // 1: dedupe table uses packed 64-bit keys (all endpoints fit 21 bits per axis)
  u32 rank, world;             // map sharding: this process owns the roots with shard_owner(root) == rank
  u32 rec_cap;                 // sharded: slots per peer block of the endpoint record exchange
  u32 par;                     // sharded, peer memory: which of the two endpoint inboxes this scan uses (xseq1 & 1)
  u32 leaf_cap2;               // sharded: slots per peer block of the leaf-mask exchange
  u32 touched2_cap;            // sharded: entries of the scratch-grid touched list
  u32 xseq1, xseq2;            // sharded, peer-memory exchange: arrival stamps of this scan's exchange 1 / exchange 2 + flags
  u32 async_id;                // pipelined insert: serial of this scan (NONE for the synchronous path)
  u32 clean16;                 // pipelined insert: 16-byte units of dedupe table k_mark zeroes for the next scan
  u32 dense;                   // 1: per-scan marks go to the dense window around the origin, 0: into the leaves
  i32 W0x, W0y, W0z;           // dense window: leaf-block coordinates (voxel >> 3) of its corner
  u32 D;                       // dense window: blocks per axis
  u32 dlist_cap;               // dense window: entries of the touched-block list
  u32 fleet;                   // sharded: 1 = every rank inserts the scan of its OWN sensor in this step (see Map::set_fleet)
  i32 fO[MAX_PEERS][3];        // fleet: origin voxel of sensor s (the points held by rank s)
  u32 use_transform;           // fused ROS pre-step: drop non-finite points, then T * p in float before classifying
  float T[12];                 // rows 0..2 of the 4x4 sensor->world matrix
};

struct ScanCounters {
  unsigned long long ray_chunk;  // (rays with >= 1 cell) << 40 | 8-cell chunks: ONE atomic orders both
  unsigned long long sum_m;      // sum of ray lengths in cells
  u32 n_endpoints;               // endpoint voxels updated this scan (E)
  u32 n_touched;                 // leaves on the touched list
  u32 n_changed;                 // cells changed by the free-space apply pass
  u32 overflow;                  // scan scratch overflow bits
  u32 n_touched2;                // sharded: scratch-grid leaves touched (cells owned by other ranks)
  u32 n_dropped;                 // fused pre-step: non-finite points removed from the scan
  GridCounters gc;               // snapshot of the grid counters taken by the last kernel of the scan
  u32 cnt1[MAX_PEERS];           // sharded: endpoint records bucketed for owner o
  u32 cnt2[MAX_PEERS];           // sharded: leaf records emitted for owner o
  u32 done1, done2;              // last-block tickets of the two producing kernels
  u32 gate_pool, gate_ovf;       // sharded: the reduced error flags the apply pass acted on
  u32 done3;                     // last-block ticket of the merge kernel
  u32 gate_fill;                 // sharded: fullest leaf-record block of the scan over all ranks
};
static_assert(sizeof(ScanCounters) % 16 == 0, "cleared in 16-byte units");

// one record per pipelined scan, written by the device into pinned host memory (zero copy) when the scan ends
struct AsyncRecord {
  u32 error, n_leaves, n_inner, n_roots;
  u32 n_endpoints, n_changed, n_touched, n_points;
  unsigned long long sum_m, ray_chunk;
  u32 n_dropped;
  u32 leaf_fill;  // sharded: fullest leaf-record block of the scan over all ranks (same value on every rank)
  u32 pad;
  volatile u32 id;  // written last
};
static_assert(sizeof(AsyncRecord) == 64, "one record per 64 bytes");

struct ScanBuffers {
  int4* ep;          // per point: endpoint voxel xyz + type (0 hit, 1 miss)
  u32* slot_of;      // per point: its slot in the dedupe table
  u32* table;        // endpoint dedupe table: lowest point index per voxel
  unsigned long long* keys;  // packed-key flavour of the table
  int4* rays;        // per ray with >= 1 cell: end voxel xyz + first chunk
  u32* tile_first;   // per 32-chunk tile: ray that owns the tile's first chunk
  u32* touched;      // leaves first touched in this scan
  int4* pending;     // queued addHitPoint/addMissPoint endpoints (xyz, type)
  u32* touched2;     // sharded: scratch-grid leaves touched in this scan
  unsigned char* ray_src;  // sharded fleet step: which sensor a ray belongs to
  unsigned long long* dense;  // dense window: per block {u64 touched[8]; u64 hit[8]} = one 128-B line, all zero between scans
  u32* dbits;        // dense window: one bit per block, set when the block gets its first mark of the scan
  u32* dhint;        // leaf hint table: hash of absolute block coordinates -> leaf index + 1 (verified on use)
  const int4* recs;  // sharded: received endpoint records, [world][rec_cap], element 0 of a block = {count}
  const u32* gate;   // sharded: all-reduced error flags; the apply kernels skip when any is set (NULL otherwise)
  const u32* my_flags;  // sharded, peer-memory exchange: this rank's mailbox flag area (NULL: exchanges run by the caller)
  const PeerBoxes* px;  // sharded: where the records of this rank go (device memory)
  ScanCounters* sc;
  AsyncRecord* ring;  // pipelined insert: mapped pinned host memory, RING entries
  const u32* poison;  // &GridCounters::error of the map's grid: non-zero freezes every scan kernel
};

class Map {
 public:
  Map() = default;
  ~Map();
  int init(double resolution);

  int insert(const void* points, i64 stride_bytes, i64 n, bool f64, const double origin[3], double max_range, int where);
  int add_point(const double p[3], bool miss);
  // the next insert / insert_async first drops non-finite points and applies this 4x4 (row-major, float) to the rest
  void set_next_transform(const float T16[16]) {
    for (int k = 0; k < 12; ++k) next_T_[k] = T16[k];
    use_next_T_ = true;
  }
  void clear_next_transform() { use_next_T_ = false; }
  // Fleet step of a sharded map: the NEXT sharded insert takes one scan per rank — rank s holds the whole cloud of sensor
  // s, origins[s] is that sensor's origin — and gives the result of inserting the scans of sensors 0..world-1 one after
  // the other (each with its own update id), all of them concurrently. Needs sensors whose reach (max_range) does not
  // overlap: then no cell is touched by two of them and the order cannot matter; otherwise BNX_ERR_UNSUPPORTED (insert
  // them one by one). origins = nullptr switches back to one scan split over the ranks.
  int set_fleet(const double* origins_world_x3);
  // where a scan keeps its per-scan marks: 1 = in the leaves ("sparse": serves every scan), 2 = in the dense window
  // around the origin when the range allows it (experimental: exact, but slower as measured), 0 = default (sparse
  // unless BNX_DENSE=1)
  int set_marking(int mode);

  // ---- root-key sharding across processes (one map shard per GPU), as stages: with caller-owned exchange buffers the
  // caller moves them between the stages; with mailboxes attached (NULL buffers) the kernels exchange by themselves.
  int shard_config(int rank, int world);
  int shard_begin(const void* points, i64 stride_bytes, i64 n, bool f64, u32 index_base, const double origin[3], double max_range,
                  void* send_records, i64 cap_records, int where);
  int shard_resolve_mark(const void* recv_records, void* send_leaves, i64 cap_leaves);
  int shard_merge(const void* recv_leaves, void* flags);
  int shard_finish(const void* flags_reduced, int* retry);
  // the same protocol driven by the library: mailboxes (peer memory) set up over NCCL, or NCCL send/recv all-to-alls +
  // all-reduce with BNX_SHARD_EXCHANGE=nccl (NCCL is resolved with dlopen). async: nothing synchronises; failures
  // freeze all ranks at the same scan.
  static int nccl_unique_id(const char* nccl_path, void* out128);
  int shard_comm_init(const char* nccl_path, const void* unique_id128, int rank, int world);
  int shard_insert(const void* points, i64 stride_bytes, i64 n, bool f64, u32 index_base, i64 n_max, const double origin[3], double max_range,
                   int where, bool async);
  // peer-memory exchange. p2p_alloc: (re)creates this rank's mailbox and returns its CUDA IPC handle (64 bytes) and
  // raw device pointer; p2p_attach: maps the mailboxes of all ranks (handles: [world][64] bytes; local_ptrs != NULL:
  // the shards live in this process and the raw pointers are used). The staged calls then take NULL buffers.
  // shard_comm_init does both itself (handles all-gathered through NCCL) unless BNX_SHARD_EXCHANGE=nccl.
  int p2p_alloc(i64 cap_records, i64 cap_leaves, void* ipc_handle64, void** local_ptr);
  int p2p_attach(const void* handles, void* const* local_ptrs);
  int exchange_kind() const { return p2p_ready_ ? 2 : (comm_ ? 1 : 0); }  // 0 caller, 1 NCCL, 2 peer memory
  // bootstrap without NCCL: the caller supplies the all-gather that hands the 64-byte mailbox handles around (any
  // transport: gloo, MPI, a socket). fn(ctx, send, recv, bytes_per_rank) gathers `bytes_per_rank` bytes of HOST memory
  // from every rank into recv[world][bytes_per_rank]; it is collective. Peer memory is the only exchange then.
  typedef int (*AllGatherFn)(void* ctx, const void* send, void* recv, i64 bytes_per_rank);
  int shard_host_init(int rank, int world, AllGatherFn fn, void* ctx);
  // what the sharded pipeline really did: {resolve_mark attempts, frozen-pipeline replays, mailbox (re)creations,
  // collective drains, synchronous retries, leaf-inbox capacity, largest leaf-inbox fill seen, scans inserted}
  i64 shard_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  i64 mailbox_cap(int which) const { return which ? mbox_cap_leaf_ : mbox_cap_rec_; }

  // ---- pipelined insert: enqueue a scan and return; drain() completes everything queued (growing pools and
  // replaying from the first scan that ran short, if any). Input buffers must stay valid until drain().
  int insert_async(const void* points, i64 stride_bytes, i64 n, bool f64, const double origin[3], double max_range, int where);
  int drain(bool report = true);
  i64 totals[4] = {0, 0, 0, 0};  // cumulative N, E, V, U over every scan inserted so far
  int query(const i32* xyz, i64 n, int kind, u8* out, int where);

  Grid grid;
  i32 options[5];
  u32 update_count = 1;
  i64 counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool profiling = false;
  double phase_us[8] = {0, 0, 0, 0, 0, 0, 0, 0};

 private:
  // ---- one scan = launch_front (clear + classify) + launch_back (resolve, mark, apply)
  int reserve_scan(i64 n, i64 stride_bytes, double max_range, i64 table_n = -1);
  size_t tile_bytes(size_t np, double max_range) const;
  int build_params(i64 n, const double origin[3], double max_range, ScanParams* out);
  int launch_front(cudaStream_t s, const void* d_points, i64 stride_bytes, bool f64, ScanParams& p);
  int launch_back(cudaStream_t s, ScanParams& p, bool first_attempt);
  int launch_scan(const void* d_points, i64 stride_bytes, bool f64, ScanParams& p, bool first_attempt);  // both, on the map's stream
  int run_scan(const void* d_points, i64 stride_bytes, bool f64, const ScanParams& base, bool reuse_classify);  // synchronous, with retries
  void account(const ScanCounters& st, i64 n, i64 pending, i64 retries);

  ScanBuffers buf_ = {};
  DevBuf b_pts_, b_rays_, b_tiles_, b_touched_, b_touched2_, b_pending_, b_q_xyz_, b_q_out_;
  DevBuf b_dense_, b_dbits_, b_dhint_, b_dlist_;
  DevBuf b_ray_src_;
  std::vector<double> fleet_origins_;  // [world][3] of the next sharded insert (empty: one scan split over the ranks)
  std::vector<double> sp_fleet_;       // ... of the sharded scan in flight  // dense marking window (allocated at the first scan that can use it)
  int marking_ = 0;
  u32 dense_D_ = 0;                      // blocks per axis the window buffers are sized (and zeroed) for
  int reserve_dense(ScanParams& p);      // decides p.dense and sizes the window
  int resume_apply(cudaStream_t s, ScanParams& p);
  int dense_internal_error(const ScanCounters& st, const ScanParams& p);
  u32 dense_dim(double max_range) const;  // blocks per axis a scan of this range needs (0: not eligible)
  bool scan_is_dense(const double origin[3], double max_range) const;
  cudaEvent_t grown_ = nullptr;  // pools grown ahead of need: the zero fill on the copy stream
  i64 grown_ahead_ = 0;          // how often that happened (statistics)
  // Scratch written BEFORE a scan touches the map (classify: endpoints, their dedupe table, the counters) exists once
  // per scan in flight: the pipelined insert classifies on its own stream, many scans ahead of the map updates.
  // The synchronous and the sharded paths use set 0.
  struct ScratchSet {
    DevBuf ep, slot, table, stage;  // table = [ScanCounters | table u32[slots] | keys u64[slots]]; stage = H2D staging
    u64 clean_slots = 0;            // dedupe-table slots known to be zero (left clean by the set's previous scan)
    bool sc_clean = false, t1_clean = false;  // sharded pipeline: counters / sender-side table known to be zero
    cudaEvent_t classified = nullptr;         // front half of the set's scan done
  };
  static constexpr int SETS = 34;  // scans in flight (<= 32) + 2: the set of scan id - SETS is free when scan id is enqueued
  ScratchSet sets_[SETS];
  int set_ = 0, sets_active_ = SETS;
  i64 sets_n_ = -1;              // points every active set is sized for (pipelined insert)
  size_t sets_stage_bytes_ = 0;  // H2D staging bytes every active set holds
  ScratchSet& S() { return sets_[set_]; }
  ScanCounters* d_sc_ = nullptr;      // head of the current set's table buffer: counters + dedupe table are cleared by ONE memset
  ScanCounters* h_status_ = nullptr;  // pinned
  u32 n_pending_ = 0;                 // queued addHitPoint / addMissPoint endpoints
  float next_T_[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  bool use_next_T_ = false;
  u32 seq_ = 0;  // scan serial number (leaf stamps)
  cudaEvent_t ev_[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // profiling ([6], [7]: inside the sharded resolve_mark stage)

  // ---- pipelined insert
  static constexpr u32 RING = 1024;
  struct Queued {
    ScanParams p;
    const void* points;
    i64 stride;
    bool f64;
    int where;
  };
  std::vector<Queued> queue_;
  int drain_queue();
  int complete_queue();
  static bool sync_via_ring();  // BNX_SYNC_RING=0 restores the memset + copy + cudaStreamSynchronize flavour of insert()
  int deferred_ = BNX_OK;  // error of a pipelined scan that was refused, reported by the next synchronising call
  std::string deferred_msg_;
  AsyncRecord* h_ring_ = nullptr;  // pinned + mapped: one record per scan, written by the device
  AsyncRecord* d_ring_ = nullptr;
  u32 async_next_ = 0;
  size_t done_upto_ = 0;        // queue_ entries whose record has been seen
  u64 max_leaf_growth_ = 2048;  // largest per-scan leaf allocation seen so far (head-room estimate)
  cudaStream_t pre_stream_ = nullptr;  // front halves: H2D copy + classify

  // ---- sharded map
  int rank_ = 0, world_ = 1;
  Grid* scratch_ = nullptr;  // staging grid for cells whose root another rank owns (masks only)
  ScanParams sp_ = {};       // parameters of the scan in flight
  i64 shard_retries_ = 0;
  bool staged_p2p_ = false;   // the scan in flight uses the mailboxes
  bool shard_async_ = false;  // the scan being enqueued is pipelined (set by shard_insert around the stages)
  u32 shard_async_id_ = 0, shard_attempt_ = 0;
  i64 shard_n_max_ = 0;
  bool t2_clean_ = false;  // pipelined: receiver table known to be zero
  bool t2_packed_ = true;  // flavour of the receiver table the last scan used
  DevBuf b_table2_;        // receiver-side dedupe table
  void shard_phase_times();
  struct ShardQueued {
    const void* points;
    i64 stride, n, n_max;
    bool f64;
    u32 index_base, async_id, c;
    double origin[3], max_range;
    int where;
    int set = 0;                // scratch set the scan's front half used
    std::vector<double> fleet;  // origins of a fleet step (empty otherwise)
  };
  std::vector<ShardQueued> squeue_;
  int shard_drain();
  // NCCL: bootstrap of the mailboxes, and the exchange itself with BNX_SHARD_EXCHANGE=nccl
  void* comm_ = nullptr;  // ncclComm_t
  AllGatherFn host_gather_ = nullptr;  // caller-supplied bootstrap (instead of NCCL)
  void* host_gather_ctx_ = nullptr;
  u64 drain_prev_[4] = {0, 0, 0, 0};  // leaves / inner nodes of the map and of the scratch grid at the previous drain
  u32 drain_max_fill_ = 0;  // largest leaf-inbox fill (max over ranks, so equal on every rank) since the last drain
  DevBuf x_send1_, x_recv1_, x_send2_, x_recv2_, x_flags_, x_handles_;
  i64 cap_rec_ = 0, cap_leaf_ = 1 << 16;  // 80-B leaf records per sender block; doubled at a collective drain when half full
  int all_to_all(const void* send, void* recv, size_t block_bytes);
  void note_leaf_fill(u32 fill);
  // peer-memory exchange
  int p2p_collective_setup(i64 cap_records, i64 cap_leaves);
  void p2p_close_peers();
  int upload_boxes();
  unsigned char* mbox_ = nullptr;  // this rank's mailbox: [flags 4 KiB | inbox1 [world][cap_rec] x 16 B | inbox2 [world][cap_leaf] x 80 B]
  i64 mbox_cap_rec_ = 0, mbox_cap_leaf_ = 0;
  std::vector<void*> mbox_retired_;  // replaced mailboxes stay allocated until the map dies (a slow peer may still map them)
  void* peer_base_[MAX_PEERS] = {};
  bool peer_ipc_[MAX_PEERS] = {};
  bool p2p_ready_ = false, want_p2p_ = false;
  u32 xseq1_ = 0, xseq2_ = 0;  // arrival stamps: every rank enqueues the same sequence of exchanges
  PeerBoxes px_host_ = {}, px_uploaded_ = {};
  bool px_uploaded_valid_ = false;
  DevBuf b_px_;
  // pipelined insert with host input: H2D staging ring on a copy stream
  static constexpr size_t SHARD_QUEUE = 64;  // scans between two collective drains
  static constexpr int SHARD_STAGES = (int)SHARD_QUEUE + 2;
  cudaStream_t copy_stream_ = nullptr;
  // pipelined + peer memory: the front half of scan k+1 (copy, classify, bucket + stores into the owners' inboxes) runs
  // on the pre-stream while scan k is still marking / merging / applying. Safe because the endpoint inboxes are double
  // buffered by the parity of the exchange serial, and the front half of exchange x+2 waits (event) for this rank's merge
  // of exchange x: that merge has seen the exchange-2 stamps of EVERY rank, which they send after they have consumed
  // inbox x.
  static constexpr int SHARD_SETS = 4;
  cudaEvent_t x_merged_[SHARD_SETS] = {};   // merge kernel of exchange x enqueued (x % SHARD_SETS)
  u32 x_merged_seq_[SHARD_SETS] = {};       // which exchange serial the event belongs to (0: never recorded)
  cudaEvent_t x_begun_[SHARD_SETS] = {};    // front half of exchange x done
  i64 shard_sets_n_ = 0;                    // points the SHARD_SETS scratch sets are sized for
  DevBuf x_stage_[SHARD_STAGES];
  cudaEvent_t x_copied_[SHARD_STAGES] = {};
  size_t x_stage_bytes_ = 0;
};

}  // namespace bnx
