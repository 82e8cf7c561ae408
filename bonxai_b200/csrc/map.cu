// ProbabilisticMap::insertPointCloud on the device.
//
// Reference: bonxai_map/include/bonxai_map/probabilistic_map.hpp:141-203 (insertPointCloud, RayIterator),
//            bonxai_map/src/probabilistic_map.cpp:30-54,77-106 (addHitPoint, addMissPoint, updateFreeCells).
//
// The reference walks points, then rays, sequentially and lets the per-cell update_id decide who updates a
// cell. The same result is produced here by four data-parallel phases:
//   1 classify   fp64 range clip + posToCoord per point; a scan-local hash finds, per endpoint voxel, the
//                LOWEST point index (the point the reference would have processed first: it decides hit/miss)
//   2 resolve    one thread per winning point: find-or-create the leaf, stale test (update_id == c against the
//                PRE-scan state), set the endpoint's bit in the leaf's per-scan HIT mask (hit) or touched mask (miss)
//                and — if the voxel differs from the origin's — emit a ray with its range in the flat space of
//                8-cell ray chunks
//   3 mark       the flat chunk space is walked by all threads: exact integer DDA restarted from the closed form
//                at the chunk's first cell, setting bits in a per-leaf 512-bit "touched" mask (test before atomicOr)
//   4 apply      one warp per touched leaf: hit endpoints get the clamped hit update, every other touched cell whose
//                update_id != c the clamped miss update; all get the stamp. Hits win over ray misses exactly like in
//                the reference, where the endpoints are stamped before any ray is cast.
// Phases 1-3 never change a cell value, so a scan that runs out of pool space is simply repeated after the
// pools have grown; phase 4 cannot fail.
// The pipelined insert runs phase 1 (and the H2D copy) on its own stream, scans ahead of phases 2-4; the sharded
// map (one shard per GPU) adds the stages around the two record exchanges, which are peer-memory stores + arrival
// stamps.
#include "map.hpp"

#include "nccl_dyn.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace bnx {

namespace {

constexpr int TPB = 256;
constexpr u32 CHUNK = 8;  // ray cells per work item
#ifndef MARK_MIN_BLOCKS
#define MARK_MIN_BLOCKS 6
#endif
constexpr unsigned long long CHUNK_FIELD = (1ull << 40) - 1ull;
constexpr u32 RING_SIZE = 1024;  // entries of the pipelined-insert record ring (Map::RING)
constexpr u32 OVF_TILES = 1u, OVF_CHUNKS = 2u, OVF_RECORDS = 4u, OVF_LEAVES = 8u, OVF_WINDOW = 16u;

inline int blocks_for(i64 n, int tpb = TPB) { return (int)std::max<i64>(1, ceil_div(n, tpb)); }

// ------------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): the kernels of a scan are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the NEXT kernel of the stream is scheduled while this one
// still runs (its launch latency and block start-up overlap this kernel's tail) and blocks in pdl_enter() until the
// previous kernel has completed and its memory operations are visible. pdl_enter() is the first statement of
// every scan kernel; in a kernel launched the ordinary way both instructions do nothing. BNX_PDL=0 turns it off.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool v = [] {
    const char* e = std::getenv("BNX_PDL");
    return !(e && std::strcmp(e, "0") == 0);
  }();
  return v;
}

// only the pipelined paths use it: between the memsets, event records and copies of the synchronous path it costs time
thread_local bool t_pdl = false;
struct PdlScope {
  bool saved;
  explicit PdlScope(bool on) : saved(t_pdl) { t_pdl = on && pdl_enabled(); }
  ~PdlScope() { t_pdl = saved; }
};

template <typename... KArgs, typename... Args>
inline void launch_scan_kernel(void (*kernel)(KArgs...), int blocks, int threads, cudaStream_t s, Args&&... args) {
  note_launch();
  if (!t_pdl) {
    kernel<<<blocks, threads, 0, s>>>(KArgs(args)...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------
// peer-memory exchange primitives (sharded map)
// ------------------------------------------------------------------------------------------------
// a dead peer becomes ERR_PEER, never a hang: 20 s by default, BNX_PEER_TIMEOUT_MS overrides (read when the map is sharded)
__device__ unsigned long long g_peer_timeout_ns = 20000000000ull;

__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
  u32 v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(u32* p, u32 v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Consumer side: threads 0..world-1 of the block spin on the arrival stamps flags[t * stride] that the peers store
// into THIS rank's mailbox (acquire, system scope), then the block synchronises. Everything a peer wrote before its
// stamp is visible afterwards (read it with __ldcg: the lines were never in this SM's L1 during this kernel).
// Stamps only grow, so ">= seq" is the test. A peer that never arrives sets ERR_PEER (sticky: later waits return at once).
__device__ __forceinline__ void wait_arrivals(const u32* flags, u32 stride, u32 world, u32 seq, u32* err) {
  if (flags == nullptr) return;
  if (threadIdx.x < world) {
    const u32* f = flags + threadIdx.x * stride;
    if ((int)(ld_acquire_sys(f) - seq) < 0) {
      const unsigned long long t0 = global_ns();
      while ((int)(ld_acquire_sys(f) - seq) < 0) {
        if (*reinterpret_cast<volatile u32*>(err) & ERR_PEER) break;
        if (global_ns() - t0 > g_peer_timeout_ns) {
          atomicOr(err, ERR_PEER);
          break;
        }
        __nanosleep(128);
      }
    }
  }
  __syncthreads();
}

// Producer side, called by every block at the end of a kernel that stored records into the owners' inboxes: the
// LAST block writes the record count of every block header and then the arrival stamp into every owner's flag area
// (release, system scope; each block fenced its own stores before taking its ticket).
__device__ __forceinline__ void publish_blocks(const PeerBoxes* px, bool leaves, const u32* cnt, u32* done, u32 flag_off, u32 rank, u32 seq, u32 world,
                                               u32 cap, size_t rec_off = 0) {
  __shared__ bool s_last;
  const bool remote = px->flag[0] != nullptr;
  __syncthreads();  // the block's stores happen before thread 0's fence (cumulative), which happens before its ticket
  if (threadIdx.x == 0) {
    if (remote) {
      __threadfence_system();
    } else {
      __threadfence();
    }
    s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *done = 0u;  // ready for the next scan (the pipelined path does not memset the counters)
  __threadfence();
  if (threadIdx.x < world) {
    const u32 o = threadIdx.x;
    const u32 c = min(*reinterpret_cast<const volatile u32*>(cnt + o), cap - 1u);
    int4* block = leaves ? px->leaf[o] : px->rec[o] + rec_off;
    block[0] = make_int4((int)c, 0, 0, 0);
    if (remote) {
      __threadfence_system();
      st_release_sys(px->flag[o] + flag_off + rank, seq);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// phase 1: classify + endpoint dedupe
// ------------------------------------------------------------------------------------------------
// The per-point body of insertPointCloud (probabilistic_map.hpp:146-158) followed by posToCoord
// (bonxai.hpp:404-410). Every fp64 operation is an explicitly rounded intrinsic: no FMA contraction, and
// squaredNorm associates as (x*x + y*y) + z*z like the oracle's Eigen stand-in.
__device__ __forceinline__ int4 classify_point(double px, double py, double pz, const ScanParams& p) {
  const double vx = __dsub_rn(px, p.ox), vy = __dsub_rn(py, p.oy), vz = __dsub_rn(pz, p.oz);
  const double sq = __dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz));
  int type = 0;
  double ex = px, ey = py, ez = pz;
  if (sq >= p.max_range_sqr) {
    const double nrm = __dsqrt_rn(sq);
    ex = __dadd_rn(p.ox, __dmul_rn(__ddiv_rn(vx, nrm), p.max_range));
    ey = __dadd_rn(p.oy, __dmul_rn(__ddiv_rn(vy, nrm), p.max_range));
    ez = __dadd_rn(p.oz, __dmul_rn(__ddiv_rn(vz, nrm), p.max_range));
    type = 1;
  }
  return make_int4(__double2int_rd(__dmul_rn(ex, p.inv_res)), __double2int_rd(__dmul_rn(ey, p.inv_res)),
                   __double2int_rd(__dmul_rn(ez, p.inv_res)), type);
}

// Endpoint dedupe table, two flavours (0 always means "empty" so that ONE memset clears counters + table):
//   PACKED   all endpoint voxels of the scan fit 21 bits per axis (guaranteed by the host from origin and
//            max_range): 64-bit key = packed xyz + 1 claimed by atomicCAS, value = max(~index) = lowest index.
//   indirect any coordinates: a slot holds (point index + 1) and its key is ep[index].xyz; needs ep[i] to be
//            visible (fence) before the slot can name i.
__device__ __forceinline__ unsigned long long pack_key(const int4& e) {
  return ((unsigned long long)(u32)(e.x + (1 << 20)) | ((unsigned long long)(u32)(e.y + (1 << 20)) << 21) |
          ((unsigned long long)(u32)(e.z + (1 << 20)) << 42)) + 1ull;
}

template <bool F64, bool VEC4, bool PACKED>
__global__ void __launch_bounds__(TPB) k_classify(const unsigned char* __restrict__ pts, u32 stride, ScanParams p, ScanBuffers b) {
  pdl_enter();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n || *b.poison) return;
  // Non-finite points leave the cloud here, in every flavour: the reference's callers remove them before
  // insertPointCloud (bonxai_ros/src/bonxai_server.cpp:148-154) because the reference itself has undefined behaviour
  // on them ((int32)floor(NaN), a ray of 2^31 cells); a defined "dropped" beats a platform-dependent voxel.
  double px, py, pz;
  if (F64) {
    const double* q = reinterpret_cast<const double*>(pts + (size_t)i * stride);
    px = q[0];
    py = q[1];
    pz = q[2];
    if (!(isfinite(px) && isfinite(py) && isfinite(pz))) {
      b.ep[i] = make_int4(0, 0, 0, 2);
      b.slot_of[i] = NONE;
      atomicAdd(&b.sc->n_dropped, 1u);
      return;
    }
  } else {
    float fx, fy, fz;
    if (VEC4) {
      const float4 q = *reinterpret_cast<const float4*>(pts + (size_t)i * 16);  // pcl::PointXYZ
      fx = q.x;
      fy = q.y;
      fz = q.z;
    } else {
      const float* q = reinterpret_cast<const float*>(pts + (size_t)i * stride);
      fx = q[0];
      fy = q[1];
      fz = q[2];
    }
    if (!(isfinite(fx) && isfinite(fy) && isfinite(fz))) {
      b.ep[i] = make_int4(0, 0, 0, 2);
      b.slot_of[i] = NONE;
      atomicAdd(&b.sc->n_dropped, 1u);
      return;
    }
    if (p.use_transform) {
      // the ROS caller's pre-step (bonxai_ros/src/bonxai_server.cpp:148-171): after the non-finite filter the cloud
      // goes through pcl::transformPointCloud in float. Association of PCL's SSE kernel on x86-64:
      // x*c0 + (y*c1 + (z*c2 + c3)), every product and sum rounded on its own (no FMA).
      const float tx = __fadd_rn(__fmul_rn(fx, p.T[0]), __fadd_rn(__fmul_rn(fy, p.T[1]), __fadd_rn(__fmul_rn(fz, p.T[2]), p.T[3])));
      const float ty = __fadd_rn(__fmul_rn(fx, p.T[4]), __fadd_rn(__fmul_rn(fy, p.T[5]), __fadd_rn(__fmul_rn(fz, p.T[6]), p.T[7])));
      const float tz = __fadd_rn(__fmul_rn(fx, p.T[8]), __fadd_rn(__fmul_rn(fy, p.T[9]), __fadd_rn(__fmul_rn(fz, p.T[10]), p.T[11])));
      fx = tx;
      fy = ty;
      fz = tz;
    }
    px = (double)fx;
    py = (double)fy;
    pz = (double)fz;
  }
  const int4 e = classify_point(px, py, pz, p);
  b.ep[i] = e;
  u32 slot = (u32)hash3(e.x, e.y, e.z) & p.hash_mask;
  if (PACKED) {
    const unsigned long long key = pack_key(e);
    for (;;) {
      const unsigned long long k = atomicCAS(&b.keys[slot], 0ull, key);  // one L2 round trip, hit or claim
      if (k == 0ull || k == key) break;
      slot = (slot + 1) & p.hash_mask;
    }
    atomicMax(&b.table[slot], ~i);
  } else {
    __threadfence();  // ep[i] must be visible before a table slot can name i
    for (;;) {
      u32 v = b.table[slot];
      if (v == 0u) {
        v = atomicCAS(&b.table[slot], 0u, i + 1u);
        if (v == 0u) break;  // claimed an empty slot
      }
      const int4 o = __ldcg(&b.ep[v - 1u]);
      if (o.x == e.x && o.y == e.y && o.z == e.z) {
        if (i + 1u < v) atomicMin(&b.table[slot], i + 1u);
        break;
      }
      slot = (slot + 1) & p.hash_mask;
    }
  }
  b.slot_of[i] = slot;
}

// ------------------------------------------------------------------------------------------------
// ray geometry shared by resolve (chunk accounting) and mark (the walk)
// ------------------------------------------------------------------------------------------------
// A ray origin -> end has m = max|delta| cells k = 0..m-1 (probabilistic_map.hpp:173-183; the end cell is
// excluded). Along the major axis (|delta| == m) cell k sits exactly at origin + sign*k, so the ray is cut into
// chunks at the 8-cell LEAF boundaries of that axis: chunk 0 = the len0 cells up to the first boundary, then 8
// cells each. Inside a chunk the major leaf coordinate is constant and each minor axis crosses at most one
// leaf boundary.
struct RayGeom {
  u32 ax, ay, az, m;
  int sx, sy, sz;
  u32 len0, chunks;
  int Ox, Oy, Oz;  // origin voxel of the ray's sensor
};

// update id and origin voxel of sensor `src`: a fleet step (sharded map) inserts the scans of sensors 0..world-1 as if one
// after the other, so sensor s uses the id the s-th of those inserts would have had (probabilistic_map.cpp:103-105)
__device__ __forceinline__ u32 scan_c(const ScanParams& p, u32 src) { return p.fleet ? (p.c - 1u + src) % 3u + 1u : p.c; }
constexpr u32 LEAF_MASK = 0x0FFFFFFFu;  // touched-list entry of a fleet step = leaf | sensor << 28

__device__ __forceinline__ RayGeom ray_geom(const ScanParams& p, int ex, int ey, int ez, u32 src = 0) {
  RayGeom r;
  r.Ox = p.fleet ? p.fO[src][0] : p.Ox;
  r.Oy = p.fleet ? p.fO[src][1] : p.Oy;
  r.Oz = p.fleet ? p.fO[src][2] : p.Oz;
  const i64 dx = (i64)ex - r.Ox, dy = (i64)ey - r.Oy, dz = (i64)ez - r.Oz;
  r.ax = (u32)(dx < 0 ? -dx : dx);
  r.ay = (u32)(dy < 0 ? -dy : dy);
  r.az = (u32)(dz < 0 ? -dz : dz);
  r.sx = dx < 0 ? -1 : 1;
  r.sy = dy < 0 ? -1 : 1;
  r.sz = dz < 0 ? -1 : 1;
  r.m = max(max(r.ax, r.ay), r.az);
  const int Oa = r.ax == r.m ? r.Ox : (r.ay == r.m ? r.Oy : r.Oz);
  const int sa = r.ax == r.m ? r.sx : (r.ay == r.m ? r.sy : r.sz);
  r.len0 = sa > 0 ? 8u - ((u32)Oa & 7u) : ((u32)Oa & 7u) + 1u;
  r.chunks = r.m == 0u ? 0u : (r.m <= r.len0 ? 1u : 1u + (r.m - r.len0 + 7u) / 8u);
  return r;
}

// ------------------------------------------------------------------------------------------------
// phase 2: resolve winners -> endpoint records + rays
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, u32 lane) {
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (u32)o) v += t;
  }
  return v;
}

// OR `bits` into word w of a per-scan leaf mask (touched or hit); defined with the mark kernel below
__device__ __forceinline__ bool mark_bits(const GridDev& G, u32 leaf, unsigned long long* word, unsigned long long bits, u32 seq);
__device__ __forceinline__ void list_leaves(bool mine, u32 entry, u32* n_list, u32* list, u32 cap);

// which root does which rank own (map sharding)? Uses the upper hash bits: the root table slot uses the lower.
__host__ __device__ __forceinline__ u32 shard_owner(int rx, int ry, int rz, u32 world) {
  return (u32)((hash3(rx, ry, rz) >> 34) % world);
}

// 128-bit compare-and-swap (atom.cas.b128): claims a whole {x, y, z, best} slot of the receiver's any-coordinate dedupe
// table at once, so a slot's key is never seen half-written
__device__ __forceinline__ int4 cas128(int4* addr, int4 cmp, int4 val) {
  const unsigned long long clo = ((unsigned long long)(u32)cmp.y << 32) | (u32)cmp.x, chi = ((unsigned long long)(u32)cmp.w << 32) | (u32)cmp.z;
  const unsigned long long vlo = ((unsigned long long)(u32)val.y << 32) | (u32)val.x, vhi = ((unsigned long long)(u32)val.w << 32) | (u32)val.z;
  unsigned long long olo, ohi;
  asm volatile(
      "{\n\t.reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.global.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(olo), "=l"(ohi)
      : "l"(clo), "l"(chi), "l"(vlo), "l"(vhi), "l"(addr)
      : "memory");
  return make_int4((int)(u32)olo, (int)(olo >> 32), (int)(u32)ohi, (int)(ohi >> 32));
}
// the best (= ~lowest w) entry of a receiver-table slot: packed flavour table[slot], any-coordinate flavour slot.w
__device__ __forceinline__ u32 shard_slot_best(const ScanParams& p, const ScanBuffers& b, u32 slot) {
  return p.packed ? b.table[slot] : b.table[(size_t)slot * 4 + 3];
}

// Sharded map, receiver side of exchange 1: the inbox is [world][rec_cap] records with the count in element 0 of every
// block. The records of all blocks are addressed as ONE dense range 0..R-1 (block after block), so that the threads of
// the receiving kernels are fully used whatever the split between the senders is.
// Called by FULL warps: lane s reads the header of block s, a warp scan turns the counts into ranges. Returns R and, for
// i < R, the record itself.
__device__ __forceinline__ u32 shard_locate(const ScanParams& p, const ScanBuffers& b, u32 i, bool& found, int4& e, u32* from = nullptr) {
  const u32 lane = threadIdx.x & 31;
  u32 incl = lane < p.world ? min((u32)__ldcg(&b.recs[(size_t)lane * p.rec_cap]).x, p.rec_cap - 1u) : 0u;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32)o) incl += t;
  }
  const u32 total = __shfl_sync(0xffffffffu, incl, 31);
  found = false;
  u32 base = 0;
  for (u32 src = 0; src < p.world; ++src) {
    const u32 end = __shfl_sync(0xffffffffu, incl, src);
    if (!found && i >= base && i < end) {
      found = true;
      e = __ldcg(&b.recs[(size_t)src * p.rec_cap + 1u + (i - base)]);
      if (from) *from = src;
    }
    base = end;
  }
  return total;
}
// the receiver's dedupe table only uses as many slots as the records that actually arrived need (load <= 1/2)
__device__ __forceinline__ u32 shard_table_mask(const ScanParams& p, u32 received) {
  u32 slots = 1024u;
  while (slots < 2u * received) slots <<= 1;
  return min(slots - 1u, p.hash_mask);
}

// MODE 0: one thread per point of the scan (winner = lowest index of its endpoint voxel).
// MODE 1: one thread per queued addHitPoint / addMissPoint endpoint (already updated and stamped when it was
//         queued; it only needs its ray).
// MODE 2: sharded map — one thread per received endpoint record (dense range over the [world][rec_cap] inbox);
//         w = global point index << 1 | type, winner = lowest w of its voxel.
// dense window: index of the leaf block (lx, ly, lz) [voxel >> 3]; NONE if it lies outside (cannot happen for cells
// within max_range of the origin: reported as an overflow, never written)
__device__ __forceinline__ u32 dense_block(const ScanParams& p, int lx, int ly, int lz) {
  const u32 bx = (u32)(lx - p.W0x), by = (u32)(ly - p.W0y), bz = (u32)(lz - p.W0z);
  if (bx >= p.D || by >= p.D || bz >= p.D) return NONE;
  return (bz * p.D + by) * p.D + bx;
}

// DENSE: the scan's marks (touched + hit bits) live in the dense window, not in the leaves: resolve only READS the map
// (no leaf is created here; a voxel without a leaf is simply unknown, hence not stale).
template <int MODE, bool DENSE>
__global__ void __launch_bounds__(TPB) k_resolve(GridDev g, ScanParams p, ScanBuffers b, u32 count) {
  pdl_enter();
  constexpr bool PENDING = MODE == 1;
  __shared__ unsigned long long s_warp[TPB / 32];
  __shared__ u32 s_warp_e[TPB / 32];
  __shared__ unsigned long long s_base, s_m;
  __shared__ u32 s_base_e;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (MODE != 1 && *b.poison) return;  // an earlier pipelined scan ran short: freeze until the host recovers
  // MODE 2 is launched for the expected number of received records and loops (block-uniformly) over the rest
  for (u32 vb = blockIdx.x;; vb += gridDim.x) {
  const u32 i = vb * blockDim.x + threadIdx.x;
  u32 received = 0;
  if (threadIdx.x == 0) s_m = 0;

  bool is_end = false, winner = false;
  int4 e = make_int4(0, 0, 0, 0);
  u32 leaf = NONE, ci = 0, m = 0, chunks = 0, src = 0;  // src: the rank a record came from = its sensor in a fleet step
  if (MODE == 2) {
    received = shard_locate(p, b, i, winner, e, &src);
    if (winner) {
      winner = shard_slot_best(p, b, b.slot_of[i]) == ~(u32)e.w;
      e.w &= 1;
    }
  } else if (i < count) {
    if (MODE == 1) {
      e = b.pending[i];
      winner = true;
    } else {
      const u32 slot = b.slot_of[i];  // NONE: the point was dropped by the fused pre-step
      e = b.ep[i];                    // unconditional (coalesced): in flight together with the table look-up
      winner = slot != NONE && b.table[slot] == (p.packed ? ~i : i + 1u);
    }
  }
  if (!PENDING) {
    // Accessor::value(coord, true), bonxai.hpp:469-494 — warp-aggregated: neighbouring points of a scan mostly fall
    // into the same few leaves, so ONE lane per distinct leaf walks (and, if needed, creates) it and broadcasts the
    // index; nobody spins on a slot that a lane of its own warp holds locked.
    const u32 want = __ballot_sync(0xffffffffu, winner);
    if (winner) {
      const u32 peers = __match_any_sync(want, e.x >> 3) & __match_any_sync(want, e.y >> 3) & __match_any_sync(want, e.z >> 3);
      const int leader = __ffs(peers) - 1;
      if ((int)lane == leader) leaf = DENSE ? leaf_find(g, e.x, e.y, e.z) : leaf_find_or_create(g, e.x, e.y, e.z);
      leaf = __shfl_sync(peers, leaf, leader);
    }
  }
  if (winner) {
    bool stale = false;
    if (!PENDING) {
      ci = ((u32)e.x & 7u) | (((u32)e.y & 7u) << 3) | (((u32)e.z & 7u) << 6);
      if (leaf != NONE) {
        // mask word and cell are loaded together (one round trip); the cell only counts if its bit is on
        const u64 act = leaf_active(g, leaf)[ci >> 6];
        const u32 raw = reinterpret_cast<const u32*>(leaf_cells(g, leaf))[ci];
        const u32 word = ((act >> (ci & 63)) & 1ull) ? raw : 0u;
        stale = (word & 0xFu) == scan_c(p, src);  // probabilistic_map.cpp:34 / :47 — skipped AND no ray is cast
      } else {
        stale = !DENSE;  // sparse: pool exhausted, the scan will be repeated; dense: no leaf = unknown cell = not stale
      }
    }
    if (!stale) {
      is_end = true;
      const RayGeom rg = ray_geom(p, e.x, e.y, e.z, src);
      m = rg.m;  // probabilistic_map.hpp:180; the ray has exactly m cells (end excluded)
      chunks = rg.chunks;
    }
  }
  if (chunks > p.max_chunks) {  // would overflow the packed counter: refuse the scan (BNX_ERR_UNSUPPORTED)
    atomicOr(&b.sc->overflow, OVF_CHUNKS);
    chunks = 0;
    m = 0;
  }
  const unsigned long long mine = is_end && m ? ((1ull << 40) | chunks) : 0ull;
  // block-wide exclusive scans: packed (rays, chunks) and endpoint count
  unsigned long long incl = warp_incl_scan(mine, lane);
  const u32 eballot = __ballot_sync(0xffffffffu, is_end && !PENDING);
  const u32 e_excl_w = __popc(eballot & ((1u << lane) - 1u));
  if (lane == 31) s_warp[warp] = incl;
  if (lane == 0) s_warp_e[warp] = __popc(eballot);
  unsigned long long msum = m;
  for (int o = 16; o; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
  __syncthreads();
  if (lane == 0 && msum) atomicAdd(&s_m, msum);
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    u32 tot_e = 0;
    for (int k = 0; k < TPB / 32; ++k) {
      const unsigned long long c = s_warp[k];
      s_warp[k] = tot;
      tot += c;
      const u32 ce = s_warp_e[k];
      s_warp_e[k] = tot_e;
      tot_e += ce;
    }
    s_base = tot ? atomicAdd(&b.sc->ray_chunk, tot) : 0ull;
    s_base_e = tot_e ? atomicAdd(&b.sc->n_endpoints, tot_e) : 0u;
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_m) atomicAdd(&b.sc->sum_m, s_m);
  if (DENSE && is_end && !PENDING) {
    // the same, with the marks in the dense window: hit bits in the second half of the block's line
    const u32 blk = dense_block(p, e.x >> 3, e.y >> 3, e.z >> 3);
    const u32 grp = __match_any_sync(eballot, blk);  // one lane per distinct block sets its bit in the bitmap
    if (blk == NONE) {
      atomicOr(&b.sc->overflow, OVF_WINDOW);
    } else {
      atomicOr(b.dense + (size_t)blk * 16 + (e.w ? 0u : 8u) + (ci >> 6), 1ull << (ci & 63));
      if ((int)lane == __ffs(grp) - 1) atomicOr(b.dbits + (blk >> 5), 1u << (blk & 31u));
    }
  }
  bool list_me = false;
  if (!DENSE && is_end && !PENDING) {
    // addHitPoint / addMissPoint (probabilistic_map.cpp:30-54) deferred to the apply pass: a hit endpoint sets its bit in
    // the leaf's HIT mask; a miss endpoint gets exactly the update of a ray cell (max(p + miss, clamp_min), stamp), so
    // it simply joins the touched mask. Either way the leaf is listed for the apply pass.
    unsigned long long* word = reinterpret_cast<unsigned long long*>(e.w ? leaf_touched(g, leaf) : leaf_hit(g, leaf)) + (ci >> 6);
    atomicOr(word, 1ull << (ci & 63));  // result unused. Resolve runs before any ray is marked: nearly every endpoint leaf
    // is touched here for the first time in this scan, so ONE lane per distinct leaf of the warp installs the stamp and
    // lists the leaf (the words themselves need no test and no return value).
    const u32 grp = __match_any_sync(eballot, leaf);
    list_me = (int)lane == __ffs(grp) - 1 && atomicExch(leaf_stamp(g, leaf), p.seq) != p.seq;
  }
  if (!DENSE && !PENDING) list_leaves(list_me, p.fleet ? leaf | (src << 28) : leaf, &b.sc->n_touched, b.touched, p.touched_cap);
  if (mine) {
    const unsigned long long at = s_base + s_warp[warp] + (incl - mine);
    const u32 ray = (u32)(at >> 40);
    const unsigned long long cb = at & CHUNK_FIELD;
    if (cb + chunks > 0xFFFFFFF0ull) {
      atomicOr(&b.sc->overflow, OVF_CHUNKS);
    } else {
      b.rays[ray] = make_int4(e.x, e.y, e.z, (int)(u32)cb);
      if (MODE == 2 && p.fleet) b.ray_src[ray] = (unsigned char)src;
      // every 32-chunk tile whose first chunk lies inside this ray learns its owner (the dense mark kernel needs no tiles)
      const u32 t0 = ((u32)cb + 31u) >> 5, t1 = DENSE ? 0u : ((u32)cb + chunks - 1u) >> 5;
      for (u32 t = t0; t <= t1 && !DENSE; ++t) {
        if (t < p.tile_cap) {
          b.tile_first[t] = ray;
        } else {
          atomicOr(&b.sc->overflow, OVF_TILES);
          break;
        }
      }
    }
  }
  if (MODE != 2 || (u64)(vb + gridDim.x) * blockDim.x >= received) break;
  __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// phase 3: mark ray cells
// ------------------------------------------------------------------------------------------------
// cells [k0, k1) of chunk j. (Chunks of several leaf blocks with the segment flushed as soon as its key changes
// were measured too: 1.6-2x slower — fewer, longer, more divergent work items. profiles/r1_notes.md)
__device__ __forceinline__ void chunk_range(const RayGeom& r, u32 j, u32& k0, u32& k1) {
  k0 = j == 0u ? 0u : r.len0 + 8u * (j - 1u);
  k1 = min(r.m, j == 0u ? r.len0 : k0 + 8u);
}

// One lane = one chunk of one ray.
//   walk   exact integer DDA restarted from the closed form at the chunk's first cell (cell k of RayIterator,
//          probabilistic_map.hpp:162-203, is origin + sign * floor((2*k*|d| + m) / (2*m)) per axis with residual
//          error k*|d| - pos*m, then the reference's own error accumulator). Pure ALU; the (leaf, mask word)
//          segments it crosses — at most 8, never repeating because every coordinate is monotone — are pushed
//          on a per-lane stack in shared memory.
//   flush  all lanes then walk their segment stacks in lock step (convergent): find-or-create the leaf, test
//          the 64-bit touched word and only atomicOr when it adds bits. The thread that turns a word from 0
//          to non-zero stamps the leaf and, if the leaf was not stamped in this scan yet, lists it.
template <typename E>
__device__ __forceinline__ u32 walk_chunk(const ScanParams& p, const RayGeom& r, u32 k0, u32 k1, int& lx0, int& ly0, int& lz0,
                                          unsigned long long (*s_bits)[TPB], unsigned char (*s_key)[TPB]) {
  u32 px, py, pz;
  if (r.m < 32768u) {
    px = (2u * k0 * r.ax + r.m) / (2u * r.m);
    py = (2u * k0 * r.ay + r.m) / (2u * r.m);
    pz = (2u * k0 * r.az + r.m) / (2u * r.m);
  } else {
    px = (u32)((2ull * k0 * r.ax + r.m) / (2ull * r.m));
    py = (u32)((2ull * k0 * r.ay + r.m) / (2ull * r.m));
    pz = (u32)((2ull * k0 * r.az + r.m) / (2ull * r.m));
  }
  E ex = (E)((i64)k0 * r.ax - (i64)px * r.m), ey = (E)((i64)k0 * r.ay - (i64)py * r.m), ez = (E)((i64)k0 * r.az - (i64)pz * r.m);
  int x = r.Ox + r.sx * (int)px, y = r.Oy + r.sy * (int)py, z = r.Oz + r.sz * (int)pz;
  lx0 = x >> 3;
  ly0 = y >> 3;
  lz0 = z >> 3;
  u32 nseg = 0, key = 0xFFFFFFFFu;
  unsigned long long bits = 0;
  const E em = (E)r.m, half = (E)((r.m + 1u) >> 1);  // (e << 1) >= m  <=>  e >= ceil(m / 2)
  const E dax = (E)r.ax, day = (E)r.ay, daz = (E)r.az;
#pragma unroll
  for (u32 c = 0; c < CHUNK; ++c) {
    if (k0 + c < k1) {
      // segment key: mask word (z & 7) + the leaf-parity bit of every axis. Inside a chunk a coordinate crosses at
      // most one leaf boundary, so equal keys <=> same leaf and same word.
      const u32 kc = ((u32)z & 7u) | ((u32)x & 8u) | (((u32)y & 8u) << 1) | (((u32)z & 8u) << 2);
      if (kc != key) {
        if (bits) {
          s_bits[nseg][threadIdx.x] = bits;
          s_key[nseg][threadIdx.x] = (unsigned char)key;
          ++nseg;
        }
        key = kc;
        bits = 0;
      }
      bits |= 1ull << (((u32)x & 7u) | (((u32)y & 7u) << 3));
      ex += dax;
      ey += day;
      ez += daz;
      if (ex >= half) {
        x += r.sx;
        ex -= em;
      }
      if (ey >= half) {
        y += r.sy;
        ey -= em;
      }
      if (ez >= half) {
        z += r.sz;
        ez -= em;
      }
    }
  }
  if (bits) {
    s_bits[nseg][threadIdx.x] = bits;
    s_key[nseg][threadIdx.x] = (unsigned char)key;
    ++nseg;
  }
  return nseg;
}

// leaf (lx,ly,lz) [leaf coordinates, default bits 2/3] in grid G, creating root / inner / leaf as needed
__device__ __forceinline__ u32 mark_leaf(const GridDev& G, u32& inner, bool new_root, int lx, int ly, int lz) {
  if (new_root) {
    const int kx = (lx >> 2) << 5, ky = (ly >> 2) << 5, kz = (lz >> 2) << 5;
    inner = root_find(G, kx, ky, kz);
    if (inner == NONE) inner = root_find_or_insert(G, kx, ky, kz);
  }
  if (inner == NONE) return NONE;
  const u32 ii = ((u32)lx & 3u) | (((u32)ly & 3u) << 2) | (((u32)lz & 3u) << 4);  // getInnerIndex, bonxai.hpp:431-438
  const u32 v = inner_ptr(G, inner)[G.inner_child_off + ii];
  return v >= 2u ? v - 2u : leaf_create_in_inner(G, inner, ii, lx << 3, ly << 3, lz << 3);
}

// OR `bits` into one 64-bit word of a leaf's per-scan mask (touched, or hit). Test first (a stale L1 line only costs a
// redundant atomic): the leaves around the sensor are hit by every ray. A word seen non-zero can never be the leaf's
// first touch, so only writers of an (apparently) empty word need the old value back; the thread that really turns
// a word non-zero stamps the leaf and, if nobody stamped it in this scan yet, appends it to the touched list.
// Returns true for the ONE thread of the scan that must append the leaf to the touched list (list_leaves).
__device__ __forceinline__ bool mark_bits(const GridDev& G, u32 leaf, unsigned long long* word, unsigned long long bits, u32 seq) {
  const unsigned long long cur = *word;
  if ((cur & bits) == bits) return false;
  if (cur != 0ull) {
    atomicOr(word, bits);  // result unused: a fire-and-forget reduction
    return false;
  }
  return atomicOr(word, bits) == 0ull && atomicExch(leaf_stamp(G, leaf), seq) != seq;
}
// Appends `entry` for every lane with mine == true; called by CONVERGED warps. The slots of the warp come from ONE atomic
// on the list counter: tens of thousands of returning atomics per scan on that single address serialise in L2.
// entry = leaf, or leaf | sensor << 28 in a fleet step (every leaf is touched by the rays of ONE sensor then).
__device__ __forceinline__ void list_leaves(bool mine, u32 entry, u32* n_list, u32* list, u32 cap) {
  const u32 m = __ballot_sync(0xffffffffu, mine);
  if (m == 0u) return;
  const u32 lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  u32 at = 0;
  if ((int)lane == leader) at = atomicAdd(n_list, (u32)__popc(m));
  at = __shfl_sync(0xffffffffu, at, leader) + __popc(m & ((1u << lane) - 1u));
  if (mine && at < cap) list[at] = entry;
}

// SHARD: cells whose root this rank does not own are marked in the scratch grid gs (same code, other pools);
// their (leaf, mask) records travel to the owner after the kernel.
template <bool SHARD>
__global__ void __launch_bounds__(TPB, MARK_MIN_BLOCKS) k_mark(GridDev g, GridDev gs, ScanParams p, ScanBuffers b) {
  pdl_enter();
  __shared__ unsigned long long s_bits[CHUNK][TPB];
  __shared__ unsigned char s_key[CHUNK][TPB];
  const unsigned long long rc = b.sc->ray_chunk;
  const u32 n_rays = (u32)(rc >> 40);
  const u32 total = (u32)(rc & CHUNK_FIELD);
  if (b.sc->overflow | *b.poison) return;
  // pipelined insert: the dedupe table was last read by k_resolve; leave it zeroed for the next scan's k_classify
  // (saves a clearing launch per scan). Skipped when the pipeline is frozen, like every other write.
  if (p.clean16) {
    if (SHARD) {  // the receiver's table: only the slots the arrived records could use (table and keys are apart)
      int4 e_;
      bool f_;
      const u32 n16 = (shard_table_mask(p, shard_locate(p, b, NONE, f_, e_)) + 1u) / 4u;
      uint4* tab = reinterpret_cast<uint4*>(b.table);
      uint4* keys = reinterpret_cast<uint4*>(b.keys);
      if (p.packed) {
        for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < 3u * n16; i += gridDim.x * blockDim.x) {
          if (i < n16) {
            tab[i] = make_uint4(0, 0, 0, 0);
          } else {
            keys[i - n16] = make_uint4(0, 0, 0, 0);
          }
        }
      } else {  // 16-byte slots
        for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < 4u * n16; i += gridDim.x * blockDim.x) tab[i] = make_uint4(0, 0, 0, 0);
      }
    } else {
      uint4* tab = reinterpret_cast<uint4*>(b.table);
      for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < p.clean16; i += gridDim.x * blockDim.x) tab[i] = make_uint4(0, 0, 0, 0);
    }
  }
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  for (u32 tile = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); (u64)tile * 32 < total; tile += warps) {
    const u32 c0 = tile * 32;
    const u32 r_first = b.tile_first[tile];
    // which of the next 32 rays start inside this tile? bit j = a ray starts at chunk c0 + j
    const u32 rb = r_first + 1 + lane;
    u32 bit = 0;
    if (rb < n_rays) {
      const u32 base = (u32)b.rays[rb].w;
      if (base - c0 < 32u) bit = 1u << (base - c0);
    }
    const u32 starts = __reduce_or_sync(0xffffffffu, bit);
    const u32 chunk = c0 + lane;
    u32 nseg = 0, tag = 0;
    int lx0 = 0, ly0 = 0, lz0 = 0, sx = 1, sy = 1, sz = 1;
    if (chunk < total) {
      const u32 r = r_first + __popc(starts & ((2u << lane) - 1u));
      const int4 ray = b.rays[r];
      if (SHARD && p.fleet) tag = (u32)b.ray_src[r] << 28;
      const RayGeom rg = ray_geom(p, ray.x, ray.y, ray.z, tag >> 28);
      u32 k0, k1;
      chunk_range(rg, chunk - (u32)ray.w, k0, k1);
      sx = rg.sx;
      sy = rg.sy;
      sz = rg.sz;
      nseg = rg.m < (1u << 29) ? walk_chunk<int>(p, rg, k0, k1, lx0, ly0, lz0, s_bits, s_key)
                               : walk_chunk<i64>(p, rg, k0, k1, lx0, ly0, lz0, s_bits, s_key);
    }
    // ---- flush: all lanes walk their segment stacks in lock step (segments of one leaf are consecutive)
    const u32 nmax = __reduce_max_sync(0xffffffffu, nseg);
    const u32 par0 = ((u32)lx0 & 1u) | (((u32)ly0 & 1u) << 1) | (((u32)lz0 & 1u) << 2);
    u32 cur_q = 0xFFu, leaf = NONE, inner = NONE;
    int rrx = 0, rry = 0, rrz = 0;
    bool have_root = false, own = true;
    for (u32 sgi = 0; sgi < nmax; ++sgi) {
      bool list_own = false, list_scratch = false;
      if (sgi < nseg) {
        const u32 key = s_key[sgi][threadIdx.x];
        const unsigned long long bits = s_bits[sgi][threadIdx.x];
        const u32 q = (key >> 3) ^ par0, w = key & 7u;  // q: which axes have crossed into the neighbouring leaf
        if (q != cur_q) {
          cur_q = q;
          const int lx = lx0 + ((q & 1u) ? sx : 0), ly = ly0 + ((q & 2u) ? sy : 0), lz = lz0 + ((q & 4u) ? sz : 0);
          const int rx = lx >> 2, ry = ly >> 2, rz = lz >> 2;
          const bool new_root = !have_root || rx != rrx || ry != rry || rz != rrz;
          if (new_root) {
            rrx = rx;
            rry = ry;
            rrz = rz;
            have_root = true;
            if (SHARD) own = shard_owner(rx, ry, rz, p.world) == p.rank;
          }
          leaf = (!SHARD || own) ? mark_leaf(g, inner, new_root, lx, ly, lz) : mark_leaf(gs, inner, new_root, lx, ly, lz);
        }
        if (leaf != NONE) {
          if (!SHARD || own) {
            list_own = mark_bits(g, leaf, reinterpret_cast<unsigned long long*>(leaf_touched(g, leaf)) + w, bits, p.seq);
          } else {
            list_scratch = mark_bits(gs, leaf, reinterpret_cast<unsigned long long*>(leaf_touched(gs, leaf)) + w, bits, p.seq);
          }
        }
      }
      list_leaves(list_own, leaf | tag, &b.sc->n_touched, b.touched, p.touched_cap);
      if (SHARD) list_leaves(list_scratch, leaf | tag, &b.sc->n_touched2, b.touched2, p.touched2_cap);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// phase 3, dense flavour: rays marked into the dense window
// ------------------------------------------------------------------------------------------------
// Work item = (group of 32 consecutive rays, stretch s): lane l walks cells [SEG s, SEG s + SEG) of ray 32 g + l with the
// reference's own integer DDA (probabilistic_map.hpp:162-203), restarted ONCE per stretch from the closed form
// cell_k = O + sign * floor((2 k |d| + m) / 2m) (residual error k |d| - pos m). While the cell stays in the same leaf
// block and z slab its bit is ORed into a register; when that (block, slab) key changes the 64-bit word goes to the
// dense window as ONE fire-and-forget reduction (red.or): the address is pure arithmetic on the block coordinates — no
// root, inner-node or leaf look-up, nothing is created here — and nothing is read back, so the walk never waits for
// memory. Which blocks were touched is recorded the same way, one bit per block in a bitmap the apply pass scans.
// Compared with the leaf-resident flavour (8-cell chunks, one closed-form restart + segment stack + leaf look-ups +
// test-and-set round trips per chunk) a visited cell costs a fraction of the instructions and no round trip.
#ifndef MARKD_MIN_BLOCKS
#define MARKD_MIN_BLOCKS 6
#endif
#ifndef APPLY_MIN_BLOCKS
#define APPLY_MIN_BLOCKS 4
#endif
#ifndef MARK_SEG
#define MARK_SEG 32
#endif
constexpr u32 SEG = MARK_SEG;  // cells per lane and work item

// One finished (block, z slab) word of a lane. The word's address is pure arithmetic; the test-before-set that keeps
// the hot words around the sensor from being hammered needs the word's current value, and waiting for that load is what
// the walk must never do: the load is issued when a word is finished and its value is only looked at when the NEXT word
// of the lane is finished, ~100 instructions later. A stale value costs a redundant red.or, never a missed bit.
struct PendingWord {
  unsigned long long* word;  // nullptr: nothing pending
  unsigned long long bits, cur;
  u32 blk;
};
__device__ __forceinline__ void commit_word(const ScanBuffers& b, const PendingWord& pw) {
  if (pw.word == nullptr || (pw.cur & pw.bits) == pw.bits) return;
  atomicOr(pw.word, pw.bits);  // result unused: red.or
  // whoever turns a word non-zero saw it zero: the block's bit in the bitmap of touched blocks is set at least once
  if (pw.cur == 0ull) atomicOr(b.dbits + (pw.blk >> 5), 1u << (pw.blk & 31u));
}
// key = bx | by << 10 | z slab << 20 | bz << 23, block coordinates relative to the window corner
__device__ __forceinline__ void issue_word(const ScanParams& p, const ScanBuffers& b, u32 key, unsigned long long bits, PendingWord& pw) {
  const u32 bx = key & 1023u, by = (key >> 10) & 1023u, w = (key >> 20) & 7u, bz = key >> 23;
  pw.word = nullptr;
  if (bx >= p.D || by >= p.D || bz >= p.D) {
    atomicOr(&b.sc->overflow, OVF_WINDOW);
    return;
  }
  pw.blk = (bz * p.D + by) * p.D + bx;
  pw.word = b.dense + (size_t)pw.blk * 16 + w;
  pw.bits = bits;
  pw.cur = *pw.word;  // in flight until the next commit
}

__global__ void __launch_bounds__(TPB, MARKD_MIN_BLOCKS) k_mark_dense(ScanParams p, ScanBuffers b, u32 smax) {
  pdl_enter();
  const u32 n_rays = (u32)(b.sc->ray_chunk >> 40);
  if (b.sc->overflow | *b.poison) return;
  if (p.clean16) {  // pipelined insert: leave the dedupe table zeroed for the next scan's k_classify (its last reader is done)
    uint4* tab = reinterpret_cast<uint4*>(b.table);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < p.clean16; i += gridDim.x * blockDim.x) tab[i] = make_uint4(0, 0, 0, 0);
  }
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  const u32 groups = (n_rays + 31u) >> 5;
  const u32 tiles = groups * smax;  // < 2^24 / 32 * 2^7
  for (u32 tile = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); tile < tiles; tile += warps) {
    // group-major order: the warps that run at the same time work at all distances from the sensor (the words next to
    // the sensor are shared by every ray: few warps should be there at once)
    const u32 grp = tile / smax, st = tile - grp * smax;
    const u32 r = grp * 32u + lane;
    const u32 k0 = st * SEG;
    u32 m = 0, ax = 0, ay = 0, az = 0;
    int sx = 1, sy = 1, sz = 1;
    if (r < n_rays) {
      const int4 ray = b.rays[r];
      const int dx = ray.x - p.Ox, dy = ray.y - p.Oy, dz = ray.z - p.Oz;  // the window bounds |d| far below 2^15
      ax = (u32)abs(dx);
      ay = (u32)abs(dy);
      az = (u32)abs(dz);
      sx = dx < 0 ? -1 : 1;
      sy = dy < 0 ? -1 : 1;
      sz = dz < 0 ? -1 : 1;
      m = max(max(ax, ay), az);  // probabilistic_map.hpp:180: the ray has exactly m cells, the end cell is excluded
    }
    const u32 k1 = min(m, k0 + SEG);
    const u32 steps = __reduce_max_sync(0xffffffffu, k1 > k0 ? k1 - k0 : 0u);
    if (steps == 0u) continue;  // every ray of the group ends before this stretch
    int x = p.Ox, y = p.Oy, z = p.Oz;
    int ex = 0, ey = 0, ez = 0;
    if (k0 && k0 < m) {
      const u32 m2 = 2u * m;
      const u32 px = (2u * k0 * ax + m) / m2, py = (2u * k0 * ay + m) / m2, pz = (2u * k0 * az + m) / m2;
      ex = (int)(k0 * ax) - (int)(px * m);
      ey = (int)(k0 * ay) - (int)(py * m);
      ez = (int)(k0 * az) - (int)(pz * m);
      x += sx * (int)px;
      y += sy * (int)py;
      z += sz * (int)pz;
    }
    const int em = (int)m, half = (int)((m + 1u) >> 1);  // (e << 1) >= m  <=>  e >= ceil(m / 2)
    const int dax = (int)ax, day = (int)ay, daz = (int)az;
    u32 key = NONE;
    unsigned long long bits = 0ull;
    PendingWord pw;
    pw.word = nullptr;
    pw.bits = pw.cur = 0ull;
    pw.blk = 0u;
    for (u32 c = 0; c < steps; ++c) {
      if (k0 + c < k1) {
        const u32 kc = ((u32)((x >> 3) - p.W0x) & 1023u) | (((u32)((y >> 3) - p.W0y) & 1023u) << 10) | (((u32)z & 7u) << 20) |
                       ((u32)((z >> 3) - p.W0z) << 23);
        if (kc != key) {
          if (bits) {
            commit_word(b, pw);
            issue_word(p, b, key, bits, pw);
          }
          key = kc;
          bits = 0ull;
        }
        bits |= 1ull << (((u32)x & 7u) | (((u32)y & 7u) << 3));
        ex += dax;
        ey += day;
        ez += daz;
        if (ex >= half) {
          x += sx;
          ex -= em;
        }
        if (ey >= half) {
          y += sy;
          ey -= em;
        }
        if (ez >= half) {
          z += sz;
          ez -= em;
        }
      }
    }
    commit_word(b, pw);
    if (bits) {
      issue_word(p, b, key, bits, pw);
      commit_word(b, pw);
    }
  }
}

// phase 3b, dense flavour: the bitmap of touched blocks becomes a list (and is cleared for the next scan). One thread
// per bitmap word, one atomic per warp.
__global__ void __launch_bounds__(TPB) k_list_blocks(ScanParams p, ScanBuffers b, u32 nwords) {
  pdl_enter();
  if (b.sc->overflow | *b.poison) return;
  const u32 lane = threadIdx.x & 31;
  for (u32 base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < nwords; base += gridDim.x * blockDim.x) {
    const u32 i = base + lane;
    u32 w = i < nwords ? b.dbits[i] : 0u;
    const u32 cnt = __popc(w);
    u32 incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o) incl += v;
    }
    const u32 total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0u) continue;
    u32 at = 0;
    if (lane == 31) at = atomicAdd(&b.sc->n_touched, total);
    at = __shfl_sync(0xffffffffu, at, 31) + incl - cnt;
    if (w) b.dbits[i] = 0u;
    while (w) {
      const u32 bit = __ffs(w) - 1u;
      w &= w - 1u;
      if (at < p.dlist_cap) b.touched[at] = i * 32u + bit;
      ++at;
    }
  }
}

// where the leaf of a block probably is: a direct-mapped hint table indexed by a hash of the block's absolute
// coordinates (4 B per entry: leaf index + 1). A hint is verified against the leaf's own header, so stale or colliding
// entries only cost the regular look-up.
constexpr u32 HINT_BITS = 20;
__device__ __forceinline__ u32 hint_slot(int lx, int ly, int lz) {
  u32 h = (u32)lx * 0x9E3779B1u ^ (u32)ly * 0x85EBCA77u ^ (u32)lz * 0xC2B2AE3Du;
  h ^= h >> 15;
  return h & ((1u << HINT_BITS) - 1u);
}

// phase 4, dense flavour: one warp per listed block. Its leaf is found through the hint table — or by the regular walk,
// or created: the only place of a dense scan where the map grows — then the block's line of marks is applied to the
// leaf's cells exactly like k_apply_leaves does, and the line is cleared. A block whose leaf cannot be created (pool
// exhausted: error bit set) keeps its marks; the host grows the pool and runs this kernel again over the same list:
// blocks that were applied have empty lines and are skipped, so every cell is still updated exactly once ("resume",
// not "repeat").
__global__ void __launch_bounds__(TPB, APPLY_MIN_BLOCKS) k_apply_dense(GridDev g, ScanParams p, ScanBuffers b, u32 resume) {
  pdl_enter();
  // frozen pipeline (an earlier scan ran short) or a scan that overflowed its scratch: nothing is applied
  const bool skip = ((resume ? 0u : g.ctr->error) | b.sc->overflow) != 0u;
  const u32 n = skip ? 0u : min(b.sc->n_touched, p.dlist_cap);
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  u32 changed = 0;
  u32 t = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
  u32 blk_n = t < n ? b.touched[t] : NONE;
  for (; t < n; t += warps) {
    const u32 blk = blk_n;
    blk_n = t + warps < n ? b.touched[t + warps] : NONE;
    const u32 bx = blk % p.D, by = (blk / p.D) % p.D, bz = blk / (p.D * p.D);
    const int lx = p.W0x + (int)bx, ly = p.W0y + (int)by, lz = p.W0z + (int)bz;
    u32* line = reinterpret_cast<u32*>(b.dense + (size_t)blk * 16);
    // lane j < 16 holds 32-bit half j of the touched and of the hit mask = the bits of cell row j
    u32 th = 0, hh = 0, ah = 0;
    if (lane < 16) {
      th = line[lane];
      hh = line[16 + lane];
    }
    // the leaf: hint first (verified against the leaf's header, which shares a line with the ON mask)
    u32* hint = b.dhint + hint_slot(lx, ly, lz);
    u32 leaf = *hint - 1u;
    int4 hdr = make_int4(0, 0, 0, 0);
    if (leaf < g.leaf_cap) {
      const unsigned char* lp0 = leaf_ptr(g, leaf);
      hdr = *reinterpret_cast<const int4*>(lp0);
      if (lane < 16) ah = reinterpret_cast<const u32*>(lp0 + g.off_active)[lane];
    }
    if (__ballot_sync(0xffffffffu, (th | hh) != 0u) == 0u) continue;  // applied by an earlier attempt
    if (!(leaf < g.leaf_cap && hdr.x == (lx << 3) && hdr.y == (ly << 3) && hdr.z == (lz << 3) && (hdr.w & 1))) {
      leaf = NONE;
      if (lane == 0) {
        leaf = leaf_find_or_create(g, lx << 3, ly << 3, lz << 3);
        if (leaf != NONE) *hint = leaf + 1u;
      }
      leaf = __shfl_sync(0xffffffffu, leaf, 0);
      if (leaf == NONE) continue;  // pool exhausted: the marks stay, the host grows the pool and resumes
      ah = 0;
      if (lane < 16) ah = reinterpret_cast<const u32*>(leaf_ptr(g, leaf) + g.off_active)[lane];
    }
    unsigned char* lp = leaf_ptr(g, leaf);
    // lane owns cells it*32 + lane, it = 0..15 (coalesced 128-B rows); bit `it` of mine/on/hit = that cell touched/ON/hit
    u32 mine = 0, on = 0, hit = 0;
#pragma unroll
    for (u32 j = 0; j < 16; ++j) {
      const u32 t32 = __shfl_sync(0xffffffffu, th, j), a32 = __shfl_sync(0xffffffffu, ah, j), h32 = __shfl_sync(0xffffffffu, hh, j);
      mine |= ((t32 >> lane) & 1u) << j;
      on |= ((a32 >> lane) & 1u) << j;
      hit |= ((h32 >> lane) & 1u) << j;
    }
    mine |= hit;
    u32* cells = reinterpret_cast<u32*>(lp + g.off_cells);
#pragma unroll
    for (u32 half = 0; half < 2; ++half) {
      u32 word[8];
#pragma unroll
      for (u32 it = 0; it < 8; ++it) word[it] = ((mine & on) >> (half * 8 + it)) & 1u ? cells[(half * 8 + it) * 32 + lane] : 0u;
#pragma unroll
      for (u32 it = 0; it < 8; ++it) {
        const u32 r = half * 8 + it;
        if ((hit >> r) & 1u) {
          const i32 prob = min(((i32)word[it] >> 4) + p.hit, p.cmax);
          cells[r * 32 + lane] = ((u32)prob << 4) | p.c;
          ++changed;
        } else if (((mine >> r) & 1u) && (word[it] & 0xFu) != p.c) {
          const i32 prob = max(((i32)word[it] >> 4) + p.miss, p.cmin);
          cells[r * 32 + lane] = ((u32)prob << 4) | p.c;
          ++changed;
        }
      }
    }
    if (lane < 16) reinterpret_cast<u32*>(lp + g.off_active)[lane] = ah | th | hh;
    line[lane] = 0u;  // 32 lanes x 4 B: the whole line of marks
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if (lane == 0 && changed) atomicAdd(&b.sc->n_changed, changed);
  // The LAST block to get here leaves the grid counters next to the scan's (the host reads both with one copy) and,
  // pipelined, publishes the scan's record to the host ring (zero copy). A pool that ran out during this kernel has set
  // the grid's error bits: every later scan in the queue skips itself until the host has grown the pool and resumed
  // this scan.
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&g.ctr->done_blocks, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  g.ctr->done_blocks = 0;
  __threadfence();
  {
    const volatile u32* src = reinterpret_cast<const volatile u32*>(g.ctr);
    u32* dst = reinterpret_cast<u32*>(&b.sc->gc);
    for (u32 k = 0; k < sizeof(GridCounters) / 4; ++k) dst[k] = src[k];
  }
  if (p.async_id == NONE || resume) return;
  const volatile ScanCounters* sc = b.sc;
  u32 err = g.ctr->error;
  if (sc->overflow && !err) {
    err = ERR_SCAN;
    atomicOr(&g.ctr->error, ERR_SCAN);
  }
  if (err && g.ctr->failed_id == NONE) g.ctr->failed_ovf = sc->overflow;
  if (err && g.ctr->failed_id == NONE) g.ctr->failed_id = p.async_id;
  AsyncRecord* r = b.ring + (p.async_id & (RING_SIZE - 1u));
  r->error = err;
  r->n_leaves = g.ctr->n_leaves;
  r->n_inner = g.ctr->n_inner;
  r->n_roots = g.ctr->n_roots;
  r->n_endpoints = sc->n_endpoints;
  r->n_changed = sc->n_changed;
  r->n_touched = sc->n_touched;
  r->n_points = p.n;
  r->n_dropped = sc->n_dropped;
  r->leaf_fill = 0;
  r->sum_m = sc->sum_m;
  r->ray_chunk = sc->ray_chunk;
  __threadfence_system();
  r->id = p.async_id;
  if (!err) {  // healthy: hand zeroed counters to the next scan (a failed scan's counters stay for the host)
    uint4* z = reinterpret_cast<uint4*>(b.sc);
    for (u32 k = 0; k < sizeof(ScanCounters) / 16; ++k) z[k] = make_uint4(0, 0, 0, 0);
  }
}

// retry path only: a failed attempt leaves touched bits behind; the list of that attempt says where
__global__ void __launch_bounds__(TPB) k_clear_touched(GridDev g, ScanBuffers b, u32 n) {
  pdl_enter();
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  for (u32 t = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); t < n; t += warps) {
    if (lane < 16) reinterpret_cast<unsigned long long*>(leaf_touched(g, b.touched[t] & LEAF_MASK))[lane] = 0ull;  // touched[8] + hit[8] are contiguous
  }
}

// ------------------------------------------------------------------------------------------------
// phase 4 / 5: apply
// ------------------------------------------------------------------------------------------------
// the reduced flags as seen by the apply pass (block-uniform; ends with a __syncthreads in the peer-memory flavour)
// fill = the largest number of leaf records any rank sent to any owner in this scan (equal on every rank: the host uses
// it to grow the leaf inboxes collectively, at a drain, BEFORE they overflow)
__device__ __forceinline__ void read_gate(const ScanParams& p, const ScanBuffers& b, u32& pool, u32& ovf, u32& fill) {
  __shared__ u32 s_gate[3];
  pool = 0;
  ovf = 0;
  fill = 0;
  if (!b.gate) return;
  if (!b.my_flags) {  // all-reduced by the caller or by NCCL
    pool = b.gate[0];
    ovf = b.gate[1];
    fill = b.gate[2];
    return;
  }
  const u32* base = b.gate + (p.xseq2 & 1u) * (MAX_PEERS * 4);
  if (threadIdx.x < 3) s_gate[threadIdx.x] = 0;
  wait_arrivals(base + 3, 4, p.world, p.xseq2, const_cast<u32*>(b.poison));
  if (threadIdx.x < p.world) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(base) + threadIdx.x);
    if (v.x) atomicOr(&s_gate[0], v.x);
    if (v.y) atomicOr(&s_gate[1], v.y);
    if (v.z) atomicMax(&s_gate[2], v.z);
  }
  __syncthreads();
  pool = s_gate[0];
  ovf = s_gate[1];
  fill = s_gate[2];
}

// ---- bulk-copy engine (TMA, 1-D): global -> shared, completion counted on an mbarrier
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// One warp per listed leaf: hit endpoints (addHitPoint, probabilistic_map.cpp:30-41) and the union of all rays + miss
// endpoints (clearPoint / addMissPoint, :43-54,81-89). Hit endpoints are never stale (resolve filtered them) and win
// over ray cells, like the reference where they are stamped before any ray is cast.
//
// TMA = true (default grid layout: 8^3 leaves of 4-byte cells): the chain list entry -> masks -> cell rows of a leaf is
// taken off the warp's critical path by the bulk-copy engine. Every warp keeps three leaves in flight: the 192 bytes of
// masks (ON, touched, hit: contiguous in the leaf) of leaf i + 2 and the needed 128-byte cell rows of leaf i + 1 land in
// shared memory (cp.async.bulk, completion counted on mbarriers) while leaf i is updated. Only rows that hold a touched
// cell which is already ON are fetched (the same bytes as the register flavour); results go back with plain coalesced
// stores.
template <bool TMA>
__global__ void __launch_bounds__(TPB, APPLY_MIN_BLOCKS) k_apply_leaves(GridDev g, ScanParams p, ScanBuffers b) {
  pdl_enter();
  // last kernel of the scan: the host reads counters + grid counters with one copy
  u32 gate_pool, gate_ovf, gate_fill;
  read_gate(p, b, gate_pool, gate_ovf, gate_fill);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    b.sc->gc = *g.ctr;
    b.sc->gate_pool = gate_pool;
    b.sc->gate_ovf = gate_ovf;
    b.sc->gate_fill = gate_fill;
  }
  const bool skip = (g.ctr->error | b.sc->overflow | gate_pool | gate_ovf) != 0u;
  const u32 n = skip ? 0u : min(b.sc->n_touched, p.touched_cap);
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  u32 changed = 0;
  if (TMA) {
    // shared memory per warp: masks of 3 leaves (192 B each), cell rows of 2 leaves (16 x 128 B each), 5 mbarriers
    __shared__ __align__(128) unsigned char s_rows[TPB / 32][2][2048];
    __shared__ __align__(16) u32 s_masks[TPB / 32][3][48];
    __shared__ __align__(8) unsigned long long s_bar[TPB / 32][5];
    const u32 wib = threadIdx.x >> 5;
    unsigned long long* mbar = s_bar[wib];       // [0..2]: masks of leaf i % 3
    unsigned long long* rbar = s_bar[wib] + 3;   // [0..1]: rows of leaf i % 2
    if (lane == 0) {
      for (int k = 0; k < 5; ++k) mbar_init(&s_bar[wib][k], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const u32 t0 = blockIdx.x * (TPB / 32) + wib;
    const u32 mine_n = t0 < n ? (n - t0 + warps - 1u) / warps : 0u;  // leaves of this warp: t0 + i * warps
    auto entry_of = [&](u32 i) { return i < mine_n ? b.touched[t0 + i * warps] : NONE; };
    auto issue_masks = [&](u32 i, u32 entry) {  // 192 B: ON @64, touched @128, hit @192 of the leaf
      if (lane == 0) {
        const unsigned char* lp = leaf_ptr(g, p.fleet ? entry & LEAF_MASK : entry);
        mbar_expect_tx(&mbar[i % 3u], 192u);
        bulk_g2s(s_masks[wib][i % 3u], lp + 64, 192u, &mbar[i % 3u]);
      }
    };
    auto issue_rows = [&](u32 i, u32 entry) {  // lane r < 16: row r if it holds a touched cell that is ON
      const u32* mk = s_masks[wib][i % 3u];
      const bool need = lane < 16 && ((mk[16 + lane] | mk[32 + lane]) & mk[lane]) != 0u;
      const u32 rows = __ballot_sync(0xffffffffu, need);
      if (lane == 0) mbar_expect_tx(&rbar[i & 1u], 128u * (u32)__popc(rows));
      __syncwarp();
      if (need) {
        const unsigned char* lp = leaf_ptr(g, p.fleet ? entry & LEAF_MASK : entry);
        bulk_g2s(s_rows[wib][i & 1u] + lane * 128u, lp + 256 + lane * 128u, 128u, &rbar[i & 1u]);
      }
    };
    u32 e0 = entry_of(0), e1 = entry_of(1), e2 = entry_of(2);
    if (mine_n > 0) issue_masks(0, e0);
    if (mine_n > 1) issue_masks(1, e1);
    if (mine_n > 0) {
      mbar_wait(&mbar[0], 0);
      issue_rows(0, e0);
    }
    for (u32 i = 0; i < mine_n; ++i) {
      const u32 e3 = entry_of(i + 3);  // list entries run three leaves ahead
      if (i + 2 < mine_n) issue_masks(i + 2, e2);
      if (i + 1 < mine_n) {
        mbar_wait(&mbar[(i + 1) % 3u], ((i + 1) / 3u) & 1u);
        issue_rows(i + 1, e1);
      }
      mbar_wait(&rbar[i & 1u], (i >> 1) & 1u);
      const u32 leaf = p.fleet ? e0 & LEAF_MASK : e0;
      const u32 c = p.fleet ? scan_c(p, e0 >> 28) : p.c;
      unsigned char* lp = leaf_ptr(g, leaf);
      const u32* mk = s_masks[wib][i % 3u];
      u32 th = 0, hh = 0, ah = 0;
      if (lane < 16) {
        ah = mk[lane];
        th = mk[16 + lane];
        hh = mk[32 + lane];
      }
      u32 mine = 0, on = 0, hit = 0;
#pragma unroll
      for (u32 j = 0; j < 16; ++j) {
        const u32 t32 = __shfl_sync(0xffffffffu, th, j), a32 = __shfl_sync(0xffffffffu, ah, j), h32 = __shfl_sync(0xffffffffu, hh, j);
        mine |= ((t32 >> lane) & 1u) << j;
        on |= ((a32 >> lane) & 1u) << j;
        hit |= ((h32 >> lane) & 1u) << j;
      }
      mine |= hit;
      const u32* srow = reinterpret_cast<const u32*>(s_rows[wib][i & 1u]);
      u32* cells = reinterpret_cast<u32*>(lp + g.off_cells);
#pragma unroll
      for (u32 r = 0; r < 16; ++r) {
        if (!((mine >> r) & 1u)) continue;
        const u32 word = ((on >> r) & 1u) ? srow[r * 32 + lane] : 0u;
        if ((hit >> r) & 1u) {
          const i32 prob = min(((i32)word >> 4) + p.hit, p.cmax);
          cells[r * 32 + lane] = ((u32)prob << 4) | c;
          ++changed;
        } else if ((word & 0xFu) != c) {
          const i32 prob = max(((i32)word >> 4) + p.miss, p.cmin);
          cells[r * 32 + lane] = ((u32)prob << 4) | c;
          ++changed;
        }
      }
      if (lane < 16 && (th | hh)) {
        reinterpret_cast<u32*>(lp + g.off_active)[lane] = ah | th | hh;
        reinterpret_cast<u32*>(lp + g.off_touched)[lane] = 0u;
        reinterpret_cast<u32*>(lp + g.off_hit)[lane] = 0u;
      }
      __syncwarp();  // every lane is done with the buffers of leaf i before the engine may overwrite them
      e0 = e1;
      e1 = e2;
      e2 = e3;
    }
  }
  // Register flavour (any grid layout). Kept at 32 registers so that all blocks are resident in ONE wave: the kernel is a
  // chain of dependent loads per leaf (list entry -> masks -> cells), so what hides the latency is the number of leaves
  // in flight, not work per thread. Only the list entry of the warp's next leaf is prefetched.
  u32 t = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
  u32 leaf_n = t < n && !TMA ? b.touched[t] : NONE;
  for (; !TMA && t < n; t += warps) {
    // fleet step: the entry also names the sensor whose rays touched the leaf, i.e. which update id stamps its cells
    const u32 leaf = p.fleet ? leaf_n & LEAF_MASK : leaf_n;
    const u32 c = p.fleet ? scan_c(p, leaf_n >> 28) : p.c;
    leaf_n = t + warps < n ? b.touched[t + warps] : NONE;
    unsigned char* lp = leaf_ptr(g, leaf);
    // lane j < 16 holds 32-bit half j of the three masks = the bits of cell row j (cells 32 j .. 32 j + 31)
    u32 th = 0, hh = 0, ah = 0;
    if (lane < 16) {
      th = reinterpret_cast<const u32*>(lp + g.off_touched)[lane];
      hh = reinterpret_cast<const u32*>(lp + g.off_hit)[lane];
      ah = reinterpret_cast<const u32*>(lp + g.off_active)[lane];
    }
    // lane owns cells it*32 + lane, it = 0..15 (coalesced 128-B rows); bit `it` of mine/on/hit = that cell touched/ON/hit
    u32 mine = 0, on = 0, hit = 0;
#pragma unroll
    for (u32 j = 0; j < 16; ++j) {
      const u32 t32 = __shfl_sync(0xffffffffu, th, j), a32 = __shfl_sync(0xffffffffu, ah, j), h32 = __shfl_sync(0xffffffffu, hh, j);
      mine |= ((t32 >> lane) & 1u) << j;
      on |= ((a32 >> lane) & 1u) << j;
      hit |= ((h32 >> lane) & 1u) << j;
    }
    mine |= hit;
    u32* cells = reinterpret_cast<u32*>(lp + g.off_cells);
#pragma unroll
    for (u32 half = 0; half < 2; ++half) {
      // the 8 loads of a half leaf are issued before the first use
      u32 word[8];
#pragma unroll
      for (u32 it = 0; it < 8; ++it) word[it] = ((mine & on) >> (half * 8 + it)) & 1u ? cells[(half * 8 + it) * 32 + lane] : 0u;
#pragma unroll
      for (u32 it = 0; it < 8; ++it) {
        const u32 r = half * 8 + it;
        if ((hit >> r) & 1u) {
          const i32 prob = min(((i32)word[it] >> 4) + p.hit, p.cmax);
          cells[r * 32 + lane] = ((u32)prob << 4) | c;
          ++changed;
        } else if (((mine >> r) & 1u) && (word[it] & 0xFu) != c) {
          const i32 prob = max(((i32)word[it] >> 4) + p.miss, p.cmin);
          cells[r * 32 + lane] = ((u32)prob << 4) | c;
          ++changed;
        }
      }
    }
    if (lane < 16 && (th | hh)) {
      reinterpret_cast<u32*>(lp + g.off_active)[lane] = ah | th | hh;
      reinterpret_cast<u32*>(lp + g.off_touched)[lane] = 0u;
      reinterpret_cast<u32*>(lp + g.off_hit)[lane] = 0u;
    }
  }
  for (int o = 16; o; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
  if (lane == 0 && changed) atomicAdd(&b.sc->n_changed, changed);
  if (p.async_id == NONE) return;
  // pipelined insert: the LAST block to get here publishes the scan's record to the host ring (zero copy) and
  // makes a failure sticky, so that every later scan in the queue skips itself until the host has recovered
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&g.ctr->done_blocks, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  g.ctr->done_blocks = 0;
  __threadfence();
  const volatile ScanCounters* sc = b.sc;
  u32 err = g.ctr->error;
  if ((sc->overflow | gate_pool | gate_ovf) && !err) {  // this rank, or (sharded) any rank, ran short
    err = ERR_SCAN;
    atomicOr(&g.ctr->error, ERR_SCAN);
  }
  if (err && g.ctr->failed_id == NONE) g.ctr->failed_ovf = gate_ovf | sc->overflow;
  if (err && g.ctr->failed_id == NONE) g.ctr->failed_id = p.async_id;
  AsyncRecord* r = b.ring + (p.async_id & (RING_SIZE - 1u));
  r->error = err;
  r->n_leaves = g.ctr->n_leaves;
  r->n_inner = g.ctr->n_inner;
  r->n_roots = g.ctr->n_roots;
  r->n_endpoints = sc->n_endpoints;
  r->n_changed = sc->n_changed;
  r->n_touched = sc->n_touched;
  r->n_points = p.n;
  r->n_dropped = sc->n_dropped;
  r->leaf_fill = gate_fill;
  r->sum_m = sc->sum_m;
  r->ray_chunk = sc->ray_chunk;
  __threadfence_system();
  r->id = p.async_id;
  if (!err) {  // healthy: hand zeroed counters to the next scan (a failed scan's counters stay for the host)
    uint4* z = reinterpret_cast<uint4*>(b.sc);
    for (u32 k = 0; k < sizeof(ScanCounters) / 16; ++k) z[k] = make_uint4(0, 0, 0, 0);
  }
}

// pipelined insert: clears the scan counters + dedupe table like the memset of the synchronous path, unless the
// pipeline is frozen (the failed scan's counters and touched list must survive until the host has seen them)
__global__ void __launch_bounds__(TPB) k_begin_scan(ScanBuffers b, uint4* base, u32 n16) {
  pdl_enter();
  if (*b.poison) return;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) base[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// sharded map: staging kernels around the two exchanges
// ------------------------------------------------------------------------------------------------
// exchange 1, sender: every locally winning endpoint goes to the rank that owns its root, straight into block [rank] of
// the owner's inbox (peer memory) or of the caller's send buffer: record = {x, y, z, global point index << 1 | type};
// slot 0 of a block carries the count. Slots are handed out per owner by a warp-aggregated atomic on a LOCAL counter.
__global__ void __launch_bounds__(TPB) k_shard_bucket(ScanParams p, ScanBuffers b, u32 index_base) {
  pdl_enter();
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 lane = threadIdx.x & 31;
  const u32 cap = p.rec_cap;
  bool win = false;
  int4 e = make_int4(0, 0, 0, 0);
  u32 o = 0;
  if (i < p.n && !*b.poison) {
    const u32 slot = b.slot_of[i];
    win = slot != NONE && b.table[slot] == (p.packed ? ~i : i + 1u);  // not dropped, and the lowest local index of its voxel
    if (win) {
      e = b.ep[i];
      o = shard_owner(e.x >> 5, e.y >> 5, e.z >> 5, p.world);
    }
  }
  // slots: warp-aggregated into shared-memory counters, then ONE global atomic per owner and block (the per-owner counters
  // are a handful of addresses: tens of thousands of returning atomics on them serialise in L2)
  __shared__ u32 s_cnt[MAX_PEERS], s_base[MAX_PEERS];
  if (threadIdx.x < MAX_PEERS) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const u32 act = __ballot_sync(0xffffffffu, win);
  u32 local = 0;
  if (win) {
    const u32 peers = __match_any_sync(act, o);
    const int leader = __ffs(peers) - 1;
    u32 base = 0;
    if ((int)lane == leader) base = atomicAdd(&s_cnt[o], (u32)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    local = base + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (threadIdx.x < p.world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&b.sc->cnt1[threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
  if (win) {
    const u32 at = s_base[o] + local + 1u;
    if (at < cap) {
      (b.px->rec[o] + (size_t)p.par * p.world * cap)[at] = make_int4(e.x, e.y, e.z, (int)(((index_base + i) << 1) | (u32)e.w));
    } else {
      atomicOr(&b.sc->overflow, OVF_RECORDS);
    }
  }
  publish_blocks(b.px, false, b.sc->cnt1, &b.sc->done1, MBOX_FLAG1 + p.par * MAX_PEERS, p.rank, p.xseq1, p.world, cap, (size_t)p.par * p.world * cap);
}

// exchange 1, receiver: lowest global index per endpoint voxel over the records of all ranks. Pipelined path: also
// zeroes the sender-side dedupe table of this scan (its last reader, k_shard_bucket, is done) for the next scan.
__global__ void __launch_bounds__(TPB) k_shard_dedupe(ScanParams p, ScanBuffers b, u32 count, uint4* clean, u32 clean16) {
  pdl_enter();
  // peer-memory exchange: the records of every rank must have arrived (the wait happens even when the pipeline is
  // frozen, so that no rank ever runs ahead of an exchange point)
  wait_arrivals(b.my_flags ? b.my_flags + MBOX_FLAG1 + p.par * MAX_PEERS : nullptr, 1, p.world, p.xseq1, const_cast<u32*>(b.poison));
  if (*b.poison) return;
  const u32 stride = gridDim.x * blockDim.x;
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  for (u32 k = i; k < clean16; k += stride) clean[k] = make_uint4(0, 0, 0, 0);
  // launched for the expected number of records; whole warps loop over the rest (count = the worst case)
  for (;;) {
    int4 e;
    bool found;
    const u32 received = shard_locate(p, b, i, found, e);
    if (found && p.packed) {
      const u32 mask = shard_table_mask(p, received);
      const unsigned long long key = pack_key(e);
      u32 slot = (u32)hash3(e.x, e.y, e.z) & mask;
      for (;;) {
        unsigned long long k = b.keys[slot];
        if (k == 0ull) k = atomicCAS(&b.keys[slot], 0ull, key);
        if (k == 0ull || k == key) break;
        slot = (slot + 1) & mask;
      }
      atomicMax(&b.table[slot], ~(u32)e.w);
      b.slot_of[i] = slot;
    } else if (found) {
      // any coordinates (max_range = inf, voxels beyond +-2^20): 16-byte slots {x, y, z, best}, claimed whole by a
      // 128-bit CAS; best = ~w is never 0, so an all-zero slot is empty
      const u32 mask = shard_table_mask(p, received);
      int4* tab = reinterpret_cast<int4*>(b.table);
      const int4 mine = make_int4(e.x, e.y, e.z, (int)~(u32)e.w);
      u32 slot = (u32)hash3(e.x, e.y, e.z) & mask;
      for (;;) {
        int4 cur = __ldcg(&tab[slot]);
        if (cur.w == 0) {
          cur = cas128(&tab[slot], make_int4(0, 0, 0, 0), mine);
          if (cur.w == 0) break;  // claimed, with this record as the best so far
        }
        if (cur.x == e.x && cur.y == e.y && cur.z == e.z) {
          atomicMax(reinterpret_cast<u32*>(&tab[slot].w), ~(u32)e.w);
          break;
        }
        slot = (slot + 1) & mask;
      }
      b.slot_of[i] = slot;
    }
    i += stride;
    if ((i & ~31u) >= min(received, count)) break;
  }
}

// exchange 2, sender: eight lanes per scratch leaf touched in this scan -> {leaf origin, 512-bit mask} into block [rank]
// of the inbox of the rank that owns its root; the scratch mask is cleared for the next scan. Record = 5 x int4 (80 B).
__global__ void __launch_bounds__(TPB) k_shard_emit(GridDev gs, ScanParams p, ScanBuffers b) {
  pdl_enter();
  const u32 n = min(b.sc->n_touched2, p.touched2_cap);
  const u32 cap = p.leaf_cap2;
  const u32 lane = threadIdx.x & 31, sub = lane & 7u, first = lane & 24u;
  constexpr u32 PER_BLOCK = TPB / 8;  // leaves per block and round
  __shared__ u32 s_cnt[MAX_PEERS], s_base[MAX_PEERS];
  for (u32 tb = blockIdx.x * PER_BLOCK; tb < n; tb += gridDim.x * PER_BLOCK) {  // block-uniform trip count
    if (threadIdx.x < MAX_PEERS) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const u32 t = tb + (threadIdx.x >> 3);
    const bool valid = t < n;
    u32 local = 0, o = 0;
    int4 hdr = make_int4(0, 0, 0, 0);
    unsigned long long m = 0;
    if (valid) {
      const u32 entry = b.touched2[t], leaf = p.fleet ? entry & LEAF_MASK : entry;
      hdr = *reinterpret_cast<const int4*>(leaf_ptr(gs, leaf));
      hdr.w = p.fleet ? (int)(entry >> 28) : 0;  // fleet step: the sensor whose rays touched the leaf travels with the record
      unsigned long long* touched = reinterpret_cast<unsigned long long*>(leaf_touched(gs, leaf));
      o = shard_owner(hdr.x >> 5, hdr.y >> 5, hdr.z >> 5, p.world);
      if (sub == 0) local = atomicAdd(&s_cnt[o], 1u);  // slot inside this block's share: shared-memory atomic
      m = touched[sub];
      touched[sub] = 0ull;
    }
    __syncthreads();
    // ONE global atomic per owner and block (tens of thousands of returning atomics on `world` addresses serialise in L2)
    if (threadIdx.x < p.world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&b.sc->cnt2[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    local = __shfl_sync(0xffffffffu, local, first);
    if (valid) {
      const u32 at = s_base[o] + local + 1u;
      int4* block = b.px->leaf[o];
      if (at < cap) {
        if (sub == 0) block[(size_t)at * 5] = make_int4(hdr.x, hdr.y, hdr.z, hdr.w);
        reinterpret_cast<unsigned long long*>(block + (size_t)at * 5 + 1)[sub] = m;
      } else if (sub == 0) {
        atomicOr(&b.sc->overflow, OVF_LEAVES);
      }
    }
    __syncthreads();
  }
  publish_blocks(b.px, true, b.sc->cnt2, &b.sc->done2, MBOX_FLAG2, p.rank, p.xseq2, p.world, cap);
}

// This rank's error flags for the reduction (OR) that gates the apply phase on every rank. Caller-run exchange and
// NCCL: written to `flags`, all-reduced afterwards. Peer memory: stored into slot [rank] of every rank's flag table
// (values first, then the stamp); the apply kernel ORs the table itself. Threads 0..world-1 of one block.
__device__ __forceinline__ void shard_flags(const GridDev& g, const GridDev& gs, const ScanParams& p, const ScanBuffers& b, u32* flags) {
  const u32 f0 = g.ctr->error | (gs.ctr->error << 8), f1 = b.sc->overflow;
  u32 f2 = 0;  // fullest leaf-record block this rank sent
  for (u32 o = 0; o < p.world; ++o) f2 = max(f2, *reinterpret_cast<const volatile u32*>(&b.sc->cnt2[o]));
  if (b.px->flag[0] == nullptr) {
    if (threadIdx.x == 0) {
      flags[0] = f0;
      flags[1] = f1;
      flags[2] = f2;
      flags[3] = 0;
    }
    return;
  }
  if (threadIdx.x < p.world) {
    u32* dst = b.px->flag[threadIdx.x] + MBOX_FLAGS4 + (p.xseq2 & 1u) * (MAX_PEERS * 4) + p.rank * 4;
    dst[0] = f0;
    dst[1] = f1;
    dst[2] = f2;
    __threadfence_system();
    st_release_sys(dst + 3, p.xseq2);
  }
}

// exchange 2, receiver: OR the remote masks into this rank's leaves. The records of all senders are addressed as one
// dense range (as in k_shard_dedupe), eight lanes per record: one finds / creates the leaf, each merges one 64-bit word.
// The last block then publishes this rank's error flags.
__global__ void __launch_bounds__(TPB) k_shard_merge(GridDev g, GridDev gs, ScanParams p, ScanBuffers b, const int4* recv, u32 cap, u32* flags) {
  pdl_enter();
  const u32 lane = threadIdx.x & 31, sub = lane & 7u, first = lane & 24u;
  const u32 groups = gridDim.x * (TPB / 8);
  wait_arrivals(b.my_flags ? b.my_flags + MBOX_FLAG2 : nullptr, 1, p.world, p.xseq2, const_cast<u32*>(b.poison));
  const bool frozen = *b.poison != 0u;
  u32 incl = lane < p.world ? min((u32)__ldcg(&recv[(size_t)lane * cap * 5]).x, cap - 1u) : 0u;
  for (int o = 1; o < 32; o <<= 1) {
    const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (u32)o) incl += v;
  }
  const u32 total = frozen ? 0u : __shfl_sync(0xffffffffu, incl, 31);
  for (u32 t0 = (blockIdx.x * (TPB / 32) + (threadIdx.x >> 5)) * 4u; t0 < total; t0 += groups) {  // warp-uniform trip count
    const u32 t = t0 + (lane >> 3);
    const int4* rec = nullptr;
    u32 base = 0;
    for (u32 src = 0; src < p.world; ++src) {
      const u32 end = __shfl_sync(0xffffffffu, incl, src);
      if (rec == nullptr && t >= base && t < end) rec = recv + ((size_t)src * cap + 1u + (t - base)) * 5;
      base = end;
    }
    u32 leaf = NONE, tag = 0;
    if (rec != nullptr && sub == 0) {
      const int4 hdr = __ldcg(rec);
      leaf = leaf_find_or_create(g, hdr.x, hdr.y, hdr.z);
      tag = p.fleet ? (u32)hdr.w << 28 : 0u;
    }
    leaf = __shfl_sync(0xffffffffu, leaf, first);
    tag = __shfl_sync(0xffffffffu, tag, first);
    bool list_it = false;
    if (leaf != NONE) {
      const unsigned long long bits = __ldcg(reinterpret_cast<const unsigned long long*>(rec + 1) + sub);
      if (bits) list_it = mark_bits(g, leaf, reinterpret_cast<unsigned long long*>(leaf_touched(g, leaf)) + sub, bits, p.seq);
    }
    list_leaves(list_it, leaf | tag, &b.sc->n_touched, b.touched, p.touched_cap);
  }
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&b.sc->done3, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) b.sc->done3 = 0u;
  __threadfence();  // acquire side: the error bits other blocks set before their tickets
  shard_flags(g, gs, p, b, flags);
}

// public addHitPoint / addMissPoint (probabilistic_map.cpp:30-54): update now, queue the ray
__global__ void k_add_point(GridDev g, ScanParams p, ScanBuffers b, int4 e, u32 at, u32* queued) {
  const u32 leaf = leaf_find_or_create(g, e.x, e.y, e.z);
  *queued = 0;
  if (leaf == NONE) return;
  const u32 ci = ((u32)e.x & 7u) | (((u32)e.y & 7u) << 3) | (((u32)e.z & 7u) << 6);
  u32* cell = reinterpret_cast<u32*>(leaf_cells(g, leaf)) + ci;
  unsigned long long* act = reinterpret_cast<unsigned long long*>(leaf_active(g, leaf)) + (ci >> 6);
  const unsigned long long bit = 1ull << (ci & 63);
  const bool on = (*act & bit) != 0;
  const u32 word = on ? *cell : 0u;
  if (!on) {
    *act |= bit;
    *cell = 0u;
  }
  if ((word & 0xFu) != p.c) {
    i32 prob = (i32)word >> 4;
    prob = e.w ? max(prob + p.miss, p.cmin) : min(prob + p.hit, p.cmax);
    *cell = ((u32)prob << 4) | p.c;
    b.pending[at] = e;
    *queued = 1;
  }
}

// isOccupied / isUnknown / isFree, probabilistic_map.cpp:56-75
__global__ void __launch_bounds__(TPB) k_query(GridDev g, const i32* __restrict__ xyz, i64 n, int kind, i32 thr, u8* __restrict__ out) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const u32 leaf = leaf_find(g, x, y, z);
    bool r = kind == BNX_UNKNOWN;  // missing cell: unknown, neither occupied nor free
    if (leaf != NONE) {
      const u32 ci = ((u32)x & 7u) | (((u32)y & 7u) << 3) | (((u32)z & 7u) << 6);
      if ((leaf_active(g, leaf)[ci >> 6] >> (ci & 63)) & 1ull) {
        const i32 prob = (i32) reinterpret_cast<const u32*>(leaf_cells(g, leaf))[ci] >> 4;
        r = kind == BNX_OCCUPIED ? prob > thr : kind == BNX_UNKNOWN ? prob == thr : prob < thr;
      }
    }
    out[i] = r;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// BNX_APPLY_TMA=1: the apply pass keeps three leaves per warp in flight through the bulk-copy engine (cp.async.bulk +
// mbarrier) instead of register-held loads. Bit-exact, measured 1.2 us SLOWER per scan on the B200 (16.9 vs 15.7 us:
// a 192-byte and a few 128-byte bulk copies per leaf have a higher latency than the LDGs they replace, and three leaves
// per warp do not cover it; profiles/r2_notes.md), so the register flavour stays the default.
static bool apply_tma() {
  static const bool v = [] {
    const char* e = std::getenv("BNX_APPLY_TMA");
    return e && std::strcmp(e, "1") == 0;
  }();
  return v;
}

static bool shard_debug() {
  static const bool v = std::getenv("BNX_DEBUG") != nullptr;
  return v;
}

Map::~Map() {
  if (grid.stream()) cudaStreamSynchronize(grid.stream());
  if (copy_stream_) {
    cudaStreamSynchronize(copy_stream_);
    cudaStreamDestroy(copy_stream_);
  }
  if (pre_stream_) {
    cudaStreamSynchronize(pre_stream_);
    cudaStreamDestroy(pre_stream_);
  }
  for (auto& st : sets_)
    if (st.classified) cudaEventDestroy(st.classified);
  for (auto& e : x_copied_)
    if (e) cudaEventDestroy(e);
  if (grown_) cudaEventDestroy(grown_);
  for (auto& e : x_merged_)
    if (e) cudaEventDestroy(e);
  for (auto& e : x_begun_)
    if (e) cudaEventDestroy(e);
  if (h_ring_) cudaFreeHost(h_ring_);
  p2p_close_peers();
  if (mbox_) cudaFree(mbox_);
  for (void* old : mbox_retired_) cudaFree(old);
  // the communicator only bootstraps (or carries the BNX_SHARD_EXCHANGE=nccl exchanges of scans that are complete by now):
  // it is ABORTED, not destroyed — ncclCommDestroy may wait for the peer ranks, and maps are not destroyed at the same
  // moment on every rank (a garbage-collected handle in one process must not dead-lock the next collective of another)
  if (comm_) {
    const NcclApi& api = nccl_api(nullptr);
    if (api.CommAbort) {
      api.CommAbort(static_cast<ncclComm_t>(comm_));
    } else {
      api.CommDestroy(static_cast<ncclComm_t>(comm_));
    }
  }
  delete scratch_;
  if (h_status_) cudaFreeHost(h_status_);
  for (auto& e : ev_)
    if (e) cudaEventDestroy(e);
}

bool Map::sync_via_ring() {
  static const bool v = [] {
    const char* e = std::getenv("BNX_SYNC_RING");
    return !(e && std::strcmp(e, "0") == 0);
  }();
  return v;
}

static i32 logods_host(float prob) {  // probabilistic_map.hpp:34-36
  return (i32)(1e6 * std::log(prob / (1.0 - prob)));
}

constexpr size_t SC_BYTES = 256;  // ScanCounters header of a set's table buffer
static_assert(sizeof(ScanCounters) <= SC_BYTES, "ScanCounters must fit its header");

static u64 table_slots(i64 n) {
  u64 slots = 1024;
  while (slots < (u64)n * 2) slots <<= 1;
  return slots;
}

int Map::init(double resolution) {
  BNX_TRY(grid.init(resolution, 2, 3, 4));  // probabilistic_map.cpp:14-16
  options[0] = logods_host(0.4f);
  options[1] = logods_host(0.7f);
  options[2] = logods_host(0.12f);
  options[3] = logods_host(0.97f);
  options[4] = logods_host(0.5f);
  BNX_CUDA(cudaMallocHost(&h_status_, sizeof(ScanCounters) + 16));
  for (auto& e : ev_) BNX_CUDA(cudaEventCreate(&e));
  BNX_TRY(b_pending_.reserve(1024 * sizeof(int4)));
  buf_.pending = b_pending_.as<int4>();
  BNX_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_ring_), sizeof(AsyncRecord) * RING, cudaHostAllocMapped));
  std::memset(h_ring_, 0xFF, sizeof(AsyncRecord) * RING);
  BNX_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_ring_), h_ring_, 0));
  BNX_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
  BNX_CUDA(cudaStreamCreateWithFlags(&pre_stream_, cudaStreamNonBlocking));
  for (auto& st : sets_) BNX_CUDA(cudaEventCreateWithFlags(&st.classified, cudaEventDisableTiming));
  for (auto& e : x_copied_) BNX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  BNX_CUDA(cudaEventCreateWithFlags(&grown_, cudaEventDisableTiming));
  for (auto& e : x_merged_) BNX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : x_begun_) BNX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  buf_.poison = &grid.dev().ctr->error;
  buf_.ring = d_ring_;
  return reserve_scan(0, 16, 1.0);
}

// 32-chunk tiles of a scan: estimated from the longest possible ray, grown on overflow
size_t Map::tile_bytes(size_t np, double max_range) const {
  double cells = std::isfinite(max_range) ? std::ceil(max_range * grid.inv_resolution) + 2.0 : 512.0;
  cells = std::min(cells, 4096.0);
  return ((size_t)((double)np * (cells / CHUNK + 2.0) / 32.0) + np / 32 + 64) * 4;
}

int Map::reserve_scan(i64 n, i64 stride_bytes, double max_range, i64 table_n) {
  (void)stride_bytes;
  const size_t np = (size_t)n + n_pending_ + 32;
  BNX_TRY(S().ep.reserve(np * sizeof(int4)));
  BNX_TRY(S().slot.reserve(np * 4));
  BNX_TRY(b_rays_.reserve(np * sizeof(int4)));
  // [ScanCounters | table u32[slots] | keys u64[slots]] — contiguous so that one memset clears all of it
  const u64 slots = table_slots(table_n >= 0 ? table_n : n);
  const void* table_before = S().table.p;
  BNX_TRY(S().table.reserve(SC_BYTES + slots * 12));
  if (S().table.p != table_before) {  // fresh allocation: nothing is known to be zero
    S().clean_slots = 0;
    S().sc_clean = S().t1_clean = false;
  }
  BNX_TRY(b_tiles_.reserve(tile_bytes(np, max_range)));
  if (b_touched_.bytes == 0) BNX_TRY(b_touched_.reserve((size_t)grid.dev().leaf_cap * 8));  // 2 x the pool; grown by the sparse launch path only
  d_sc_ = S().table.as<ScanCounters>();
  buf_.sc = d_sc_;
  buf_.table = reinterpret_cast<u32*>(S().table.as<unsigned char>() + SC_BYTES);
  buf_.keys = reinterpret_cast<unsigned long long*>(S().table.as<unsigned char>() + SC_BYTES + slots * 4);
  buf_.ep = S().ep.as<int4>();
  buf_.slot_of = S().slot.as<u32>();
  buf_.rays = b_rays_.as<int4>();
  buf_.tile_first = b_tiles_.as<u32>();
  buf_.touched = b_touched_.as<u32>();
  return BNX_OK;
}

int Map::build_params(i64 n, const double origin[3], double max_range, ScanParams* out) {
  ScanParams p = {};
  p.ox = origin[0];
  p.oy = origin[1];
  p.oz = origin[2];
  p.max_range = max_range;
  p.max_range_sqr = max_range * max_range;  // probabilistic_map.hpp:145
  p.inv_res = grid.inv_resolution;
  // posToCoord(origin), probabilistic_map.cpp:91 — one fp64 multiply then floor, identical on host and device
  p.Ox = (i32)std::floor(origin[0] * grid.inv_resolution);
  p.Oy = (i32)std::floor(origin[1] * grid.inv_resolution);
  p.Oz = (i32)std::floor(origin[2] * grid.inv_resolution);
  p.miss = options[0];
  p.hit = options[1];
  p.cmin = options[2];
  p.cmax = options[3];
  p.c = update_count;
  p.n = (u32)n;
  p.world = 1;
  p.async_id = NONE;
  p.clean16 = 0;
  p.use_transform = use_next_T_ ? 1u : 0u;
  for (int k = 0; k < 12; ++k) p.T[k] = next_T_[k];
  use_next_T_ = false;
  p.max_chunks = (u32)std::min<u64>(((1ull << 40) - 1) / (u64)std::max<i64>(n + n_pending_, 1), 1ull << 28);
  // every endpoint lies within max_range of the origin (hits by the range test, misses by truncation): if that
  // ball fits 21 bits per axis the dedupe table can use packed keys
  p.packed = 0;
  if (std::isfinite(max_range) && max_range >= 0.0) {
    const double reach = std::ceil(max_range * grid.inv_resolution) + 4.0;
    const double lim = (double)(1 << 20) - 1.0;
    if (std::fabs((double)p.Ox) + reach < lim && std::fabs((double)p.Oy) + reach < lim && std::fabs((double)p.Oz) + reach < lim) p.packed = 1;
  }
  p.hash_mask = (u32)(table_slots(n) - 1);
  BNX_TRY(reserve_dense(p));
  *out = p;
  return BNX_OK;
}

// bytes of host memory a strided cloud really occupies: up to the z of the last point, not n * stride (x may sit at a
// non-zero offset inside the caller's point type, and the base pointer handed in is &points[0].x)
static size_t cloud_bytes(i64 n, i64 stride_bytes, bool f64) { return n > 0 ? (size_t)(n - 1) * (size_t)stride_bytes + (f64 ? 24u : 12u) : 0u; }

// Dense marking window (EXPERIMENTAL, off by default): every endpoint and every ray cell of a scan lies
// within max_range of the origin, so when that ball — in 8^3 leaf blocks — is small enough, the per-scan marks can go to
// a dense array addressed by arithmetic (one 128-B line per block) instead of into the leaves. Bit-exact like the
// leaf-resident ("sparse") marks, but measured SLOWER on the B200 in all four variants tried (profiles/r2_notes.md), so
// it only runs when asked for: BNX_DENSE=1 or bnx_map_set_marking(m, 2). The sparse marks serve every scan, including
// max_range = inf, huge ranges and grids with other inner/leaf bits. The buffers are allocated (and zeroed) once per
// window size; the apply pass leaves every line it used zeroed again.
static bool dense_by_default() {
  static const bool v = [] {
    const char* e = std::getenv("BNX_DENSE");
    return e && std::strcmp(e, "1") == 0;
  }();
  return v;
}
static u32 dense_max_blocks_per_axis() {
  static const u32 v = [] {
    size_t mb = 2048;
    if (const char* m = std::getenv("BNX_DENSE_MAX_MB")) mb = (size_t)std::strtoull(m, nullptr, 10);
    u32 d = 8;
    while ((size_t)(d + 8) * (d + 8) * (d + 8) * 128 <= (mb << 20)) d += 8;
    return d;
  }();
  return v;
}

int Map::set_fleet(const double* origins_world_x3) {
  BNX_REQUIRE(world_ > 1, "set_fleet: the map is not sharded");
  fleet_origins_.clear();
  if (origins_world_x3) fleet_origins_.assign(origins_world_x3, origins_world_x3 + (size_t)world_ * 3);
  return BNX_OK;
}

int Map::set_marking(int mode) {
  BNX_REQUIRE(mode >= 0 && mode <= 2, "set_marking: 0 (default), 1 (sparse) or 2 (dense window where the range allows it)");
  BNX_TRY(drain());
  marking_ = mode;
  return BNX_OK;
}

u32 Map::dense_dim(double max_range) const {
  const bool want_dense = marking_ == 2 || (marking_ == 0 && dense_by_default());
  const u32 dmax = want_dense ? dense_max_blocks_per_axis() : 0u;
  const GridDev g = grid.dev();
  if (!dmax || g.ib != 2 || g.lb != 3 || !std::isfinite(max_range) || max_range < 0.0) return 0;
  const double reach_d = std::ceil(max_range * grid.inv_resolution) + 3.0;  // |endpoint voxel - origin voxel| stays below this
  if (reach_d > 8.0 * dmax) return 0;
  const u32 D = (u32)((2 * (i64)reach_d) / 8 + 2);  // blocks per axis that cover [O - reach, O + reach] wherever O sits in its block
  return D <= dmax ? D : 0;
}

bool Map::scan_is_dense(const double origin[3], double max_range) const {
  if (!dense_dim(max_range)) return false;
  const double reach = std::ceil(max_range * grid.inv_resolution) + 3.0, lim = (double)(1 << 30);
  for (int k = 0; k < 3; ++k)
    if (!(std::fabs(std::floor(origin[k] * grid.inv_resolution)) + reach < lim)) return false;
  return true;
}

int Map::reserve_dense(ScanParams& p) {
  p.dense = 0;
  const double o[3] = {p.ox, p.oy, p.oz};
  if (!scan_is_dense(o, p.max_range)) return BNX_OK;
  const u32 D = dense_dim(p.max_range);
  const i64 reach = (i64)(std::ceil(p.max_range * p.inv_res) + 3.0);
  if (D > dense_D_) {
    // (re)allocation: nothing may be in flight (the callers drain first when dense_need() says so)
    const u32 Da = (D + 7u) & ~7u;
    const size_t blocks = (size_t)Da * Da * Da;
    b_dense_.release();
    b_dbits_.release();
    b_dlist_.release();
    BNX_TRY(b_dense_.reserve(blocks * 128));
    BNX_TRY(b_dbits_.reserve(blocks / 8 + 128));
    BNX_TRY(b_dlist_.reserve(blocks * 4));
    BNX_CUDA(cudaMemsetAsync(b_dense_.p, 0, b_dense_.bytes, grid.stream()));
    BNX_CUDA(cudaMemsetAsync(b_dbits_.p, 0, b_dbits_.bytes, grid.stream()));
    if (!b_dhint_.p) {
      BNX_TRY(b_dhint_.reserve((size_t)4 << HINT_BITS));
      BNX_CUDA(cudaMemsetAsync(b_dhint_.p, 0, b_dhint_.bytes, grid.stream()));
    }
    dense_D_ = Da;
  }
  p.dense = 1;
  p.D = D;
  p.W0x = (i32)((p.Ox - reach) >> 3);
  p.W0y = (i32)((p.Oy - reach) >> 3);
  p.W0z = (i32)((p.Oz - reach) >> 3);
  p.dlist_cap = (u32)std::min<size_t>(b_dlist_.bytes / 4, 0xFFFFFFF0ull);
  return BNX_OK;
}

static int check_insert_args(const void* points, i64 stride_bytes, i64 n, bool f64, const double origin[3], i64 pending) {
  BNX_REQUIRE(n >= 0 && n + pending < (1ll << 24), "insert: at most 2^24-1 points per scan");
  BNX_REQUIRE(n == 0 || points != nullptr, "insert: null points");
  BNX_REQUIRE(origin != nullptr, "insert: null origin");
  if (f64) {
    BNX_REQUIRE(stride_bytes >= 24 && stride_bytes % 8 == 0, "insert_f64: stride must be a multiple of 8, >= 24");
  } else {
    BNX_REQUIRE(stride_bytes >= 12 && stride_bytes % 4 == 0, "insert_f32: stride must be a multiple of 4, >= 12");
  }
  return BNX_OK;
}

// The synchronous insertPointCloud (what the drop-in C++ header calls). Fast path: the scan goes through the pipelined
// machinery — PDL launches, no memset, no counter copy — and the host then waits for the scan's record in the pinned
// ring instead of synchronising the stream; a scan that ran short falls into the usual drain (grow + replay).
// Queued addHitPoint/addMissPoint rays, a sharded map and per-phase profiling take the classic path below.
int Map::insert(const void* points, i64 stride_bytes, i64 n, bool f64, const double origin[3], double max_range, int where) {
  BNX_TRY(check_insert_args(points, stride_bytes, n, f64, origin, n_pending_));
  BNX_TRY(drain());
  if (!n_pending_ && world_ == 1 && !profiling && sync_via_ring()) {
    if (where == BNX_DEVICE) {
      // the synchronous call reads a device buffer in the order of the map's stream (include/bonxai_b200.h): the front
      // half runs on the internal stream, so that stream waits for what the map's stream has been given so far
      BNX_CUDA(cudaEventRecord(ev_[0], grid.stream()));
      BNX_CUDA(cudaStreamWaitEvent(pre_stream_, ev_[0], 0));
    }
    BNX_TRY(insert_async(points, stride_bytes, n, f64, origin, max_range, where));
    return complete_queue();
  }
  set_ = 0;
  cudaStream_t s = grid.stream();
  if (profiling) cudaEventRecord(ev_[0], s);
  BNX_TRY(reserve_scan(n, stride_bytes, max_range));
  const void* d_points = points;
  if (where == BNX_HOST && n > 0) {
    BNX_TRY(b_pts_.reserve((size_t)n * stride_bytes));
    BNX_CUDA(cudaMemcpyAsync(b_pts_.p, points, cloud_bytes(n, stride_bytes, f64), cudaMemcpyHostToDevice, s));
    d_points = b_pts_.p;
  }
  ScanParams p;
  BNX_TRY(build_params(n, origin, max_range, &p));
  BNX_TRY(run_scan(d_points, stride_bytes, f64, p, false));
  if (++update_count == 4) update_count = 1;  // probabilistic_map.cpp:103-105
  return BNX_OK;
}

template <bool F64, bool VEC4>
static void launch_classify(bool packed, int blocks, cudaStream_t s, const unsigned char* pts, u32 stride, const ScanParams& p, const ScanBuffers& b) {
  if (packed) {
    launch_scan_kernel(k_classify<F64, VEC4, true>, blocks, TPB, s, pts, stride, p, b);
  } else {
    launch_scan_kernel(k_classify<F64, VEC4, false>, blocks, TPB, s, pts, stride, p, b);
  }
}

int Map::launch_front(cudaStream_t s, const void* d_points, i64 stride_bytes, bool f64, ScanParams& p) {
  const i64 n = p.n;
  const u64 slots = (u64)p.hash_mask + 1;
  const int persistent = sm_count() * 8;
  buf_.poison = &grid.dev().ctr->error;
  buf_.ring = d_ring_;
  S().sc_clean = S().t1_clean = false;
  // counters + dedupe table (+ packed keys) in one clear
  const size_t bytes = n > 0 ? SC_BYTES + slots * (p.packed ? 12 : 4) : SC_BYTES;
  if (p.async_id == NONE) {
    BNX_CUDA(cudaMemsetAsync(d_sc_, 0, bytes, s));
    S().clean_slots = 0;  // the synchronous path leaves a used table behind
  } else {
    if (slots > S().clean_slots) {
      launch_scan_kernel(k_begin_scan, std::min<int>(persistent, blocks_for((i64)(bytes / 16))), TPB, s, buf_, reinterpret_cast<uint4*>(d_sc_), (u32)(bytes / 16));
    }
    p.clean16 = (u32)(slots * 12 / 16);  // k_mark zeroes table + keys again, the epilogue the counters
    S().clean_slots = slots;
  }
  if (n > 0) {
    const unsigned char* pts = static_cast<const unsigned char*>(d_points);
    const int blocks = blocks_for(n);
    if (f64) {
      launch_classify<true, false>(p.packed, blocks, s, pts, (u32)stride_bytes, p, buf_);
    } else if (stride_bytes == 16 && (reinterpret_cast<uintptr_t>(pts) & 15u) == 0) {
      launch_classify<false, true>(p.packed, blocks, s, pts, 16u, p, buf_);
    } else {
      launch_classify<false, false>(p.packed, blocks, s, pts, (u32)stride_bytes, p, buf_);
    }
  }
  BNX_CUDA(cudaGetLastError());
  return BNX_OK;
}

int Map::launch_back(cudaStream_t s, ScanParams& p, bool first_attempt) {
  const i64 n = p.n;
  const int persistent = sm_count() * 8;
  const GridDev g = grid.dev();
  buf_.poison = &g.ctr->error;
  buf_.ring = d_ring_;
  p.seq = ++seq_;
  p.tile_cap = (u32)std::min<size_t>(b_tiles_.bytes / 4, 0xFFFFFFFFull);
  p.touched_cap = (u32)std::min<size_t>(b_touched_.bytes / 4, 0xFFFFFFFFull);
  if (p.dense) {
    buf_.dense = b_dense_.as<unsigned long long>();
    buf_.dbits = b_dbits_.as<u32>();
    buf_.dhint = b_dhint_.as<u32>();
    buf_.touched = b_dlist_.as<u32>();
    if (n_pending_) launch_scan_kernel(k_resolve<1, true>, blocks_for(n_pending_), TPB, s, g, p, buf_, n_pending_);
    if (n > 0) launch_scan_kernel(k_resolve<0, true>, blocks_for(n), TPB, s, g, p, buf_, (u32)n);
    if (profiling && first_attempt) cudaEventRecord(ev_[3], s);
    // stretches of SEG cells a ray of this window can have
    const u32 smax = ((u32)(std::ceil(p.max_range * p.inv_res) + 3.0) + SEG - 1u) / SEG;
    launch_scan_kernel(k_mark_dense, sm_count() * MARKD_MIN_BLOCKS, TPB, s, p, buf_, smax);
    if (profiling && first_attempt) cudaEventRecord(ev_[4], s);
    const u32 nwords = (p.D * p.D * p.D + 31u) / 32u;
    launch_scan_kernel(k_list_blocks, std::min<int>(sm_count() * 4, blocks_for(nwords)), TPB, s, p, buf_, nwords);
    launch_scan_kernel(k_apply_dense, sm_count() * APPLY_MIN_BLOCKS, TPB, s, g, p, buf_, 0u);
    BNX_CUDA(cudaGetLastError());
    if (profiling && first_attempt) cudaEventRecord(ev_[5], s);
    buf_.touched = b_touched_.as<u32>();
    return BNX_OK;
  }
  if ((size_t)g.leaf_cap * 4 > b_touched_.bytes) {  // pipelined callers have drained before a scan that needs this
    BNX_TRY(b_touched_.reserve((size_t)g.leaf_cap * 8));
    buf_.touched = b_touched_.as<u32>();
    p.touched_cap = (u32)std::min<size_t>(b_touched_.bytes / 4, 0xFFFFFFFFull);
  }
  if (n_pending_) launch_scan_kernel(k_resolve<1, false>, blocks_for(n_pending_), TPB, s, g, p, buf_, n_pending_);
  if (n > 0) launch_scan_kernel(k_resolve<0, false>, blocks_for(n), TPB, s, g, p, buf_, (u32)n);
  if (profiling && first_attempt) cudaEventRecord(ev_[3], s);
  launch_scan_kernel(k_mark<false>, sm_count() * MARK_MIN_BLOCKS, TPB, s, g, g, p, buf_);
  if (profiling && first_attempt) cudaEventRecord(ev_[4], s);
  launch_scan_kernel(apply_tma() ? k_apply_leaves<true> : k_apply_leaves<false>, sm_count() * APPLY_MIN_BLOCKS, TPB, s, g, p, buf_);
  BNX_CUDA(cudaGetLastError());
  if (profiling && first_attempt) cudaEventRecord(ev_[5], s);
  return BNX_OK;
}

// a dense scan reported a scratch overflow (a mark outside its window: cannot happen by construction). Nothing was
// applied; wipe the window so that later scans stay exact, and fail loudly.
int Map::dense_internal_error(const ScanCounters& st, const ScanParams&) {
  cudaStream_t s = grid.stream();
  BNX_CUDA(cudaMemsetAsync(b_dense_.p, 0, b_dense_.bytes, s));
  BNX_CUDA(cudaMemsetAsync(b_dbits_.p, 0, b_dbits_.bytes, s));
  BNX_CUDA(cudaStreamSynchronize(s));
  if (st.gc.error) BNX_TRY(grid.recover(st.gc));
  set_error("insert: internal error, a ray cell fell outside the dense marking window (overflow bits " + std::to_string(st.overflow) + ")");
  return BNX_ERR_CUDA;
}

// dense scans only: the apply pass ran out of leaves / inner nodes / root slots part-way. The pools have been grown;
// run the pass again over the same list: blocks that were applied have empty lines, the others still hold their marks.
int Map::resume_apply(cudaStream_t s, ScanParams& p) {
  const GridDev g = grid.dev();
  buf_.dense = b_dense_.as<unsigned long long>();
  buf_.dbits = b_dbits_.as<u32>();
  buf_.dhint = b_dhint_.as<u32>();
  buf_.touched = b_dlist_.as<u32>();
  buf_.poison = &g.ctr->error;
  note_launch(), k_apply_dense<<<sm_count() * APPLY_MIN_BLOCKS, TPB, 0, s>>>(g, p, buf_, 1u);
  BNX_CUDA(cudaGetLastError());
  buf_.touched = b_touched_.as<u32>();
  return BNX_OK;
}

// the kernel sequence of ONE attempt at a scan on the map's stream (no synchronisation)
int Map::launch_scan(const void* d_points, i64 stride_bytes, bool f64, ScanParams& p, bool first_attempt) {
  cudaStream_t s = grid.stream();
  if (first_attempt) {
    if (profiling) cudaEventRecord(ev_[1], s);
    BNX_TRY(launch_front(s, d_points, stride_bytes, f64, p));
    if (profiling) cudaEventRecord(ev_[2], s);
  } else {
    BNX_CUDA(cudaMemsetAsync(d_sc_, 0, SC_BYTES, s));
  }
  return launch_back(s, p, first_attempt);
}

void Map::account(const ScanCounters& st, i64 n, i64 pending, i64 retries) {
  n -= st.n_dropped;  // the reference's caller removes non-finite points before insertPointCloud sees the cloud
  counters[0] = n;
  counters[1] = (i64)st.n_endpoints + pending;
  counters[2] = (i64)st.sum_m + n;
  counters[3] = (i64)st.n_changed;  // endpoints are applied (and counted) by the leaf pass too
  counters[4] = st.n_touched;
  counters[5] = retries;
  counters[6] = (i64)(st.ray_chunk >> 40);
  counters[7] = (i64)(st.ray_chunk & CHUNK_FIELD);
  for (int k = 0; k < 4; ++k) totals[k] += counters[k];
}

// synchronous scan: attempts until the pools were large enough (each failed attempt changed nothing)
int Map::run_scan(const void* d_points, i64 stride_bytes, bool f64, const ScanParams& base, bool) {
  cudaStream_t s = grid.stream();
  ScanParams p = base;
  p.async_id = NONE;
  p.clean16 = 0;  // a retry re-reads the dedupe table: the synchronous path clears it with its own memset
  const int persistent = sm_count() * 8;
  i64 retries = 0;
  if (p.dense) {
    // dense marks: only the apply pass can run short (it is the one place where leaves are created); it is RESUMED
    // after the pools have grown — blocks that were applied have left the bitmap
    BNX_TRY(launch_scan(d_points, stride_bytes, f64, p, true));
    for (;; ++retries) {
      BNX_CUDA(cudaMemcpyAsync(h_status_, d_sc_, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
      BNX_CUDA(cudaStreamSynchronize(s));
      const ScanCounters st = *h_status_;
      if (st.overflow) return dense_internal_error(st, p);
      if (st.gc.error == 0) break;
      if (retries > 48) {
        set_error("insert: node pools could not be grown enough for this scan");
        return BNX_ERR_NOMEM;
      }
      BNX_TRY(grid.recover(st.gc));
      BNX_TRY(resume_apply(s, p));
    }
  }
  for (; !p.dense; ++retries) {
    if (retries > 48) {
      set_error("insert: node pools could not be grown enough for this scan");
      return BNX_ERR_NOMEM;
    }
    BNX_TRY(launch_scan(d_points, stride_bytes, f64, p, retries == 0));
    BNX_CUDA(cudaMemcpyAsync(h_status_, d_sc_, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
    BNX_CUDA(cudaStreamSynchronize(s));
    const ScanCounters st = *h_status_;
    if (st.gc.error == 0 && st.overflow == 0) break;
    // phases 4/5 skipped themselves: nothing was applied. Drop the marks of the failed attempt (they must never reach a
    // later scan's apply pass), grow what was short and repeat from phase 2.
    if (st.n_touched) {
      note_launch(), k_clear_touched<<<persistent, TPB, 0, s>>>(grid.dev(), buf_, std::min(st.n_touched, p.touched_cap));
      BNX_CUDA(cudaGetLastError());
      BNX_CUDA(cudaStreamSynchronize(s));
    }
    if (st.gc.error) BNX_TRY(grid.recover(st.gc));
    if (st.overflow & OVF_CHUNKS) {  // the map is unchanged and clean: the scan is refused, later scans are not affected
      set_error("insert: more than 2^32 ray chunks in one scan");
      return BNX_ERR_UNSUPPORTED;
    }
    if (st.overflow & OVF_TILES) {
      const u64 chunks = st.ray_chunk & CHUNK_FIELD;
      BNX_TRY(b_tiles_.reserve((size_t)(chunks / 32 + 64) * 4));
      buf_.tile_first = b_tiles_.as<u32>();
    }
    BNX_TRY(b_touched_.reserve((size_t)grid.dev().leaf_cap * 4));
    buf_.touched = b_touched_.as<u32>();
  }
  const ScanCounters st = *h_status_;
  account(st, p.n, n_pending_, retries);
  n_pending_ = 0;
  if (profiling) {
    float ms;
    for (int k = 0; k < 5; ++k) {
      cudaEventElapsedTime(&ms, ev_[k], ev_[k + 1]);
      phase_us[k] = ms * 1e3;
    }
    cudaEventElapsedTime(&ms, ev_[0], ev_[5]);
    phase_us[5] = ms * 1e3;
  }
  return grid.maintain(st.gc);
}

// ------------------------------------------------------------------------------------------------
// pipelined insert
// ------------------------------------------------------------------------------------------------
int Map::insert_async(const void* points, i64 stride_bytes, i64 n, bool f64, const double origin[3], double max_range, int where) {
  BNX_TRY(check_insert_args(points, stride_bytes, n, f64, origin, n_pending_));
  if (n_pending_ || world_ > 1) return insert(points, stride_bytes, n, f64, origin, max_range, where);  // rare paths stay synchronous
  cudaStream_t s = grid.stream();
  if (queue_.size() >= RING / 2) BNX_TRY(drain(false));
  // how many scratch sets (= scans in flight + 2) does this scan size allow? (2 GiB of scratch at most)
  {
    const size_t np = (size_t)n + 32;
    const size_t set_bytes = np * 20 + SC_BYTES + table_slots(n) * 12 + (size_t)n * stride_bytes;  // staging counted for any input: host and device scans may alternate
    const int want = (int)std::min<size_t>(SETS, std::max<size_t>(4, (2ull << 30) / std::max<size_t>(set_bytes, 1)));
    if (want != sets_active_) {
      BNX_TRY(drain(false));  // the id -> set mapping changes: nothing may be in flight
      sets_active_ = want;
    }
  }
  const size_t max_in_flight = (size_t)sets_active_ - 2;
  // Flow control without a CUDA sync: at most max_in_flight scans are queued ahead of the newest record the device
  // has published in the host ring; that bounds how stale the pool head-room information below can be, and it makes
  // the scratch set of scan (id - sets_active_) free when scan id is enqueued.
  if (queue_.size() > done_upto_ + max_in_flight) {
    const u32 wait_id = queue_[queue_.size() - 1 - max_in_flight].p.async_id;
    const volatile AsyncRecord* r = &h_ring_[wait_id & (RING - 1)];
    unsigned spins = 0;
    bool seen = true;
    while (r->id != wait_id) {
      if (++spins > (1u << 14)) {  // never spin on a wedged device: fall back to a real synchronisation
        if (cudaStreamQuery(s) != cudaErrorNotReady) {
          seen = r->id == wait_id;
          break;
        }
        spins = 0;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (!seen) BNX_TRY(drain(false));  // the stream is idle but the record never came (a frozen or failed pipeline)
  }
  // head-room check on the newest published record: grow early, so that a queued scan (almost) never runs short
  if (!queue_.empty()) {
    for (size_t k = queue_.size(); k-- > done_upto_;) {
      const AsyncRecord& r = h_ring_[queue_[k].p.async_id & (RING - 1)];
      if (r.id != queue_[k].p.async_id) continue;
      done_upto_ = k + 1;
      const GridDev g = grid.dev();
      const u64 ahead = queue_.size() - k;  // scans that may still allocate before we look again
      const u64 leaves_ahead = (u64)r.n_leaves + ahead * max_leaf_growth_, inner_ahead = (u64)r.n_inner + ahead * 64;
      const u64 roots_ahead = (u64)r.n_roots + ahead * 64;
      // what the scans already queued may still allocate must fit the pools THEY were launched with; one more pipeline
      // depth of head-room is the trigger for mapping the next step
      const u64 full = (u64)sets_active_ * max_leaf_growth_, full_inner = (u64)sets_active_ * 64;
      const bool short_of_pool = leaves_ahead + full > g.leaf_cap || inner_ahead + full_inner > g.inner_cap;
      const u64 want_leaves = (u64)r.n_leaves + 2 * full + grid.leaf_step(r.n_leaves);
      const u64 want_inner = (u64)r.n_inner + 2 * full_inner + grid.inner_step(r.n_inner);
      // the touched-leaf list of the sparse marks holds one entry per leaf of the pool: it is allocated for twice the pool,
      // so the pool can grow behind running scans until it has doubled
      const bool list_fits = scan_is_dense(origin, max_range) || (want_leaves + want_leaves / 8) * 4 <= b_touched_.bytes;
      if (r.error || roots_ahead * 2 > (u64)g.root_mask + 1 || (short_of_pool && !list_fits)) {
        // the root table is rehashed, or the leaf-sized list re-allocated: nothing may be in flight
        BNX_TRY(drain(false));
        BNX_TRY(grid.ensure_leaf_capacity(want_leaves));
        BNX_TRY(grid.ensure_inner_capacity(want_inner));
        if (roots_ahead * 2 > (u64)g.root_mask + 1) BNX_TRY(grid.grow_root_table(((u64)g.root_mask + 1) * 4));
        BNX_TRY(b_touched_.reserve((size_t)grid.dev().leaf_cap * 8));
      } else if (short_of_pool) {
        // mapping memory behind the pools does not wait for the kernels in flight (VMM) — grow AHEAD, in the background:
        // the zero fill runs on the copy stream and only the scans enqueued from now on (which see the larger pool) wait
        // for it
        BNX_TRY(grid.ensure_leaf_capacity(want_leaves, copy_stream_));
        BNX_TRY(grid.ensure_inner_capacity(want_inner, copy_stream_));
        BNX_CUDA(cudaEventRecord(grown_, copy_stream_));
        BNX_CUDA(cudaStreamWaitEvent(s, grown_, 0));
        ++grown_ahead_;
      } else if (k > 0) {
        const AsyncRecord& q = h_ring_[queue_[k - 1].p.async_id & (RING - 1)];
        // per-scan leaf growth: the largest recent value (decays by 2 % per look, floor 2048): the first scans of a map
        // create far more leaves than the steady state, and a head-room estimate that never forgets them asks for
        // gigabytes of free pool for ever (city run: a growth step on almost every insert)
        max_leaf_growth_ = std::max<u64>(2048, max_leaf_growth_ - max_leaf_growth_ / 50);
        if (q.id == queue_[k - 1].p.async_id && r.n_leaves > q.n_leaves) max_leaf_growth_ = std::max<u64>(max_leaf_growth_, r.n_leaves - q.n_leaves);
      }
      break;
    }
  }
  // buffers shared by all scans in flight (stream-ordered use) must not be reallocated under them (the leaf-sized list
  // belongs to the sparse marks only)
  if ((!scan_is_dense(origin, max_range) && (size_t)grid.dev().leaf_cap * 4 > b_touched_.bytes) || ((size_t)n + 32) * sizeof(int4) > b_rays_.bytes ||
      tile_bytes((size_t)n + 32, max_range) > b_tiles_.bytes) {
    BNX_TRY(drain(false));
  }
  if (dense_dim(max_range) > dense_D_) BNX_TRY(drain(false));  // the dense window is (re)allocated by build_params
  // all scratch sets are sized together, up front: an allocation in the middle of the pipeline would synchronise the device
  if (n > sets_n_ || (where == BNX_HOST && (size_t)n * stride_bytes > sets_stage_bytes_)) {
    BNX_TRY(drain(false));
    sets_n_ = std::max<i64>(sets_n_, n + n / 8);
    if (where == BNX_HOST) sets_stage_bytes_ = std::max<size_t>(sets_stage_bytes_, (size_t)(n + n / 8) * stride_bytes);
    const int keep = set_;
    for (set_ = 0; set_ < sets_active_; ++set_) {
      BNX_TRY(reserve_scan(sets_n_, stride_bytes, max_range));
      if (sets_stage_bytes_) BNX_TRY(S().stage.reserve(sets_stage_bytes_));
    }
    set_ = keep;
  }
  Queued q;
  q.p.async_id = async_next_;
  set_ = (int)(async_next_ % (u32)sets_active_);  // free: its previous scan (id - sets_active_) has published its record
  ++async_next_;
  const u32 my_id = q.p.async_id;
  BNX_TRY(reserve_scan(n, stride_bytes, max_range));
  BNX_TRY(build_params(n, origin, max_range, &q.p));
  q.p.async_id = my_id;
  q.points = points;
  q.stride = stride_bytes;
  q.f64 = f64;
  q.where = where;
  // front half on the pre-stream, any number of scans ahead of the map updates: H2D copy (host input) + classify.
  // It only touches this scan's scratch set, never the map.
  const void* d_points = points;
  if (where == BNX_HOST && n > 0) {
    BNX_TRY(S().stage.reserve((size_t)n * stride_bytes));
    BNX_CUDA(cudaMemcpyAsync(S().stage.p, points, cloud_bytes(n, stride_bytes, f64), cudaMemcpyHostToDevice, pre_stream_));
    d_points = S().stage.p;
  }
  {
    PdlScope pdl(true);
    BNX_TRY(launch_front(pre_stream_, d_points, stride_bytes, f64, q.p));
    BNX_CUDA(cudaEventRecord(S().classified, pre_stream_));
    // back half on the map's stream: resolve -> mark -> apply, one scan after the other
    BNX_CUDA(cudaStreamWaitEvent(s, S().classified, 0));
    BNX_TRY(launch_back(s, q.p, true));
  }
  queue_.push_back(q);
  if (++update_count == 4) update_count = 1;
  return BNX_OK;
}

// the synchronous insert's wait: spin on the record of the newest queued scan (the device writes it into pinned host
// memory when the scan's last kernel ends), then account for the queue without touching the streams
int Map::complete_queue() {
  if (queue_.empty()) return BNX_OK;
  cudaStream_t s = grid.stream();
  const u32 last_id = queue_.back().p.async_id;
  const volatile AsyncRecord* r = &h_ring_[last_id & (RING - 1)];
  unsigned spins = 0;
  while (r->id != last_id) {
    if (++spins > (1u << 12)) {  // never spin on a wedged device
      if (cudaStreamQuery(s) != cudaErrorNotReady) break;
      spins = 0;
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  bool healthy = r->id == last_id && r->error == 0;
  for (size_t k = 0; healthy && k < queue_.size(); ++k) {
    const AsyncRecord& q = h_ring_[queue_[k].p.async_id & (RING - 1)];
    healthy = q.id == queue_[k].p.async_id && q.error == 0;
  }
  if (!healthy) return drain();  // ran short (or refused): grow, replay, report
  GridCounters gc = {};
  for (size_t k = 0; k < queue_.size(); ++k) {
    const AsyncRecord& q = h_ring_[queue_[k].p.async_id & (RING - 1)];
    ScanCounters st = {};
    st.n_endpoints = q.n_endpoints;
    st.n_changed = q.n_changed;
    st.n_touched = q.n_touched;
    st.n_dropped = q.n_dropped;
    st.sum_m = q.sum_m;
    st.ray_chunk = q.ray_chunk;
    account(st, queue_[k].p.n, 0, 0);
    gc.n_leaves = q.n_leaves;
    gc.n_inner = q.n_inner;
    gc.n_roots = q.n_roots;
  }
  queue_.clear();
  done_upto_ = 0;
  // pool head-room for the next scans, from the record (no device read); growing synchronises, but only when it grows
  const GridDev g = grid.dev();
  if ((u64)gc.n_roots * 2 > (u64)g.root_mask + 1 || (u64)gc.n_leaves + grid.leaf_step(gc.n_leaves) / 2 > g.leaf_cap ||
      (u64)gc.n_inner + grid.inner_step(gc.n_inner) / 2 > g.inner_cap) {  // the thresholds of Grid::maintain
    BNX_CUDA(cudaStreamSynchronize(s));
    BNX_CUDA(cudaStreamSynchronize(pre_stream_));
    return grid.maintain(gc);
  }
  return BNX_OK;
}

// report: hand a deferred error (a scan the pipeline refused) to the caller. insert_async drains with report = false —
// the scan being enqueued must not be lost to the failure of an earlier one; the next synchronising call reports it.
int Map::drain(bool report) {
  if (!squeue_.empty()) return shard_drain();
  const int st = drain_queue();
  if (st != BNX_OK) return st;
  if (report && deferred_ != BNX_OK) {
    const int d = deferred_;
    deferred_ = BNX_OK;
    set_error(deferred_msg_);
    return d;
  }
  return BNX_OK;
}

int Map::drain_queue() {
  if (queue_.empty()) return BNX_OK;
  cudaStream_t s = grid.stream();
  GridCounters gc;
  BNX_TRY(grid.read_counters(&gc));  // synchronises the map's stream
  BNX_CUDA(cudaStreamSynchronize(pre_stream_));
  std::vector<Queued> q;
  q.swap(queue_);
  done_upto_ = 0;
  if (gc.error) {
    // the counters of the scan that failed are still in ITS scratch set; the front halves that ran ahead of the freeze
    // left used tables behind, so no set is known to be clean any more
    set_ = (int)(gc.failed_id % (u32)sets_active_);
    BNX_CUDA(cudaMemcpyAsync(h_status_, sets_[set_].table.p, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
    BNX_CUDA(cudaStreamSynchronize(s));
    for (auto& st : sets_) {
      st.clean_slots = 0;
      st.sc_clean = st.t1_clean = false;
    }
  }
  size_t done = q.size();
  if (gc.error) {
    // frozen at the first scan that ran short: everything before it is applied, nothing after it is
    done = 0;
    while (done < q.size() && q[done].p.async_id != gc.failed_id) ++done;
  }
  for (size_t k = 0; k < done; ++k) {
    const AsyncRecord& r = h_ring_[q[k].p.async_id & (RING - 1)];
    ScanCounters st = {};
    st.n_endpoints = r.n_endpoints;
    st.n_changed = r.n_changed;
    st.n_touched = r.n_touched;
    st.n_dropped = r.n_dropped;
    st.sum_m = r.sum_m;
    st.ray_chunk = r.ray_chunk;
    account(st, q[k].p.n, 0, 0);
  }
  if (!gc.error) return grid.maintain(gc);
  if (q[done].p.dense) {
    // dense marks: the failed scan's apply pass ran out of pool space part-way (later scans skipped themselves, so its
    // window is intact): grow, RESUME the pass — applied blocks have left the bitmap — then replay the rest
    ScanParams fp = q[done].p;
    buf_.sc = d_sc_ = sets_[set_].table.as<ScanCounters>();
    ScanCounters st = *h_status_;
    if (st.overflow) {
      BNX_TRY(dense_internal_error(st, fp));
    }
    GridCounters cur = gc;
    i64 resumes = 0;
    for (int attempt = 0;; ++attempt) {
      ++resumes;
      if (attempt > 48) {
        set_error("insert: node pools could not be grown enough for this scan");
        return BNX_ERR_NOMEM;
      }
      BNX_TRY(grid.recover(cur));
      BNX_TRY(resume_apply(s, fp));
      BNX_CUDA(cudaMemcpyAsync(h_status_, d_sc_, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
      BNX_CUDA(cudaStreamSynchronize(s));
      st = *h_status_;
      if (st.gc.error == 0) break;
      cur = st.gc;
    }
    account(st, fp.n, 0, resumes);
    set_ = 0;
    for (size_t k = done + 1; k < q.size(); ++k) {
      const Queued& e = q[k];
      BNX_TRY(reserve_scan(e.p.n, e.stride, e.p.max_range));
      const void* d_points = e.points;
      if (e.where == BNX_HOST && e.p.n > 0) {
        BNX_TRY(b_pts_.reserve((size_t)e.p.n * e.stride));
        BNX_CUDA(cudaMemcpyAsync(b_pts_.p, e.points, cloud_bytes(e.p.n, e.stride, e.f64), cudaMemcpyHostToDevice, s));
        d_points = b_pts_.p;
      }
      BNX_TRY(run_scan(d_points, e.stride, e.f64, e.p, false));  // keeps the update_id (and the window) it was queued with
    }
    return BNX_OK;
  }
  // the failed scan's touched list is still intact (later scans skipped themselves): drop its marks, grow, replay
  const ScanCounters st = *h_status_;
  if (st.n_touched) {
    note_launch(), k_clear_touched<<<sm_count() * 8, TPB, 0, s>>>(grid.dev(), buf_, std::min<u32>(st.n_touched, (u32)(b_touched_.bytes / 4)));
    BNX_CUDA(cudaGetLastError());
    BNX_CUDA(cudaStreamSynchronize(s));
  }
  BNX_TRY(grid.recover(gc));
  if (st.overflow & OVF_TILES) {
    const u64 chunks = st.ray_chunk & CHUNK_FIELD;
    BNX_TRY(b_tiles_.reserve((size_t)(chunks / 32 + 64) * 4));
    buf_.tile_first = b_tiles_.as<u32>();
  }
  set_ = 0;
  // a scan with more ray chunks than the counters can hold is refused (reported below); the scans queued behind it are
  // replayed like after any other failure
  const bool refused = (st.overflow & OVF_CHUNKS) != 0u;
  for (size_t k = done + (refused ? 1 : 0); k < q.size(); ++k) {
    const Queued& e = q[k];
    BNX_TRY(reserve_scan(e.p.n, e.stride, e.p.max_range));
    const void* d_points = e.points;
    if (e.where == BNX_HOST && e.p.n > 0) {
      BNX_TRY(b_pts_.reserve((size_t)e.p.n * e.stride));
      BNX_CUDA(cudaMemcpyAsync(b_pts_.p, e.points, cloud_bytes(e.p.n, e.stride, e.f64), cudaMemcpyHostToDevice, s));
      d_points = b_pts_.p;
    }
    BNX_TRY(run_scan(d_points, e.stride, e.f64, e.p, false));  // keeps the update_id this scan was queued with
    if (k == done) ++counters[5];  // the attempt that froze the pipeline counts as a retry of this scan
  }
  if (refused) {
    deferred_ = BNX_ERR_UNSUPPORTED;
    deferred_msg_ = "insert: more than 2^32 ray chunks in one scan (that scan was dropped, the others are applied)";
  }
  return BNX_OK;
}

// ------------------------------------------------------------------------------------------------
// sharded map (one shard per process / GPU) — host side of the stages
// ------------------------------------------------------------------------------------------------
int Map::shard_config(int rank, int world) {
  BNX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "shard_config: bad rank/world");
  if (const char* e = std::getenv("BNX_PEER_TIMEOUT_MS")) {
    const unsigned long long ns = std::strtoull(e, nullptr, 10) * 1000000ull;
    if (ns) BNX_CUDA(cudaMemcpyToSymbol(g_peer_timeout_ns, &ns, sizeof(ns)));
  }
  rank_ = rank;
  world_ = world;
  if (world > 1 && !scratch_) {
    scratch_ = new Grid();
    const int st = scratch_->init(grid.resolution, 2, 3, 4);
    if (st != BNX_OK) {
      delete scratch_;
      scratch_ = nullptr;
      return st;
    }
  }
  return BNX_OK;
}

// where do this rank's records go? peer memory: block [rank] of every owner's inbox; otherwise the caller's send buffers
int Map::upload_boxes() {
  BNX_TRY(b_px_.reserve(sizeof(PeerBoxes)));
  buf_.px = b_px_.as<PeerBoxes>();
  if (px_uploaded_valid_ && std::memcmp(&px_uploaded_, &px_host_, sizeof(PeerBoxes)) == 0) return BNX_OK;
  px_uploaded_ = px_host_;
  px_uploaded_valid_ = true;
  BNX_CUDA(cudaMemcpyAsync(b_px_.p, &px_host_, sizeof(PeerBoxes), cudaMemcpyHostToDevice, grid.stream()));
  buf_.px = b_px_.as<PeerBoxes>();
  return BNX_OK;
}

int Map::shard_begin(const void* points, i64 stride_bytes, i64 n, bool f64, u32 index_base, const double origin[3], double max_range,
                     void* send_records, i64 cap_records, int where) {
  BNX_REQUIRE(world_ > 1 && scratch_, "shard_begin: call shard_config(rank, world > 1) first");
  BNX_REQUIRE(n >= 0 && n < (1ll << 24) && (u64)index_base + (u64)n < (1ull << 31), "shard_begin: point count / index out of range");
  BNX_REQUIRE(n_pending_ == 0, "shard_begin: addHitPoint/addMissPoint queues are not supported on a sharded map");
  staged_p2p_ = send_records == nullptr;
  if (staged_p2p_) {
    BNX_REQUIRE(p2p_ready_, "shard_begin: NULL send buffer but no mailboxes attached (bnx_map_shard_p2p_attach)");
    cap_records = mbox_cap_rec_;
    BNX_REQUIRE(n + 2 <= cap_records, "shard_begin: the mailbox holds fewer endpoint records than this slice has points");
  }
  BNX_REQUIRE(origin && cap_records >= 2, "shard_begin: null argument");
  // fleet step (Map::set_fleet): consumed here; nested drains below replay other scans with their own origins
  std::vector<double> fleet;
  fleet.swap(fleet_origins_);
  if (!fleet.empty()) {
    BNX_REQUIRE(std::isfinite(max_range) && max_range >= 0.0, "fleet step: needs a finite max_range");
    // no leaf may be touched by the rays of two sensors: reach balls apart by more than two leaf blocks
    const double need = 2.0 * max_range + 40.0 * grid.resolution;
    for (int a = 0; a < world_; ++a)
      for (int c = a + 1; c < world_; ++c) {
        double d2 = 0.0;
        for (int k = 0; k < 3; ++k) d2 += (fleet[a * 3 + k] - fleet[c * 3 + k]) * (fleet[a * 3 + k] - fleet[c * 3 + k]);
        if (!(d2 > need * need)) {
          set_error("fleet step: the reach of sensors " + std::to_string(a) + " and " + std::to_string(c) +
                    " overlaps; insert their scans one after the other");
          return BNX_ERR_UNSUPPORTED;
        }
      }
    origin = &fleet[(size_t)rank_ * 3];
  }
  BNX_REQUIRE(f64 ? (stride_bytes >= 24 && stride_bytes % 8 == 0) : (stride_bytes >= 12 && stride_bytes % 4 == 0), "shard_begin: bad stride");
  if (!queue_.empty()) BNX_TRY(drain());  // single-GPU pipeline first; the sharded queue is drained collectively
  cudaStream_t s = grid.stream();
  scratch_->set_stream(s);
  const i64 slots = (i64)world_ * cap_records;
  const bool lean = shard_async_;  // pipelined: the kernels leave tables and counters zeroed for the next scan
  // pipelined + peer memory: the front half runs on the pre-stream, one scratch set per exchange in flight (see map.hpp)
  const bool overlap = lean && staged_p2p_;
  const u32 xnext = xseq1_ + 1;
  if (overlap && slots > shard_sets_n_) {
    // all sets are sized together, up front (same decision on every rank: slots = world * mailbox capacity)
    const u32 keep_id = shard_async_id_;
    const i64 keep_n_max = shard_n_max_;
    BNX_TRY(drain());
    shard_async_ = true;
    shard_async_id_ = keep_id;
    shard_n_max_ = keep_n_max;
    staged_p2p_ = send_records == nullptr;
    for (set_ = 0; set_ < SHARD_SETS; ++set_) BNX_TRY(reserve_scan(slots, stride_bytes, max_range, cap_records));
    shard_sets_n_ = slots;
  }
  set_ = overlap ? (int)(xnext % (u32)SHARD_SETS) : 0;
  cudaStream_t fs = overlap ? pre_stream_ : s;
  // sender-side dedupe table: sized from the exchange capacity (>= n + 2), so it does not change from scan to scan
  BNX_TRY(reserve_scan(std::max<i64>(n, slots), stride_bytes, max_range, cap_records));
  const void* d_points = points;
  int stage_slot = -1;
  if (where == BNX_HOST && n > 0) {
    if (lean) {
      // staging ring on a copy stream: at most SHARD_QUEUE scans are in flight between two (collective) drains, so
      // slot id mod (SHARD_QUEUE + 2) is free when scan id is enqueued and the copy needs no dependency at all: it runs
      // as early as the host enqueues it. Sized from n_max, which is the same on every rank (growing it drains).
      stage_slot = (int)(shard_async_id_ % (u32)SHARD_STAGES);
      const size_t need = (size_t)std::max<i64>(n, shard_n_max_) * stride_bytes;
      if (need > x_stage_bytes_) {
        // draining may replay queued scans through nested (synchronous) shard_insert calls: keep this call's state
        const u32 keep_id = shard_async_id_;
        const i64 keep_n_max = shard_n_max_;
        BNX_TRY(drain());
        shard_async_ = true;
        shard_async_id_ = keep_id;
        shard_n_max_ = keep_n_max;
        staged_p2p_ = send_records == nullptr;
        BNX_TRY(reserve_scan(std::max<i64>(n, slots), stride_bytes, max_range, cap_records));  // buf_ points at the sender-side table again
        for (auto& st : x_stage_) BNX_TRY(st.reserve(need));
        x_stage_bytes_ = x_stage_[0].bytes;
      }
      BNX_CUDA(cudaMemcpyAsync(x_stage_[stage_slot].p, points, cloud_bytes(n, stride_bytes, f64), cudaMemcpyHostToDevice, copy_stream_));
      BNX_CUDA(cudaEventRecord(x_copied_[stage_slot], copy_stream_));
      BNX_CUDA(cudaStreamWaitEvent(fs, x_copied_[stage_slot], 0));
      d_points = x_stage_[stage_slot].p;
    } else {
      BNX_TRY(b_pts_.reserve((size_t)n * stride_bytes));
      BNX_CUDA(cudaMemcpyAsync(b_pts_.p, points, cloud_bytes(n, stride_bytes, f64), cudaMemcpyHostToDevice, s));
      d_points = b_pts_.p;
    }
  }
  ScanParams p = {};
  p.ox = origin[0];
  p.oy = origin[1];
  p.oz = origin[2];
  p.max_range = max_range;
  p.max_range_sqr = max_range * max_range;
  p.inv_res = grid.inv_resolution;
  p.Ox = (i32)std::floor(origin[0] * grid.inv_resolution);
  p.Oy = (i32)std::floor(origin[1] * grid.inv_resolution);
  p.Oz = (i32)std::floor(origin[2] * grid.inv_resolution);
  p.miss = options[0];
  p.hit = options[1];
  p.cmin = options[2];
  p.cmax = options[3];
  p.c = update_count;
  p.n = (u32)n;
  p.rank = (u32)rank_;
  p.world = (u32)world_;
  p.async_id = NONE;
  p.rec_cap = (u32)cap_records;
  p.xseq1 = ++xseq1_;
  p.par = staged_p2p_ ? (p.xseq1 & 1u) : 0u;
  p.max_chunks = (u32)std::min<u64>(((1ull << 40) - 1) / (u64)std::max<i64>(slots, 1), 1ull << 28);
  const double reach = std::ceil(max_range * grid.inv_resolution) + 4.0, lim = (double)(1 << 20) - 1.0;
  p.packed = std::isfinite(max_range) && max_range >= 0.0 && std::fabs((double)p.Ox) + reach < lim &&
             std::fabs((double)p.Oy) + reach < lim && std::fabs((double)p.Oz) + reach < lim;
  if (!fleet.empty()) {
    p.fleet = 1;
    for (int g = 0; g < world_; ++g)
      for (int k = 0; k < 3; ++k) {
        p.fO[g][k] = (i32)std::floor(fleet[(size_t)g * 3 + k] * grid.inv_resolution);
        if (!(std::fabs((double)p.fO[g][k]) + reach < lim)) p.packed = 0;  // the receiver sees the voxels of every sensor
      }
  }
  const u64 tslots = table_slots(cap_records);
  p.hash_mask = (u32)(tslots - 1);
  sp_ = p;
  sp_fleet_ = fleet;  // (set here: the nested drains above run other scans through this function)
  shard_retries_ = 0;
  shard_attempt_ = 0;
  if (profiling) cudaEventRecord(ev_[0], s);
  if (!staged_p2p_) {
    BNX_REQUIRE(world_ <= MAX_PEERS, "sharded insert: at most 16 ranks");
    px_host_ = PeerBoxes{};
    for (int o = 0; o < world_; ++o) px_host_.rec[o] = static_cast<int4*>(send_records) + (size_t)o * cap_records;
    BNX_TRY(upload_boxes());
    buf_.my_flags = nullptr;
  } else {
    buf_.my_flags = reinterpret_cast<const u32*>(mbox_);
  }
  if (overlap) {
    // inbox parity and scratch set of exchange x were last used by exchange x - 2 / x - 4: this rank's merge of x - 2 has
    // seen every rank's exchange-2 stamp, i.e. every owner is done with the inbox; everything older is done a fortiori
    const u32 prev = p.xseq1 - 2u;
    if (p.xseq1 >= 3u && x_merged_seq_[prev % SHARD_SETS] == prev) BNX_CUDA(cudaStreamWaitEvent(fs, x_merged_[prev % SHARD_SETS], 0));
  }
  if (!(lean && S().sc_clean && S().t1_clean)) BNX_CUDA(cudaMemsetAsync(d_sc_, 0, SC_BYTES + tslots * 12, fs));
  S().sc_clean = S().t1_clean = lean;  // pipelined: the apply epilogue zeroes the counters, k_shard_dedupe this table
  S().clean_slots = 0;
  const int blocks = blocks_for(n);
  if (n > 0) {
    const unsigned char* pts = static_cast<const unsigned char*>(d_points);
    if (f64) {
      launch_classify<true, false>(p.packed, blocks, fs, pts, (u32)stride_bytes, p, buf_);
    } else if (stride_bytes == 16 && (reinterpret_cast<uintptr_t>(pts) & 15u) == 0) {
      launch_classify<false, true>(p.packed, blocks, fs, pts, 16u, p, buf_);
    } else {
      launch_classify<false, false>(p.packed, blocks, fs, pts, (u32)stride_bytes, p, buf_);
    }
  }
  // always launched: its last block writes the block headers (counts) and, with mailboxes, the arrival stamps
  launch_scan_kernel(k_shard_bucket, blocks, TPB, fs, p, buf_, index_base);
  BNX_CUDA(cudaGetLastError());
  if (overlap) {  // the back half (map's stream) starts once this rank's own front half is complete
    BNX_CUDA(cudaEventRecord(x_begun_[p.xseq1 % SHARD_SETS], fs));
    BNX_CUDA(cudaStreamWaitEvent(s, x_begun_[p.xseq1 % SHARD_SETS], 0));
  }
  if (profiling) cudaEventRecord(ev_[1], s);
  return BNX_OK;
}

int Map::shard_resolve_mark(const void* recv_records, void* send_leaves, i64 cap_leaves) {
  BNX_REQUIRE(world_ > 1 && scratch_, "shard_resolve_mark: bad argument");
  if (staged_p2p_) {
    BNX_REQUIRE(p2p_ready_, "shard_resolve_mark: no mailboxes attached");
    recv_records = mbox_ + MBOX_HEADER + (size_t)sp_.par * world_ * mbox_cap_rec_ * 16;
    cap_leaves = mbox_cap_leaf_;
  } else {
    BNX_REQUIRE(recv_records && send_leaves && cap_leaves >= 2, "shard_resolve_mark: bad argument");
    for (int o = 0; o < world_; ++o) px_host_.leaf[o] = static_cast<int4*>(send_leaves) + (size_t)o * cap_leaves * 5;
    BNX_TRY(upload_boxes());
  }
  cudaStream_t s = grid.stream();
  ScanParams& p = sp_;
  const u32 slots = p.world * p.rec_cap;
  const int persistent = sm_count() * 8;
  BNX_TRY(b_touched_.reserve((size_t)grid.dev().leaf_cap * 4));
  BNX_TRY(b_touched2_.reserve((size_t)scratch_->dev().leaf_cap * 4));
  BNX_TRY(b_ray_src_.reserve((size_t)slots + 64));
  buf_.touched = b_touched_.as<u32>();
  buf_.touched2 = b_touched2_.as<u32>();
  buf_.ray_src = b_ray_src_.as<unsigned char>();
  buf_.recs = static_cast<const int4*>(recv_records);
  buf_.gate = nullptr;
  p.seq = ++seq_;
  p.xseq2 = ++xseq2_;
  p.tile_cap = (u32)std::min<size_t>(b_tiles_.bytes / 4, 0xFFFFFFFFull);
  p.touched_cap = (u32)std::min<size_t>(b_touched_.bytes / 4, 0xFFFFFFFFull);
  p.touched2_cap = (u32)std::min<size_t>(b_touched2_.bytes / 4, 0xFFFFFFFFull);
  p.leaf_cap2 = (u32)cap_leaves;
  // receiver-side dedupe table [table u32[t2] | keys u64[t2]]: large enough for the worst case (every record of every
  // rank lands here); the kernels only use — and, pipelined, zero again — the part the arrived records need
  const u64 t1 = table_slots(p.rec_cap), t2 = table_slots(slots);
  // packed keys: [table u32[t2] | keys u64[t2]] (12 B per slot); any coordinates: int4[t2] (16 B per slot)
  const void* before = b_table2_.p;
  BNX_TRY(b_table2_.reserve(t2 * 16));
  if (b_table2_.p != before || t2_packed_ != (p.packed != 0u)) t2_clean_ = false;
  t2_packed_ = p.packed != 0u;
  uint4* table1 = reinterpret_cast<uint4*>(S().table.as<unsigned char>() + SC_BYTES);
  buf_.table = b_table2_.as<u32>();
  buf_.keys = reinterpret_cast<unsigned long long*>(b_table2_.as<unsigned char>() + t2 * 4);
  p.hash_mask = (u32)(t2 - 1);
  const bool lean = shard_async_ && shard_attempt_ == 0;
  if (shard_attempt_ > 0) BNX_CUDA(cudaMemsetAsync(d_sc_, 0, SC_BYTES, s));  // a retry: counters of the failed attempt
  if (!(lean && t2_clean_)) BNX_CUDA(cudaMemsetAsync(b_table2_.p, 0, t2 * 16, s));
  t2_clean_ = lean;
  p.clean16 = lean ? 1u : 0u;
  ++shard_attempt_;
  ++shard_stats[0];
  const GridDev g = grid.dev(), gs = scratch_->dev();
  // a rank receives about one slice worth of records; the kernels loop if it is (much) more
  const int rblocks = blocks_for(std::min<i64>(slots, 2 * (i64)p.rec_cap));
  launch_scan_kernel(k_shard_dedupe, rblocks, TPB, s, p, buf_, slots, table1, lean ? (u32)(t1 * 12 / 16) : 0u);
  launch_scan_kernel(k_resolve<2, false>, rblocks, TPB, s, g, p, buf_, slots);
  if (profiling) cudaEventRecord(ev_[6], s);
  launch_scan_kernel(k_mark<true>, sm_count() * MARK_MIN_BLOCKS, TPB, s, g, gs, p, buf_);
  if (profiling) cudaEventRecord(ev_[7], s);
  launch_scan_kernel(k_shard_emit, persistent, TPB, s, gs, p, buf_);
  BNX_CUDA(cudaGetLastError());
  if (profiling) cudaEventRecord(ev_[2], s);
  return BNX_OK;
}

int Map::shard_merge(const void* recv_leaves, void* flags) {
  BNX_REQUIRE(world_ > 1 && scratch_, "shard_merge: bad argument");
  if (staged_p2p_) {
    recv_leaves = mbox_ + MBOX_HEADER + (size_t)world_ * mbox_cap_rec_ * 16 * 2;
  } else {
    BNX_REQUIRE(recv_leaves && flags, "shard_merge: bad argument");
  }
  cudaStream_t s = grid.stream();
  const GridDev g = grid.dev(), gs = scratch_->dev();
  launch_scan_kernel(k_shard_merge, sm_count() * 8, TPB, s, g, gs, sp_, buf_, static_cast<const int4*>(recv_leaves), sp_.leaf_cap2, static_cast<u32*>(flags));
  BNX_CUDA(cudaGetLastError());
  x_merged_seq_[sp_.xseq1 % SHARD_SETS] = 0;
  if (shard_async_ && staged_p2p_) {
    BNX_CUDA(cudaEventRecord(x_merged_[sp_.xseq1 % SHARD_SETS], s));
    x_merged_seq_[sp_.xseq1 % SHARD_SETS] = sp_.xseq1;
  }
  if (profiling) cudaEventRecord(ev_[3], s);
  return BNX_OK;
}

int Map::shard_finish(const void* flags_reduced, int* retry) {
  BNX_REQUIRE(world_ > 1 && scratch_ && retry, "shard_finish: bad argument");
  if (staged_p2p_) {
    flags_reduced = mbox_ + MBOX_FLAGS4 * 4;  // the apply pass reduces the flag table itself
  } else {
    BNX_REQUIRE(flags_reduced != nullptr, "shard_finish: bad argument");
  }
  cudaStream_t s = grid.stream();
  const int persistent = sm_count() * 8;
  const GridDev g = grid.dev();
  buf_.gate = static_cast<const u32*>(flags_reduced);
  launch_scan_kernel(apply_tma() ? k_apply_leaves<true> : k_apply_leaves<false>, sm_count() * APPLY_MIN_BLOCKS, TPB, s, g, sp_, buf_);
  BNX_CUDA(cudaGetLastError());
  if (profiling) cudaEventRecord(ev_[4], s);
  BNX_CUDA(cudaMemcpyAsync(h_status_, d_sc_, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
  BNX_CUDA(cudaStreamSynchronize(s));
  buf_.gate = nullptr;
  shard_phase_times();
  const ScanCounters st = *h_status_;
  const u32 any_pool = st.gate_pool, any_ovf = st.gate_ovf;
  *retry = 0;
  if ((st.gc.error | any_pool) & ERR_PEER) {
    set_error("sharded insert: a peer rank did not reach an exchange point in time");
    return BNX_ERR_CUDA;
  }
  if (any_pool | any_ovf) {
    // some rank ran short: nobody applied. Every rank drops this attempt's marks; the short ones grow.
    ++shard_stats[4];
    if (++shard_retries_ > 48) {
      set_error("sharded insert: node pools could not be grown enough for this scan");
      return BNX_ERR_NOMEM;
    }
    if (st.n_touched) {
      note_launch(), k_clear_touched<<<persistent, TPB, 0, s>>>(g, buf_, std::min(st.n_touched, sp_.touched_cap));
      BNX_CUDA(cudaGetLastError());
      BNX_CUDA(cudaStreamSynchronize(s));
    }
    if (st.gc.error) BNX_TRY(grid.recover(st.gc));
    GridCounters sgc;
    BNX_TRY(scratch_->read_counters(&sgc));
    if (sgc.error) BNX_TRY(scratch_->recover(sgc));
    if (any_ovf & OVF_CHUNKS) {  // refused on every rank (the flags are reduced); the shards are unchanged and clean
      set_error("insert: more than 2^32 ray chunks in one scan");
      return BNX_ERR_UNSUPPORTED;
    }
    if (st.overflow & OVF_TILES) {
      const u64 chunks = st.ray_chunk & CHUNK_FIELD;
      BNX_TRY(b_tiles_.reserve((size_t)(chunks / 32 + 64) * 4));
      buf_.tile_first = b_tiles_.as<u32>();
    }
    // bit 0: repeat from shard_resolve_mark; bits 8..: exchange buffers that were too small (caller grows them)
    *retry = 1 | (int)((any_ovf & (OVF_RECORDS | OVF_LEAVES)) << 8);
    return BNX_OK;
  }
  counters[0] = sp_.n;
  counters[1] = st.n_endpoints;
  counters[2] = (i64)st.sum_m;  // ray cells of the rays this rank cast; the caller adds N once over all ranks
  counters[3] = (i64)st.n_changed;
  counters[4] = st.n_touched;
  counters[5] = shard_retries_;
  counters[6] = (i64)(st.ray_chunk >> 40);
  counters[7] = (i64)(st.ray_chunk & CHUNK_FIELD);
  for (int k = 0; k < (sp_.fleet ? world_ : 1); ++k)  // a fleet step stands for world_ inserts (probabilistic_map.cpp:103-105)
    if (++update_count == 4) update_count = 1;
  ++shard_stats[7];
  note_leaf_fill(st.gate_fill);
  GridCounters sgc;
  BNX_TRY(scratch_->read_counters(&sgc));
  BNX_TRY(scratch_->maintain(sgc));
  return grid.maintain(st.gc);
}

// device time of the stages of the last sharded scan: {0, begin, resolve_mark (waits for exchange 1), merge (waits for
// exchange 2), apply (waits for the flags), total}; valid once the stream has been synchronised
void Map::shard_phase_times() {
  if (!profiling) return;
  float ms = 0.f;
  phase_us[0] = 0.0;
  for (int k = 0; k < 4; ++k) {
    if (cudaEventElapsedTime(&ms, ev_[k], ev_[k + 1]) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    phase_us[k + 1] = ms * 1e3;
  }
  cudaEventElapsedTime(&ms, ev_[0], ev_[4]);
  phase_us[5] = ms * 1e3;
  // inside resolve_mark: [6] = wait(exchange 1) + dedupe + resolve, [7] = mark (the rest of the stage is the emit kernel)
  if (cudaEventElapsedTime(&ms, ev_[1], ev_[6]) == cudaSuccess) phase_us[6] = ms * 1e3;
  if (cudaEventElapsedTime(&ms, ev_[6], ev_[7]) == cudaSuccess) phase_us[7] = ms * 1e3;
  cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// peer-memory exchange: mailboxes
// ------------------------------------------------------------------------------------------------
void Map::p2p_close_peers() {
  for (int o = 0; o < MAX_PEERS; ++o) {
    if (peer_ipc_[o] && peer_base_[o]) cudaIpcCloseMemHandle(peer_base_[o]);
    peer_base_[o] = nullptr;
    peer_ipc_[o] = false;
  }
  p2p_ready_ = false;
}

int Map::p2p_alloc(i64 cap_records, i64 cap_leaves, void* ipc_handle64, void** local_ptr) {
  BNX_REQUIRE(world_ > 1 && world_ <= MAX_PEERS, "p2p_alloc: call shard_config(rank, 2..16) first");
  BNX_REQUIRE(cap_records >= 2 && cap_leaves >= 2, "p2p_alloc: capacities must be >= 2");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  BNX_TRY(drain());
  BNX_CUDA(cudaStreamSynchronize(grid.stream()));
  p2p_close_peers();
  if (mbox_) mbox_retired_.push_back(mbox_);
  mbox_ = nullptr;
  const size_t bytes = MBOX_HEADER + (size_t)world_ * ((size_t)cap_records * 16 * 2 + (size_t)cap_leaves * 80);  // two endpoint inboxes (parity)
  void* ptr = nullptr;
  BNX_CUDA(cudaMalloc(&ptr, bytes));
  BNX_CUDA(cudaMemset(ptr, 0, MBOX_HEADER));
  BNX_CUDA(cudaDeviceSynchronize());
  mbox_ = static_cast<unsigned char*>(ptr);
  mbox_cap_rec_ = cap_records;
  mbox_cap_leaf_ = cap_leaves;
  if (ipc_handle64) {
    cudaIpcMemHandle_t h;
    BNX_CUDA(cudaIpcGetMemHandle(&h, ptr));
    std::memcpy(ipc_handle64, &h, 64);
  }
  if (local_ptr) *local_ptr = ptr;
  return BNX_OK;
}

int Map::p2p_attach(const void* handles, void* const* local_ptrs) {
  BNX_REQUIRE(mbox_ != nullptr, "p2p_attach: call p2p_alloc first");
  BNX_REQUIRE(handles != nullptr || local_ptrs != nullptr, "p2p_attach: null handles");
  p2p_close_peers();
  for (int o = 0; o < world_; ++o) {
    if (o == rank_) {
      peer_base_[o] = mbox_;
    } else if (local_ptrs) {
      BNX_REQUIRE(local_ptrs[o] != nullptr, "p2p_attach: null mailbox pointer");
      peer_base_[o] = local_ptrs[o];
    } else {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)o * 64, 64);
      void* ptr = nullptr;
      BNX_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      peer_base_[o] = ptr;
      peer_ipc_[o] = true;
    }
  }
  px_host_ = PeerBoxes{};
  for (int o = 0; o < world_; ++o) {
    unsigned char* base = static_cast<unsigned char*>(peer_base_[o]);
    px_host_.flag[o] = reinterpret_cast<u32*>(base);
    px_host_.rec[o] = reinterpret_cast<int4*>(base + MBOX_HEADER) + (size_t)rank_ * mbox_cap_rec_;
    px_host_.leaf[o] = reinterpret_cast<int4*>(base + MBOX_HEADER + (size_t)world_ * mbox_cap_rec_ * 16 * 2) + (size_t)rank_ * mbox_cap_leaf_ * 5;
  }
  BNX_TRY(upload_boxes());
  BNX_CUDA(cudaStreamSynchronize(grid.stream()));
  p2p_ready_ = true;
  return BNX_OK;
}

// native driver: every rank (re)creates its mailbox and the IPC handles travel through one NCCL all-gather.
// Collective: all ranks call it at the same point of the protocol with the same capacities.
int Map::p2p_collective_setup(i64 cap_records, i64 cap_leaves) {
  cudaStream_t s = grid.stream();
  if (shard_debug()) std::fprintf(stderr, "[bnx rank %d] mailbox setup: %lld records, %lld leaves per sender\n", rank_, (long long)cap_records, (long long)cap_leaves);
  unsigned char mine[64];
  BNX_TRY(p2p_alloc(cap_records, cap_leaves, mine, nullptr));
  ++shard_stats[2];
  shard_stats[5] = cap_leaves;
  if (host_gather_) {
    std::vector<unsigned char> all((size_t)world_ * 64);
    if (host_gather_(host_gather_ctx_, mine, all.data(), 64) != 0) {
      set_error("sharded map: the caller's all-gather failed while exchanging the mailbox handles");
      return BNX_ERR_CUDA;
    }
    return p2p_attach(all.data(), nullptr);
  }
  const NcclApi& api = nccl_api(nullptr);
  BNX_TRY(x_handles_.reserve((size_t)world_ * 64));
  BNX_CUDA(cudaMemcpyAsync(x_handles_.as<unsigned char>() + (size_t)rank_ * 64, mine, 64, cudaMemcpyHostToDevice, s));
  BNX_NCCL(api, api.AllGather(x_handles_.as<unsigned char>() + (size_t)rank_ * 64, x_handles_.p, 64, ncclChar, static_cast<ncclComm_t>(comm_), s));
  std::vector<unsigned char> all((size_t)world_ * 64);
  BNX_CUDA(cudaMemcpyAsync(all.data(), x_handles_.p, all.size(), cudaMemcpyDeviceToHost, s));
  BNX_CUDA(cudaStreamSynchronize(s));
  return p2p_attach(all.data(), nullptr);
}

// ---- native driver of the sharded protocol (NCCL resolved at run time)
int Map::nccl_unique_id(const char* nccl_path, void* out128) {
  const NcclApi& api = nccl_api(nccl_path);
  if (!api.ok) {
    set_error("NCCL could not be loaded (libnccl.so.2)");
    return BNX_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  BNX_NCCL(api, api.GetUniqueId(&id));
  std::memcpy(out128, &id, 128);
  return BNX_OK;
}

int Map::shard_comm_init(const char* nccl_path, const void* unique_id128, int rank, int world) {
  BNX_REQUIRE(unique_id128 != nullptr, "shard_comm_init: null id");
  const NcclApi& api = nccl_api(nccl_path);
  if (!api.ok) {
    set_error("NCCL could not be loaded (libnccl.so.2)");
    return BNX_ERR_UNSUPPORTED;
  }
  BNX_TRY(shard_config(rank, world));
  ncclUniqueId id;
  std::memcpy(&id, unique_id128, 128);
  ncclComm_t comm = nullptr;
  BNX_NCCL(api, api.CommInitRank(&comm, world, id, rank));
  comm_ = comm;
  BNX_TRY(x_flags_.reserve(64));
  // data path: peer memory over NVLink unless told otherwise; the mailboxes are created by the first insert
  const char* ex = std::getenv("BNX_SHARD_EXCHANGE");
  want_p2p_ = !(ex && std::strcmp(ex, "nccl") == 0) && world <= MAX_PEERS;
  return BNX_OK;
}

int Map::shard_host_init(int rank, int world, AllGatherFn fn, void* ctx) {
  BNX_REQUIRE(fn != nullptr, "shard_host_init: null all-gather callback");
  BNX_REQUIRE(world >= 2 && world <= MAX_PEERS, "shard_host_init: 2..16 ranks");
  BNX_TRY(shard_config(rank, world));
  host_gather_ = fn;
  host_gather_ctx_ = ctx;
  want_p2p_ = true;
  BNX_TRY(x_flags_.reserve(64));
  return BNX_OK;
}

// the leaf inboxes are doubled as soon as a scan filled one of them half way (decided from a value that is the same on
// every rank, at a point every rank reaches with the same scans behind it: the mailboxes are replaced collectively by
// the next insert), so that an overflow — a frozen pipeline and a synchronous replay — stays the exception
void Map::note_leaf_fill(u32 fill) {
  shard_stats[6] = std::max<i64>(shard_stats[6], fill);
  while ((i64)fill * 2 > cap_leaf_) cap_leaf_ *= 2;
}

// block o of `send` goes to rank o, block r of `recv` comes from rank r: grouped ncclSend/ncclRecv over NVLink
int Map::all_to_all(const void* send, void* recv, size_t block_bytes) {
  const NcclApi& api = nccl_api(nullptr);
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  cudaStream_t s = grid.stream();
  BNX_NCCL(api, api.GroupStart());
  for (int peer = 0; peer < world_; ++peer) {
    BNX_NCCL(api, api.Send(static_cast<const char*>(send) + (size_t)peer * block_bytes, block_bytes, ncclChar, peer, comm, s));
    BNX_NCCL(api, api.Recv(static_cast<char*>(recv) + (size_t)peer * block_bytes, block_bytes, ncclChar, peer, comm, s));
  }
  BNX_NCCL(api, api.GroupEnd());
  return BNX_OK;
}

int Map::shard_insert(const void* points, i64 stride_bytes, i64 n, bool f64, u32 index_base, i64 n_max, const double origin[3], double max_range,
                      int where, bool async) {
  BNX_REQUIRE((comm_ != nullptr || host_gather_ != nullptr) && world_ > 1, "shard_insert: call shard_comm_init / shard_host_init first");
  BNX_REQUIRE(n_max >= n, "shard_insert: n_max must be the largest slice of the scan over all ranks");
  // the fleet origins armed for THIS scan are set aside first: the drains below may replay queued scans, which arm (and
  // consume) their own
  std::vector<double> armed;
  armed.swap(fleet_origins_);
  const NcclApi& api = nccl_api(nullptr);
  cudaStream_t s = grid.stream();
  if (!async) BNX_TRY(drain());
  if (async && squeue_.size() >= SHARD_QUEUE) BNX_TRY(drain());  // same count on every rank: draining stays collective
  // equal-split exchange buffers (all ranks compute the same capacities from n_max)
  const i64 want_rec = std::max<i64>(cap_rec_, n_max + 2);
  if (want_p2p_) {
    // mailboxes: created (and, after an overflow, replaced by larger ones) by all ranks together
    if (!p2p_ready_ || want_rec > mbox_cap_rec_ || cap_leaf_ > mbox_cap_leaf_) {
      if (async) BNX_TRY(drain());
      cap_rec_ = want_rec;
      BNX_TRY(p2p_collective_setup(cap_rec_, cap_leaf_));
    }
  } else {
    if (want_rec != cap_rec_ || !x_send1_.p) {
      if (async) BNX_TRY(drain());
      cap_rec_ = want_rec;
      BNX_TRY(x_send1_.reserve((size_t)world_ * cap_rec_ * 16));
      BNX_TRY(x_recv1_.reserve((size_t)world_ * cap_rec_ * 16));
    }
    if ((size_t)world_ * cap_leaf_ * 80 > x_send2_.bytes) {
      if (async) BNX_TRY(drain());
      BNX_TRY(x_send2_.reserve((size_t)world_ * cap_leaf_ * 80));
      BNX_TRY(x_recv2_.reserve((size_t)world_ * cap_leaf_ * 80));
    }
  }
  const bool p2p = want_p2p_;
  PdlScope pdl(async);
  u32* flags = x_flags_.as<u32>();
  const u32 my_async = async ? async_next_++ : NONE;
  shard_async_ = async;
  shard_async_id_ = my_async;
  shard_n_max_ = n_max;
  fleet_origins_.swap(armed);
  BNX_TRY(shard_begin(points, stride_bytes, n, f64, index_base, origin, max_range, p2p ? nullptr : x_send1_.p, cap_rec_, where));
  sp_.async_id = my_async;
  if (!p2p) BNX_TRY(all_to_all(x_send1_.p, x_recv1_.p, (size_t)cap_rec_ * 16));
  for (;;) {
    BNX_TRY(shard_resolve_mark(p2p ? nullptr : x_recv1_.p, p2p ? nullptr : x_send2_.p, cap_leaf_));
    if (!p2p) BNX_TRY(all_to_all(x_send2_.p, x_recv2_.p, (size_t)cap_leaf_ * 80));
    BNX_TRY(shard_merge(p2p ? nullptr : x_recv2_.p, flags));
    if (!p2p) BNX_NCCL(api, api.AllReduce(flags, flags, 4, ncclUint32, ncclMax, static_cast<ncclComm_t>(comm_), s));
    if (async) {
      buf_.gate = p2p ? reinterpret_cast<const u32*>(mbox_) + MBOX_FLAGS4 : flags;
      launch_scan_kernel(apply_tma() ? k_apply_leaves<true> : k_apply_leaves<false>, sm_count() * APPLY_MIN_BLOCKS, TPB, s, grid.dev(), sp_, buf_);
      BNX_CUDA(cudaGetLastError());
      if (profiling) cudaEventRecord(ev_[4], s);
      buf_.gate = nullptr;
      shard_async_ = false;
      ShardQueued q;
      q.points = points;
      q.stride = stride_bytes;
      q.n = n;
      q.n_max = n_max;
      q.f64 = f64;
      q.index_base = index_base;
      q.async_id = my_async;
      q.c = sp_.c;
      std::memcpy(q.origin, origin, sizeof(q.origin));
      q.max_range = max_range;
      q.where = where;
      q.set = set_;
      q.fleet = sp_fleet_;
      squeue_.push_back(q);
      for (int k = 0; k < (sp_.fleet ? world_ : 1); ++k)
        if (++update_count == 4) update_count = 1;
      return BNX_OK;
    }
    int retry = 0;
    shard_async_ = false;
    BNX_TRY(shard_finish(p2p ? nullptr : flags, &retry));
    if (!retry) return BNX_OK;
    if (retry & (int)(OVF_RECORDS << 8)) {
      set_error("sharded insert: endpoint record exchange overflowed (n_max too small)");
      return BNX_ERR_INVALID;
    }
    if (retry & (int)(OVF_LEAVES << 8)) {  // every rank saw the same reduced flags: same growth everywhere
      cap_leaf_ *= 4;
      if (p2p) {
        // the mailboxes are replaced, the received endpoint records with them: start the scan over (nothing was applied)
        fleet_origins_ = sp_fleet_;
        return shard_insert(points, stride_bytes, n, f64, index_base, n_max, origin, max_range, where, false);
      }
      BNX_TRY(x_send2_.reserve((size_t)world_ * cap_leaf_ * 80));
      BNX_TRY(x_recv2_.reserve((size_t)world_ * cap_leaf_ * 80));
    }
  }
}

int Map::shard_drain() {
  cudaStream_t s = grid.stream();
  if (shard_debug()) std::fprintf(stderr, "[bnx rank %d] shard_drain: %zu queued, next id %u\n", rank_, squeue_.size(), async_next_);
  GridCounters gc;
  BNX_TRY(grid.read_counters(&gc));  // synchronises the stream
  BNX_CUDA(cudaStreamSynchronize(pre_stream_));
  {
    // the counters of the scan that froze the pipeline are in ITS scratch set (a healthy queue: the newest scan's)
    const void* sc = d_sc_;
    if (gc.error)
      for (const ShardQueued& e : squeue_)
        if (e.async_id == gc.failed_id) sc = sets_[e.set].table.p;
    BNX_CUDA(cudaMemcpyAsync(h_status_, sc, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s));
    BNX_CUDA(cudaStreamSynchronize(s));
  }
  shard_phase_times();
  if (gc.error) {  // frozen kernels cleaned nothing
    t2_clean_ = false;
    for (auto& st : sets_) st.sc_clean = st.t1_clean = false;
  }
  std::vector<ShardQueued> q;
  q.swap(squeue_);
  size_t done = q.size();
  if (gc.error) {
    done = 0;
    while (done < q.size() && q[done].async_id != gc.failed_id) ++done;
  }
  ++shard_stats[3];
  shard_stats[7] += (i64)done;
  for (size_t k = 0; k < done; ++k) {
    const AsyncRecord& r = h_ring_[q[k].async_id & (RING - 1)];
    drain_max_fill_ = std::max(drain_max_fill_, r.leaf_fill);
    counters[0] = q[k].n;
    counters[1] = r.n_endpoints;
    counters[2] = (i64)r.sum_m;
    counters[3] = r.n_changed;
    counters[4] = r.n_touched;
    counters[5] = 0;
    counters[6] = (i64)(r.ray_chunk >> 40);
    counters[7] = (i64)(r.ray_chunk & CHUNK_FIELD);
    for (int j = 0; j < 4; ++j) totals[j] += counters[j];
  }
  note_leaf_fill(drain_max_fill_);  // the records carry the max over ranks: every rank takes the same decision here
  if (shard_debug())
    std::fprintf(stderr, "[bnx rank %d] shard_drain: synced, error %u, done %zu, fill %u, cap_leaf %lld\n", rank_, gc.error, done, drain_max_fill_, (long long)cap_leaf_);
  drain_max_fill_ = 0;
  GridCounters sgc;
  BNX_TRY(scratch_->read_counters(&sgc));
  if (!gc.error) {
    // head-room until the next collective drain: three times what the window that just ended allocated (a fleet step adds
    // the leaves of `world` scans; the scratch grid mirrors most of the foreign map), so that a queued scan running short —
    // a frozen pipeline and a synchronous replay on every rank — stays the exception
    const u64 w_leaves = gc.n_leaves > drain_prev_[0] ? gc.n_leaves - drain_prev_[0] : 0, w_inner = gc.n_inner > drain_prev_[1] ? gc.n_inner - drain_prev_[1] : 0;
    const u64 s_leaves = sgc.n_leaves > drain_prev_[2] ? sgc.n_leaves - drain_prev_[2] : 0, s_inner = sgc.n_inner > drain_prev_[3] ? sgc.n_inner - drain_prev_[3] : 0;
    drain_prev_[0] = gc.n_leaves;
    drain_prev_[1] = gc.n_inner;
    drain_prev_[2] = sgc.n_leaves;
    drain_prev_[3] = sgc.n_inner;
    BNX_TRY(scratch_->maintain(sgc, 3 * s_leaves, 3 * s_inner));
    return grid.maintain(gc, 3 * w_leaves, 3 * w_inner);
  }
  ++shard_stats[1];
  if (gc.error & ERR_PEER) {
    set_error("sharded insert: a peer rank did not reach an exchange point in time");
    return BNX_ERR_CUDA;
  }
  // every rank is frozen at the same scan (the flags were all-reduced): drop its marks, grow what was short on
  // this rank, then replay the rest of the queue with synchronous (collective) inserts
  const ScanCounters st = *h_status_;
  if (st.n_touched) {
    note_launch(), k_clear_touched<<<sm_count() * 8, TPB, 0, s>>>(grid.dev(), buf_, std::min<u32>(st.n_touched, (u32)(b_touched_.bytes / 4)));
    BNX_CUDA(cudaGetLastError());
    BNX_CUDA(cudaStreamSynchronize(s));
  }
  BNX_TRY(grid.recover(gc));
  if (sgc.error) BNX_TRY(scratch_->recover(sgc));
  if ((st.overflow | gc.failed_ovf) & OVF_CHUNKS) {  // refused on every rank; the scans queued behind it are dropped with it
    set_error("insert: more than 2^32 ray chunks in one scan");
    return BNX_ERR_UNSUPPORTED;
  }
  if ((st.overflow | gc.failed_ovf) & OVF_TILES) {
    const u64 chunks = st.ray_chunk & CHUNK_FIELD;
    BNX_TRY(b_tiles_.reserve((size_t)(std::max<u64>(chunks, b_tiles_.bytes / 4 * 32) * 2 / 32 + 64) * 4));
    buf_.tile_first = b_tiles_.as<u32>();
  }
  if (gc.failed_ovf & OVF_LEAVES) cap_leaf_ *= 4;
  const u32 resume = update_count;
  for (size_t k = done; k < q.size(); ++k) {
    const ShardQueued& e = q[k];
    update_count = e.c;
    fleet_origins_ = e.fleet;
    BNX_TRY(shard_insert(e.points, e.stride, e.n, e.f64, e.index_base, e.n_max, e.origin, e.max_range, e.where, false));
  }
  update_count = resume;
  return BNX_OK;
}

int Map::add_point(const double pt[3], bool miss) {
  BNX_TRY(drain());
  cudaStream_t s = grid.stream();
  if ((size_t)(n_pending_ + 1) * sizeof(int4) > b_pending_.bytes) {
    // grow, keeping the queue
    DevBuf bigger;
    BNX_TRY(bigger.reserve(b_pending_.bytes * 2));
    BNX_CUDA(cudaMemcpyAsync(bigger.p, b_pending_.p, (size_t)n_pending_ * sizeof(int4), cudaMemcpyDeviceToDevice, s));
    BNX_CUDA(cudaStreamSynchronize(s));
    std::swap(bigger.p, b_pending_.p);
    std::swap(bigger.bytes, b_pending_.bytes);
    buf_.pending = b_pending_.as<int4>();
  }
  ScanParams p = {};
  p.miss = options[0];
  p.hit = options[1];
  p.cmin = options[2];
  p.cmax = options[3];
  p.c = update_count;
  const int4 e = make_int4((i32)std::floor(pt[0] * grid.inv_resolution), (i32)std::floor(pt[1] * grid.inv_resolution),
                           (i32)std::floor(pt[2] * grid.inv_resolution), miss ? 1 : 0);
  u32* d_flag = reinterpret_cast<u32*>(d_sc_);  // scratch word, rewritten by the next scan anyway
  for (int attempt = 0; attempt < 8; ++attempt) {
    note_launch(), k_add_point<<<1, 1, 0, s>>>(grid.dev(), p, buf_, e, n_pending_, d_flag);
    BNX_CUDA(cudaGetLastError());
    BNX_CUDA(cudaMemcpyAsync(h_status_, d_sc_, sizeof(u32), cudaMemcpyDeviceToHost, s));
    GridCounters gc;
    BNX_TRY(grid.read_counters(&gc));
    if (gc.error == 0) {
      u32 queued;
      std::memcpy(&queued, h_status_, 4);
      n_pending_ += queued;
      return BNX_OK;
    }
    BNX_TRY(grid.recover(gc));
  }
  set_error("add_point: node pools could not be grown");
  return BNX_ERR_NOMEM;
}

int Map::query(const i32* xyz, i64 n, int kind, u8* out, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || (xyz && out)), "query: null input");
  BNX_REQUIRE(kind == BNX_OCCUPIED || kind == BNX_UNKNOWN || kind == BNX_FREE, "query: unknown kind");
  if (n == 0) return BNX_OK;
  BNX_TRY(drain());
  cudaStream_t s = grid.stream();
  const i32* dx = xyz;
  u8* dout = out;
  if (where == BNX_HOST) {
    BNX_TRY(b_q_xyz_.reserve((size_t)n * 12));
    BNX_TRY(b_q_out_.reserve((size_t)n));
    BNX_CUDA(cudaMemcpyAsync(b_q_xyz_.p, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, s));
    dx = b_q_xyz_.as<i32>();
    dout = b_q_out_.as<u8>();
  }
  note_launch(), k_query<<<std::min(blocks_for(n), sm_count() * 8), TPB, 0, s>>>(grid.dev(), dx, n, kind, options[4], dout);
  BNX_CUDA(cudaGetLastError());
  if (where == BNX_HOST) {
    BNX_CUDA(cudaMemcpyAsync(out, dout, (size_t)n, cudaMemcpyDeviceToHost, s));
    BNX_CUDA(cudaStreamSynchronize(s));
  }
  return BNX_OK;
}

}  // namespace bnx
