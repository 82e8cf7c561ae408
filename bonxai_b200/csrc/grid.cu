// VoxelGrid<DataT> on the device: batched accessor operations, iteration, maintenance.
// Reference behaviour: bonxai_core/include/bonxai/bonxai.hpp (cited per function below).
#include "grid.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace bnx {

// ------------------------------------------------------------------------------------------------
// error text / device info
// ------------------------------------------------------------------------------------------------
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
const char* get_error() { return t_error.c_str(); }

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace {

constexpr int TPB = 256;
constexpr i64 SUB_BATCH = 1ll << 22;  // elements per dedupe pass (bounds the batch-local hash)

inline int grid_for(i64 n, int tpb = TPB, int waves = 8) {
  const i64 blocks = ceil_div(n, tpb);
  const i64 cap = (i64)sm_count() * waves;
  return (int)std::max<i64>(1, std::min(blocks, cap));
}

size_t env_mb(const char* name, size_t dflt_mb) {
  const char* s = std::getenv(name);
  if (!s || !*s) return dflt_mb << 20;
  return (size_t)std::strtoull(s, nullptr, 10) << 20;
}

__device__ __forceinline__ void copy_cell(void* dst, const void* src, u32 bytes) {
  switch (bytes) {
    case 4: *static_cast<u32*>(dst) = *static_cast<const u32*>(src); break;
    case 8: *static_cast<u64*>(dst) = *static_cast<const u64*>(src); break;
    case 16: *static_cast<uint4*>(dst) = *static_cast<const uint4*>(src); break;
    case 2: *static_cast<uint16_t*>(dst) = *static_cast<const uint16_t*>(src); break;
    case 1: *static_cast<u8*>(dst) = *static_cast<const u8*>(src); break;
    default:
      for (u32 b = 0; b < bytes; ++b) static_cast<u8*>(dst)[b] = static_cast<const u8*>(src)[b];
  }
}
__device__ __forceinline__ void zero_cell(void* dst, u32 bytes) {
  switch (bytes) {
    case 4: *static_cast<u32*>(dst) = 0u; break;
    case 8: *static_cast<u64*>(dst) = 0ull; break;
    case 16: *static_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0); break;
    default:
      for (u32 b = 0; b < bytes; ++b) static_cast<u8*>(dst)[b] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// kernels: locate / dedupe
// ------------------------------------------------------------------------------------------------
// Accessor::getLeafGrid(coord, create_if_missing), bonxai.hpp:588-621, one coordinate per thread.
template <bool CREATE>
__global__ void __launch_bounds__(TPB) k_locate(GridDev g, const i32* __restrict__ xyz, i64 n, u32* __restrict__ loc) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const int x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    loc[i] = CREATE ? leaf_find_or_create(g, x, y, z) : leaf_find(g, x, y, z);
  }
}

// Batch-local hash keyed by (leaf, cell): finds, for every distinct cell of the batch, the first and the
// last batch index that names it — the information sequential Accessor semantics depend on.
__global__ void __launch_bounds__(TPB) k_dedupe(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc, u32 n,
                                                 unsigned long long* keys, u32* first, u32* last, u32* slot_of, u32 mask) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 leaf = loc[i];
    if (leaf == NONE) {
      slot_of[i] = NONE;
      continue;
    }
    const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const unsigned long long key = ((unsigned long long)leaf << 12) | ci;
    u32 slot = (u32)mix64(key) & mask;
    for (;;) {
      unsigned long long k = keys[slot];
      if (k == key) break;
      if (k == ~0ull) {
        k = atomicCAS(&keys[slot], ~0ull, key);
        if (k == ~0ull || k == key) break;
      }
      slot = (slot + 1) & mask;
    }
    atomicMin(&first[slot], i);
    atomicMax(&last[slot], i);
    slot_of[i] = slot;
  }
}

// ------------------------------------------------------------------------------------------------
// kernels: accessor operations
// ------------------------------------------------------------------------------------------------
// Accessor::setValue, bonxai.hpp:449-466. The thread holding the LAST index of a cell stores the value
// and turns the cell on; it also reports the pre-batch state for the FIRST index. Every other repeat
// sees an ON cell, as it would sequentially.
__global__ void __launch_bounds__(TPB) k_set_values(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc,
                                                     const u32* __restrict__ slot_of, const u32* __restrict__ first,
                                                     const u32* __restrict__ last, const u8* __restrict__ vals,
                                                     u8* __restrict__ was_on, u32 n) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of[i];
    const u32 f = first[slot], l = last[slot];
    if (i == l) {
      const u32 leaf = loc[i];
      const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      copy_cell(leaf_cells(g, leaf) + (size_t)ci * g.cell_bytes, vals + (size_t)i * g.cell_bytes, g.cell_bytes);
      const unsigned long long bit = 1ull << (ci & 63);
      const unsigned long long old =
          atomicOr(reinterpret_cast<unsigned long long*>(leaf_active(g, leaf) + (ci >> 6)), bit);
      if (was_on) was_on[f] = (old & bit) != 0;
    }
    if (was_on && i != f) was_on[i] = 1;
  }
}

// Accessor::setCellOn(coord, default_value), bonxai.hpp:537-554
__global__ void __launch_bounds__(TPB) k_set_on(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc,
                                                 const u32* __restrict__ slot_of, const u32* __restrict__ first,
                                                 const u8* __restrict__ dflt, u8* __restrict__ was_on, u32 n) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of[i];
    if (first[slot] == i) {
      const u32 leaf = loc[i];
      const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      const unsigned long long bit = 1ull << (ci & 63);
      const unsigned long long old =
          atomicOr(reinterpret_cast<unsigned long long*>(leaf_active(g, leaf) + (ci >> 6)), bit);
      if (!(old & bit)) copy_cell(leaf_cells(g, leaf) + (size_t)ci * g.cell_bytes, dflt, g.cell_bytes);
      if (was_on) was_on[i] = (old & bit) != 0;
    } else if (was_on) {
      was_on[i] = 1;
    }
  }
}

// Accessor::setCellOff, bonxai.hpp:557-569
__global__ void __launch_bounds__(TPB) k_set_off(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc,
                                                  const u32* __restrict__ slot_of, const u32* __restrict__ first,
                                                  u8* __restrict__ was_on, u32 n) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of[i];
    u8 w = 0;
    if (slot != NONE && first[slot] == i) {
      const u32 leaf = loc[i];
      const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      const unsigned long long bit = 1ull << (ci & 63);
      const unsigned long long old =
          atomicAnd(reinterpret_cast<unsigned long long*>(leaf_active(g, leaf) + (ci >> 6)), ~bit);
      w = (old & bit) != 0;
    }
    if (was_on) was_on[i] = w;
  }
}

// Accessor::value(coord, true), bonxai.hpp:469-494 — creation half (first index of each cell)
__global__ void __launch_bounds__(TPB) k_create_cells(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc,
                                                       const u32* __restrict__ slot_of, const u32* __restrict__ first, u32 n) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (first[slot_of[i]] != i) continue;
    const u32 leaf = loc[i];
    const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    const unsigned long long bit = 1ull << (ci & 63);
    const unsigned long long old = atomicOr(reinterpret_cast<unsigned long long*>(leaf_active(g, leaf) + (ci >> 6)), bit);
    if (!(old & bit)) zero_cell(leaf_cells(g, leaf) + (size_t)ci * g.cell_bytes, g.cell_bytes);
  }
}

// ConstAccessor::value / isCellOn, bonxai.hpp:496-534 (loc == NONE: leaf missing)
__global__ void __launch_bounds__(TPB) k_read_values(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc,
                                                      u8* __restrict__ vals, u8* __restrict__ found, i64 n) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const u32 leaf = loc[i];
    bool on = false;
    if (leaf != NONE) {
      const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      on = (leaf_active(g, leaf)[ci >> 6] >> (ci & 63)) & 1ull;
      if (on && vals) copy_cell(vals + (size_t)i * g.cell_bytes, leaf_cells(g, leaf) + (size_t)ci * g.cell_bytes, g.cell_bytes);
    }
    if (found) found[i] = on;
  }
}

// write through a previously returned value pointer: only ON cells, last index wins
__global__ void __launch_bounds__(TPB) k_update_values(GridDev g, const i32* __restrict__ xyz, const u32* __restrict__ loc,
                                                        const u32* __restrict__ slot_of, const u32* __restrict__ last,
                                                        const u8* __restrict__ vals, u32 n) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const u32 slot = slot_of[i];
    if (slot == NONE || last[slot] != i) continue;
    const u32 leaf = loc[i];
    const u32 ci = leaf_index(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if ((leaf_active(g, leaf)[ci >> 6] >> (ci & 63)) & 1ull)
      copy_cell(leaf_cells(g, leaf) + (size_t)ci * g.cell_bytes, vals + (size_t)i * g.cell_bytes, g.cell_bytes);
  }
}

// posToCoord / coordToPos, bonxai.hpp:404-417 — one rounded fp64 multiply, floor, cast
__global__ void __launch_bounds__(TPB) k_pos_to_coord(const double* __restrict__ p, i64 n3, double inv_res, i32* __restrict__ out) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n3; i += (i64)gridDim.x * blockDim.x)
    out[i] = __double2int_rd(__dmul_rn(p[i], inv_res));
}
__global__ void __launch_bounds__(TPB) k_coord_to_pos(const i32* __restrict__ c, i64 n3, double res, double* __restrict__ out) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n3; i += (i64)gridDim.x * blockDim.x)
    out[i] = __dmul_rn((double)c[i], res);
}

// ------------------------------------------------------------------------------------------------
// kernels: whole-grid passes (one warp per leaf; masks walked with popc, bonxai.hpp:689-743)
// ------------------------------------------------------------------------------------------------
// activeCellsCount, bonxai.hpp:689-701
__global__ void __launch_bounds__(TPB) k_count_active(GridDev g, u32 n_leaves, unsigned long long* total) {
  unsigned long long local = 0;
  const u64 items = (u64)n_leaves * g.mask_words;
  for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < items; t += (u64)gridDim.x * blockDim.x) {
    const u32 leaf = (u32)(t / g.mask_words), w = (u32)(t % g.mask_words);
    const unsigned char* lp = leaf_ptr(g, leaf);
    if (reinterpret_cast<const int4*>(lp)->w & 1) local += __popcll(reinterpret_cast<const u64*>(lp + g.off_active)[w]);
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ unsigned long long s[TPB / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long b = 0;
    for (int k = 0; k < TPB / 32; ++k) b += s[k];
    if (b) atomicAdd(total, b);
  }
}

// Order-independent digest of forEachCell's (coord, value) pairs: {sum, xor, count} of
// mix64(hash3(x, y, z) + FNV-1a(value bytes) * 0x9E3779B97F4A7C15) over every ON cell. Equal digests <=> equal dumps (up to a
// 2^-64 collision), and digests of disjoint shards add up (sum, count) / xor (xor): full-size parity checks and the
// sharded bench compare maps without moving them to the host. One warp per leaf, coalesced cell rows.
__global__ void __launch_bounds__(TPB) k_digest(GridDev g, u32 n_leaves, unsigned long long* out3) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  const u32 cells = 1u << (3 * g.lb), lm = (1u << g.lb) - 1u;
  unsigned long long sum = 0, x = 0, cnt = 0;
  for (u32 leaf = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); leaf < n_leaves; leaf += warps) {
    const unsigned char* lp = leaf_ptr(g, leaf);
    const int4 hdr = *reinterpret_cast<const int4*>(lp);
    if (!(hdr.w & 1)) continue;
    const u64* act = reinterpret_cast<const u64*>(lp + g.off_active);
    for (u32 ci = lane; ci < cells; ci += 32) {
      if (!((act[ci >> 6] >> (ci & 63)) & 1ull)) continue;
      const unsigned char* v = lp + g.off_cells + (size_t)ci * g.cell_bytes;
      u64 f = 0xCBF29CE484222325ull;
      if (g.cell_bytes == 4) {
        const u32 w = *reinterpret_cast<const u32*>(v);
        for (int k = 0; k < 4; ++k) f = (f ^ ((w >> (8 * k)) & 0xFFu)) * 0x100000001B3ull;
      } else {
        for (u32 k = 0; k < g.cell_bytes; ++k) f = (f ^ v[k]) * 0x100000001B3ull;
      }
      const i32 cx = hdr.x | (i32)(ci & lm), cy = hdr.y | (i32)((ci >> g.lb) & lm), cz = hdr.z | (i32)((ci >> (2 * g.lb)) & lm);
      const u64 h = mix64(hash3(cx, cy, cz) + f * 0x9E3779B97F4A7C15ull);
      sum += h;
      x ^= h;
      ++cnt;
    }
  }
  for (int o = 16; o; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    x ^= __shfl_xor_sync(0xffffffffu, x, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0 && cnt) {
    atomicAdd(&out3[0], sum);
    atomicXor(&out3[1], x);
    atomicAdd(&out3[2], cnt);
  }
}

// forEachCell (bonxai.hpp:704-743) / getOccupiedVoxels / getFreeVoxels (probabilistic_map.cpp:108-126)
// as a compaction: per-leaf predicate masks -> popcount -> block prefix -> one atomicAdd per block.
// pred < 0: every ON cell. pred 0/2: CellT word with probability_log > / < thr.
// fpos: the publisher post-step of the ROS caller fused in (bonxai_ros/src/bonxai_server.cpp:217-251): voxel corner
// coord*resolution in fp64, kept if zmin <= z <= zmax, stored as float xyz (+ w = 1.0f for a 16-byte pcl::PointXYZ).
struct DumpFilter {
  float* fpos;
  u32 fstride;  // floats per output point: 3 or 4
  int zfilter;
  double zmin, zmax;
};

__global__ void __launch_bounds__(TPB) k_dump(GridDev g, u32 n_leaves, int pred, i32 thr, double res, i32* __restrict__ xyz,
                                               double* __restrict__ pos, u8* __restrict__ vals, unsigned long long cap,
                                               unsigned long long* total, DumpFilter flt) {
  __shared__ u32 s_cnt[TPB / 32];
  __shared__ unsigned long long s_base;
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 W = g.mask_words;
  const u32 rounds = (n_leaves + gridDim.x * (TPB / 32) - 1) / (gridDim.x * (TPB / 32));
  for (u32 r = 0; r < rounds; ++r) {
    const u32 leaf = (r * gridDim.x + blockIdx.x) * (TPB / 32) + warp;
    u64 mine0 = 0, mine1 = 0;  // predicate mask of word `lane` and `lane+32`
    int4 hdr = make_int4(0, 0, 0, 0);
    const bool live = leaf < n_leaves && ((hdr = *reinterpret_cast<const int4*>(leaf_ptr(g, leaf))).w & 1);
    if (live) {
      const u64* act = leaf_active(g, leaf);
      if (pred < 0) {
        if (lane < W) mine0 = act[lane];
        if (lane + 32 < W) mine1 = act[lane + 32];
      } else {
        const u32* cells = reinterpret_cast<const u32*>(leaf_cells(g, leaf));
        for (u32 w = 0; w < W; ++w) {
          const u64 m = act[w];
          if (m == 0) continue;
          u64 pm = 0;
          for (u32 half = 0; half < 2; ++half) {
            const u32 bit = half * 32 + lane;
            bool p = false;
            if ((m >> bit) & 1ull) {
              const i32 prob = (i32)cells[w * 64 + bit] >> 4;
              p = pred == BNX_OCCUPIED ? prob > thr : prob < thr;
              if (p && flt.zfilter) {
                const u32 ci = w * 64 + bit;
                const double z = __dmul_rn((double)(hdr.z | (i32)((ci >> (2 * g.lb)) & ((1u << g.lb) - 1u))), res);
                p = z >= flt.zmin && z <= flt.zmax;
              }
            }
            pm |= (u64)__ballot_sync(0xffffffffu, p) << (half * 32);
          }
          if (lane == (w & 31)) (w < 32 ? mine0 : mine1) = pm;
        }
      }
    }
    u32 cnt = __popcll(mine0) + __popcll(mine1);
    u32 incl = cnt;  // inclusive prefix over lanes (word order)
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (u32)o) incl += v;
    }
    const u32 leaf_total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) s_cnt[warp] = leaf_total;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 sum = 0;
      for (int k = 0; k < TPB / 32; ++k) {
        const u32 c = s_cnt[k];
        s_cnt[k] = sum;
        sum += c;
      }
      s_base = sum ? atomicAdd(total, (unsigned long long)sum) : 0ull;
    }
    __syncthreads();
    if (leaf_total && (xyz || pos || vals || flt.fpos)) {
      // output order inside a leaf is lane-major (word l, word l+32, word l+1, ...): the order of a dump
      // is unspecified, only the set of (coord, value) pairs matters.
      const unsigned long long base = s_base + s_cnt[warp];
      const u32 excl0 = incl - cnt;               // cells of this leaf written before word `lane`
      const u32 excl1 = excl0 + __popcll(mine0);  // ... before word `lane + 32`
      for (u32 w = 0; w < W; ++w) {
        const u64 pm = __shfl_sync(0xffffffffu, w < 32 ? mine0 : mine1, w & 31);
        if (pm == 0) continue;
        const u32 wbase = __shfl_sync(0xffffffffu, w < 32 ? excl0 : excl1, w & 31);
        for (u32 half = 0; half < 2; ++half) {
          const u32 bit = half * 32 + lane;
          if (!((pm >> bit) & 1ull)) continue;
          const unsigned long long o = base + wbase + __popcll(pm & ((1ull << bit) - 1ull));
          if (o >= cap) continue;
          const u32 ci = w * 64 + bit;
          const u32 lm = (1u << g.lb) - 1u;
          const i32 cx = hdr.x | (i32)(ci & lm), cy = hdr.y | (i32)((ci >> g.lb) & lm), cz = hdr.z | (i32)((ci >> (2 * g.lb)) & lm);
          if (xyz) {
            xyz[3 * o] = cx;
            xyz[3 * o + 1] = cy;
            xyz[3 * o + 2] = cz;
          }
          if (pos) {  // coordToPos, bonxai.hpp:412-417
            pos[3 * o] = __dmul_rn((double)cx, res);
            pos[3 * o + 1] = __dmul_rn((double)cy, res);
            pos[3 * o + 2] = __dmul_rn((double)cz, res);
          }
          if (flt.fpos) {  // PCLPoint(voxel.x(), voxel.y(), voxel.z()): double -> float, round to nearest
            float* q = flt.fpos + (size_t)o * flt.fstride;
            q[0] = __double2float_rn(__dmul_rn((double)cx, res));
            q[1] = __double2float_rn(__dmul_rn((double)cy, res));
            q[2] = __double2float_rn(__dmul_rn((double)cz, res));
            if (flt.fstride == 4) q[3] = 1.0f;
          }
          if (vals) copy_cell(vals + (size_t)o * g.cell_bytes, leaf_cells(g, leaf) + (size_t)ci * g.cell_bytes, g.cell_bytes);
        }
      }
    }
    __syncthreads();
  }
}

// clear(SET_ALL_CELLS_OFF), bonxai.hpp:678-687: every ON bit drops, values stay
__global__ void __launch_bounds__(TPB) k_masks_off(GridDev g, u32 n_leaves) {
  const u64 items = (u64)n_leaves * g.mask_words;
  for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < items; t += (u64)gridDim.x * blockDim.x)
    leaf_active(g, (u32)(t / g.mask_words))[t % g.mask_words] = 0ull;
}

// re-insert every published root slot of the old table into a bigger one
__global__ void __launch_bounds__(TPB) k_rehash(const int4* __restrict__ old_tab, u64 old_slots, int4* new_tab, u32 new_mask, int shift) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < old_slots; i += (u64)gridDim.x * blockDim.x) {
    const int4 v = old_tab[i];
    if ((u32)v.w < 2u) continue;
    u32 slot = (u32)hash3(v.x >> shift, v.y >> shift, v.z >> shift) & new_mask;
    for (;;) {
      u32* state = reinterpret_cast<u32*>(&new_tab[slot].w);
      if (atomicCAS(state, 0u, (u32)v.w) == 0u) {
        new_tab[slot].x = v.x;
        new_tab[slot].y = v.y;
        new_tab[slot].z = v.z;
        break;
      }
      slot = (slot + 1) & new_mask;
    }
  }
}

// releaseUnusedMemory, bonxai.hpp:367-387 — step 1: one thread per (inner node, child): leaves whose mask is
// all OFF are unlinked and pushed on the free list
__global__ void __launch_bounds__(TPB) k_release_leaves(GridDev g, u32 n_inner, u32 children) {
  const u64 items = (u64)n_inner * children;
  for (u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; t < items; t += (u64)gridDim.x * blockDim.x) {
    const u32 inner = (u32)(t / children), ii = (u32)(t % children);
    u32* node = inner_ptr(g, inner);
    if (!(node[3] & 1u)) continue;
    const u32 v = node[g.inner_child_off + ii];
    if (v < 2u) continue;
    const u32 leaf = v - 2u;
    const u64* act = leaf_active(g, leaf);
    u64 any = 0;
    for (u32 w = 0; w < g.mask_words; ++w) any |= act[w];
    if (any) continue;
    node[g.inner_child_off + ii] = 0u;
    atomicAnd(reinterpret_cast<unsigned long long*>(node + 4) + (ii >> 6), ~(1ull << (ii & 63)));
    reinterpret_cast<int4*>(leaf_ptr(g, leaf))->w = 2;  // marked: to be zeroed + listed by step 2
  }
}
// step 2: zero the released leaves (a recycled leaf must look freshly mapped) and list them
__global__ void __launch_bounds__(TPB) k_release_zero(GridDev g, u32 n_leaves) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  for (u32 leaf = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); leaf < n_leaves; leaf += warps) {
    unsigned char* lp = leaf_ptr(g, leaf);
    if (reinterpret_cast<const int4*>(lp)->w != 2) continue;
    __syncwarp();
    uint4* q = reinterpret_cast<uint4*>(lp);
    for (u32 k = lane; k < g.leaf_stride / 16; k += 32) q[k] = make_uint4(0, 0, 0, 0);
    if (lane == 0) g.free_list[atomicAdd(&g.ctr->n_free, 1)] = leaf;
  }
}
// step 3: roots whose inner node lost every child are dropped while the table is rebuilt
__global__ void __launch_bounds__(TPB) k_rebuild_roots(GridDev g, const int4* __restrict__ old_tab, u64 old_slots, int4* new_tab,
                                                        u32 new_mask, u32 inner_mask_words) {
  const int shift = g.ib + g.lb;
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < old_slots; i += (u64)gridDim.x * blockDim.x) {
    const int4 v = old_tab[i];
    if ((u32)v.w < 2u) continue;
    u32* node = inner_ptr(g, (u32)v.w - 2u);
    const u64* im = reinterpret_cast<const u64*>(node + 4);
    u64 any = 0;
    for (u32 w = 0; w < inner_mask_words; ++w) any |= im[w];
    if (!any) {
      node[3] = 0u;  // dead inner node (not recycled; 288 B)
      atomicSub(&g.ctr->n_roots, 1u);
      continue;
    }
    u32 slot = (u32)hash3(v.x >> shift, v.y >> shift, v.z >> shift) & new_mask;
    for (;;) {
      u32* state = reinterpret_cast<u32*>(&new_tab[slot].w);
      if (atomicCAS(state, 0u, (u32)v.w) == 0u) {
        new_tab[slot].x = v.x;
        new_tab[slot].y = v.y;
        new_tab[slot].z = v.z;
        break;
      }
      slot = (slot + 1) & new_mask;
    }
  }
}

// Serialize, bonxai_core/include/bonxai/serialization.hpp:77-116 — one warp per root (inner node). The body of a
// root is {int32 key[3]; u64 inner_mask[Wi]; per ON child in ascending index: u64 leaf_mask[W], then the ON cells'
// raw bytes in ascending leaf index}. Root order is unspecified in the reference (unordered_map order); here it
// is the order in which warps reserve their range with one atomicAdd. out == nullptr: size query.
__device__ __forceinline__ void put_bytes(u8* dst, const void* src, u32 n) {
  const u8* s8 = static_cast<const u8*>(src);
  for (u32 i = 0; i < n; ++i) dst[i] = s8[i];
}

__global__ void __launch_bounds__(TPB) k_serialize(GridDev g, u32 n_inner, u32 children, u32 Wi, u8* out, unsigned long long* total,
                                                    unsigned long long* n_roots) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = gridDim.x * (TPB / 32);
  const u32 W = g.mask_words;
  for (u32 inner = blockIdx.x * (TPB / 32) + (threadIdx.x >> 5); inner < n_inner; inner += warps) {
    const u32* node = inner_ptr(g, inner);
    if (!(node[3] & 1u)) continue;
    // pass 1: bytes of this root
    unsigned long long bytes = 0;
    for (u32 c = lane; c < children; c += 32) {
      const u32 v = node[g.inner_child_off + c];
      if (v >= 2u) {
        const u64* act = leaf_active(g, v - 2u);
        u32 cnt = 0;
        for (u32 w = 0; w < W; ++w) cnt += __popcll(act[w]);
        bytes += (unsigned long long)W * 8 + (unsigned long long)cnt * g.cell_bytes;
      }
    }
    for (int o = 16; o; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    bytes += 12 + (unsigned long long)Wi * 8;
    unsigned long long base = 0;
    if (lane == 0) {
      base = atomicAdd(total, bytes);
      atomicAdd(n_roots, 1ull);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!out) continue;
    // pass 2: write. Children are laid out in ascending index: running offset over rounds of 32 children.
    u8* dst = out + base;
    if (lane == 0) put_bytes(dst, node, 12);
    if (lane < Wi * 2) put_bytes(dst + 12 + lane * 4, node + 4 + lane, 4);
    for (u32 k = lane + 32; k < Wi * 2; k += 32) put_bytes(dst + 12 + k * 4, node + 4 + k, 4);
    unsigned long long run = 12 + (unsigned long long)Wi * 8;
    for (u32 c0 = 0; c0 < children; c0 += 32) {
      const u32 c = c0 + lane;
      u32 leaf = NONE, cnt = 0;
      if (c < children) {
        const u32 v = node[g.inner_child_off + c];
        if (v >= 2u) {
          leaf = v - 2u;
          const u64* act = leaf_active(g, leaf);
          for (u32 w = 0; w < W; ++w) cnt += __popcll(act[w]);
        }
      }
      const unsigned long long mine = leaf != NONE ? (unsigned long long)W * 8 + (unsigned long long)cnt * g.cell_bytes : 0ull;
      unsigned long long incl = mine;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += t;
      }
      if (leaf != NONE) {
        u8* d = dst + run + (incl - mine);
        const u64* act = leaf_active(g, leaf);
        put_bytes(d, act, W * 8);
        d += W * 8;
        const u8* cells = leaf_cells(g, leaf);
        for (u32 w = 0; w < W; ++w) {
          u64 m = act[w];
          while (m) {
            const u32 bit = __ffsll((long long)m) - 1;
            m &= m - 1;
            put_bytes(d, cells + (size_t)(w * 64 + bit) * g.cell_bytes, g.cell_bytes);
            d += g.cell_bytes;
          }
        }
      }
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------
Grid::~Grid() {
  if (own_stream_) cudaStreamSynchronize(own_stream_);
  if (root_) cudaFree(root_);
  if (d_ctr_) cudaFree(d_ctr_);
  if (h_ctr_) cudaFreeHost(h_ctr_);
  if (d_count_) cudaFree(d_count_);
  if (h_count_) cudaFreeHost(h_count_);
  if (free_list_) cudaFree(free_list_);
  if (own_stream_) cudaStreamDestroy(own_stream_);
}

int Grid::init(double voxel_size, int ib, int lb, int cbytes) {
  // VoxelGrid ctor, bonxai.hpp:389-402
  BNX_REQUIRE(ib >= 1 && lb >= 1, "The minimum value of the inner_bits and leaf_bits should be 1");
  BNX_REQUIRE(lb <= 4 && ib <= 5,
              "leaf_bits <= 4 and inner_bits <= 5 are supported (the reference accepts any value >= 1, bonxai.hpp:138, but its own "
              "heap-backed Mask double-frees on destruction for leaf_bits >= 4, so larger leaves are untested upstream as well)");
  BNX_REQUIRE(cbytes >= 1 && cbytes <= 64, "cell_bytes must be in 1..64");
  BNX_REQUIRE(voxel_size > 0.0, "voxel_size must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: bonxai_b200 has no CPU fallback");
    return BNX_ERR_CUDA;
  }
  BNX_CUDA(cudaGetDevice(&device));
  resolution = voxel_size;
  inv_resolution = 1.0 / voxel_size;  // bonxai.hpp:395
  inner_bits = ib;
  leaf_bits = lb;
  cell_bytes = cbytes;
  BNX_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
  stream_ = own_stream_;

  const u32 cells = 1u << (3 * lb), children = 1u << (3 * ib);
  const u32 W = std::max(1u, cells / 64u), Wi = std::max(1u, children / 64u);
  GridDev& d = dev_;
  d.ib = ib;
  d.lb = lb;
  d.cell_bytes = (u32)cbytes;
  d.mask_words = W;
  d.off_active = 64;
  d.off_stamp = 16;
  d.off_touched = 64 + (u32)round_up(W * 8, 64);
  d.off_hit = d.off_touched + W * 8;
  d.off_cells = (u32)round_up(d.off_hit + W * 8, 128);
  d.leaf_stride = (u32)round_up(d.off_cells + (size_t)cells * cbytes, 128);
  d.inner_child_off = 4 + 2 * Wi;
  d.inner_stride = (u32)round_up(d.inner_child_off + children, 4);

  size_t free_b = 0, total_b = 0;
  BNX_CUDA(cudaMemGetInfo(&free_b, &total_b));
  BNX_TRY(leaf_arena_.init(total_b));
  BNX_TRY(inner_arena_.init(std::max<size_t>(total_b / 4, 1ull << 30)));
  BNX_CUDA(cudaMalloc(&d_ctr_, sizeof(GridCounters)));
  BNX_CUDA(cudaMallocHost(&h_ctr_, sizeof(GridCounters)));
  std::memset(h_ctr_, 0, sizeof(GridCounters));
  h_ctr_->failed_id = NONE;
  BNX_CUDA(cudaMemcpyAsync(d_ctr_, h_ctr_, sizeof(GridCounters), cudaMemcpyHostToDevice, stream_));
  BNX_CUDA(cudaStreamSynchronize(stream_));
  BNX_CUDA(cudaMalloc(&d_count_, 64));
  BNX_CUDA(cudaMallocHost(&h_count_, 64));
  d.ctr = d_ctr_;
  d.leaf = static_cast<unsigned char*>(leaf_arena_.base());
  d.inner = static_cast<u32*>(inner_arena_.base());
  BNX_TRY(grow_root_table(1ull << 14));
  BNX_TRY(ensure_inner_capacity(env_mb("BNX_INIT_INNER_MB", 16) / (d.inner_stride * 4)));
  BNX_TRY(ensure_leaf_capacity(env_mb("BNX_INIT_LEAF_MB", 1024) / d.leaf_stride));
  return sync();
}

int Grid::sync() {
  BNX_CUDA(cudaStreamSynchronize(stream_));
  return BNX_OK;
}

u64 Grid::leaf_step(u64 live_leaves) const {
  static const size_t step_bytes = env_mb("BNX_GROW_MB", 256);
  return std::max<u64>(step_bytes / dev_.leaf_stride, live_leaves / 8) + 1;
}
u64 Grid::inner_step(u64 live_inner) const { return std::max<u64>((16u << 20) / (dev_.inner_stride * 4), live_inner / 8) + 1; }

int Grid::ensure_leaf_capacity(u64 leaves, cudaStream_t zero_stream) {
  if (leaves <= dev_.leaf_cap) return BNX_OK;
  leaves = std::min<u64>(leaves, 0xFFFFFFF0ull);
  BNX_TRY(leaf_arena_.grow_to((size_t)leaves * dev_.leaf_stride, zero_stream ? zero_stream : stream_));
  dev_.leaf_cap = (u32)std::min<u64>(leaf_arena_.mapped() / dev_.leaf_stride, 0xFFFFFFF0ull);
  return BNX_OK;
}

int Grid::ensure_inner_capacity(u64 inner, cudaStream_t zero_stream) {
  if (inner <= dev_.inner_cap) return BNX_OK;
  BNX_TRY(inner_arena_.grow_to((size_t)inner * dev_.inner_stride * 4, zero_stream ? zero_stream : stream_));
  dev_.inner_cap = (u32)std::min<u64>(inner_arena_.mapped() / (dev_.inner_stride * 4), 0xFFFFFFF0ull);
  return BNX_OK;
}

int Grid::grow_root_table(u64 min_slots) {
  u64 slots = std::max<u64>(root_slots_, 1ull << 10);
  while (slots < min_slots) slots <<= 1;
  if (slots == root_slots_) return BNX_OK;
  BNX_REQUIRE(slots <= (1ull << 31), "root table too large");
  int4* fresh = nullptr;
  BNX_CUDA(cudaMalloc(&fresh, slots * sizeof(int4)));
  BNX_CUDA(cudaMemsetAsync(fresh, 0, slots * sizeof(int4), stream_));
  if (root_) {
    note_launch(), k_rehash<<<grid_for((i64)root_slots_), TPB, 0, stream_>>>(root_, root_slots_, fresh, (u32)(slots - 1), inner_bits + leaf_bits);
    BNX_CUDA(cudaGetLastError());
    BNX_CUDA(cudaStreamSynchronize(stream_));
    BNX_CUDA(cudaFree(root_));
  }
  root_ = fresh;
  root_slots_ = slots;
  dev_.root = root_;
  dev_.root_mask = (u32)(slots - 1);
  return BNX_OK;
}

int Grid::read_counters(GridCounters* out) {
  BNX_CUDA(cudaMemcpyAsync(h_ctr_, d_ctr_, sizeof(GridCounters), cudaMemcpyDeviceToHost, stream_));
  BNX_CUDA(cudaStreamSynchronize(stream_));
  *out = *h_ctr_;
  return BNX_OK;
}

int Grid::recover(const GridCounters& seen) {
  // counters that ran past their pool are clamped: the overshooting allocations were never linked
  GridCounters fix = seen;
  fix.n_leaves = std::min(seen.n_leaves, dev_.leaf_cap);
  fix.n_inner = std::min(seen.n_inner, dev_.inner_cap);
  fix.error = 0;
  fix.failed_id = NONE;
  fix.failed_ovf = 0;
  fix.done_blocks = 0;
  if (fix.n_free < 0) fix.n_free = 0;
  *h_ctr_ = fix;
  BNX_CUDA(cudaMemcpyAsync(d_ctr_, h_ctr_, sizeof(GridCounters), cudaMemcpyHostToDevice, stream_));
  BNX_CUDA(cudaStreamSynchronize(stream_));
  // a pool that actually ran out: two steps at once (the caller repeats until the work fits)
  if (seen.error & ERR_LEAF_POOL) BNX_TRY(ensure_leaf_capacity((u64)dev_.leaf_cap + std::max<u64>(2 * leaf_step(dev_.leaf_cap), dev_.leaf_cap / 2)));
  if (seen.error & ERR_INNER_POOL) BNX_TRY(ensure_inner_capacity((u64)dev_.inner_cap + std::max<u64>(2 * inner_step(dev_.inner_cap), dev_.inner_cap / 2)));
  if (seen.error & ERR_ROOT_TABLE) BNX_TRY(grow_root_table(root_slots_ * 4));
  return BNX_OK;
}

int Grid::maintain(const GridCounters& seen, u64 extra_leaves, u64 extra_inner) {
  // root table: under 50 % load now, and still under ~75 % if as many roots arrive again as there are inner nodes of head-room asked for
  while (((u64)seen.n_roots + extra_inner) * 4 > root_slots_ * 3 || (u64)seen.n_roots * 2 > root_slots_) BNX_TRY(grow_root_table(root_slots_ * 4));
  // keep half a growth step of each pool free so that the next batch rarely needs the retry path
  const u64 ls = leaf_step(seen.n_leaves), is = inner_step(seen.n_inner);
  if ((u64)seen.n_leaves + ls / 2 + extra_leaves > dev_.leaf_cap) BNX_TRY(ensure_leaf_capacity((u64)seen.n_leaves + ls + extra_leaves));
  if ((u64)seen.n_inner + is / 2 + extra_inner > dev_.inner_cap) BNX_TRY(ensure_inner_capacity((u64)seen.n_inner + is + extra_inner));
  return BNX_OK;
}

int Grid::free_list_reserve() {
  if (free_list_cap_ >= dev_.leaf_cap) return BNX_OK;
  u32* fresh = nullptr;
  BNX_CUDA(cudaMalloc(&fresh, (size_t)dev_.leaf_cap * 4));
  if (free_list_) {
    BNX_CUDA(cudaMemcpyAsync(fresh, free_list_, free_list_cap_ * 4, cudaMemcpyDeviceToDevice, stream_));
    BNX_CUDA(cudaStreamSynchronize(stream_));
    BNX_CUDA(cudaFree(free_list_));
  }
  free_list_ = fresh;
  free_list_cap_ = dev_.leaf_cap;
  dev_.free_list = free_list_;
  return BNX_OK;
}

// ------------------------------------------------------------------------------------------------
// batched accessor operations
// ------------------------------------------------------------------------------------------------
int Grid::stage_in(const void* src, size_t bytes, int where, DevBuf& buf, const void** dptr) {
  if (where == BNX_DEVICE || src == nullptr) {
    *dptr = src;
    return BNX_OK;
  }
  BNX_TRY(buf.reserve(bytes));
  BNX_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, stream_));
  *dptr = buf.p;
  return BNX_OK;
}

template <bool CREATE>
int Grid::locate(const i32* d_xyz, i64 n, u32* d_loc) {
  for (int attempt = 0; attempt < 40; ++attempt) {
    note_launch(), k_locate<CREATE><<<grid_for(n), TPB, 0, stream_>>>(dev_, d_xyz, n, d_loc);
    BNX_CUDA(cudaGetLastError());
    if constexpr (!CREATE) return BNX_OK;
    GridCounters c;
    BNX_TRY(read_counters(&c));
    if (c.error == 0) return maintain(c);
    BNX_TRY(recover(c));
  }
  set_error("node pools could not be grown enough for this batch");
  return BNX_ERR_NOMEM;
}

int Grid::dedupe(const i32* d_xyz, const u32* d_loc, i64 n) {
  u64 slots = 1024;
  while (slots < (u64)n * 2) slots <<= 1;
  BNX_TRY(b_keys_.reserve(slots * 8));
  BNX_TRY(b_first_.reserve(slots * 4));
  BNX_TRY(b_last_.reserve(slots * 4));
  BNX_TRY(b_slot_.reserve((size_t)n * 4));
  BNX_CUDA(cudaMemsetAsync(b_keys_.p, 0xFF, slots * 8, stream_));
  BNX_CUDA(cudaMemsetAsync(b_first_.p, 0xFF, slots * 4, stream_));
  BNX_CUDA(cudaMemsetAsync(b_last_.p, 0x00, slots * 4, stream_));
  note_launch(), k_dedupe<<<grid_for(n), TPB, 0, stream_>>>(dev_, d_xyz, d_loc, (u32)n, b_keys_.as<unsigned long long>(), b_first_.as<u32>(),
                                              b_last_.as<u32>(), b_slot_.as<u32>(), (u32)(slots - 1));
  BNX_CUDA(cudaGetLastError());
  return BNX_OK;
}

// common driver: stage inputs of one sub-batch, locate, optionally dedupe, run `body`, copy flags back
#define BNX_SUBBATCH_LOOP(n) for (i64 off = 0, cnt = 0; off < (n) && ((cnt = std::min<i64>(SUB_BATCH, (n)-off)), true); off += cnt)

int Grid::set_values(const i32* xyz, const void* values, i64 n, u8* was_on, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || (xyz && values)), "set_values: null input");
  BNX_SUBBATCH_LOOP(n) {
    const void *dx, *dv;
    BNX_TRY(stage_in(xyz + 3 * off, (size_t)cnt * 12, where, b_xyz_, &dx));
    BNX_TRY(stage_in(static_cast<const u8*>(values) + off * cell_bytes, (size_t)cnt * cell_bytes, where, b_val_, &dv));
    u8* dflag = nullptr;
    if (was_on) {
      if (where == BNX_DEVICE) {
        dflag = was_on + off;
      } else {
        BNX_TRY(b_flag_.reserve((size_t)cnt));
        dflag = b_flag_.as<u8>();
      }
    }
    BNX_TRY(b_loc_.reserve((size_t)cnt * 4));
    BNX_TRY(locate<true>(static_cast<const i32*>(dx), cnt, b_loc_.as<u32>()));
    BNX_TRY(dedupe(static_cast<const i32*>(dx), b_loc_.as<u32>(), cnt));
    note_launch(), k_set_values<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), b_slot_.as<u32>(),
                                                      b_first_.as<u32>(), b_last_.as<u32>(), static_cast<const u8*>(dv), dflag, (u32)cnt);
    BNX_CUDA(cudaGetLastError());
    if (was_on && where == BNX_HOST) BNX_CUDA(cudaMemcpyAsync(was_on + off, dflag, (size_t)cnt, cudaMemcpyDeviceToHost, stream_));
    if (where == BNX_HOST) BNX_TRY(sync());
  }
  return BNX_OK;
}

int Grid::get_values(const i32* xyz, i64 n, void* values, u8* found, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || xyz), "get_values: null input");
  BNX_SUBBATCH_LOOP(n) {
    const void* dx;
    BNX_TRY(stage_in(xyz + 3 * off, (size_t)cnt * 12, where, b_xyz_, &dx));
    u8 *dval = nullptr, *dflag = nullptr;
    if (where == BNX_DEVICE) {
      dval = values ? static_cast<u8*>(values) + off * cell_bytes : nullptr;
      dflag = found ? found + off : nullptr;
    } else {
      if (values) {
        BNX_TRY(b_out_.reserve((size_t)cnt * cell_bytes));
        dval = b_out_.as<u8>();
        // cells that are missing keep the caller's bytes: start from them
        BNX_CUDA(cudaMemcpyAsync(dval, static_cast<u8*>(values) + off * cell_bytes, (size_t)cnt * cell_bytes, cudaMemcpyHostToDevice, stream_));
      }
      if (found) {
        BNX_TRY(b_flag_.reserve((size_t)cnt));
        dflag = b_flag_.as<u8>();
      }
    }
    BNX_TRY(b_loc_.reserve((size_t)cnt * 4));
    BNX_TRY(locate<false>(static_cast<const i32*>(dx), cnt, b_loc_.as<u32>()));
    note_launch(), k_read_values<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), dval, dflag, cnt);
    BNX_CUDA(cudaGetLastError());
    if (where == BNX_HOST) {
      if (values) BNX_CUDA(cudaMemcpyAsync(static_cast<u8*>(values) + off * cell_bytes, dval, (size_t)cnt * cell_bytes, cudaMemcpyDeviceToHost, stream_));
      if (found) BNX_CUDA(cudaMemcpyAsync(found + off, dflag, (size_t)cnt, cudaMemcpyDeviceToHost, stream_));
      BNX_TRY(sync());
    }
  }
  return BNX_OK;
}

int Grid::is_on(const i32* xyz, i64 n, u8* out, int where) { return get_values(xyz, n, nullptr, out, where); }

int Grid::get_or_create(const i32* xyz, i64 n, void* values, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || (xyz && values)), "get_or_create: null input");
  BNX_SUBBATCH_LOOP(n) {
    const void* dx;
    BNX_TRY(stage_in(xyz + 3 * off, (size_t)cnt * 12, where, b_xyz_, &dx));
    u8* dval;
    if (where == BNX_DEVICE) {
      dval = static_cast<u8*>(values) + off * cell_bytes;
    } else {
      BNX_TRY(b_out_.reserve((size_t)cnt * cell_bytes));
      dval = b_out_.as<u8>();
    }
    BNX_TRY(b_loc_.reserve((size_t)cnt * 4));
    BNX_TRY(locate<true>(static_cast<const i32*>(dx), cnt, b_loc_.as<u32>()));
    BNX_TRY(dedupe(static_cast<const i32*>(dx), b_loc_.as<u32>(), cnt));
    note_launch(), k_create_cells<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), b_slot_.as<u32>(), b_first_.as<u32>(), (u32)cnt);
    note_launch(), k_read_values<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), dval, nullptr, cnt);
    BNX_CUDA(cudaGetLastError());
    if (where == BNX_HOST) {
      BNX_CUDA(cudaMemcpyAsync(static_cast<u8*>(values) + off * cell_bytes, dval, (size_t)cnt * cell_bytes, cudaMemcpyDeviceToHost, stream_));
      BNX_TRY(sync());
    }
  }
  return BNX_OK;
}

int Grid::update_values(const i32* xyz, const void* values, i64 n, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || (xyz && values)), "update_values: null input");
  BNX_SUBBATCH_LOOP(n) {
    const void *dx, *dv;
    BNX_TRY(stage_in(xyz + 3 * off, (size_t)cnt * 12, where, b_xyz_, &dx));
    BNX_TRY(stage_in(static_cast<const u8*>(values) + off * cell_bytes, (size_t)cnt * cell_bytes, where, b_val_, &dv));
    BNX_TRY(b_loc_.reserve((size_t)cnt * 4));
    BNX_TRY(locate<false>(static_cast<const i32*>(dx), cnt, b_loc_.as<u32>()));
    BNX_TRY(dedupe(static_cast<const i32*>(dx), b_loc_.as<u32>(), cnt));
    note_launch(), k_update_values<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), b_slot_.as<u32>(), b_last_.as<u32>(),
                                                         static_cast<const u8*>(dv), (u32)cnt);
    BNX_CUDA(cudaGetLastError());
    if (where == BNX_HOST) BNX_TRY(sync());
  }
  return BNX_OK;
}

int Grid::set_on(const i32* xyz, i64 n, const void* default_value, u8* was_on, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || xyz), "set_on: null input");
  // the default value is a host scalar in both modes
  std::vector<u8> zero((size_t)cell_bytes, 0);
  BNX_TRY(b_val_.reserve(64));
  BNX_CUDA(cudaMemcpyAsync(b_val_.p, default_value ? default_value : zero.data(), (size_t)cell_bytes, cudaMemcpyHostToDevice, stream_));
  BNX_CUDA(cudaStreamSynchronize(stream_));  // `zero` / caller scalar may die after return
  BNX_SUBBATCH_LOOP(n) {
    const void* dx;
    BNX_TRY(stage_in(xyz + 3 * off, (size_t)cnt * 12, where, b_xyz_, &dx));
    u8* dflag = nullptr;
    if (was_on) {
      if (where == BNX_DEVICE) {
        dflag = was_on + off;
      } else {
        BNX_TRY(b_flag_.reserve((size_t)cnt));
        dflag = b_flag_.as<u8>();
      }
    }
    BNX_TRY(b_loc_.reserve((size_t)cnt * 4));
    BNX_TRY(locate<true>(static_cast<const i32*>(dx), cnt, b_loc_.as<u32>()));
    BNX_TRY(dedupe(static_cast<const i32*>(dx), b_loc_.as<u32>(), cnt));
    note_launch(), k_set_on<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), b_slot_.as<u32>(), b_first_.as<u32>(),
                                                  b_val_.as<u8>(), dflag, (u32)cnt);
    BNX_CUDA(cudaGetLastError());
    if (was_on && where == BNX_HOST) BNX_CUDA(cudaMemcpyAsync(was_on + off, dflag, (size_t)cnt, cudaMemcpyDeviceToHost, stream_));
    if (where == BNX_HOST) BNX_TRY(sync());
  }
  return BNX_OK;
}

int Grid::set_off(const i32* xyz, i64 n, u8* was_on, int where) {
  BNX_REQUIRE(n >= 0 && (n == 0 || xyz), "set_off: null input");
  BNX_SUBBATCH_LOOP(n) {
    const void* dx;
    BNX_TRY(stage_in(xyz + 3 * off, (size_t)cnt * 12, where, b_xyz_, &dx));
    u8* dflag = nullptr;
    if (was_on) {
      if (where == BNX_DEVICE) {
        dflag = was_on + off;
      } else {
        BNX_TRY(b_flag_.reserve((size_t)cnt));
        dflag = b_flag_.as<u8>();
      }
    }
    BNX_TRY(b_loc_.reserve((size_t)cnt * 4));
    BNX_TRY(locate<false>(static_cast<const i32*>(dx), cnt, b_loc_.as<u32>()));
    BNX_TRY(dedupe(static_cast<const i32*>(dx), b_loc_.as<u32>(), cnt));
    note_launch(), k_set_off<<<grid_for(cnt), TPB, 0, stream_>>>(dev_, static_cast<const i32*>(dx), b_loc_.as<u32>(), b_slot_.as<u32>(), b_first_.as<u32>(), dflag, (u32)cnt);
    BNX_CUDA(cudaGetLastError());
    if (was_on && where == BNX_HOST) BNX_CUDA(cudaMemcpyAsync(was_on + off, dflag, (size_t)cnt, cudaMemcpyDeviceToHost, stream_));
    if (where == BNX_HOST) BNX_TRY(sync());
  }
  return BNX_OK;
}

int Grid::pos_to_coord(const double* xyz, i64 n, i32* out, int where) const {
  BNX_REQUIRE(n >= 0 && (n == 0 || (xyz && out)), "pos_to_coord: null input");
  if (n == 0) return BNX_OK;
  if (where == BNX_DEVICE) {
    note_launch(), k_pos_to_coord<<<grid_for(3 * n), TPB, 0, stream_>>>(xyz, 3 * n, inv_resolution, out);
    BNX_CUDA(cudaGetLastError());
    return BNX_OK;
  }
  double* din = nullptr;
  i32* dout = nullptr;
  BNX_CUDA(cudaMalloc(&din, (size_t)n * 24));
  BNX_CUDA(cudaMalloc(&dout, (size_t)n * 12));
  cudaMemcpyAsync(din, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, stream_);
  note_launch(), k_pos_to_coord<<<grid_for(3 * n), TPB, 0, stream_>>>(din, 3 * n, inv_resolution, dout);
  cudaMemcpyAsync(out, dout, (size_t)n * 12, cudaMemcpyDeviceToHost, stream_);
  cudaError_t e = cudaStreamSynchronize(stream_);
  cudaFree(din);
  cudaFree(dout);
  BNX_CUDA(e);
  return BNX_OK;
}

int Grid::coord_to_pos(const i32* xyz, i64 n, double* out, int where) const {
  BNX_REQUIRE(n >= 0 && (n == 0 || (xyz && out)), "coord_to_pos: null input");
  if (n == 0) return BNX_OK;
  if (where == BNX_DEVICE) {
    note_launch(), k_coord_to_pos<<<grid_for(3 * n), TPB, 0, stream_>>>(xyz, 3 * n, resolution, out);
    BNX_CUDA(cudaGetLastError());
    return BNX_OK;
  }
  i32* din = nullptr;
  double* dout = nullptr;
  BNX_CUDA(cudaMalloc(&din, (size_t)n * 12));
  BNX_CUDA(cudaMalloc(&dout, (size_t)n * 24));
  cudaMemcpyAsync(din, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, stream_);
  note_launch(), k_coord_to_pos<<<grid_for(3 * n), TPB, 0, stream_>>>(din, 3 * n, resolution, dout);
  cudaMemcpyAsync(out, dout, (size_t)n * 24, cudaMemcpyDeviceToHost, stream_);
  cudaError_t e = cudaStreamSynchronize(stream_);
  cudaFree(din);
  cudaFree(dout);
  BNX_CUDA(e);
  return BNX_OK;
}

// ------------------------------------------------------------------------------------------------
// whole-grid operations
// ------------------------------------------------------------------------------------------------
int Grid::active_count(i64* count) {
  BNX_REQUIRE(count != nullptr, "active_count: null output");
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_leaves = std::min(c.n_leaves, dev_.leaf_cap);
  BNX_CUDA(cudaMemsetAsync(d_count_, 0, 8, stream_));
  if (n_leaves) {
    note_launch(), k_count_active<<<grid_for((i64)n_leaves * dev_.mask_words), TPB, 0, stream_>>>(dev_, n_leaves, reinterpret_cast<unsigned long long*>(d_count_));
    BNX_CUDA(cudaGetLastError());
  }
  BNX_CUDA(cudaMemcpyAsync(h_count_, d_count_, 8, cudaMemcpyDeviceToHost, stream_));
  BNX_TRY(sync());
  *count = (i64)h_count_[0];
  return BNX_OK;
}

int Grid::digest(u64 out[3]) {
  BNX_REQUIRE(out != nullptr, "digest: null output");
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_leaves = std::min(c.n_leaves, dev_.leaf_cap);
  BNX_CUDA(cudaMemsetAsync(d_count_, 0, 24, stream_));
  if (n_leaves) {
    const int blocks = std::max(1, std::min<int>((int)ceil_div(n_leaves, TPB / 32), sm_count() * 8));
    note_launch(), k_digest<<<blocks, TPB, 0, stream_>>>(dev_, n_leaves, reinterpret_cast<unsigned long long*>(d_count_));
    BNX_CUDA(cudaGetLastError());
  }
  BNX_CUDA(cudaMemcpyAsync(h_count_, d_count_, 24, cudaMemcpyDeviceToHost, stream_));
  BNX_TRY(sync());
  for (int k = 0; k < 3; ++k) out[k] = h_count_[k];
  return BNX_OK;
}

int Grid::dump_points_f32(float* out, i64 stride_floats, int zfilter, double zmin, double zmax, i64 cap, i64* count, int where, i32 thr) {
  BNX_REQUIRE(count != nullptr && cap >= 0, "occupied points: bad argument");
  BNX_REQUIRE(stride_floats == 3 || stride_floats == 4, "occupied points: stride must be 3 or 4 floats");
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_leaves = std::min(c.n_leaves, dev_.leaf_cap);
  const int blocks = std::max(1, std::min<int>((int)ceil_div(n_leaves, TPB / 32), sm_count() * 8));
  auto run = [&](float* dout, u64 dcap) -> int {
    BNX_CUDA(cudaMemsetAsync(d_count_, 0, 8, stream_));
    if (n_leaves) {
      const DumpFilter f = {dout, (u32)stride_floats, zfilter, zmin, zmax};
      note_launch(), k_dump<<<blocks, TPB, 0, stream_>>>(dev_, n_leaves, BNX_OCCUPIED, thr, resolution, nullptr, nullptr, nullptr, dcap,
                                                  reinterpret_cast<unsigned long long*>(d_count_), f);
      BNX_CUDA(cudaGetLastError());
    }
    BNX_CUDA(cudaMemcpyAsync(h_count_, d_count_, 8, cudaMemcpyDeviceToHost, stream_));
    return sync();
  };
  if (where == BNX_DEVICE && out && cap > 0) {
    BNX_TRY(run(out, (u64)cap));
    *count = (i64)h_count_[0];
    if (*count > cap) {
      set_error("occupied points: output capacity too small");
      return BNX_ERR_CAPACITY;
    }
    return BNX_OK;
  }
  if (!out || cap == 0) {  // count only
    BNX_TRY(run(nullptr, 0));
    *count = (i64)h_count_[0];
    return BNX_OK;
  }
  // host output: ONE pass into a device staging buffer of the caller's capacity (the kernel counts every point and
  // writes the first `cap`), then one copy of what was found. A caller that keeps some head-room over the previous
  // count (the publisher runs after every scan) never needs the counting pass.
  BNX_TRY(b_out_.reserve((size_t)cap * stride_floats * 4));
  BNX_TRY(run(b_out_.as<float>(), (u64)cap));
  const i64 total = (i64)h_count_[0];
  *count = total;
  if (total > cap) {
    set_error("occupied points: output capacity too small");
    return BNX_ERR_CAPACITY;
  }
  if (total == 0) return BNX_OK;
  BNX_CUDA(cudaMemcpyAsync(out, b_out_.p, (size_t)total * stride_floats * 4, cudaMemcpyDeviceToHost, stream_));
  return sync();
}

int Grid::dump(i32* xyz, double* pos, void* values, i64 cap, i64* count, int where, int pred, i32 thr) {
  BNX_REQUIRE(count != nullptr, "dump: null count");
  BNX_REQUIRE(cap >= 0, "dump: negative capacity");
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_leaves = std::min(c.n_leaves, dev_.leaf_cap);
  const int blocks = std::max(1, std::min<int>((int)ceil_div(n_leaves, TPB / 32), sm_count() * 8));
  auto run = [&](i32* dxyz, double* dpos, u8* dval, u64 dcap) -> int {
    BNX_CUDA(cudaMemsetAsync(d_count_, 0, 8, stream_));
    if (n_leaves) {
      const DumpFilter nofilter = {nullptr, 3, 0, 0.0, 0.0};
      note_launch(), k_dump<<<blocks, TPB, 0, stream_>>>(dev_, n_leaves, pred, thr, resolution, dxyz, dpos, dval, dcap, reinterpret_cast<unsigned long long*>(d_count_), nofilter);
      BNX_CUDA(cudaGetLastError());
    }
    BNX_CUDA(cudaMemcpyAsync(h_count_, d_count_, 8, cudaMemcpyDeviceToHost, stream_));
    return sync();
  };
  const bool want_out = (xyz || pos || values) && cap > 0;
  if (!want_out) {
    BNX_TRY(run(nullptr, nullptr, nullptr, 0));
    *count = (i64)h_count_[0];
    return BNX_OK;
  }
  if (where == BNX_DEVICE) {
    BNX_TRY(run(xyz, pos, static_cast<u8*>(values), (u64)cap));
    *count = (i64)h_count_[0];
    if (*count > cap) {
      set_error("dump: output capacity too small");
      return BNX_ERR_CAPACITY;
    }
    return BNX_OK;
  }
  // host: count first, then stage exactly that many
  BNX_TRY(run(nullptr, nullptr, nullptr, 0));
  const i64 total = (i64)h_count_[0];
  *count = total;
  if (total > cap) {
    set_error("dump: output capacity too small");
    return BNX_ERR_CAPACITY;
  }
  if (total == 0) return BNX_OK;
  i32* dxyz = nullptr;
  double* dpos = nullptr;
  u8* dval = nullptr;
  if (xyz) {
    BNX_TRY(b_xyz_.reserve((size_t)total * 12));
    dxyz = b_xyz_.as<i32>();
  }
  if (pos) {
    BNX_TRY(b_out_.reserve((size_t)total * 24));
    dpos = b_out_.as<double>();
  }
  if (values) {
    BNX_TRY(b_val_.reserve((size_t)total * cell_bytes));
    dval = b_val_.as<u8>();
  }
  BNX_TRY(run(dxyz, dpos, dval, (u64)total));
  if (xyz) BNX_CUDA(cudaMemcpyAsync(xyz, dxyz, (size_t)total * 12, cudaMemcpyDeviceToHost, stream_));
  if (pos) BNX_CUDA(cudaMemcpyAsync(pos, dpos, (size_t)total * 24, cudaMemcpyDeviceToHost, stream_));
  if (values) BNX_CUDA(cudaMemcpyAsync(values, dval, (size_t)total * cell_bytes, cudaMemcpyDeviceToHost, stream_));
  return sync();
}

int Grid::clear(int option) {
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_leaves = std::min(c.n_leaves, dev_.leaf_cap);
  if (option == BNX_SET_ALL_CELLS_OFF) {
    if (n_leaves) {
      note_launch(), k_masks_off<<<grid_for((i64)n_leaves * dev_.mask_words), TPB, 0, stream_>>>(dev_, n_leaves);
      BNX_CUDA(cudaGetLastError());
    }
    return sync();
  }
  BNX_REQUIRE(option == BNX_CLEAR_MEMORY, "clear: unknown ClearOption");
  // CLEAR_MEMORY: every node goes back to the pools (kept mapped, re-zeroed)
  const u32 n_inner = std::min(c.n_inner, dev_.inner_cap);
  if (n_leaves) BNX_CUDA(cudaMemsetAsync(dev_.leaf, 0, (size_t)n_leaves * dev_.leaf_stride, stream_));
  if (n_inner) BNX_CUDA(cudaMemsetAsync(dev_.inner, 0, (size_t)n_inner * dev_.inner_stride * 4, stream_));
  BNX_CUDA(cudaMemsetAsync(root_, 0, root_slots_ * sizeof(int4), stream_));
  std::memset(h_ctr_, 0, sizeof(GridCounters));
  h_ctr_->failed_id = NONE;
  BNX_CUDA(cudaMemcpyAsync(d_ctr_, h_ctr_, sizeof(GridCounters), cudaMemcpyHostToDevice, stream_));
  return sync();
}

int Grid::release_unused() {
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_leaves = std::min(c.n_leaves, dev_.leaf_cap), n_inner = std::min(c.n_inner, dev_.inner_cap);
  if (n_inner == 0) return BNX_OK;
  BNX_TRY(free_list_reserve());
  const u32 children = 1u << (3 * inner_bits);
  note_launch(), k_release_leaves<<<grid_for((i64)n_inner * children), TPB, 0, stream_>>>(dev_, n_inner, children);
  note_launch(), k_release_zero<<<std::max(1, std::min<int>((int)ceil_div(n_leaves, TPB / 32), sm_count() * 8)), TPB, 0, stream_>>>(dev_, n_leaves);
  BNX_CUDA(cudaGetLastError());
  int4* fresh = nullptr;
  BNX_CUDA(cudaMalloc(&fresh, root_slots_ * sizeof(int4)));
  BNX_CUDA(cudaMemsetAsync(fresh, 0, root_slots_ * sizeof(int4), stream_));
  note_launch(), k_rebuild_roots<<<grid_for((i64)root_slots_), TPB, 0, stream_>>>(dev_, root_, root_slots_, fresh, dev_.root_mask, std::max(1u, children / 64u));
  BNX_CUDA(cudaGetLastError());
  BNX_TRY(sync());
  BNX_CUDA(cudaFree(root_));
  root_ = fresh;
  dev_.root = root_;
  return BNX_OK;
}

int Grid::mem_usage(i64* bytes) {
  BNX_REQUIRE(bytes != nullptr, "mem_usage: null output");
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const i64 leaves = (i64)std::min(c.n_leaves, dev_.leaf_cap) - std::max(0, c.n_free);
  *bytes = (i64)root_slots_ * 16 + (i64)c.n_roots * dev_.inner_stride * 4 + leaves * (i64)dev_.leaf_stride;
  return BNX_OK;
}

int Grid::stats(i64 out[8]) {
  GridCounters c;
  BNX_TRY(read_counters(&c));
  out[0] = c.n_roots;
  out[1] = std::min(c.n_inner, dev_.inner_cap);
  out[2] = (i64)std::min(c.n_leaves, dev_.leaf_cap) - std::max(0, c.n_free);
  out[3] = std::max(0, c.n_free);
  out[4] = (i64)root_slots_;
  out[5] = dev_.leaf_cap;
  out[6] = (i64)(leaf_arena_.mapped() + inner_arena_.mapped() + root_slots_ * 16);
  out[7] = 0;
  return BNX_OK;
}

int Grid::serialize(const char* type_name, u8* buffer, i64 cap, i64* size) {
  BNX_REQUIRE(type_name && size, "serialize: null argument");
  GridCounters c;
  BNX_TRY(read_counters(&c));
  const u32 n_inner = std::min(c.n_inner, dev_.inner_cap);
  const u32 children = 1u << (3 * inner_bits), Wi = std::max(1u, children / 64u);
  char header[256];
  std::snprintf(header, sizeof(header), "Bonxai::VoxelGrid<%s,%d,%d>(%lf)\n", type_name, inner_bits, leaf_bits, resolution);  // serialization.hpp:84-88
  const size_t hlen = std::strlen(header);
  const int blocks = std::max(1, std::min<int>((int)ceil_div(std::max(1u, n_inner), TPB / 32), sm_count() * 8));
  auto run = [&](u8* d_out) -> int {
    BNX_CUDA(cudaMemsetAsync(d_count_, 0, 16, stream_));
    if (n_inner) {
      note_launch(), k_serialize<<<blocks, TPB, 0, stream_>>>(dev_, n_inner, children, Wi, d_out, reinterpret_cast<unsigned long long*>(d_count_),
                                                       reinterpret_cast<unsigned long long*>(d_count_) + 1);
      BNX_CUDA(cudaGetLastError());
    }
    BNX_CUDA(cudaMemcpyAsync(h_count_, d_count_, 16, cudaMemcpyDeviceToHost, stream_));
    return sync();
  };
  BNX_TRY(run(nullptr));
  const u64 body = h_count_[0];
  const u32 roots = (u32)h_count_[1];
  *size = (i64)(hlen + 4 + body);
  if (!buffer) return BNX_OK;
  if (cap < *size) {
    set_error("serialize: output capacity too small");
    return BNX_ERR_CAPACITY;
  }
  std::memcpy(buffer, header, hlen);
  std::memcpy(buffer + hlen, &roots, 4);  // serialization.hpp:91
  if (body) {
    BNX_TRY(b_out_.reserve(body));
    BNX_TRY(run(b_out_.as<u8>()));
    BNX_CUDA(cudaMemcpyAsync(buffer + hlen + 4, b_out_.p, body, cudaMemcpyDeviceToHost, stream_));
    BNX_TRY(sync());
  }
  return BNX_OK;
}

// Deserialize, serialization.hpp:118-199: header "Bonxai::VoxelGrid<TYPE,IB,LB>(RES)\n", u32 root count, root bodies.
// The stream is decoded on the host into (coord, value) pairs of the ON cells and inserted with the batched
// setValue path (a leaf whose mask is all OFF leaves no trace, which no API can observe).
int Grid::deserialize(const u8* data, i64 len, int cell_bytes, const char* expect_type, Grid** out) {
  BNX_REQUIRE(data && out && len > 0, "deserialize: null argument");
  const char* txt = reinterpret_cast<const char*>(data);
  const void* nl = std::memchr(txt, '\n', (size_t)len);
  BNX_REQUIRE(nl != nullptr, "Header wasn't recognized");
  const std::string header(txt, static_cast<const char*>(nl));
  const std::string prefix = "Bonxai::VoxelGrid<";
  BNX_REQUIRE(header.rfind(prefix, 0) == 0, "Header wasn't recognized");  // serialization.hpp:127-130
  const size_t gt = header.rfind(">(");
  BNX_REQUIRE(gt != std::string::npos && header.back() == ')', "Header wasn't recognized");
  const std::string inside = header.substr(prefix.size(), gt - prefix.size());  // TYPE,IB,LB (TYPE may contain commas)
  const size_t c2 = inside.rfind(','), c1 = inside.rfind(',', c2 == std::string::npos ? 0 : c2 - 1);
  BNX_REQUIRE(c2 != std::string::npos && c1 != std::string::npos, "Header wasn't recognized");
  const std::string type_name = inside.substr(0, c1);
  const int ib = std::atoi(inside.substr(c1 + 1, c2 - c1 - 1).c_str()), lb = std::atoi(inside.substr(c2 + 1).c_str());
  const double res = std::atof(header.substr(gt + 2, header.size() - gt - 3).c_str());
  if (expect_type && type_name != expect_type) {
    set_error("DataT does not match");  // serialization.hpp:155-158
    return BNX_ERR_INVALID;
  }
  Grid* g = new Grid();
  int st = g->init(res, ib, lb, cell_bytes);
  if (st != BNX_OK) {
    delete g;
    return st;
  }
  const u8* p = static_cast<const u8*>(nl) + 1;
  const u8* end = data + len;
  auto fail = [&](const char* msg) {
    delete g;
    set_error(msg);
    return BNX_ERR_INVALID;
  };
  if (end - p < 4) return fail("deserialize: truncated stream");
  u32 roots;
  std::memcpy(&roots, p, 4);
  p += 4;
  const u32 children = 1u << (3 * ib), Wi = std::max(1u, children / 64u);
  const u32 cells = 1u << (3 * lb), W = std::max(1u, cells / 64u);
  const i32 lmask = (1 << lb) - 1, imask = (1 << ib) - 1;
  std::vector<i32> xyz;
  std::vector<u8> vals;
  auto flush = [&]() -> int {
    if (xyz.empty()) return BNX_OK;
    const int r = g->set_values(xyz.data(), vals.data(), (i64)(xyz.size() / 3), nullptr, BNX_HOST);
    xyz.clear();
    vals.clear();
    return r;
  };
  std::vector<u64> imaskw(Wi), lmaskw(W);
  for (u32 r = 0; r < roots; ++r) {
    if (end - p < (ptrdiff_t)(12 + Wi * 8)) return fail("deserialize: truncated stream");
    i32 key[3];
    std::memcpy(key, p, 12);
    p += 12;
    std::memcpy(imaskw.data(), p, Wi * 8);
    p += Wi * 8;
    for (u32 ci = 0; ci < children; ++ci) {
      if (!((imaskw[ci >> 6] >> (ci & 63)) & 1ull)) continue;
      if (end - p < (ptrdiff_t)(W * 8)) return fail("deserialize: truncated stream");
      std::memcpy(lmaskw.data(), p, W * 8);
      p += W * 8;
      const i32 bx = key[0] | ((i32)(ci & imask) << lb), by = key[1] | ((i32)((ci >> ib) & imask) << lb), bz = key[2] | ((i32)((ci >> (2 * ib)) & imask) << lb);
      for (u32 li = 0; li < cells; ++li) {
        if (!((lmaskw[li >> 6] >> (li & 63)) & 1ull)) continue;
        if (end - p < cell_bytes) return fail("deserialize: truncated stream");
        xyz.push_back(bx | (i32)(li & lmask));
        xyz.push_back(by | (i32)((li >> lb) & lmask));
        xyz.push_back(bz | (i32)((li >> (2 * lb)) & lmask));
        vals.insert(vals.end(), p, p + cell_bytes);
        p += cell_bytes;
      }
    }
    if (xyz.size() >= (size_t)3 << 21) {
      st = flush();
      if (st != BNX_OK) {
        delete g;
        return st;
      }
    }
  }
  st = flush();
  if (st != BNX_OK) {
    delete g;
    return st;
  }
  *out = g;
  return BNX_OK;
}

}  // namespace bnx
