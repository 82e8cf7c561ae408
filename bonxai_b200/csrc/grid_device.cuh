// Device-side layout of the sparse voxel grid and the lock-free find / find-or-create walks.
//
// Replaces:
//   RootMap = std::unordered_map<CoordT, InnerGrid>   bonxai_core/include/bonxai/bonxai.hpp:130
//   InnerGrid / LeafGrid (Grid<T> + Mask)              bonxai.hpp:39-103,128-129 ; mask.hpp
//   Accessor::getLeafGrid                              bonxai.hpp:588-621
//   key math getRootKey/getInnerIndex/getLeafIndex     bonxai.hpp:419-447
//
// HBM layout (all pools are zero-filled when mapped; 0 always means "empty"):
//   root table   open addressing, 16-B slots {int32 kx,ky,kz ; u32 state}. state 0 = EMPTY, 1 = LOCKED,
//                v >= 2 -> inner node (v-2). kx,ky,kz = coord & ~(2^(ib+lb)-1) exactly like getRootKey.
//                The hash is a 64-bit mixer over coord >> (ib+lb) (NOT the reference's 20-bit hash, whose
//                low bits are dead for root keys). Small enough to live in the 126 MB L2.
//   inner pool   node = {int32 key[3]; u32 flags; u64 mask[Wi]; u32 child[8^ib]}; child 0 = EMPTY,
//                1 = LOCKED, v >= 2 -> leaf (v-2). mask bit i mirrors child[i] >= 2 (kept for iteration
//                and the Serialize stream).
//   leaf pool    node = line0 {int32 origin[3]; u32 flags; u32 stamp @16; pad; u64 active[W] @64}
//                       line1 {u64 touched[W] @off_touched; u64 hit[W] @off_hit}   (per-scan scratch of the map)
//                       cells[8^lb] @off_cells, cell_bytes each.
//                Default bits (2,3), 4-byte cells: 2304 B per leaf = 18 x 128-B lines.
//
// Publication protocol: a creator CAS-locks the slot, initialises the node, __threadfence()s, then stores
// the final state. Readers take ONE 16-B (root) or 4-B (child) load: a state >= 2 observed in any cache
// level post-dates the key writes, and a published slot never changes during insert kernels, so cached
// (L1) hits are safe; only EMPTY/LOCKED observations fall through to L2 (atomicCAS / volatile re-load).
#pragma once

#include "common.cuh"

namespace bnx {

constexpr u32 NONE = 0xFFFFFFFFu;

struct GridCounters {
  u32 n_roots;   // root slots in use
  u32 n_inner;   // inner nodes bump-allocated
  u32 n_leaves;  // leaves bump-allocated (may overshoot leaf_cap after a failed scan; host clamps)
  u32 error;     // ERR_* bits
  int n_free;    // entries on the leaf free list
  u32 failed_id;    // async pipeline: id of the first scan that failed (NONE when healthy)
  u32 done_blocks;  // last-block ticket of the scan epilogue
  u32 failed_ovf;   // sharded pipeline: the all-reduced overflow bits of the scan that failed
};

struct GridDev {
  int4* root;
  u32 root_mask;  // slots - 1
  u32* inner;
  u32 inner_stride;  // u32 words per inner node
  u32 inner_cap;
  u32 inner_child_off;  // u32 word offset of child[0]
  unsigned char* leaf;
  u32 leaf_stride;  // bytes
  u32 leaf_cap;
  u32* free_list;
  GridCounters* ctr;
  int ib, lb;  // INNER_BITS, LEAF_BITS
  u32 cell_bytes;
  u32 mask_words;  // u64 words of a leaf mask
  u32 off_active, off_touched, off_hit, off_stamp, off_cells;
};

// ---------------------------------------------------------------------------------------------
// key math (bonxai.hpp:419-447). Arithmetic shifts / two's-complement masks work for negatives.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 inner_index(const GridDev& g, int x, int y, int z) {
  const u32 m = (1u << g.ib) - 1u;
  return ((u32)(x >> g.lb) & m) | (((u32)(y >> g.lb) & m) << g.ib) | (((u32)(z >> g.lb) & m) << (2 * g.ib));
}
__host__ __device__ __forceinline__ u32 leaf_index(const GridDev& g, int x, int y, int z) {
  const u32 m = (1u << g.lb) - 1u;
  return ((u32)x & m) | (((u32)y & m) << g.lb) | (((u32)z & m) << (2 * g.lb));
}

__host__ __device__ __forceinline__ u64 mix64(u64 h) {
  h ^= h >> 33;
  h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33;
  h *= 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 33;
  return h;
}
__host__ __device__ __forceinline__ u64 hash3(int x, int y, int z) {
  u64 h = (u64)(u32)x * 0x9E3779B97F4A7C15ull;
  h ^= (u64)(u32)y * 0xC2B2AE3D27D4EB4Full + (h >> 29);
  h ^= (u64)(u32)z * 0x165667B19E3779F9ull + (h << 7);
  return mix64(h);
}

// ---------------------------------------------------------------------------------------------
// node accessors
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char* leaf_ptr(const GridDev& g, u32 leaf) {
  return g.leaf + (size_t)leaf * g.leaf_stride;
}
__device__ __forceinline__ u64* leaf_active(const GridDev& g, u32 leaf) {
  return reinterpret_cast<u64*>(leaf_ptr(g, leaf) + g.off_active);
}
__device__ __forceinline__ u64* leaf_touched(const GridDev& g, u32 leaf) {
  return reinterpret_cast<u64*>(leaf_ptr(g, leaf) + g.off_touched);
}
__device__ __forceinline__ u64* leaf_hit(const GridDev& g, u32 leaf) {
  return reinterpret_cast<u64*>(leaf_ptr(g, leaf) + g.off_hit);
}
__device__ __forceinline__ u32* leaf_stamp(const GridDev& g, u32 leaf) {
  return reinterpret_cast<u32*>(leaf_ptr(g, leaf) + g.off_stamp);
}
__device__ __forceinline__ unsigned char* leaf_cells(const GridDev& g, u32 leaf) {
  return leaf_ptr(g, leaf) + g.off_cells;
}
__device__ __forceinline__ u32* inner_ptr(const GridDev& g, u32 inner) {
  return g.inner + (size_t)inner * g.inner_stride;
}

__device__ __forceinline__ u32 ld_volatile(const u32* p) { return *reinterpret_cast<const volatile u32*>(p); }
__device__ __forceinline__ int4 ld_cg_int4(const int4* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------------------
// root table
// ---------------------------------------------------------------------------------------------
// returns the inner node index or NONE. Read-only walk (also correct while other threads insert).
__device__ __forceinline__ u32 root_find(const GridDev& g, int kx, int ky, int kz) {
  const int s = g.ib + g.lb;
  u32 slot = (u32)hash3(kx >> s, ky >> s, kz >> s) & g.root_mask;
  for (u32 probe = 0; probe <= g.root_mask; ++probe) {
    int4 v = g.root[slot];
    if ((u32)v.w < 2u) {
      v = ld_cg_int4(g.root + slot);  // EMPTY/LOCKED may be a stale L1 line: confirm in L2
      while ((u32)v.w == 1u) {
        __nanosleep(32);
        v = ld_cg_int4(g.root + slot);
      }
      if ((u32)v.w == 0u) return NONE;
    }
    if (v.x == kx && v.y == ky && v.z == kz) return (u32)v.w - 2u;
    slot = (slot + 1) & g.root_mask;
  }
  return NONE;
}

// find-or-insert; NONE only on pool/table exhaustion (error bit set)
static __device__ __forceinline__ u32 root_find_or_insert(const GridDev& g, int kx, int ky, int kz) {
  const int s = g.ib + g.lb;
  u32 slot = (u32)hash3(kx >> s, ky >> s, kz >> s) & g.root_mask;
  for (u32 probe = 0; probe <= g.root_mask;) {
    int4 v = g.root[slot];
    if ((u32)v.w < 2u) v = ld_cg_int4(g.root + slot);
    u32 st = (u32)v.w;
    if (st >= 2u) {
      if (v.x == kx && v.y == ky && v.z == kz) return st - 2u;
      slot = (slot + 1) & g.root_mask;
      ++probe;
      continue;
    }
    if (st == 0u) {
      u32* state = reinterpret_cast<u32*>(&g.root[slot].w);
      const u32 old = atomicCAS(state, 0u, 1u);
      if (old == 0u) {
        // we own the slot: allocate + initialise the inner node, then publish
        const u32 inner = atomicAdd(&g.ctr->n_inner, 1u);
        if (inner >= g.inner_cap) {
          atomicOr(&g.ctr->error, ERR_INNER_POOL);
          atomicExch(state, 0u);
          return NONE;
        }
        atomicAdd(&g.ctr->n_roots, 1u);  // the host keeps the table under 50 % load between batches
        u32* node = inner_ptr(g, inner);
        node[0] = (u32)kx;
        node[1] = (u32)ky;
        node[2] = (u32)kz;
        node[3] = 1u;
        g.root[slot].x = kx;
        g.root[slot].y = ky;
        g.root[slot].z = kz;
        __threadfence();
        atomicExch(state, inner + 2u);
        return inner;
      }
      // lost the race: fall through and re-read (LOCKED or published)
    }
    __nanosleep(32);  // LOCKED by another thread: its key is unknown until published
  }
  atomicOr(&g.ctr->error, ERR_ROOT_TABLE);
  return NONE;
}

// ---------------------------------------------------------------------------------------------
// leaves
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 leaf_of_inner(const GridDev& g, u32 inner, u32 ii) {
  const u32* child = inner_ptr(g, inner) + g.inner_child_off + ii;
  u32 v = *child;
  if (v < 2u) {
    v = ld_volatile(child);
    while (v == 1u) {
      __nanosleep(32);
      v = ld_volatile(child);
    }
    if (v == 0u) return NONE;
  }
  return v - 2u;
}

__device__ __forceinline__ u32 leaf_find(const GridDev& g, int x, int y, int z) {
  const int m = ~((1 << (g.ib + g.lb)) - 1);
  const u32 inner = root_find(g, x & m, y & m, z & m);
  if (inner == NONE) return NONE;
  return leaf_of_inner(g, inner, inner_index(g, x, y, z));
}

static __device__ __forceinline__ u32 leaf_create_in_inner(const GridDev& g, u32 inner, u32 ii, int x, int y, int z) {
  u32* child = inner_ptr(g, inner) + g.inner_child_off + ii;
  for (;;) {
    u32 v = ld_volatile(child);
    if (v >= 2u) return v - 2u;
    if (v == 0u) {
      const u32 old = atomicCAS(child, 0u, 1u);
      if (old == 0u) {
        u32 leaf = NONE;
        // recycled leaves first (pop-only while insert kernels run; pushes happen in release kernels)
        if (ld_volatile(reinterpret_cast<const u32*>(&g.ctr->n_free)) != 0u) {
          const int k = atomicSub(&g.ctr->n_free, 1);
          if (k > 0) {
            leaf = g.free_list[k - 1];
          } else {
            atomicAdd(&g.ctr->n_free, 1);
          }
        }
        if (leaf == NONE) {
          leaf = atomicAdd(&g.ctr->n_leaves, 1u);
          if (leaf >= g.leaf_cap) {
            atomicOr(&g.ctr->error, ERR_LEAF_POOL);
            atomicExch(child, 0u);
            return NONE;
          }
        }
        const int lm = ~((1 << g.lb) - 1);
        int4* hdr = reinterpret_cast<int4*>(leaf_ptr(g, leaf));
        *hdr = make_int4(x & lm, y & lm, z & lm, 1);
        u64* imask = reinterpret_cast<u64*>(inner_ptr(g, inner) + 4);
        atomicOr(reinterpret_cast<unsigned long long*>(imask + (ii >> 6)), 1ull << (ii & 63));
        __threadfence();
        atomicExch(child, leaf + 2u);
        return leaf;
      }
      continue;
    }
    __nanosleep(32);
  }
}

__device__ __forceinline__ u32 leaf_in_inner_or_create(const GridDev& g, u32 inner, int x, int y, int z) {
  const u32 ii = inner_index(g, x, y, z);
  const u32 v = inner_ptr(g, inner)[g.inner_child_off + ii];
  if (v >= 2u) return v - 2u;
  return leaf_create_in_inner(g, inner, ii, x, y, z);
}

__device__ __forceinline__ u32 leaf_find_or_create(const GridDev& g, int x, int y, int z) {
  const int m = ~((1 << (g.ib + g.lb)) - 1);
  const int kx = x & m, ky = y & m, kz = z & m;
  u32 inner = root_find(g, kx, ky, kz);  // cheap read-only walk first
  if (inner == NONE) inner = root_find_or_insert(g, kx, ky, kz);
  if (inner == NONE) return NONE;
  return leaf_in_inner_or_create(g, inner, x, y, z);
}

}  // namespace bnx
