/* TEST INFRASTRUCTURE ONLY — not part of the product, never linked or loaded by bonxai_b200/.
 *
 * Plain-C CPU restatement ("port") of the reference's occupancy-mapping hot path, behind
 * oracle/oracle_api.h. It restates the *algorithm* of
 *   bonxai_map/include/bonxai_map/probabilistic_map.hpp:141-203  (insertPointCloud, RayIterator)
 *   bonxai_map/src/probabilistic_map.cpp:30-126                  (addHit/MissPoint, updateFreeCells,
 *                                                                 isOccupied/..., getOccupiedVoxels)
 *   bonxai_core/include/bonxai/bonxai.hpp:404-569,678-743        (posToCoord, Accessor ops, forEachCell)
 * on its own storage (an open-addressing table of 8x8x8 blocks), not the reference's
 * unordered_map/InnerGrid/LeafGrid classes: everything the reference exposes through its API
 * (cell value + ON state per coordinate) is independent of the node layout.
 *
 * PARITY PINNING: the reference ships no golden vectors for this path, so this
 * port is pinned against outputs of the reference itself: tests/test_oracle.py runs it side by side
 * with oracle/_ref/libbonxai_ref.so (the unmodified reference, compiled here) and against the
 * fixtures in tests/golden/ that tests/golden/make_golden.py generated from that library.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march, no -ffast-math) — fp64 classification must round once
 * per operation exactly like the reference's baseline x86-64 build (root CMakeLists.txt:37-39).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle_api.h"

/* ------------------------------------------------------------------------------------------------
 * storage: table of 8^3 blocks
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t bx, by, bz; /* block coordinate = cell coordinate >> 3 */
  uint64_t mask[8];   /* ON bits, index = x | y<<3 | z<<6 (bonxai.hpp:440-447 with LEAF_BITS 3) */
  uint32_t cell[512];
} Block;

typedef struct {
  double resolution;
  double inv_resolution; /* bonxai.hpp:395: 1.0 / resolution, computed once */
  int inner_bits, leaf_bits;
  Block** blocks; /* dense list of live blocks */
  int64_t n_blocks, cap_blocks;
  int32_t* slots; /* open addressing: index into blocks, -1 empty */
  int64_t n_slots; /* power of two */
  Block* last;     /* one-entry cache, same role as the accessor's prev_leaf_ptr_ */
} Grid;

static uint64_t mix3(int32_t x, int32_t y, int32_t z) {
  uint64_t h = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ull;
  h ^= h >> 32;
  h += (uint64_t)(uint32_t)y * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h += (uint64_t)(uint32_t)z * 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  h *= 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 31);
}

static void grid_init(Grid* g, double res, int ib, int lb) {
  memset(g, 0, sizeof(*g));
  g->resolution = res;
  g->inv_resolution = 1.0 / res;
  g->inner_bits = ib;
  g->leaf_bits = lb;
  g->n_slots = 1024;
  g->slots = (int32_t*)malloc(sizeof(int32_t) * (size_t)g->n_slots);
  memset(g->slots, 0xFF, sizeof(int32_t) * (size_t)g->n_slots);
}

static void grid_free_storage(Grid* g) {
  for (int64_t i = 0; i < g->n_blocks; ++i) free(g->blocks[i]);
  free(g->blocks);
  free(g->slots);
  g->blocks = NULL;
  g->slots = NULL;
  g->n_blocks = g->cap_blocks = 0;
  g->last = NULL;
}

static void grid_rehash(Grid* g, int64_t n_slots) {
  free(g->slots);
  g->n_slots = n_slots;
  g->slots = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_slots);
  memset(g->slots, 0xFF, sizeof(int32_t) * (size_t)n_slots);
  for (int64_t i = 0; i < g->n_blocks; ++i) {
    const Block* b = g->blocks[i];
    uint64_t s = mix3(b->bx, b->by, b->bz) & (uint64_t)(n_slots - 1);
    while (g->slots[s] >= 0) s = (s + 1) & (uint64_t)(n_slots - 1);
    g->slots[s] = (int32_t)i;
  }
}

static Block* grid_block(Grid* g, int32_t x, int32_t y, int32_t z, int create) {
  const int32_t bx = x >> 3, by = y >> 3, bz = z >> 3;
  Block* l = g->last;
  if (l && l->bx == bx && l->by == by && l->bz == bz) return l;
  uint64_t s = mix3(bx, by, bz) & (uint64_t)(g->n_slots - 1);
  for (;;) {
    const int32_t idx = g->slots[s];
    if (idx < 0) break;
    Block* b = g->blocks[idx];
    if (b->bx == bx && b->by == by && b->bz == bz) {
      g->last = b;
      return b;
    }
    s = (s + 1) & (uint64_t)(g->n_slots - 1);
  }
  if (!create) return NULL;
  if (g->n_blocks == g->cap_blocks) {
    g->cap_blocks = g->cap_blocks ? g->cap_blocks * 2 : 256;
    g->blocks = (Block**)realloc(g->blocks, sizeof(Block*) * (size_t)g->cap_blocks);
  }
  Block* b = (Block*)calloc(1, sizeof(Block));
  b->bx = bx;
  b->by = by;
  b->bz = bz;
  g->blocks[g->n_blocks] = b;
  g->slots[s] = (int32_t)g->n_blocks;
  g->n_blocks++;
  if (g->n_blocks * 2 > g->n_slots) grid_rehash(g, g->n_slots * 2);
  g->last = b;
  return b;
}

static inline uint32_t cell_index(int32_t x, int32_t y, int32_t z) {
  return (uint32_t)((x & 7) | ((y & 7) << 3) | ((z & 7) << 6));
}

/* Accessor::value(coord, create_if_missing) — bonxai.hpp:469-494 */
static uint32_t* grid_value(Grid* g, int32_t x, int32_t y, int32_t z, int create) {
  Block* b = grid_block(g, x, y, z, create);
  if (!b) return NULL;
  const uint32_t i = cell_index(x, y, z);
  const uint64_t bit = 1ull << (i & 63);
  if (b->mask[i >> 6] & bit) return &b->cell[i];
  if (create) {
    b->mask[i >> 6] |= bit;
    b->cell[i] = 0; /* DataT{} */
    return &b->cell[i];
  }
  return NULL;
}

/* ------------------------------------------------------------------------------------------------
 * scalar helpers
 * ---------------------------------------------------------------------------------------------- */
const char* orc_kind(void) { return "port"; }

/* probabilistic_map.hpp:34-36: float argument, arithmetic in double, truncation toward zero */
int32_t orc_logods(float prob) { return (int32_t)(1e6 * log(prob / (1.0 - prob))); }

/* probabilistic_map.hpp:39-42 */
float orc_prob(int32_t logods_fixed) {
  float logods = (float)((double)(float)logods_fixed * 1e-6);
  return (float)(1.0 - 1.0 / (1.0 + (double)expf(logods))); /* std::exp(float) is the float overload */
}

/* bonxai.hpp:404-410 / grid_coord.hpp:80-85 */
static inline void pos_to_coord(double inv_res, double x, double y, double z, int32_t out[3]) {
  out[0] = (int32_t)floor(x * inv_res);
  out[1] = (int32_t)floor(y * inv_res);
  out[2] = (int32_t)floor(z * inv_res);
}

void orc_pos_to_coord(double resolution, const double* xyz, int64_t n, int32_t* out) {
  const double inv = 1.0 / resolution;
  for (int64_t i = 0; i < n; ++i) pos_to_coord(inv, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], out + 3 * i);
}

/* bonxai.hpp:412-417 */
void orc_coord_to_pos(double resolution, const int32_t* xyz, int64_t n, double* out) {
  for (int64_t i = 0; i < 3 * n; ++i) out[i] = (double)xyz[i] * resolution;
}

/* RayIterator, probabilistic_map.hpp:162-203. visit() returns nothing: every functor on the hot
 * path returns true. Returns the number of cells visited. */
typedef void (*VisitFn)(void* ctx, int32_t x, int32_t y, int32_t z);

static int64_t ray_iterate(const int32_t a[3], const int32_t b[3], VisitFn visit, void* ctx) {
  if (a[0] == b[0] && a[1] == b[1] && a[2] == b[2]) return 0; /* :164-166 */
  visit(ctx, a[0], a[1], a[2]);                              /* :167-169 */
  int32_t c[3] = {a[0], a[1], a[2]};
  int32_t err[3] = {0, 0, 0};
  int32_t d[3], st[3];
  for (int k = 0; k < 3; ++k) {
    const int32_t delta = b[k] - a[k];
    st[k] = delta < 0 ? -1 : 1;
    d[k] = delta < 0 ? -delta : delta;
  }
  int32_t m = d[0] > d[1] ? d[0] : d[1];
  if (d[2] > m) m = d[2];
  int64_t n = 1;
  for (int32_t i = 0; i < m - 1; ++i) { /* :183 */
    for (int k = 0; k < 3; ++k) {
      err[k] += d[k];
      if ((err[k] << 1) >= m) {
        c[k] += st[k];
        err[k] -= m;
      }
    }
    visit(ctx, c[0], c[1], c[2]);
    ++n;
  }
  return n;
}

typedef struct {
  int32_t* out;
  int64_t cap, n;
} RayBuf;

static void ray_collect(void* ctx, int32_t x, int32_t y, int32_t z) {
  RayBuf* r = (RayBuf*)ctx;
  if (r->n < r->cap) {
    r->out[3 * r->n] = x;
    r->out[3 * r->n + 1] = y;
    r->out[3 * r->n + 2] = z;
  }
  r->n++;
}

int64_t orc_compute_ray(const int32_t a[3], const int32_t b[3], int32_t* out_xyz, int64_t cap) {
  RayBuf r = {out_xyz, cap, 0};
  ray_iterate(a, b, ray_collect, &r);
  return r.n;
}

/* ------------------------------------------------------------------------------------------------
 * VoxelGrid<uint32_t>
 * ---------------------------------------------------------------------------------------------- */
void* orc_grid_create(double voxel_size, int inner_bits, int leaf_bits) {
  if (inner_bits < 1 || leaf_bits < 1) return NULL; /* bonxai.hpp:399-401 throws */
  Grid* g = (Grid*)malloc(sizeof(Grid));
  grid_init(g, voxel_size, inner_bits, leaf_bits);
  return g;
}

void orc_grid_destroy(void* gp) {
  Grid* g = (Grid*)gp;
  if (!g) return;
  grid_free_storage(g);
  free(g);
}

/* Accessor::setValue, bonxai.hpp:449-466 */
void orc_grid_set_values(void* gp, const int32_t* xyz, const uint32_t* vals, int64_t n, uint8_t* was_on) {
  Grid* g = (Grid*)gp;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    Block* b = grid_block(g, x, y, z, 1);
    const uint32_t ci = cell_index(x, y, z);
    const uint64_t bit = 1ull << (ci & 63);
    const int w = (b->mask[ci >> 6] & bit) != 0;
    b->mask[ci >> 6] |= bit;
    b->cell[ci] = vals[i];
    if (was_on) was_on[i] = (uint8_t)w;
  }
}

/* ConstAccessor::value, bonxai.hpp:496-516 */
void orc_grid_get_values(void* gp, const int32_t* xyz, int64_t n, uint32_t* out, uint8_t* found) {
  Grid* g = (Grid*)gp;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t* v = grid_value(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0);
    found[i] = v != NULL;
    if (v) out[i] = *v;
  }
}

void orc_grid_get_or_create(void* gp, const int32_t* xyz, int64_t n, uint32_t* out) {
  Grid* g = (Grid*)gp;
  for (int64_t i = 0; i < n; ++i) out[i] = *grid_value(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 1);
}

/* Accessor::setCellOn, bonxai.hpp:537-554 */
void orc_grid_set_on(void* gp, const int32_t* xyz, int64_t n, uint32_t default_value, uint8_t* was_on) {
  Grid* g = (Grid*)gp;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    Block* b = grid_block(g, x, y, z, 1);
    const uint32_t ci = cell_index(x, y, z);
    const uint64_t bit = 1ull << (ci & 63);
    const int w = (b->mask[ci >> 6] & bit) != 0;
    b->mask[ci >> 6] |= bit;
    if (!w) b->cell[ci] = default_value;
    if (was_on) was_on[i] = (uint8_t)w;
  }
}

/* Accessor::setCellOff, bonxai.hpp:557-569: the value is kept, only the bit drops */
void orc_grid_set_off(void* gp, const int32_t* xyz, int64_t n, uint8_t* was_on) {
  Grid* g = (Grid*)gp;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    Block* b = grid_block(g, x, y, z, 0);
    int w = 0;
    if (b) {
      const uint32_t ci = cell_index(x, y, z);
      const uint64_t bit = 1ull << (ci & 63);
      w = (b->mask[ci >> 6] & bit) != 0;
      b->mask[ci >> 6] &= ~bit;
    }
    if (was_on) was_on[i] = (uint8_t)w;
  }
}

/* ConstAccessor::isCellOn, bonxai.hpp:518-534 */
void orc_grid_is_on(void* gp, const int32_t* xyz, int64_t n, uint8_t* out) {
  Grid* g = (Grid*)gp;
  for (int64_t i = 0; i < n; ++i) out[i] = grid_value(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0) != NULL;
}

/* activeCellsCount, bonxai.hpp:689-701 */
int64_t orc_grid_active_count(void* gp) {
  Grid* g = (Grid*)gp;
  int64_t n = 0;
  for (int64_t i = 0; i < g->n_blocks; ++i)
    for (int w = 0; w < 8; ++w) n += __builtin_popcountll(g->blocks[i]->mask[w]);
  return n;
}

/* forEachCell, bonxai.hpp:704-743 (order is unspecified in the reference too) */
int64_t orc_grid_dump(void* gp, int32_t* xyz, uint32_t* vals, int64_t cap) {
  Grid* g = (Grid*)gp;
  int64_t n = 0;
  for (int64_t i = 0; i < g->n_blocks; ++i) {
    const Block* b = g->blocks[i];
    for (int w = 0; w < 8; ++w) {
      uint64_t m = b->mask[w];
      while (m) {
        const int bit = __builtin_ctzll(m);
        m &= m - 1;
        const int ci = w * 64 + bit;
        if (n < cap) {
          xyz[3 * n] = (int32_t)((uint32_t)b->bx << 3) | (ci & 7);
          xyz[3 * n + 1] = (int32_t)((uint32_t)b->by << 3) | ((ci >> 3) & 7);
          xyz[3 * n + 2] = (int32_t)((uint32_t)b->bz << 3) | ((ci >> 6) & 7);
          vals[n] = b->cell[ci];
        }
        ++n;
      }
    }
  }
  return n;
}

/* clear, bonxai.hpp:678-687 */
void orc_grid_clear(void* gp, int opt) {
  Grid* g = (Grid*)gp;
  if (opt == 0) {
    const double res = g->resolution;
    const int ib = g->inner_bits, lb = g->leaf_bits;
    grid_free_storage(g);
    grid_init(g, res, ib, lb);
    return;
  }
  for (int64_t i = 0; i < g->n_blocks; ++i) memset(g->blocks[i]->mask, 0, sizeof(g->blocks[i]->mask));
}

/* releaseUnusedMemory, bonxai.hpp:367-387: drops nodes whose cells are all OFF. Values stored in
 * OFF cells become unreachable either way (value(create) re-initialises an OFF cell), so on this
 * storage it only needs to drop empty blocks. */
void orc_grid_release_unused(void* gp) {
  Grid* g = (Grid*)gp;
  int64_t k = 0;
  for (int64_t i = 0; i < g->n_blocks; ++i) {
    Block* b = g->blocks[i];
    uint64_t any = 0;
    for (int w = 0; w < 8; ++w) any |= b->mask[w];
    if (any)
      g->blocks[k++] = b;
    else
      free(b);
  }
  g->n_blocks = k;
  g->last = NULL;
  grid_rehash(g, g->n_slots);
}

/* The stream format depends on the root/inner node layout, which this storage does not model:
 * serialisation parity is checked against the reference build only (oracle/_ref). */
int64_t orc_grid_serialize(void* g, uint8_t* out, int64_t cap) {
  (void)g;
  (void)out;
  (void)cap;
  return -1;
}
void* orc_grid_deserialize(const uint8_t* data, int64_t len) {
  (void)data;
  (void)len;
  return NULL;
}

/* ------------------------------------------------------------------------------------------------
 * ProbabilisticMap
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t* v;
  int64_t n, cap;
} CoordList;

static void list_push(CoordList* l, const int32_t c[3]) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 1024;
    l->v = (int32_t*)realloc(l->v, sizeof(int32_t) * 3 * (size_t)l->cap);
  }
  memcpy(l->v + 3 * l->n, c, 12);
  l->n++;
}

typedef struct {
  Grid grid;
  int32_t miss, hit, cmin, cmax, thr; /* probabilistic_map.hpp:56-64 */
  uint8_t update_count;               /* probabilistic_map.hpp:129, cycles 1,2,3 */
  CoordList hit_coords, miss_coords;  /* probabilistic_map.hpp:131-132 */
  int64_t counters[4];
  int64_t cur_visits, cur_updates;
  double last_seconds;
} PMap;

/* CellT image: bits 0-3 update_id, bits 4-31 probability_log (signed) — probabilistic_map.hpp:44-53 */
static inline int32_t word_prob(uint32_t w) { return (int32_t)w >> 4; }
static inline uint32_t word_id(uint32_t w) { return w & 0xFu; }
static inline uint32_t make_word(int32_t prob, uint32_t id) { return ((uint32_t)prob << 4) | (id & 0xFu); }

void* orc_map_create(double resolution) {
  PMap* m = (PMap*)calloc(1, sizeof(PMap));
  grid_init(&m->grid, resolution, 2, 3);
  m->miss = orc_logods(0.4f);
  m->hit = orc_logods(0.7f);
  m->cmin = orc_logods(0.12f);
  m->cmax = orc_logods(0.97f);
  m->thr = orc_logods(0.5f);
  m->update_count = 1;
  return m;
}

void orc_map_destroy(void* mp) {
  PMap* m = (PMap*)mp;
  if (!m) return;
  grid_free_storage(&m->grid);
  free(m->hit_coords.v);
  free(m->miss_coords.v);
  free(m);
}

void orc_map_set_options(void* mp, const int32_t o[5]) {
  PMap* m = (PMap*)mp;
  m->miss = o[0];
  m->hit = o[1];
  m->cmin = o[2];
  m->cmax = o[3];
  m->thr = o[4];
}

void orc_map_get_options(void* mp, int32_t o[5]) {
  PMap* m = (PMap*)mp;
  o[0] = m->miss;
  o[1] = m->hit;
  o[2] = m->cmin;
  o[3] = m->cmax;
  o[4] = m->thr;
}

/* addHitPoint, probabilistic_map.cpp:30-41 */
static void add_hit(PMap* m, double x, double y, double z) {
  int32_t c[3];
  pos_to_coord(m->grid.inv_resolution, x, y, z, c);
  uint32_t* cell = grid_value(&m->grid, c[0], c[1], c[2], 1);
  if (word_id(*cell) != m->update_count) {
    int32_t p = word_prob(*cell) + m->hit;
    if (p > m->cmax) p = m->cmax;
    *cell = make_word(p, m->update_count);
    list_push(&m->hit_coords, c);
    m->cur_updates++;
  }
}

/* addMissPoint, probabilistic_map.cpp:43-54 */
static void add_miss(PMap* m, double x, double y, double z) {
  int32_t c[3];
  pos_to_coord(m->grid.inv_resolution, x, y, z, c);
  uint32_t* cell = grid_value(&m->grid, c[0], c[1], c[2], 1);
  if (word_id(*cell) != m->update_count) {
    int32_t p = word_prob(*cell) + m->miss;
    if (p < m->cmin) p = m->cmin;
    *cell = make_word(p, m->update_count);
    list_push(&m->miss_coords, c);
    m->cur_updates++;
  }
}

/* clearPoint lambda, probabilistic_map.cpp:81-89 */
static void clear_point(void* ctx, int32_t x, int32_t y, int32_t z) {
  PMap* m = (PMap*)ctx;
  uint32_t* cell = grid_value(&m->grid, x, y, z, 1);
  m->cur_visits++;
  if (word_id(*cell) != m->update_count) {
    int32_t p = word_prob(*cell) + m->miss;
    if (p < m->cmin) p = m->cmin;
    *cell = make_word(p, m->update_count);
    m->cur_updates++;
  }
}

/* updateFreeCells, probabilistic_map.cpp:77-106 */
static void update_free_cells(PMap* m, double ox, double oy, double oz) {
  int32_t o[3];
  pos_to_coord(m->grid.inv_resolution, ox, oy, oz, o);
  for (int64_t i = 0; i < m->hit_coords.n; ++i) ray_iterate(o, m->hit_coords.v + 3 * i, clear_point, m);
  m->hit_coords.n = 0;
  for (int64_t i = 0; i < m->miss_coords.n; ++i) ray_iterate(o, m->miss_coords.v + 3 * i, clear_point, m);
  m->miss_coords.n = 0;
  if (++m->update_count == 4) m->update_count = 1;
}

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* the per-point body of insertPointCloud, probabilistic_map.hpp:146-158. All operations are
 * individually rounded fp64; squaredNorm is (x*x + y*y) + z*z (see oracle/shim). */
static inline void insert_point(PMap* m, double px, double py, double pz, double fx, double fy, double fz,
                                double max_range, double max_range_sqr) {
  const double vx = px - fx, vy = py - fy, vz = pz - fz;
  const double sq = (vx * vx + vy * vy) + vz * vz;
  if (sq >= max_range_sqr) {
    const double nrm = sqrt(sq);
    const double nx = fx + ((vx / nrm) * max_range);
    const double ny = fy + ((vy / nrm) * max_range);
    const double nz = fz + ((vz / nrm) * max_range);
    add_miss(m, nx, ny, nz);
  } else {
    add_hit(m, px, py, pz);
  }
}

static void begin_insert(PMap* m) {
  m->cur_visits = 0;
  m->cur_updates = 0;
}

static void end_insert(PMap* m, int64_t n, int64_t updates_before) {
  (void)updates_before;
  m->counters[0] = n;
  m->counters[2] = m->cur_visits + n;
  m->counters[3] = m->cur_updates;
}

void orc_map_insert_f32(void* mp, const void* pts, int64_t stride_bytes, int64_t n, const float origin[3],
                        double max_range) {
  PMap* m = (PMap*)mp;
  const double t0 = now_seconds();
  begin_insert(m);
  const double fx = origin[0], fy = origin[1], fz = origin[2];
  const double max_range_sqr = max_range * max_range;
  const char* p = (const char*)pts;
  for (int64_t i = 0; i < n; ++i) {
    const float* q = (const float*)(p + i * stride_bytes);
    insert_point(m, (double)q[0], (double)q[1], (double)q[2], fx, fy, fz, max_range, max_range_sqr);
  }
  m->counters[1] = m->hit_coords.n + m->miss_coords.n;
  update_free_cells(m, fx, fy, fz);
  end_insert(m, n, 0);
  m->last_seconds = now_seconds() - t0;
}

void orc_map_insert_f64(void* mp, const double* pts, int64_t n, const double origin[3], double max_range) {
  PMap* m = (PMap*)mp;
  const double t0 = now_seconds();
  begin_insert(m);
  const double fx = origin[0], fy = origin[1], fz = origin[2];
  const double max_range_sqr = max_range * max_range;
  for (int64_t i = 0; i < n; ++i)
    insert_point(m, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], fx, fy, fz, max_range, max_range_sqr);
  m->counters[1] = m->hit_coords.n + m->miss_coords.n;
  update_free_cells(m, fx, fy, fz);
  end_insert(m, n, 0);
  m->last_seconds = now_seconds() - t0;
}

/* public addHitPoint/addMissPoint: stay queued until the next insertPointCloud
 * (probabilistic_map.hpp:95-103; updateFreeCells is private, :136) */
void orc_map_add_hit(void* mp, const double p[3]) { add_hit((PMap*)mp, p[0], p[1], p[2]); }
void orc_map_add_miss(void* mp, const double p[3]) { add_miss((PMap*)mp, p[0], p[1], p[2]); }

/* isOccupied / isUnknown / isFree, probabilistic_map.cpp:56-75 */
void orc_map_query(void* mp, const int32_t* xyz, int64_t n, int kind, uint8_t* out) {
  PMap* m = (PMap*)mp;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t* cell = grid_value(&m->grid, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0);
    if (!cell) {
      out[i] = kind == 1;
    } else {
      const int32_t p = word_prob(*cell);
      out[i] = kind == 0 ? p > m->thr : kind == 1 ? p == m->thr : p < m->thr;
    }
  }
}

/* getOccupiedVoxels / getFreeVoxels, probabilistic_map.cpp:108-126 */
int64_t orc_map_get_voxels(void* mp, int kind, int32_t* xyz, int64_t cap) {
  PMap* m = (PMap*)mp;
  int64_t n = 0;
  for (int64_t i = 0; i < m->grid.n_blocks; ++i) {
    const Block* b = m->grid.blocks[i];
    for (int w = 0; w < 8; ++w) {
      uint64_t mk = b->mask[w];
      while (mk) {
        const int bit = __builtin_ctzll(mk);
        mk &= mk - 1;
        const int ci = w * 64 + bit;
        const int32_t p = word_prob(b->cell[ci]);
        if (kind == 0 ? p > m->thr : p < m->thr) {
          if (n < cap) {
            xyz[3 * n] = (int32_t)((uint32_t)b->bx << 3) | (ci & 7);
            xyz[3 * n + 1] = (int32_t)((uint32_t)b->by << 3) | ((ci >> 3) & 7);
            xyz[3 * n + 2] = (int32_t)((uint32_t)b->bz << 3) | ((ci >> 6) & 7);
          }
          ++n;
        }
      }
    }
  }
  return n;
}

int64_t orc_map_active_count(void* mp) { return orc_grid_active_count(&((PMap*)mp)->grid); }

int64_t orc_map_dump(void* mp, int32_t* xyz, uint32_t* words, int64_t cap) {
  return orc_grid_dump(&((PMap*)mp)->grid, xyz, words, cap);
}

void orc_map_counters(void* mp, int64_t out[4]) { memcpy(out, ((PMap*)mp)->counters, sizeof(int64_t) * 4); }
void orc_map_track_updates(void* mp, int enable) {
  (void)mp;
  (void)enable; /* always counted */
}
double orc_map_last_insert_seconds(void* mp) { return ((PMap*)mp)->last_seconds; }

/* ---- test helper, not a restatement of anything in the reference: the order-independent digest of a dump
 * ({sum, xor, count} of mix64(hash3(x,y,z) + FNV1a64(word bytes) * 0x9E3779B97F4A7C15), see include/bonxai_b200.h
 * bnx_grid_digest) so that full-size maps of either oracle can be compared with the GPU's without sorting. */
static uint64_t dg_mix64(uint64_t h) {
  h ^= h >> 33;
  h *= 0xFF51AFD7ED558CCDull;
  h ^= h >> 33;
  h *= 0xC4CEB9FE1A85EC53ull;
  h ^= h >> 33;
  return h;
}
void orc_digest_pairs(const int32_t* xyz, const uint32_t* words, int64_t n, uint64_t out[3]) {
  uint64_t sum = 0, x = 0;
  for (int64_t i = 0; i < n; ++i) {
    uint64_t h = (uint64_t)(uint32_t)xyz[3 * i] * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)(uint32_t)xyz[3 * i + 1] * 0xC2B2AE3D27D4EB4Full + (h >> 29);
    h ^= (uint64_t)(uint32_t)xyz[3 * i + 2] * 0x165667B19E3779F9ull + (h << 7);
    h = dg_mix64(h);
    uint64_t f = 0xCBF29CE484222325ull;
    for (int k = 0; k < 4; ++k) f = (f ^ ((words[i] >> (8 * k)) & 0xFFu)) * 0x100000001B3ull;
    h = dg_mix64(h + f * 0x9E3779B97F4A7C15ull);
    sum += h;
    x ^= h;
  }
  out[0] = sum;
  out[1] = x;
  out[2] = (uint64_t)n;
}
