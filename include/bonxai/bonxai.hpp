// bonxai_b200 drop-in for bonxai_core/include/bonxai/bonxai.hpp.
//
// Bonxai::VoxelGrid<DataT> with the reference's public names (bonxai.hpp:114-333), implemented as a thin
// inline wrapper over the C ABI in include/bonxai_b200.h: the grid itself lives in B200 HBM.
//
// What is identical for a caller: construction (throws std::runtime_error for bits < 1), posToCoord /
// coordToPos, createAccessor / createConstAccessor, Accessor::setValue / value / setCellOn / setCellOff,
// ConstAccessor::value / isCellOn, forEachCell (const and mutable visitors), activeCellsCount, memUsage,
// clear, releaseUnusedMemory, getRootKey / getInnerKey / getInnerIndex / getLeafIndex.
//
// What differs (the storage is not host memory):
//   * every single-cell accessor call is a device round trip. Hot loops should use the batched members
//     setValues / getValues / setCellsOn / setCellsOff / isCellsOn added below (one kernel per batch).
//   * Accessor::value() returns a pointer into a per-accessor staging slot, not into the leaf. Writing
//     through it works as in the reference (examples/tutorial.cpp:42-51): the slot is written back on the
//     accessor's next call, on sync() and on destruction.
//   * forEachCell hands the visitor references into a host copy of the cells; values changed by a mutable
//     visitor are written back (cells the visitor switched off keep their old value).
//   * rootMap(), allocateLeafGrid(), lastInnerGrid()/lastLeafGrid()/getLeafGrid() expose host node
//     objects in the reference and have no counterpart here.
//   * DataT must be trivially copyable, at most 64 bytes; DataT{} must be all-zero bytes.
#pragma once

#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "bonxai/grid_coord.hpp"
#include "bonxai_b200.h"

namespace Bonxai {

// empty payload of BinaryVoxelGrid (bonxai.hpp:29-30)
struct EmptyVoxel {};

// bonxai.hpp:107-112
enum ClearOption {
  CLEAR_MEMORY = BNX_CLEAR_MEMORY,
  SET_ALL_CELLS_OFF = BNX_SET_ALL_CELLS_OFF
};

namespace detail {
inline void check(int status) {
  if (status == BNX_OK) return;
  const std::string msg = std::string("bonxai_b200: ") + bnx_last_error();
  if (status == BNX_ERR_NOMEM) throw std::bad_alloc();
  throw std::runtime_error(msg);
}
}  // namespace detail

template <typename DataT>
class VoxelGrid {
  static_assert(std::is_trivially_copyable_v<DataT>, "bonxai_b200: DataT must be trivially copyable");
  static_assert(sizeof(DataT) <= 64, "bonxai_b200: DataT must be at most 64 bytes");

 public:
  explicit VoxelGrid(double voxel_size, uint8_t inner_bits = 2, uint8_t leaf_bits = 3)
      : INNER_BITS(inner_bits), LEAF_BITS(leaf_bits), Log2N(inner_bits + leaf_bits), resolution(voxel_size),
        inv_resolution(1.0 / voxel_size) {
    if (LEAF_BITS < 1 || INNER_BITS < 1) {
      throw std::runtime_error("The minimum value of the inner_bits and leaf_bits should be 1");
    }
    bnx_grid_t* h = nullptr;
    detail::check(bnx_grid_create(voxel_size, inner_bits, leaf_bits, (int)sizeof(DataT), &h));
    handle_ = h;
    owned_ = true;
  }

  // view of a grid owned by someone else (ProbabilisticMap::grid())
  VoxelGrid(bnx_grid_t* borrowed, double voxel_size, uint8_t inner_bits, uint8_t leaf_bits)
      : INNER_BITS(inner_bits), LEAF_BITS(leaf_bits), Log2N(inner_bits + leaf_bits), resolution(voxel_size),
        inv_resolution(1.0 / voxel_size), handle_(borrowed), owned_(false) {}

  VoxelGrid(const VoxelGrid&) = delete;
  VoxelGrid& operator=(const VoxelGrid&) = delete;
  VoxelGrid(VoxelGrid&& o) noexcept { *this = std::move(o); }
  VoxelGrid& operator=(VoxelGrid&& o) noexcept {
    if (this != &o) {
      release();
      INNER_BITS = o.INNER_BITS;
      LEAF_BITS = o.LEAF_BITS;
      Log2N = o.Log2N;
      resolution = o.resolution;
      inv_resolution = o.inv_resolution;
      handle_ = o.handle_;
      owned_ = o.owned_;
      o.handle_ = nullptr;
    }
    return *this;
  }
  ~VoxelGrid() { release(); }

  uint32_t innetBits() const { return INNER_BITS; }  // (sic) the reference's spelling, bonxai.hpp:146
  uint32_t leafBits() const { return LEAF_BITS; }
  double voxelSize() const { return resolution; }
  bnx_grid_t* handle() const { return handle_; }
  // take ownership of a handle passed to the view constructor (used by Deserialize)
  void adopt() { owned_ = true; }

  [[nodiscard]] size_t memUsage() const {
    int64_t b = 0;
    detail::check(bnx_grid_mem_usage(handle_, &b));
    return (size_t)b;
  }
  void releaseUnusedMemory() { detail::check(bnx_grid_release_unused(handle_)); }

  // Node-object API of the reference (bonxai.hpp:156-161 rootMap, :233-243 lastInnerGrid / lastLeafGrid / getLeafGrid,
  // :572-586 allocateLeafGrid): it hands out host pointers to std::unordered_map / Grid<> nodes. The nodes of this grid
  // live in device pools, so the members exist only to turn "no member named rootMap" into a message that says what to
  // use instead; they fire when (and only when) a caller instantiates them.
  template <class Dummy = void>
  void rootMap() const {
    static_assert(sizeof(Dummy) == 0, "bonxai_b200: rootMap() exposes host node objects; the nodes live in device pools. "
                                      "Use forEachCell / bnx_grid_dump (cells), bnx_grid_stats (node counts) or Serialize().");
  }
  template <class Dummy = void>
  void lastInnerGrid() const {
    static_assert(sizeof(Dummy) == 0, "bonxai_b200: Accessor::lastInnerGrid()/lastLeafGrid()/getLeafGrid() return host node pointers; "
                                      "use the batched accessor calls (getValues / setValues / isCellsOn) instead.");
  }
  template <class Dummy = void>
  void allocateLeafGrid() const {
    static_assert(sizeof(Dummy) == 0, "bonxai_b200: leaves are allocated by the device-side pools (bump allocator + free list); "
                                      "there is no host-side LeafGrid to allocate.");
  }

  [[nodiscard]] size_t activeCellsCount() const {
    int64_t n = 0;
    detail::check(bnx_grid_active_count(handle_, &n));
    return (size_t)n;
  }

  [[nodiscard]] CoordT posToCoord(double x, double y, double z) const { return PosToCoord({x, y, z}, inv_resolution); }
  [[nodiscard]] CoordT posToCoord(const Point3D& pos) const { return PosToCoord(pos, inv_resolution); }
  [[nodiscard]] Point3D coordToPos(const CoordT& coord) const { return CoordToPos(coord, resolution); }

  // ---- iteration (bonxai.hpp:704-743). Order is unspecified, as in the reference (unordered_map order).
  template <class VisitorFunction>
  void forEachCell(VisitorFunction func) const {
    std::vector<CoordT> coords;
    std::vector<DataT> values;
    snapshot(coords, values);
    for (size_t i = 0; i < coords.size(); ++i) func(values[i], coords[i]);
  }
  template <class VisitorFunction>
  void forEachCell(VisitorFunction func) {
    std::vector<CoordT> coords;
    std::vector<DataT> values;
    snapshot(coords, values);
    std::vector<DataT> before(values);
    for (size_t i = 0; i < coords.size(); ++i) func(values[i], coords[i]);
    // write back what the visitor changed through its DataT& (only cells that are still ON take it)
    std::vector<CoordT> dc;
    std::vector<DataT> dv;
    for (size_t i = 0; i < coords.size(); ++i) {
      if (std::memcmp(&values[i], &before[i], sizeof(DataT)) != 0) {
        dc.push_back(coords[i]);
        dv.push_back(values[i]);
      }
    }
    if (!dc.empty()) {
      detail::check(bnx_grid_update_values(handle_, &dc[0].x, dv.data(), (int64_t)dc.size(), BNX_HOST));
    }
  }

  void clear(ClearOption opt) { detail::check(bnx_grid_clear(handle_, (int)opt)); }

  // ---- batched extensions: the fast path (one kernel per call, sequential semantics inside the batch)
  void setValues(const CoordT* coords, const DataT* values, size_t n, uint8_t* was_on = nullptr) {
    detail::check(bnx_grid_set_values(handle_, n ? &coords[0].x : nullptr, values, (int64_t)n, was_on, BNX_HOST));
  }
  void getValues(const CoordT* coords, size_t n, DataT* values, uint8_t* found) const {
    detail::check(bnx_grid_get_values(handle_, n ? &coords[0].x : nullptr, (int64_t)n, values, found, BNX_HOST));
  }
  void setCellsOn(const CoordT* coords, size_t n, const DataT& default_value = DataT(), uint8_t* was_on = nullptr) {
    detail::check(bnx_grid_set_on(handle_, n ? &coords[0].x : nullptr, (int64_t)n, &default_value, was_on, BNX_HOST));
  }
  void setCellsOff(const CoordT* coords, size_t n, uint8_t* was_on = nullptr) {
    detail::check(bnx_grid_set_off(handle_, n ? &coords[0].x : nullptr, (int64_t)n, was_on, BNX_HOST));
  }
  void isCellsOn(const CoordT* coords, size_t n, uint8_t* out) const {
    detail::check(bnx_grid_is_on(handle_, n ? &coords[0].x : nullptr, (int64_t)n, out, BNX_HOST));
  }

  // ---- accessors (bonxai.hpp:215-313)
  class ConstAccessor {
   public:
    explicit ConstAccessor(const VoxelGrid& grid) : grid_(&grid) {}

    // nullptr when the cell is OFF or missing. The pointer stays valid until this accessor's next call.
    [[nodiscard]] const DataT* value(const CoordT& coord) const {
      uint8_t found = 0;
      detail::check(bnx_grid_get_values(grid_->handle_, &coord.x, 1, &slot_, &found, BNX_HOST));
      return found ? &slot_ : nullptr;
    }
    [[nodiscard]] bool isCellOn(const CoordT& coord) const {
      uint8_t on = 0;
      detail::check(bnx_grid_is_on(grid_->handle_, &coord.x, 1, &on, BNX_HOST));
      return on != 0;
    }

   protected:
    const VoxelGrid* grid_;
    mutable DataT slot_{};
  };

  class Accessor : public ConstAccessor {
   public:
    explicit Accessor(VoxelGrid& grid) : ConstAccessor(grid), mutable_grid_(&grid) {}
    Accessor(const Accessor& o) : ConstAccessor(o), mutable_grid_(o.mutable_grid_) {}  // staging slot is not shared
    Accessor& operator=(const Accessor& o) {
      sync();
      ConstAccessor::operator=(o);
      mutable_grid_ = o.mutable_grid_;
      pending_ = false;
      return *this;
    }
    ~Accessor() {
      try {
        sync();
      } catch (...) {
      }
    }

    // bonxai.hpp:449-466 — returns the previous state of the cell (ON = true)
    bool setValue(const CoordT& coord, const DataT& value) {
      sync();
      uint8_t was_on = 0;
      detail::check(bnx_grid_set_values(mutable_grid_->handle_, &coord.x, &value, 1, &was_on, BNX_HOST));
      return was_on != 0;
    }

    // bonxai.hpp:469-494 — pointer to the value, nullptr if missing and !create_if_missing. Writes through the
    // pointer reach the grid at the next accessor call / sync() / destruction.
    [[nodiscard]] DataT* value(const CoordT& coord, bool create_if_missing = false) {
      sync();
      if (create_if_missing) {
        detail::check(bnx_grid_get_or_create(mutable_grid_->handle_, &coord.x, 1, &this->slot_, BNX_HOST));
      } else {
        uint8_t found = 0;
        detail::check(bnx_grid_get_values(mutable_grid_->handle_, &coord.x, 1, &this->slot_, &found, BNX_HOST));
        if (!found) return nullptr;
      }
      fetched_ = this->slot_;
      pending_coord_ = coord;
      pending_ = true;
      return &this->slot_;
    }

    // bonxai.hpp:537-554
    bool setCellOn(const CoordT& coord, const DataT& default_value = DataT()) {
      sync();
      uint8_t was_on = 0;
      detail::check(bnx_grid_set_on(mutable_grid_->handle_, &coord.x, 1, &default_value, &was_on, BNX_HOST));
      return was_on != 0;
    }

    // bonxai.hpp:557-569 — the value is kept
    bool setCellOff(const CoordT& coord) {
      sync();
      uint8_t was_on = 0;
      detail::check(bnx_grid_set_off(mutable_grid_->handle_, &coord.x, 1, &was_on, BNX_HOST));
      return was_on != 0;
    }

    // push a value modified through the pointer returned by value()
    void sync() {
      if (!pending_) return;
      pending_ = false;
      if (std::memcmp(&this->slot_, &fetched_, sizeof(DataT)) != 0) {
        detail::check(bnx_grid_update_values(mutable_grid_->handle_, &pending_coord_.x, &this->slot_, 1, BNX_HOST));
      }
    }

   private:
    VoxelGrid* mutable_grid_;
    DataT fetched_{};
    CoordT pending_coord_{0, 0, 0};
    bool pending_ = false;
  };

  Accessor createAccessor() { return Accessor(*this); }
  ConstAccessor createConstAccessor() const { return ConstAccessor(*this); }

  // ---- key math, bonxai.hpp:419-447
  [[nodiscard]] CoordT getRootKey(const CoordT& c) const {
    const int32_t mask = ~((1 << Log2N) - 1);
    return {c.x & mask, c.y & mask, c.z & mask};
  }
  [[nodiscard]] CoordT getInnerKey(const CoordT& c) const {
    const int32_t mask = ~((1 << LEAF_BITS) - 1);
    return {c.x & mask, c.y & mask, c.z & mask};
  }
  [[nodiscard]] uint32_t getInnerIndex(const CoordT& c) const {
    const uint32_t m = (1u << INNER_BITS) - 1u;
    return ((uint32_t)(c.x >> LEAF_BITS) & m) | (((uint32_t)(c.y >> LEAF_BITS) & m) << INNER_BITS) |
           (((uint32_t)(c.z >> LEAF_BITS) & m) << (2 * INNER_BITS));
  }
  [[nodiscard]] uint32_t getLeafIndex(const CoordT& c) const {
    const uint32_t m = (1u << LEAF_BITS) - 1u;
    return ((uint32_t)c.x & m) | (((uint32_t)c.y & m) << LEAF_BITS) | (((uint32_t)c.z & m) << (2 * LEAF_BITS));
  }

 private:
  void release() {
    if (handle_ && owned_) bnx_grid_destroy(handle_);
    handle_ = nullptr;
  }
  void snapshot(std::vector<CoordT>& coords, std::vector<DataT>& values) const {
    int64_t n = 0;
    detail::check(bnx_grid_active_count(handle_, &n));
    coords.resize((size_t)n);
    values.resize((size_t)n);
    if (n == 0) return;
    int64_t got = 0;
    detail::check(bnx_grid_dump(handle_, &coords[0].x, values.data(), n, &got, BNX_HOST));
    coords.resize((size_t)got);
    values.resize((size_t)got);
  }

  uint32_t INNER_BITS = 2;
  uint32_t LEAF_BITS = 3;
  uint32_t Log2N = 5;
  double resolution = 0.0;
  double inv_resolution = 0.0;
  bnx_grid_t* handle_ = nullptr;
  bool owned_ = false;
};

using BinaryVoxelGrid = VoxelGrid<EmptyVoxel>;

}  // namespace Bonxai
