// Growable device memory region with a STABLE base address.
//
// Node pools must grow while kernels hold indices (and the drop-in API holds pointers) into them, and a
// city-scale map is tens of GB: copy-on-grow would need 2x the memory. The arena reserves a large virtual
// range once (cuMemAddressReserve) and maps physical 2 MiB-granular chunks behind it on demand
// (cuMemCreate + cuMemMap), zero-filling each new chunk. Driver entry points are resolved through
// cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda.
#pragma once

#include <cuda.h>

#include <vector>

#include "common.cuh"

namespace bnx {

class Arena {
 public:
  Arena() = default;
  ~Arena() { destroy(); }
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;

  // reserve `reserve_bytes` of virtual address space on the current device
  int init(size_t reserve_bytes);
  // make at least `bytes` usable (mapped + zeroed); keeps the base pointer. Stream-ordered zero fill.
  int grow_to(size_t bytes, cudaStream_t stream);
  // unmap everything but keep the reservation
  int reset();
  void destroy();

  void* base() const { return reinterpret_cast<void*>(base_); }
  size_t mapped() const { return mapped_; }
  size_t reserved() const { return reserved_; }

 private:
  CUdeviceptr base_ = 0;
  size_t reserved_ = 0;
  size_t mapped_ = 0;
  size_t gran_ = 0;
  int device_ = 0;
  struct Chunk {
    CUmemGenericAllocationHandle handle;
    size_t offset, size;
  };
  std::vector<Chunk> chunks_;
};

}  // namespace bnx
