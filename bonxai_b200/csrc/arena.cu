#include "arena.hpp"

#include <mutex>

namespace bnx {

namespace {

struct DriverApi {
  CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*getErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

DriverApi g_drv;
std::once_flag g_drv_once;

template <class F>
bool load_sym(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

const DriverApi& drv() {
  std::call_once(g_drv_once, [] {
    bool ok = true;
    ok &= load_sym("cuMemAddressReserve", g_drv.memAddressReserve);
    ok &= load_sym("cuMemAddressFree", g_drv.memAddressFree);
    ok &= load_sym("cuMemCreate", g_drv.memCreate);
    ok &= load_sym("cuMemRelease", g_drv.memRelease);
    ok &= load_sym("cuMemMap", g_drv.memMap);
    ok &= load_sym("cuMemUnmap", g_drv.memUnmap);
    ok &= load_sym("cuMemSetAccess", g_drv.memSetAccess);
    ok &= load_sym("cuMemGetAllocationGranularity", g_drv.memGetAllocationGranularity);
    ok &= load_sym("cuGetErrorString", g_drv.getErrorString);
    g_drv.ok = ok;
  });
  return g_drv;
}

int drv_fail(const char* what, CUresult r) {
  const char* s = nullptr;
  if (g_drv.getErrorString) g_drv.getErrorString(r, &s);
  set_error(std::string(what) + ": " + (s ? s : "unknown driver error"));
  return r == CUDA_ERROR_OUT_OF_MEMORY ? BNX_ERR_NOMEM : BNX_ERR_CUDA;
}

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  return prop;
}

}  // namespace

int Arena::init(size_t reserve_bytes) {
  BNX_CUDA(cudaFree(0));  // make sure the primary context exists
  const DriverApi& d = drv();
  if (!d.ok) {
    set_error("CUDA virtual memory management entry points unavailable");
    return BNX_ERR_CUDA;
  }
  BNX_CUDA(cudaGetDevice(&device_));
  CUmemAllocationProp prop = alloc_prop(device_);
  CUresult r = d.memGetAllocationGranularity(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS) return drv_fail("cuMemGetAllocationGranularity", r);
  reserved_ = round_up(reserve_bytes, gran_);
  r = d.memAddressReserve(&base_, reserved_, 0, 0, 0);
  if (r != CUDA_SUCCESS) {
    base_ = 0;
    reserved_ = 0;
    return drv_fail("cuMemAddressReserve", r);
  }
  mapped_ = 0;
  return BNX_OK;
}

int Arena::grow_to(size_t bytes, cudaStream_t stream) {
  if (bytes <= mapped_) return BNX_OK;
  const DriverApi& d = drv();
  size_t target = round_up(bytes, gran_);
  if (target > reserved_) {
    set_error("arena: request exceeds the reserved address range");
    return BNX_ERR_NOMEM;
  }
  const size_t add = target - mapped_;
  CUmemAllocationProp prop = alloc_prop(device_);
  Chunk c;
  c.offset = mapped_;
  c.size = add;
  CUresult r = d.memCreate(&c.handle, add, &prop, 0);
  if (r != CUDA_SUCCESS) return drv_fail("cuMemCreate", r);
  r = d.memMap(base_ + c.offset, add, 0, c.handle, 0);
  if (r != CUDA_SUCCESS) {
    d.memRelease(c.handle);
    return drv_fail("cuMemMap", r);
  }
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.memSetAccess(base_ + c.offset, add, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d.memUnmap(base_ + c.offset, add);
    d.memRelease(c.handle);
    return drv_fail("cuMemSetAccess", r);
  }
  chunks_.push_back(c);
  mapped_ = target;
  BNX_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(base_ + c.offset), 0, add, stream));
  return BNX_OK;
}

int Arena::reset() {
  const DriverApi& d = drv();
  for (auto& c : chunks_) {
    d.memUnmap(base_ + c.offset, c.size);
    d.memRelease(c.handle);
  }
  chunks_.clear();
  mapped_ = 0;
  return BNX_OK;
}

void Arena::destroy() {
  if (!base_) return;
  reset();
  drv().memAddressFree(base_, reserved_);
  base_ = 0;
  reserved_ = 0;
}

}  // namespace bnx
