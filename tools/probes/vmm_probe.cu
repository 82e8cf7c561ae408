// Probe: do cuMemCreate / cuMemMap / cuMemSetAccess block while a kernel runs on another stream, and what do they cost
// per GB? (decides whether the node pools can be grown behind a running scan pipeline). Build:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/vmm_probe tools/probes/vmm_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>

__global__ void spin(unsigned long long ns, unsigned* sink) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
  if (sink && t == 1) *sink = 1;
}
__global__ void touch(unsigned char* p, size_t n) {
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4096; i < n; i += (size_t)gridDim.x * blockDim.x * 4096) p[i] = 1;
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x)                                                    \
  do {                                                           \
    CUresult r_ = (x);                                           \
    if (r_ != CUDA_SUCCESS) {                                    \
      const char* s_;                                            \
      cuGetErrorString(r_, &s_);                                 \
      printf("%s failed: %s\n", #x, s_);                         \
      return 1;                                                  \
    }                                                            \
  } while (0)

int main() {
  cudaFree(0);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = 0;
  size_t gran = 0;
  CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  printf("granularity %zu\n", gran);
  CUdeviceptr base;
  const size_t reserve = 64ull << 30;
  CK(cuMemAddressReserve(&base, reserve, 0, 0, 0));
  size_t off = 0;
  for (int busy = 0; busy < 2; ++busy) {
    for (size_t mb : {64, 256, 1024, 4096}) {
      const size_t bytes = mb << 20;
      if (busy) spin<<<148, 256, 0, s1>>>(200000000ull, nullptr);  // 200 ms of a busy GPU on stream 1
      const double t0 = now_ms();
      CUmemGenericAllocationHandle h;
      CK(cuMemCreate(&h, bytes, &prop, 0));
      const double t1 = now_ms();
      CK(cuMemMap(base + off, bytes, 0, h, 0));
      const double t2 = now_ms();
      CUmemAccessDesc acc = {};
      acc.location = prop.location;
      acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      CK(cuMemSetAccess(base + off, bytes, &acc, 1));
      const double t3 = now_ms();
      cudaMemsetAsync((void*)(base + off), 0, bytes, s2);
      cudaStreamSynchronize(s2);
      const double t4 = now_ms();
      const bool still_busy = busy && cudaStreamQuery(s1) == cudaErrorNotReady;
      cudaStreamSynchronize(s1);
      printf("busy=%d %5zu MB: create %.2f ms, map %.2f ms, setaccess %.2f ms, memset %.2f ms; spin kernel still running afterwards: %d\n", busy, mb, t1 - t0,
             t2 - t1, t3 - t2, t4 - t3, (int)still_busy);
      off += bytes;
    }
  }
  // many small chunks vs one: 1 GB as 16 x 64 MB
  {
    const double t0 = now_ms();
    for (int k = 0; k < 16; ++k) {
      CUmemGenericAllocationHandle h;
      CK(cuMemCreate(&h, 64ull << 20, &prop, 0));
      CK(cuMemMap(base + off, 64ull << 20, 0, h, 0));
      off += 64ull << 20;
    }
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(base + off - (1ull << 30), 1ull << 30, &acc, 1));
    printf("1 GB as 16 x 64 MB chunks + one setaccess: %.2f ms\n", now_ms() - t0);
  }
  touch<<<148, 256, 0, s2>>>((unsigned char*)base, off);
  printf("touch: %s\n", cudaGetErrorString(cudaStreamSynchronize(s2)));
  return 0;
}
