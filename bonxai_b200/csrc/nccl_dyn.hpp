// NCCL entry points resolved at run time (dlopen): the library keeps loading on hosts without NCCL or a GPU, and a
// process that already carries NCCL (torch's bundled copy) shares that copy instead of a second one.
#pragma once

#include <dlfcn.h>
#include <nccl.h>  // types only; no link-time dependency

#include "common.cuh"

namespace bnx {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;  // optional: frees a communicator without waiting for the peers
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  void* handle = nullptr;
  bool ok = false;
};

// path may be NULL/empty: "libnccl.so.2" is looked up the usual way (an already loaded copy wins)
inline const NcclApi& nccl_api(const char* path) {
  static NcclApi api;
  if (api.ok) return api;
  const char* candidates[3] = {path && *path ? path : nullptr, "libnccl.so.2", "libnccl.so"};
  for (const char* c : candidates) {
    if (!c) continue;
    api.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
  auto sym = [&](const char* n) { return dlsym(api.handle, n); };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.CommAbort = reinterpret_cast<decltype(api.CommAbort)>(sym("ncclCommAbort"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv &&
           api.AllReduce && api.AllGather && api.GetErrorString;
  return api;
}

#define BNX_NCCL(api, expr)                                                                               \
  do {                                                                                                    \
    ncclResult_t _r = (expr);                                                                             \
    if (_r != ncclSuccess) {                                                                              \
      ::bnx::set_error(std::string(#expr) + ": " + ((api).GetErrorString ? (api).GetErrorString(_r) : "NCCL error")); \
      return BNX_ERR_CUDA;                                                                                \
    }                                                                                                     \
  } while (0)

}  // namespace bnx
