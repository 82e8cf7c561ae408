// Shared host/device definitions of the bonxai_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "bonxai_b200.h"

namespace bnx {

using u8 = uint8_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using i64 = int64_t;

// thread-local error text behind bnx_last_error()
void set_error(const std::string& msg);
const char* get_error();

struct StatusError {
  int code;
};

#define BNX_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::bnx::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                       std::to_string(__LINE__) + ")");                                         \
      return (_e == cudaErrorMemoryAllocation) ? BNX_ERR_NOMEM : BNX_ERR_CUDA;                  \
    }                                                                                           \
  } while (0)

#define BNX_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != BNX_OK) return _s; \
  } while (0)

#define BNX_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      ::bnx::set_error(std::string("invalid argument: ") + (msg)); \
      return BNX_ERR_INVALID;                          \
    }                                                  \
  } while (0)

// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
void note_launch();
long long launch_count();

// The B200 has 148 SMs; grids of the grid-stride kernels are sized in multiples of the SM count.
int sm_count();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// device-side error bits (GridCounters::error)
enum : u32 {
  ERR_LEAF_POOL = 1u,   // leaf pool exhausted
  ERR_INNER_POOL = 2u,  // inner-node pool exhausted
  ERR_ROOT_TABLE = 4u,  // root hash table full
  ERR_RAY_LIST = 8u,    // scan scratch overflow (should not happen: sized from n)
  ERR_SCAN = 16u,       // async pipeline: a scan's scratch lists overflowed; sticky until the host recovers
  ERR_PEER = 32u,       // sharded map, peer-memory exchange: a peer's arrival stamp did not show up in time (fatal)
};

}  // namespace bnx
