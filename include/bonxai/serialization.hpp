// bonxai_b200 drop-in for bonxai_core/include/bonxai/serialization.hpp.
//
// Same names (Serialize, HeaderInfo, GetHeaderInfo, Deserialize) and the same stream format
// (serialization.hpp:77-116): a text line `Bonxai::VoxelGrid<TYPE,INNER_BITS,LEAF_BITS>(RESOLUTION)\n`, the u32
// root count, then per root its key, inner mask, and per ON leaf the leaf mask followed by the ON cells' raw
// bytes. The bytes are produced by a compaction kernel over the device node pools (bnx_grid_serialize); streams
// written by the reference load here and vice versa. Root order inside a stream is unspecified on both sides.
#pragma once

#include <cstdlib>
#include <istream>
#include <iterator>
#include <memory>
#include <ostream>
#include <string>
#include <typeinfo>
#include <vector>

#include "bonxai/bonxai.hpp"

#ifdef __GNUG__
#include <cxxabi.h>
#endif

namespace Bonxai {

struct HeaderInfo {
  std::string type_name;
  int inner_bits = 0;
  int leaf_bits = 0;
  double resolution = 0;
};

namespace details {
// the header carries the demangled name of DataT, like the reference's (serialization.hpp:50-66)
inline std::string demangle(const char* name) {
#ifdef __GNUG__
  int status = -4;
  std::unique_ptr<char, void (*)(void*)> res{abi::__cxa_demangle(name, nullptr, nullptr, &status), std::free};
  return status == 0 ? res.get() : name;
#else
  return name;
#endif
}
}  // namespace details

template <typename DataT>
inline void Serialize(std::ostream& out, const VoxelGrid<DataT>& grid) {
  const std::string type_name = details::demangle(typeid(DataT).name());
  int64_t size = 0;
  detail::check(bnx_grid_serialize(grid.handle(), type_name.c_str(), nullptr, 0, &size));
  std::vector<uint8_t> bytes((size_t)size);
  detail::check(bnx_grid_serialize(grid.handle(), type_name.c_str(), bytes.data(), size, &size));
  out.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)size);
}

// parses the first line of a stream (without or with the trailing newline)
inline HeaderInfo GetHeaderInfo(std::string header) {
  const std::string prefix = "Bonxai::VoxelGrid<";
  if (header.rfind(prefix, 0) != 0) throw std::runtime_error("Header wasn't recognized");
  while (!header.empty() && (header.back() == '\n' || header.back() == '\r')) header.pop_back();
  const size_t gt = header.rfind(">(");
  if (gt == std::string::npos || header.back() != ')') throw std::runtime_error("Header wasn't recognized");
  const std::string inside = header.substr(prefix.size(), gt - prefix.size());
  const size_t c2 = inside.rfind(',');
  const size_t c1 = c2 == std::string::npos || c2 == 0 ? std::string::npos : inside.rfind(',', c2 - 1);
  if (c1 == std::string::npos) throw std::runtime_error("Header wasn't recognized");
  HeaderInfo info;
  info.type_name = inside.substr(0, c1);
  info.inner_bits = std::stoi(inside.substr(c1 + 1, c2 - c1 - 1));
  info.leaf_bits = std::stoi(inside.substr(c2 + 1));
  info.resolution = std::stod(header.substr(gt + 2, header.size() - gt - 3));
  return info;
}

// `input` is positioned after the header line, as in the reference (examples/test_serialization.cpp:31-36)
template <typename DataT>
inline VoxelGrid<DataT> Deserialize(std::istream& input, HeaderInfo info) {
  const std::string type_name = details::demangle(typeid(DataT).name());
  if (type_name != info.type_name) throw std::runtime_error("DataT does not match");
  char header[300];
  std::snprintf(header, sizeof(header), "Bonxai::VoxelGrid<%s,%d,%d>(%lf)\n", info.type_name.c_str(), info.inner_bits, info.leaf_bits,
                info.resolution);
  std::vector<uint8_t> bytes(header, header + std::strlen(header));
  bytes.insert(bytes.end(), std::istreambuf_iterator<char>(input), std::istreambuf_iterator<char>());
  bnx_grid_t* h = nullptr;
  detail::check(bnx_grid_deserialize(bytes.data(), (int64_t)bytes.size(), (int)sizeof(DataT), type_name.c_str(), &h));
  VoxelGrid<DataT> grid(h, info.resolution, (uint8_t)info.inner_bits, (uint8_t)info.leaf_bits);
  grid.adopt();
  return grid;
}

}  // namespace Bonxai
