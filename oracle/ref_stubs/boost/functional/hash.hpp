// Stand-in for boost/functional/hash.hpp: hash_combine as Boost published it up to
// 1.80 (seed ^= hash(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2), integers hash to
// themselves).  In the reference it only orders an unordered_map, i.e. the
// NUMBERING of clusters, never the partition.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <cstddef>

namespace boost {

template <class T> inline void hash_combine(std::size_t &seed, const T &v) {
  seed ^= static_cast<std::size_t>(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}

}  // namespace boost
