// oracle/shim: the slice of thrust the reference's proposal / NMS files use, over host memory.
// stable_sort_by_key(policy, keys, keys_end, values, comp): keys and values permuted together, equal keys keep
// their order - the documented thrust semantics.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../mxnet_shim.h"

namespace thrust {
struct shim_policy {};
static const shim_policy device = shim_policy(), host = shim_policy();
template <typename T>
struct greater {
  bool operator()(const T& a, const T& b) const { return a > b; }
};
template <typename T>
struct less {
  bool operator()(const T& a, const T& b) const { return a < b; }
};
template <typename Policy, typename K, typename V, typename Cmp>
inline void stable_sort_by_key(const Policy&, K* kfirst, K* klast, V* vfirst, Cmp cmp) {
  const size_t n = (size_t)(klast - kfirst);
  std::vector<size_t> idx(n);
  std::iota(idx.begin(), idx.end(), (size_t)0);
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return cmp(kfirst[a], kfirst[b]); });
  std::vector<K> k(n);
  std::vector<V> v(n);
  for (size_t i = 0; i < n; ++i) { k[i] = kfirst[idx[i]]; v[i] = vfirst[idx[i]]; }
  std::copy(k.begin(), k.end(), kfirst);
  std::copy(v.begin(), v.end(), vfirst);
}
template <typename Policy, typename K, typename V>
inline void stable_sort_by_key(const Policy& p, K* kfirst, K* klast, V* vfirst) {
  stable_sort_by_key(p, kfirst, klast, vfirst, less<K>());
}
template <typename Policy, typename T>
inline void sequence(const Policy&, T* first, T* last) { std::iota(first, last, T(0)); }
}  // namespace thrust
