// Shim: boost::unordered_map -> std::unordered_map.  Iteration order differs from Boost's; the reference's results
// depend on it only through float summation order and .pss line order.
#pragma once
#include <unordered_map>
namespace boost {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>, class A = std::allocator<std::pair<const K, V>>>
using unordered_map = std::unordered_map<K, V, H, E, A>;
}
