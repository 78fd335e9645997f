// Minimal stand-in for <boost/functional/hash.hpp> used ONLY to compile the unmodified reference
// sources into oracle/_ref (test infrastructure, not product code). The reference only uses
// hash_combine to pick unordered_set buckets (src/hash_dup_remover.hpp:49,60-64); the mixer never
// influences which records survive, so any reasonable mixer is behaviour-preserving.
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <class T>
inline void hash_combine(std::size_t& seed, const T& v) {
    std::size_t h = std::hash<T>{}(v);
    h *= 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    seed ^= h + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
}  // namespace boost
