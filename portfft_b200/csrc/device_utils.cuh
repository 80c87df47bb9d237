// Small device helpers shared by the kernels.
#pragma once
#include "dft.cuh"

namespace pfft {

template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  using type = float2;
};
template <>
struct VecOf<double> {
  using type = double2;
};

// read-only (L1-cached) load of one complex from a device-resident table
template <typename T>
__device__ __forceinline__ cx<T> ldg_cx(const void* base, long long i) {
  const typename VecOf<T>::type v = __ldg(reinterpret_cast<const typename VecOf<T>::type*>(base) + i);
  return {v.x, v.y};
}

}  // namespace pfft
