// Host-side launch helpers shared by the kernel files.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

namespace pfft {

// Raise the dynamic shared-memory limit of `kern` on the current device -- once per (kernel, device) and size, not on
// every launch (cudaFuncSetAttribute costs about as much as the launch itself, which matters for the launch-bound
// small problems such as BASELINE config C1).
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kern, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const auto key = std::make_pair(reinterpret_cast<const void*>(kern), dev);
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done[key] = bytes;
  return e;
}

}  // namespace pfft
