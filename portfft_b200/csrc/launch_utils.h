// Host-side launch helpers shared by the kernel files.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>

namespace pfft {

constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// SM count of the current device, cached per device ordinal (a process may drive several, possibly different, GPUs)
inline int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

namespace detail {
// (kernel, device) -> value table for the launch path: lock-free look-ups (open addressing over atomics), the rare
// insertions serialised by a mutex.  Replaces a std::map under a mutex on every launch (launch-bound problems such
// as BASELINE config C1 pay for every host instruction between two launches).
struct KernelDeviceTable {
  static constexpr unsigned kSlots = 4096;
  std::atomic<uintptr_t> key[kSlots];
  std::atomic<long long> val[kSlots];
  std::mutex mu;
  static uintptr_t make_key(const void* kern, int dev) {
    return (reinterpret_cast<uintptr_t>(kern) << 6) ^ (uintptr_t)(dev + 1);  // device < 64; kernels are >= 64 B apart
  }
  static unsigned hash(uintptr_t k) { return (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 52); }
  bool find(uintptr_t k, long long* v) const {
    for (unsigned i = hash(k), n = 0; n < kSlots; i = (i + 1) % kSlots, ++n) {
      const uintptr_t kk = key[i].load(std::memory_order_acquire);
      if (kk == k) {
        *v = val[i].load(std::memory_order_relaxed);
        return true;
      }
      if (kk == 0) return false;
    }
    return false;
  }
  void put(uintptr_t k, long long v) {
    std::lock_guard<std::mutex> lock(mu);
    for (unsigned i = hash(k), n = 0; n < kSlots; i = (i + 1) % kSlots, ++n) {
      const uintptr_t kk = key[i].load(std::memory_order_relaxed);
      if (kk == k || kk == 0) {
        val[i].store(v, std::memory_order_relaxed);
        key[i].store(k, std::memory_order_release);
        return;
      }
    }
  }
};
inline KernelDeviceTable& smem_table() {
  static KernelDeviceTable t;
  return t;
}
inline KernelDeviceTable& slot_table() {
  static KernelDeviceTable t;
  return t;
}
}  // namespace detail

// Raise the dynamic shared-memory limit of `kern` on the current device -- once per (kernel, device) and size, not on
// every launch (cudaFuncSetAttribute costs about as much as the launch itself).
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kern, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  const uintptr_t key = detail::KernelDeviceTable::make_key(reinterpret_cast<const void*>(kern), current_device());
  long long have = 0;
  if (detail::smem_table().find(key, &have) && (size_t)have >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) detail::smem_table().put(key, (long long)bytes);
  return e;
}

// Resident CTAs of `kern` on the whole current device (occupancy x SMs) for a persistent grid; cached per
// (kernel, device).  0: the kernel cannot run with this configuration.
template <typename K>
inline int persistent_slots(K kern, int threads, size_t smem) {
  const uintptr_t key = detail::KernelDeviceTable::make_key(reinterpret_cast<const void*>(kern), current_device());
  long long have = 0;
  if (detail::slot_table().find(key, &have)) return (int)have;
  int occ = 0;
  if (ensure_dynamic_smem(kern, smem) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess) return 0;
  const int slots = occ * sm_count();
  if (slots > 0) detail::slot_table().put(key, slots);
  return slots;
}

}  // namespace pfft
