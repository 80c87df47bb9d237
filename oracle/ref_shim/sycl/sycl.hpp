// Minimal host stand-in for <sycl/sycl.hpp>: just enough surface for the reference's header-only device math
// (common/workitem.hpp, common/subgroup.hpp) and host logic (descriptor_validation.hpp, utils.hpp) to compile with
// g++ and run on the CPU.  TEST INFRASTRUCTURE ONLY (used by oracle/ref_shim/ref_build.cpp).  Written for this
// repository; it contains no reference code.  A sub-group is emulated by 32 host threads that run in lock step:
// every cross-lane operation publishes the lane's value, waits on a barrier, reads the peer's value, waits again.
#pragma once
#include <algorithm>
#include <array>
#include <numeric>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

namespace sycl {

namespace access {
enum class address_space { global_space, local_space, private_space, generic_space };
enum class decorated { no, yes, legacy };
enum class mode { read, write, read_write };
}  // namespace access

template <access::address_space, access::decorated, typename T>
T* address_space_cast(T* p) {
  return p;
}

class handler {};
class event {
 public:
  void wait() {}
};
class device {};
class context {};
class queue {
 public:
  template <typename... A>
  event submit(A&&...) { return {}; }
  void wait() {}
};
class kernel_id {};
template <typename K>
kernel_id get_kernel_id() { return {}; }
template <typename T>
struct specialization_id {
  T v;
  constexpr specialization_id() : v{} {}
  constexpr explicit specialization_id(T x) : v(x) {}
};
template <typename T>
T* malloc_device(std::size_t n, queue&) { return static_cast<T*>(std::malloc(n * sizeof(T))); }
inline void free(void* p, queue&) { std::free(p); }

template <typename T, int D = 1>
class buffer {
 public:
  std::size_t size() const { return 0; }
  template <access::mode M>
  T* get_access(handler&) { return nullptr; }
  template <typename U, int E = 1>
  buffer<U, E> reinterpret(std::size_t) const { return {}; }
};
template <int D>
struct range {
  std::size_t v[D];
};
template <int D>
struct id {
  std::size_t v[D];
};
template <typename T, int N>
struct vec {
  T s[N];
  T& operator[](int i) { return s[i]; }
  const T& operator[](int i) const { return s[i]; }
};

inline float cospi(float x) { return static_cast<float>(std::cos(M_PI * static_cast<double>(x))); }
inline float sinpi(float x) { return static_cast<float>(std::sin(M_PI * static_cast<double>(x))); }
inline double cospi(double x) { return static_cast<double>(std::cos(3.14159265358979323846264338327950288L * (long double)x)); }
inline double sinpi(double x) { return static_cast<double>(std::sin(3.14159265358979323846264338327950288L * (long double)x)); }

// ---- lock-step sub-group emulation --------------------------------------------------------------------------
struct sg_shared_state {
  static constexpr int lanes = 32;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  unsigned long generation = 0;
  double slot[lanes];
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    unsigned long gen = generation;
    if (++waiting == lanes) {
      waiting = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != generation; });
    }
  }
};

class sub_group {
 public:
  using linear_id_type = std::uint32_t;
  sub_group() = default;
  sub_group(sg_shared_state* s, linear_id_type lane, linear_id_type group = 0) : state(s), lane_id(lane), group_id(group) {}
  linear_id_type get_local_linear_id() const { return lane_id; }
  linear_id_type get_local_range_size() const { return sg_shared_state::lanes; }
  linear_id_type get_group_id() const { return group_id; }  // index of the sub-group inside its work-group
  linear_id_type get_group_linear_range() const { return 1; }
  // Intel block loads / stores (transfers.hpp:217-314, compiled out by PORTFFT_USE_SG_TRANSFERS = off): declared only
  template <typename P>
  auto load(P p) const { return *p; }
  template <int N, typename P>
  auto load(P p) const { return *p; }
  template <typename P, typename V>
  void store(P, const V&) const {}
  template <int N, typename P, typename V>
  void store(P, const V&) const {}
  sg_shared_state* state = nullptr;
  linear_id_type lane_id = 0;
  linear_id_type group_id = 0;
};

// ---- work-group emulation: every work-item is a host thread; one barrier object per work-group ---------------------
struct wg_shared_state {
  std::mutex m;
  std::condition_variable cv;
  int size = 1, waiting = 0;
  unsigned long generation = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    unsigned long gen = generation;
    if (++waiting == size) {
      waiting = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != generation; });
    }
  }
};
template <int D>
struct group {
  wg_shared_state* state = nullptr;
};
template <int D>
inline void group_barrier(group<D> g) {
  if (g.state) g.state->barrier();
}

template <typename T>
T select_from_group(sub_group sg, T value, std::size_t source_lane) {
  sg.state->slot[sg.lane_id] = static_cast<double>(value);
  sg.state->barrier();
  T r = static_cast<T>(sg.state->slot[source_lane % sg_shared_state::lanes]);
  sg.state->barrier();
  return r;
}
template <typename T>
T permute_group_by_xor(sub_group sg, T value, sub_group::linear_id_type mask) {
  return select_from_group(sg, value, (sg.lane_id ^ mask) % sg_shared_state::lanes);
}
inline void group_barrier(sub_group sg) { sg.state->barrier(); }

template <int D>
class nd_item {
 public:
  nd_item() = default;
  nd_item(wg_shared_state* wg, sub_group sg, std::size_t local_id, std::size_t local_range)
      : wg_(wg), sg_(sg), local_id_(local_id), local_range_(local_range) {}
  sub_group get_sub_group() const { return sg_; }
  group<D> get_group() const { return {wg_}; }
  std::size_t get_group(int) const { return 0; }
  std::size_t get_group_range(int) const { return 1; }
  std::size_t get_local_linear_id() const { return local_id_; }
  std::size_t get_local_id(int) const { return local_id_; }
  std::size_t get_global_linear_id() const { return local_id_; }
  std::size_t get_local_range(int) const { return local_range_; }
  std::size_t get_global_range(int) const { return local_range_; }
  std::size_t get_group_linear_id() const { return 0; }

 private:
  wg_shared_state* wg_ = nullptr;
  sub_group sg_;
  std::size_t local_id_ = 0, local_range_ = 1;
};
class stream {
 public:
  stream(std::size_t, std::size_t, handler&) {}
  template <typename T>
  const stream& operator<<(const T&) const { return *this; }
};
inline constexpr struct flush_t {
} flush{};
inline constexpr struct endl_t {
} endl{};

}  // namespace sycl
