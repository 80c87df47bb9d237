// portfft::committed_descriptor<Scalar, Domain>: compute_forward / compute_backward on device (USM) pointers, the same
// overload set as /root/reference/src/portfft/committed_descriptor.hpp:171-310.  The SYCL buffer overloads (:58-162)
// have no CUDA meaning and are not provided.  A committed descriptor is copyable with the reference's semantics
// (committed_descriptor_impl.hpp:774-803): a copy shares the immutable device tables and owns its own workspaces
// (pfft_clone), so copies may compute concurrently on different queues.
#ifndef PFFT_B200_PORTFFT_COMMITTED_DESCRIPTOR_HPP
#define PFFT_B200_PORTFFT_COMMITTED_DESCRIPTOR_HPP

#include <complex>
#include <memory>
#include <utility>
#include <vector>

#include "descriptor.hpp"

namespace portfft {

template <typename Scalar, domain Domain>
class committed_descriptor {
  friend struct descriptor<Scalar, Domain>;

  descriptor<Scalar, Domain> params;
  queue queue_;
  pfft_plan* plan_ = nullptr;

  committed_descriptor(const descriptor<Scalar, Domain>& d, queue& q) : params(d), queue_(q) {
    pfft_desc c = params.to_c();
    detail::throw_on_status(pfft_commit(&c, q.device(), q.stream(), &plan_));
  }

  event run(direction dir, const void* in, const void* in_imag, void* out, void* out_imag,
            const std::vector<event>& dependencies) {
    for (const event& e : dependencies)
      if (e.native()) cudaStreamWaitEvent(queue_.stream(), e.native(), 0);
    detail::throw_on_status(pfft_compute(plan_, static_cast<int>(dir), in, in_imag, out, out_imag, queue_.stream()));
    return event::record(queue_.stream(), queue_.device());  // one event per call, owned by the returned object and its copies
  }

 public:
  committed_descriptor(const committed_descriptor& other) : params(other.params), queue_(other.queue_) {
    detail::throw_on_status(pfft_clone(other.plan_, &plan_));
  }
  committed_descriptor(committed_descriptor&& other) noexcept
      : params(std::move(other.params)), queue_(other.queue_), plan_(other.plan_) {
    other.plan_ = nullptr;
  }
  committed_descriptor& operator=(const committed_descriptor& other) {
    if (this != &other) {
      committed_descriptor tmp(other);
      swap(tmp);
    }
    return *this;
  }
  committed_descriptor& operator=(committed_descriptor&& other) noexcept {
    if (this != &other) {
      committed_descriptor tmp(std::move(other));
      swap(tmp);
    }
    return *this;
  }
  /// Waits for the queue's outstanding work before releasing the plan (committed_descriptor_impl.hpp:825-828).
  ~committed_descriptor() {
    if (plan_) pfft_destroy(plan_);
  }
  void swap(committed_descriptor& other) noexcept {
    std::swap(params, other.params);
    std::swap(queue_, other.queue_);
    std::swap(plan_, other.plan_);
  }

  using scalar_type = Scalar;
  using complex_type = std::complex<Scalar>;

  static_assert(Domain == domain::COMPLEX || Domain == domain::REAL, "unknown domain");

  const descriptor<Scalar, Domain>& get_descriptor() const noexcept { return params; }
  /// Level chosen for one dimension (thread / warp / block / multi-kernel).
  detail::level get_level(std::size_t dimension = 0) const {
    return static_cast<detail::level>(pfft_plan_level(plan_, dimension));
  }
  std::size_t get_workspace_bytes() const { return pfft_workspace_bytes(plan_); }
  /// Extension (no reference counterpart: a SYCL queue may run independent commands concurrently, a CUDA stream is
  /// in order): the queue this object submits to.  A copy bound to another queue of the same device computes
  /// concurrently with the original; each owns its workspaces.
  void set_queue(const queue& q) { queue_ = q; }
  const queue& get_queue() const noexcept { return queue_; }

  // ---- in-place ------------------------------------------------------------------------------------------------
  event compute_forward(complex_type* inout, const std::vector<event>& dependencies = {}) {
    return compute_forward(inout, inout, dependencies);
  }
  event compute_forward(scalar_type* inout_real, scalar_type* inout_imag, const std::vector<event>& dependencies = {}) {
    return compute_forward(inout_real, inout_imag, inout_real, inout_imag, dependencies);
  }
  event compute_backward(complex_type* inout, const std::vector<event>& dependencies = {}) {
    if constexpr (Domain == domain::REAL)
      return compute_backward(inout, reinterpret_cast<scalar_type*>(inout), dependencies);
    else
      return compute_backward(inout, inout, dependencies);
  }
  event compute_backward(scalar_type* inout_real, scalar_type* inout_imag,
                         const std::vector<event>& dependencies = {}) {
    return compute_backward(inout_real, inout_imag, inout_real, inout_imag, dependencies);
  }
  // ---- out-of-place ----------------------------------------------------------------------------------------------
  event compute_forward(const complex_type* in, complex_type* out, const std::vector<event>& dependencies = {}) {
    return run(direction::FORWARD, in, nullptr, out, nullptr, dependencies);
  }
  event compute_forward(const scalar_type* in_real, const scalar_type* in_imag, scalar_type* out_real,
                        scalar_type* out_imag, const std::vector<event>& dependencies = {}) {
    require_split();
    return run(direction::FORWARD, in_real, in_imag, out_real, out_imag, dependencies);
  }
  event compute_backward(const complex_type* in, complex_type* out, const std::vector<event>& dependencies = {}) {
    return run(direction::BACKWARD, in, nullptr, out, nullptr, dependencies);
  }
  event compute_backward(const scalar_type* in_real, const scalar_type* in_imag, scalar_type* out_real,
                         scalar_type* out_imag, const std::vector<event>& dependencies = {}) {
    require_split();
    return run(direction::BACKWARD, in_real, in_imag, out_real, out_imag, dependencies);
  }
  // ---- REAL domain: the overloads the reference reserves and leaves unimplemented (:134-137, :201-206, :273-278).
  // Forward: real array -> half spectrum (lengths[last] / 2 + 1 complex elements along the last dimension);
  // backward: half spectrum -> real array.
  event compute_forward(const scalar_type* in, complex_type* out, const std::vector<event>& dependencies = {}) {
    require_real();
    return run(direction::FORWARD, in, nullptr, out, nullptr, dependencies);
  }
  event compute_forward(scalar_type* inout, const std::vector<event>& dependencies = {}) {
    return compute_forward(inout, reinterpret_cast<complex_type*>(inout), dependencies);
  }
  event compute_forward(const scalar_type* in, scalar_type* out_real, scalar_type* out_imag,
                        const std::vector<event>& dependencies = {}) {
    require_real();
    require_split();
    return run(direction::FORWARD, in, nullptr, out_real, out_imag, dependencies);
  }
  event compute_backward(const complex_type* in, scalar_type* out, const std::vector<event>& dependencies = {}) {
    require_real();
    return run(direction::BACKWARD, in, nullptr, out, nullptr, dependencies);
  }
  event compute_backward(const scalar_type* in_real, const scalar_type* in_imag, scalar_type* out,
                         const std::vector<event>& dependencies = {}) {
    require_real();
    require_split();
    return run(direction::BACKWARD, in_real, in_imag, out, nullptr, dependencies);
  }

 private:
  void require_real() const {
    if (Domain != domain::REAL)
      throw invalid_configuration("real-to-complex / complex-to-real overloads need a descriptor of domain::REAL");
  }
  void require_split() const {
    if (params.complex_storage != complex_storage::SPLIT_COMPLEX)
      throw invalid_configuration(
          "To use split data layout, please set the storage in the descriptor to SPLIT_COMPLEX");
  }
};

}  // namespace portfft
#endif
