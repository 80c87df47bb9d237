// Multi-GPU extension of the portFFT API mirror (no reference counterpart: the reference commits to one sycl::queue,
// /root/reference/src/portfft/committed_descriptor_impl.hpp:109).  One process drives several GPUs:
//
//   portfft::distributed::sharded_descriptor<float, domain::COMPLEX> plan(desc, {0, 1, 2, 3});   // batch sharding
//   plan.compute_forward(host_in, host_out);                                   // un-sharded pinned host buffers
//
//   portfft::distributed::slab_descriptor<float> fft3d(desc3d, {0, 1, 2, 3});                    // one 3-D transform
//   fft3d.compute_forward(x_slabs);          // x_slabs[r]: [n0 / W][n1][n2] on GPU r; returns the y-slabs of the spectrum
//
// Thin wrappers over pfft_commit_multi / pfft_slab_commit_local (include/pfft.h, csrc/multi.cu).  With one process per
// GPU, use the C entry points directly (pfft_commit_shard, pfft_slab_commit + pfft_slab_export / pfft_slab_import).
#ifndef PFFT_B200_PORTFFT_DISTRIBUTED_HPP
#define PFFT_B200_PORTFFT_DISTRIBUTED_HPP

#include <complex>
#include <utility>
#include <vector>

#include "descriptor.hpp"

namespace portfft {
namespace distributed {

/// Batch sharding: GPU r of `devices` transforms a contiguous range of the descriptor's transforms (pfft_partition).
template <typename Scalar, domain Domain>
class sharded_descriptor {
  descriptor<Scalar, Domain> params;
  pfft_multi* multi_ = nullptr;

 public:
  using complex_type = std::complex<Scalar>;
  struct shard {
    std::size_t first, count, forward_start, backward_start;
  };

  sharded_descriptor(const descriptor<Scalar, Domain>& d, const std::vector<int>& devices) : params(d) {
    pfft_desc c = params.to_c();
    detail::throw_on_status(pfft_commit_multi(&c, static_cast<int>(devices.size()), devices.data(), nullptr, &multi_));
  }
  sharded_descriptor(const sharded_descriptor&) = delete;
  sharded_descriptor& operator=(const sharded_descriptor&) = delete;
  sharded_descriptor(sharded_descriptor&& o) noexcept : params(std::move(o.params)), multi_(o.multi_) { o.multi_ = nullptr; }
  ~sharded_descriptor() {
    if (multi_) pfft_multi_destroy(multi_);
  }

  int size() const { return pfft_multi_size(multi_); }
  shard get_shard(int r) const {
    pfft_shard_info i{};
    detail::throw_on_status(pfft_multi_shard(multi_, r, &i, nullptr));
    return shard{i.first, i.count, i.forward_start, i.backward_start};
  }

  /// Rank-local device buffers, one per GPU; asynchronous (wait()).  Interleaved storage.
  void compute_forward(const std::vector<const complex_type*>& in, const std::vector<complex_type*>& out) {
    run(PFFT_FORWARD, in, out);
  }
  void compute_backward(const std::vector<const complex_type*>& in, const std::vector<complex_type*>& out) {
    run(PFFT_BACKWARD, in, out);
  }
  /// Un-sharded HOST buffers (batch-major layouts): scatter, transform and gather pipelined per GPU; blocking.
  void compute_forward(const complex_type* host_in, complex_type* host_out) {
    detail::throw_on_status(pfft_multi_compute_host(multi_, PFFT_FORWARD, host_in, nullptr, host_out, nullptr));
  }
  void compute_backward(const complex_type* host_in, complex_type* host_out) {
    detail::throw_on_status(pfft_multi_compute_host(multi_, PFFT_BACKWARD, host_in, nullptr, host_out, nullptr));
  }
  void wait() { detail::throw_on_status(pfft_multi_sync(multi_)); }

 private:
  void run(int dir, const std::vector<const complex_type*>& in, const std::vector<complex_type*>& out) {
    if (static_cast<int>(in.size()) != size() || static_cast<int>(out.size()) != size())
      throw invalid_configuration("one buffer per GPU expected");
    std::vector<const void*> i(in.begin(), in.end());
    std::vector<void*> o(out.begin(), out.end());
    detail::throw_on_status(pfft_multi_compute(multi_, dir, i.data(), nullptr, o.data(), nullptr));
  }
};

/// One 3-D complex transform slab-decomposed over `devices` (entries may repeat): rank r holds the x-planes
/// [r n0/W, (r+1) n0/W) of the input and receives the y-rows [r n1/W, (r+1) n1/W) of the spectrum.
template <typename Scalar>
class slab_descriptor {
  std::vector<pfft_slab*> slabs_;

 public:
  using complex_type = std::complex<Scalar>;

  slab_descriptor(const descriptor<Scalar, domain::COMPLEX>& d, const std::vector<int>& devices)
      : slabs_(devices.size(), nullptr) {
    pfft_desc c = d.to_c();
    detail::throw_on_status(
        pfft_slab_commit_local(&c, static_cast<int>(devices.size()), devices.data(), nullptr, slabs_.data()));
  }
  slab_descriptor(const slab_descriptor&) = delete;
  slab_descriptor& operator=(const slab_descriptor&) = delete;
  ~slab_descriptor() {
    for (pfft_slab* s : slabs_)
      if (s) pfft_slab_destroy(s);
  }

  int size() const { return static_cast<int>(slabs_.size()); }
  /// complex elements of every rank-local buffer
  std::size_t get_slab_count() const { return pfft_slab_elems(slabs_[0]); }

  /// Asynchronous; the returned pointers ([n0][n1 / W][n2] on GPU r) stay valid until the next call.
  std::vector<complex_type*> compute_forward(const std::vector<const complex_type*>& x_slabs) {
    if (x_slabs.size() != slabs_.size()) throw invalid_configuration("one x-slab per GPU expected");
    std::vector<complex_type*> out(slabs_.size(), nullptr);
    for (std::size_t r = 0; r < slabs_.size(); ++r) {
      void* p = nullptr;
      detail::throw_on_status(pfft_slab_forward(slabs_[r], x_slabs[r], &p));
      out[r] = static_cast<complex_type*>(p);
    }
    return out;
  }
  void compute_backward(const std::vector<const complex_type*>& y_slabs, const std::vector<complex_type*>& x_slabs) {
    if (y_slabs.size() != slabs_.size() || x_slabs.size() != slabs_.size())
      throw invalid_configuration("one slab per GPU expected");
    for (std::size_t r = 0; r < slabs_.size(); ++r)
      detail::throw_on_status(pfft_slab_backward(slabs_[r], y_slabs[r], x_slabs[r]));
  }
  void wait() {
    for (pfft_slab* s : slabs_) detail::throw_on_status(pfft_slab_sync(s));
  }
};

}  // namespace distributed
}  // namespace portfft
#endif
