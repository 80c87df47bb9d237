// portfft::descriptor<Scalar, Domain>: the user-facing problem description, same public mutable fields, defaults,
// getters and commit() as /root/reference/src/portfft/descriptor.hpp:43-271, implemented over the C ABI.
#ifndef PFFT_B200_PORTFFT_DESCRIPTOR_HPP
#define PFFT_B200_PORTFFT_DESCRIPTOR_HPP

#include <cstddef>
#include <type_traits>
#include <vector>

#include "../pfft.h"
#include "enums.hpp"
#include "exceptions.hpp"
#include "queue.hpp"

namespace portfft {

template <typename Scalar, domain Domain>
class committed_descriptor;

namespace detail {
/// Row-major strides with the last dimension contiguous (reference: utils.hpp:190-201).
inline std::vector<std::size_t> get_default_strides(const std::vector<std::size_t>& lengths) {
  std::vector<std::size_t> strides(lengths.size());
  std::size_t running = 1;
  for (std::size_t i = lengths.size(); i-- > 0;) {
    strides[i] = running;
    running *= lengths[i];
  }
  return strides;
}
}  // namespace detail

template <typename DescScalar, domain DescDomain>
struct descriptor {
  using Scalar = DescScalar;
  static_assert(std::is_same_v<Scalar, float> || std::is_same_v<Scalar, double>, "Scalar must be float or double");
  static constexpr domain Domain = DescDomain;

  std::vector<std::size_t> lengths;
  Scalar forward_scale = 1;
  Scalar backward_scale = 1;
  std::size_t number_of_transforms = 1;
  portfft::complex_storage complex_storage = portfft::complex_storage::INTERLEAVED_COMPLEX;
  portfft::placement placement = portfft::placement::OUT_OF_PLACE;
  std::vector<std::size_t> forward_strides;
  std::vector<std::size_t> backward_strides;
  std::size_t forward_distance = 1;
  std::size_t backward_distance = 1;
  std::size_t forward_offset = 0;
  std::size_t backward_offset = 0;

  explicit descriptor(const std::vector<std::size_t>& lens)
      : lengths(lens), forward_strides(detail::get_default_strides(lens)), backward_strides(forward_strides) {
    forward_distance = backward_distance = get_flattened_length();
  }

  /// Validate and build the plan (twiddles, workspace, kernels) on the queue's device.
  committed_descriptor<Scalar, Domain> commit(queue& q) { return committed_descriptor<Scalar, Domain>(*this, q); }
  committed_descriptor<Scalar, Domain> commit(queue&& q) { return committed_descriptor<Scalar, Domain>(*this, q); }

  std::size_t get_flattened_length() const noexcept {
    std::size_t n = 1;
    for (std::size_t l : lengths) n *= l;
    return n;
  }
  std::size_t get_input_count(direction dir) const noexcept {
    pfft_desc c = to_c();
    return pfft_get_buffer_count(&c, static_cast<int>(dir));
  }
  std::size_t get_output_count(direction dir) const noexcept { return get_input_count(inv(dir)); }

  const std::vector<std::size_t>& get_strides(direction dir) const noexcept {
    return dir == direction::FORWARD ? forward_strides : backward_strides;
  }
  std::vector<std::size_t>& get_strides(direction dir) noexcept {
    return dir == direction::FORWARD ? forward_strides : backward_strides;
  }
  std::size_t get_distance(direction dir) const noexcept {
    return dir == direction::FORWARD ? forward_distance : backward_distance;
  }
  std::size_t& get_distance(direction dir) noexcept {
    return dir == direction::FORWARD ? forward_distance : backward_distance;
  }
  std::size_t get_offset(direction dir) const noexcept {
    return dir == direction::FORWARD ? forward_offset : backward_offset;
  }
  std::size_t& get_offset(direction dir) noexcept { return dir == direction::FORWARD ? forward_offset : backward_offset; }
  Scalar get_scale(direction dir) const noexcept { return dir == direction::FORWARD ? forward_scale : backward_scale; }
  Scalar& get_scale(direction dir) noexcept { return dir == direction::FORWARD ? forward_scale : backward_scale; }

  /// POD view for the C ABI (pointers alias this object's vectors).
  pfft_desc to_c() const noexcept {
    pfft_desc c{};
    c.precision = std::is_same_v<Scalar, double> ? PFFT_DOUBLE : PFFT_FLOAT;
    c.domain = Domain == domain::COMPLEX ? PFFT_DOMAIN_COMPLEX : PFFT_DOMAIN_REAL;
    c.rank = lengths.size();
    c.lengths = lengths.data();
    c.forward_scale = static_cast<double>(forward_scale);
    c.backward_scale = static_cast<double>(backward_scale);
    c.number_of_transforms = number_of_transforms;
    c.complex_storage = static_cast<int>(complex_storage);
    c.placement = static_cast<int>(placement);
    c.n_forward_strides = forward_strides.size();
    c.forward_strides = forward_strides.data();
    c.n_backward_strides = backward_strides.size();
    c.backward_strides = backward_strides.data();
    c.forward_distance = forward_distance;
    c.backward_distance = backward_distance;
    c.forward_offset = forward_offset;
    c.backward_offset = backward_offset;
    return c;
  }
};

namespace detail {
template <typename Descriptor>
layout get_layout(const Descriptor& desc, direction dir) {
  pfft_desc c = desc.to_c();
  switch (pfft_get_layout(&c, static_cast<int>(dir))) {
    case PFFT_LAYOUT_PACKED: return layout::PACKED;
    case PFFT_LAYOUT_BATCH_INTERLEAVED: return layout::BATCH_INTERLEAVED;
    default: return layout::UNPACKED;
  }
}
}  // namespace detail

}  // namespace portfft
#endif
