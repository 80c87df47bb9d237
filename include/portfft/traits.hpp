// Type traits of the portFFT API: `get_real<T>::type` and `get_domain<T>::value`, the names user code and the tests
// of the reference use (/root/reference/src/portfft/traits.hpp:31-51).  Both are views of one classification of T.
#ifndef PFFT_B200_PORTFFT_TRAITS_HPP
#define PFFT_B200_PORTFFT_TRAITS_HPP

#include <complex>

#include "enums.hpp"

namespace portfft {

namespace detail {
// scalar type and complexness of an element type: T itself, or the value type of std::complex<T>
template <typename Element>
struct element_traits {
  using scalar = Element;
  static constexpr bool is_complex = false;
};
template <typename Scalar>
struct element_traits<std::complex<Scalar>> {
  using scalar = Scalar;
  static constexpr bool is_complex = true;
};
}  // namespace detail

/// get_real<float>::type == get_real<std::complex<float>>::type == float
template <typename T>
struct get_real {
  using type = typename detail::element_traits<T>::scalar;
};
template <typename T>
using get_real_t = typename get_real<T>::type;

/// get_domain<float>::value == domain::REAL, get_domain<std::complex<float>>::value == domain::COMPLEX
template <typename T>
struct get_domain {
  static constexpr domain value = detail::element_traits<T>::is_complex ? domain::COMPLEX : domain::REAL;
};
template <typename T>
inline constexpr domain get_domain_v = get_domain<T>::value;

}  // namespace portfft
#endif
