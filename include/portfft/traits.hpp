// Type traits of the portFFT API (/root/reference/src/portfft/traits.hpp:31-51).
#ifndef PFFT_B200_PORTFFT_TRAITS_HPP
#define PFFT_B200_PORTFFT_TRAITS_HPP

#include <complex>

#include "enums.hpp"

namespace portfft {

template <typename T>
struct get_real {
  using type = T;
};
template <typename T>
struct get_real<std::complex<T>> {
  using type = T;
};
template <typename T>
using get_real_t = typename get_real<T>::type;

template <typename T>
struct get_domain {
  static constexpr domain value = domain::REAL;
};
template <typename T>
struct get_domain<std::complex<T>> {
  static constexpr domain value = domain::COMPLEX;
};
template <typename T>
inline constexpr domain get_domain_v = get_domain<T>::value;

}  // namespace portfft
#endif
