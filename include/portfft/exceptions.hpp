// Exception types of the portFFT API (/root/reference/src/portfft/common/exceptions.hpp:32-77), raised from the
// status codes of the C ABI (include/pfft.h).
#ifndef PFFT_B200_PORTFFT_EXCEPTIONS_HPP
#define PFFT_B200_PORTFFT_EXCEPTIONS_HPP

#include <sstream>
#include <stdexcept>
#include <string>

#include "../pfft.h"

namespace portfft {

class base_error : public std::runtime_error {
  template <typename... Ts>
  static std::string join(const Ts&... parts) {
    std::ostringstream os;
    (void)std::initializer_list<int>{((os << parts), 0)...};
    return os.str();
  }

 public:
  template <typename... Ts>
  explicit base_error(const Ts&... parts) : std::runtime_error(join(parts...)) {}
};

struct internal_error : base_error {
  template <typename... Ts>
  explicit internal_error(const Ts&... parts) : base_error(parts...) {}
};
struct invalid_configuration : base_error {
  template <typename... Ts>
  explicit invalid_configuration(const Ts&... parts) : base_error(parts...) {}
};
struct unsupported_configuration : base_error {
  template <typename... Ts>
  explicit unsupported_configuration(const Ts&... parts) : base_error(parts...) {}
};
struct out_of_local_memory_error : unsupported_configuration {
  template <typename... Ts>
  explicit out_of_local_memory_error(const Ts&... parts) : unsupported_configuration(parts...) {}
};
/// CUDA / NCCL runtime failure (no SYCL counterpart: the reference would surface a sycl::exception).
struct device_error : base_error {
  template <typename... Ts>
  explicit device_error(const Ts&... parts) : base_error(parts...) {}
};

namespace detail {
inline void throw_on_status(pfft_status st) {
  if (st == PFFT_OK) return;
  const char* msg = pfft_last_error();
  switch (st) {
    case PFFT_INVALID_CONFIGURATION: throw invalid_configuration(msg);
    case PFFT_UNSUPPORTED_CONFIGURATION: throw unsupported_configuration(msg);
    case PFFT_OUT_OF_LOCAL_MEMORY: throw out_of_local_memory_error(msg);
    case PFFT_INTERNAL_ERROR: throw internal_error(msg);
    default: throw device_error(msg);
  }
}
}  // namespace detail
}  // namespace portfft
#endif
