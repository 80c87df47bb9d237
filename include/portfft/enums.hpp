// Public enums of the portFFT API (names and enumerator order as /root/reference/src/portfft/enums.hpp:26-69).
#ifndef PFFT_B200_PORTFFT_ENUMS_HPP
#define PFFT_B200_PORTFFT_ENUMS_HPP

#include <string>

namespace portfft {

enum class domain { REAL, COMPLEX };
enum class complex_storage { INTERLEAVED_COMPLEX, SPLIT_COMPLEX };
enum class placement { IN_PLACE, OUT_OF_PLACE };
enum class direction { FORWARD, BACKWARD };

/// The opposite direction.
inline direction inv(direction dir) { return dir == direction::FORWARD ? direction::BACKWARD : direction::FORWARD; }

namespace detail {
/// Which level of the hierarchy computes one dimension (thread / warp / block / multi-kernel).
enum class level { WORKITEM, SUBGROUP, WORKGROUP, GLOBAL };
/// Classification of a (strides, distance) pair.
enum class layout { PACKED, UNPACKED, BATCH_INTERLEAVED };

inline std::string layout_to_string(layout l) {
  switch (l) {
    case layout::PACKED: return "PACKED";
    case layout::UNPACKED: return "UNPACKED";
    case layout::BATCH_INTERLEAVED: return "BATCH_INTERLEAVED";
  }
  return "UNKNOWN";
}
}  // namespace detail
}  // namespace portfft
#endif
