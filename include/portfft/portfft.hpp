// Umbrella header (same name as /root/reference/include/portfft/portfft.hpp): B200-native implementation of the
// portFFT descriptor -> commit -> compute API. Link with libpfft_b200.so and the CUDA runtime.
#ifndef PFFT_B200_PORTFFT_HPP
#define PFFT_B200_PORTFFT_HPP

#include "committed_descriptor.hpp"
#include "descriptor.hpp"
#include "enums.hpp"
#include "exceptions.hpp"
#include "queue.hpp"
#include "traits.hpp"

#endif
