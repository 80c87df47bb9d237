// Host-side generation of the device-resident tables of a plan (twiddles, inter-factor twiddles, Bluestein chirps).
// Pure C++ (no CUDA) so that tests can read the same tables on a machine without a GPU (pfft_debug_mod_table).
//
// Replaces /root/reference/scripts/generate_twiddles.py:60-92 (generated 65x65 table) and the commit-time device
// twiddle kernels (/root/reference/src/portfft/dispatcher/subgroup_dispatcher.hpp:666-693,
// workgroup_dispatcher.hpp:382-443, global_dispatcher.hpp:107-256): every entry is evaluated in long double with an
// exact octant reduction and rounded once to the plan's scalar type.
#pragma once
#include <vector>

namespace pfft {

enum ModTable : int {
  MODT_NONE = 0,
  MODT_CHIRP = 1,         // w[j] = exp(-i*pi*j^2/L), j in [0, L)
  MODT_CHIRP_OVER_M = 2,  // w[k] / M, k in [0, L)
  MODT_CONV = 3           // FFT_M(b), b[m mod M] = conj(w[|m|]) for |m| < L, zero elsewhere; M entries
};

long double cos2pi_ld(long long p, long long q);
long double sin2pi_ld(long long p, long long q);

// table[i] = w_n^{i * mult} = exp(-2*pi*i * i*mult / n), i in [0, count); interleaved (re, im)
template <typename T>
std::vector<T> make_twiddles(long long n, long long count, long long mult);

// Bluestein tables for transform length L and convolution length M (power of two >= 2L - 1); interleaved (re, im)
template <typename T>
std::vector<T> make_mod_table(int kind, long long L, long long M);

}  // namespace pfft
