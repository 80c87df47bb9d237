// See tables.h.
#include "tables.h"

#include <cmath>
#include <complex>
#include <type_traits>

namespace pfft {

// cos(2*pi*p/q), sin(2*pi*p/q) in long double with exact octant reduction
long double cos2pi_ld(long long p, long long q) {
  p %= q;
  if (p < 0) p += q;
  if (2 * p > q) p = q - p;
  if (4 * p > q) return -cos2pi_ld(q - 2 * p, 2 * q);
  if (8 * p > q) return sin2pi_ld(q - 4 * p, 4 * q);
  return cosl(6.283185307179586476925286766559L * (long double)p / (long double)q);
}
long double sin2pi_ld(long long p, long long q) {
  p %= q;
  if (p < 0) p += q;
  if (2 * p > q) return -sin2pi_ld(q - p, q);
  if (4 * p > q) return sin2pi_ld(q - 2 * p, 2 * q);
  if (8 * p > q) return cos2pi_ld(q - 4 * p, 4 * q);
  return sinl(6.283185307179586476925286766559L * (long double)p / (long double)q);
}

template <typename T>
std::vector<T> make_twiddles(long long n, long long count, long long mult) {
  std::vector<T> t((size_t)count * 2);
  for (long long i = 0; i < count; ++i) {
    const long long k = (long long)(((__int128)i * mult) % n);
    t[2 * i] = (T)cos2pi_ld(k, n);
    t[2 * i + 1] = (T)(-sin2pi_ld(k, n));
  }
  return t;
}

namespace {

// in-place iterative radix-2 transform of length m (power of two) in precision R, forward sign
template <typename R>
void host_fft_pow2(std::vector<std::complex<R>>& a) {
  const size_t m = a.size();
  for (size_t i = 1, j = 0; i < m; ++i) {
    size_t bit = m >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  std::vector<std::complex<R>> w(m / 2 ? m / 2 : 1);
  for (size_t k = 0; k < m / 2; ++k) w[k] = {(R)cos2pi_ld((long long)k, (long long)m), (R)(-sin2pi_ld((long long)k, (long long)m))};
  for (size_t len = 2; len <= m; len <<= 1) {
    const size_t half = len / 2, step = m / len;
    for (size_t i = 0; i < m; i += len)
      for (size_t k = 0; k < half; ++k) {
        const std::complex<R> u = a[i + k], v = a[i + k + half] * w[k * step];
        a[i + k] = u + v;
        a[i + k + half] = u - v;
      }
  }
}

// exp(-i*pi*j^2/L) = exp(-2*pi*i * (j^2 mod 2L) / 2L)
inline std::complex<long double> chirp(long long j, long long L) {
  const long long e = (long long)(((__int128)j * j) % (2 * L));
  return {cos2pi_ld(e, 2 * L), -sin2pi_ld(e, 2 * L)};
}

}  // namespace

template <typename T>
std::vector<T> make_mod_table(int kind, long long L, long long M) {
  std::vector<T> t;
  if (kind == MODT_CHIRP || kind == MODT_CHIRP_OVER_M) {
    t.resize((size_t)L * 2);
    const long double s = kind == MODT_CHIRP ? 1.0L : 1.0L / (long double)M;
    for (long long j = 0; j < L; ++j) {
      const std::complex<long double> w = chirp(j, L);
      t[2 * j] = (T)(w.real() * s);
      t[2 * j + 1] = (T)(w.imag() * s);
    }
  } else if (kind == MODT_CONV) {
    // extended precision relative to T: long double for double plans, double for float plans
    using R = typename std::conditional<sizeof(T) == 8, long double, double>::type;
    std::vector<std::complex<R>> b((size_t)M, std::complex<R>(0, 0));
    for (long long m = 0; m < L; ++m) {
      const std::complex<long double> w = std::conj(chirp(m, L));
      b[(size_t)m] = {(R)w.real(), (R)w.imag()};
      if (m > 0) b[(size_t)(M - m)] = b[(size_t)m];
    }
    host_fft_pow2<R>(b);
    t.resize((size_t)M * 2);
    for (long long k = 0; k < M; ++k) {
      t[2 * k] = (T)b[(size_t)k].real();
      t[2 * k + 1] = (T)b[(size_t)k].imag();
    }
  }
  return t;
}

template std::vector<float> make_twiddles<float>(long long, long long, long long);
template std::vector<double> make_twiddles<double>(long long, long long, long long);
template std::vector<float> make_mod_table<float>(int, long long, long long);
template std::vector<double> make_mod_table<double>(int, long long, long long);

}  // namespace pfft
