// Register-resident DFT butterflies for one CUDA thread (the WORKITEM building block of every level).
//
// Replaces the reference's run-time recursive `wi_dft` / `cooley_tukey_dft` / `naive_dft`
// (/root/reference/src/portfft/common/workitem.hpp:64-127,200-219) and its generated 65x65 twiddle table
// (/root/reference/src/portfft/common/twiddle.hpp, scripts/generate_twiddles.py:60-92).  Here every size is a
// compile-time template: indices are static so the data never leaves registers, twiddles are constexpr literals
// (exact 0/+-1 at quarter turns, as generate_twiddles.py:67-79 does) and trivial rotations cost no multiplies.
#pragma once
#include <cuda_runtime.h>

namespace pfft {

template <typename T>
struct alignas(2 * sizeof(T)) cx {
  T x, y;
};

template <typename T>
__host__ __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__host__ __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T>
__host__ __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__host__ __device__ __forceinline__ cx<T> cscale(cx<T> a, T s) { return {a.x * s, a.y * s}; }
// multiply by -i / +i
template <typename T>
__host__ __device__ __forceinline__ cx<T> mul_mi(cx<T> a) { return {a.y, -a.x}; }
template <typename T>
__host__ __device__ __forceinline__ cx<T> mul_pi(cx<T> a) { return {-a.y, a.x}; }

// ---------------------------------------------------------------------------------------------------------------
// constexpr cos / sin of 2*pi*p/q with exact symmetry reduction (argument of the series <= pi/4).
// ---------------------------------------------------------------------------------------------------------------
namespace ct {
constexpr double kTwoPi = 6.283185307179586476925286766559;

__host__ __device__ constexpr double series_cos(double x) {
  // Horner on x^2, 13 terms: |x| <= pi/4 -> truncation < 1e-22
  double x2 = x * x;
  double r = 1.0;
  for (int k = 24; k >= 2; k -= 2) r = 1.0 - r * x2 / double(k * (k - 1));
  return r;
}
__host__ __device__ constexpr double series_sin(double x) {
  double x2 = x * x;
  double r = 1.0;
  for (int k = 25; k >= 3; k -= 2) r = 1.0 - r * x2 / double(k * (k - 1));
  return r * x;
}
__host__ __device__ constexpr double sin2pi(long long p, long long q);
__host__ __device__ constexpr double cos2pi(long long p, long long q) {
  p %= q;
  if (p < 0) p += q;
  if (2 * p > q) p = q - p;
  if (4 * p > q) return -cos2pi(q - 2 * p, 2 * q);
  if (8 * p > q) return sin2pi(q - 4 * p, 4 * q);
  return series_cos(kTwoPi * double(p) / double(q));
}
__host__ __device__ constexpr double sin2pi(long long p, long long q) {
  p %= q;
  if (p < 0) p += q;
  if (2 * p > q) return -sin2pi(q - p, q);
  if (4 * p > q) return sin2pi(q - 2 * p, 2 * q);
  if (8 * p > q) return cos2pi(q - 4 * p, 4 * q);
  return series_sin(kTwoPi * double(p) / double(q));
}
__host__ __device__ constexpr int smallest_prime_factor(int n) {
  for (int i = 2; i * i <= n; ++i)
    if (n % i == 0) return i;
  return n;
}
__host__ __device__ constexpr bool is_prime(int n) { return n >= 2 && smallest_prime_factor(n) == n; }
// first factor A of the in-register Cooley-Tukey split N = A * B (A-point DFTs run last)
__host__ __device__ constexpr int split_factor(int n) {
  if (n % 4 == 0 && n > 4) return 4;
  if (n % 2 == 0) return 2;
  return smallest_prime_factor(n);
}
}  // namespace ct

// a * w_Q^P  (w_Q = exp(-2*pi*i/Q)), compile-time P, Q; trivial rotations are free.
template <int P, int Q, typename T>
__host__ __device__ __forceinline__ cx<T> mul_w(cx<T> a) {
  constexpr int p = ((P % Q) + Q) % Q;
  if constexpr (p == 0) {
    return a;
  } else if constexpr (4 * p == Q) {
    return mul_mi(a);
  } else if constexpr (2 * p == Q) {
    return {-a.x, -a.y};
  } else if constexpr (4 * p == 3 * Q) {
    return mul_pi(a);
  } else if constexpr ((8 * p) % Q == 0) {
    constexpr T h = T(0.70710678118654752440084436210485);
    constexpr int o = (8 * p) / Q;  // odd: 1, 3, 5, 7
    if constexpr (o == 1) return {(a.x + a.y) * h, (a.y - a.x) * h};
    if constexpr (o == 3) return {(a.y - a.x) * h, -(a.x + a.y) * h};
    if constexpr (o == 5) return {-(a.x + a.y) * h, (a.x - a.y) * h};
    return {(a.x - a.y) * h, (a.x + a.y) * h};
  } else {
    constexpr T c = T(ct::cos2pi(p, Q));
    constexpr T s = T(-ct::sin2pi(p, Q));
    return {a.x * c - a.y * s, a.x * s + a.y * c};
  }
}

template <int N, typename T>
struct DFT;

// Odd-prime DFT using the conjugate-pair symmetry: (P-1)^2/2 real FMAs per component instead of P^2 complex MACs.
template <int P, typename T>
struct DFTPrime {
  static constexpr int H = (P - 1) / 2;
  template <int K, int J>
  static __host__ __device__ __forceinline__ void acc(const cx<T>* a, const cx<T>* b, cx<T>& A, cx<T>& B) {
    if constexpr (J <= H) {
      constexpr T c = T(ct::cos2pi((long long)J * K, P));
      constexpr T s = T(ct::sin2pi((long long)J * K, P));
      A.x += a[J - 1].x * c;
      A.y += a[J - 1].y * c;
      B.x += b[J - 1].x * s;
      B.y += b[J - 1].y * s;
      acc<K, J + 1>(a, b, A, B);
    }
  }
  template <int K>
  static __host__ __device__ __forceinline__ void outk(const cx<T>& x0, const cx<T>* a, const cx<T>* b, cx<T>* v) {
    if constexpr (K <= H) {
      cx<T> A = x0, B = {T(0), T(0)};
      acc<K, 1>(a, b, A, B);
      // X[K] = A - i*B ; X[P-K] = A + i*B
      v[K] = {A.x + B.y, A.y - B.x};
      v[P - K] = {A.x - B.y, A.y + B.x};
      outk<K + 1>(x0, a, b, v);
    }
  }
  static __host__ __device__ __forceinline__ void run(cx<T>* v) {
    cx<T> a[H], b[H];
    cx<T> x0 = v[0];
    cx<T> sum = v[0];
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      a[j - 1] = v[j] + v[P - j];
      b[j - 1] = v[j] - v[P - j];
      sum = sum + a[j - 1];
    }
    outk<1>(x0, a, b, v);
    v[0] = sum;
  }
};

template <typename T>
struct DFT<1, T> {
  static __host__ __device__ __forceinline__ void run(cx<T>*) {}
};
template <typename T>
struct DFT<2, T> {
  static __host__ __device__ __forceinline__ void run(cx<T>* v) {
    cx<T> a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
  }
};
template <typename T>
struct DFT<4, T> {
  static __host__ __device__ __forceinline__ void run(cx<T>* v) {
    cx<T> s02 = v[0] + v[2], d02 = v[0] - v[2];
    cx<T> s13 = v[1] + v[3], d13 = mul_mi(v[1] - v[3]);
    v[0] = s02 + s13;
    v[1] = d02 + d13;
    v[2] = s02 - s13;
    v[3] = d02 - d13;
  }
};

// Generic size: primes use DFTPrime, composites an in-register Cooley-Tukey split N = A*B with constexpr twiddles:
//   X[kb + B*ka] = sum_ja w_A^{ja ka} [ w_N^{ja kb} sum_jb x[ja + A*jb] w_B^{jb kb} ]
template <int N, typename T>
struct DFT {
  static constexpr int A = ct::split_factor(N);
  static constexpr int B = N / A;
  template <int JA, int KB>
  static __host__ __device__ __forceinline__ void twiddle_row(cx<T>* t) {
    if constexpr (KB < B) {
      t[JA * B + KB] = mul_w<JA * KB, N>(t[JA * B + KB]);
      twiddle_row<JA, KB + 1>(t);
    }
  }
  template <int JA>
  static __host__ __device__ __forceinline__ void step1(const cx<T>* v, cx<T>* t) {
    if constexpr (JA < A) {
#pragma unroll
      for (int jb = 0; jb < B; ++jb) t[JA * B + jb] = v[JA + A * jb];
      DFT<B, T>::run(t + JA * B);
      twiddle_row<JA, 0>(t);
      step1<JA + 1>(v, t);
    }
  }
  static __host__ __device__ __forceinline__ void run(cx<T>* v) {
    if constexpr (ct::is_prime(N)) {
      DFTPrime<N, T>::run(v);
    } else {
      cx<T> t[N];
      step1<0>(v, t);
#pragma unroll
      for (int kb = 0; kb < B; ++kb) {
        cx<T> u[A];
#pragma unroll
        for (int ja = 0; ja < A; ++ja) u[ja] = t[ja * B + kb];
        DFT<A, T>::run(u);
#pragma unroll
        for (int ka = 0; ka < A; ++ka) v[kb + B * ka] = u[ka];
      }
    }
  }
};

}  // namespace pfft
