// REAL-domain pre / post-processing on data a transform kernel holds in registers / shared memory (the formulas of
// real.cu), shared by the kernels that fuse it: wg_cube.cu (tile kernels), wi_tma.cu (thread-level kernel).
#pragma once
#include <type_traits>

#include "device_utils.cuh"

namespace pfft {

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// X_k = E_k + w^k O_k from a = Z_k, b = Z_{N-k}, w = w_{2N}^k
template <typename T>
__device__ __forceinline__ cx<T> r2c_combine(cx<T> a, cx<T> b, cx<T> w) {
  b.y = -b.y;
  const cx<T> ev{(a.x + b.x) * T(0.5), (a.y + b.y) * T(0.5)};
  const cx<T> od{(a.y - b.y) * T(0.5), -(a.x - b.x) * T(0.5)};  // (a - b) / (2i)
  return ev + cmul(w, od);
}

// z'_j (the input of the plain forward transform that yields the unnormalised inverse) from a = X_{N-j}, b = X_j,
// w = w_{2N}^j:  (a + conj b) - i w (a - conj b); j = 0: a = Re X_0, b = Re X_N, (a + b) + i (a - b)
template <typename T>
__device__ __forceinline__ cx<T> c2r_combine(cx<T> a, cx<T> b, cx<T> w, bool first) {
  if (first) return cx<T>{a.x + b.x, a.x - b.x};
  b.y = -b.y;
  const cx<T> s = a + b, d = a - b;
  const cx<T> t = cmul(w, d);
  return cx<T>{s.x + t.y, s.y - t.x};  // s - i t
}

// w_Q^P as a compile-time constant
template <int P, int Q, typename T>
__device__ __forceinline__ constexpr cx<T> const_w() {
  return cx<T>{T(ct::cos2pi(P, Q)), T(-ct::sin2pi(P, Q))};
}

}  // namespace pfft
