// REAL domain (real-to-complex forward, complex-to-real backward), the pre- / post-processing passes around the
// complex transform of half the length.
//
// The reference reserves this API and throws (`compute_forward(const Scalar*, complex_type*)`:
// /root/reference/src/portfft/committed_descriptor.hpp:134-137,201-206,273-278; validate_descriptor:
// /root/reference/src/portfft/descriptor_validation.hpp:268-270); its test generator already defines the expected
// result as numpy.fft.rfftn (/root/reference/test/common/reference_data_wrangler.hpp:136-137,196).  Scheme (N even,
// H = N/2, w = exp(-2 pi i / N)):
//   forward : z_j = x_{2j} + i x_{2j+1}, Z = DFT_H(z),  X_k = E_k + w^k O_k,  k = 0..H,
//             E_k = (Z_k + conj Z_{H-k}) / 2,  O_k = (Z_k - conj Z_{H-k}) / (2i)                      (r2c_post)
//   backward: Z'_k = (X_k + conj X_{H-k}) + i conj(w^k) (X_k - conj X_{H-k}),  k = 0..H-1, stored in REVERSED index
//             order so that the plain forward DFT_H of the scratch row is the unnormalised inverse,
//             x_{2j} = Re z_j, x_{2j+1} = Im z_j                                                         (c2r_pre)
// When the real buffer has unit stride and even offsets the complex passes read / write it directly as interleaved
// pairs (no pack / unpack pass).  Odd N: the real row is widened to a complex row (real_pack), transformed at full
// length and truncated to N/2 + 1 outputs (r2c_post, variant 1); backward mirrors it (c2r_pre / real_unpack variant 1).
// Plan-internal rows are packed interleaved complex; the user's complex side follows the descriptor's storage.
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "pass.h"

namespace pfft {

namespace {

// batch multi-index of row g -> input / output base offsets
__device__ __forceinline__ void row_bases(const PassParams& p, long long g, long long& ib, long long& ob) {
  if (single_batch_dim(p)) {  // the common case costs no 64-bit division
    ib = p.ioff + g * p.ibd[0];
    ob = p.ooff + g * p.obd[0];
    return;
  }
  ib = p.ioff;
  ob = p.ooff;
#pragma unroll
  for (int d = 0; d < kMaxBatchDims; ++d) {
    const long long q = g / p.nb[d];
    const long long b = g - q * p.nb[d];
    g = q;
    ib += b * p.ibd[d];
    ob += b * p.obd[d];
  }
}

// Work distribution shared by the four kernels: a CTA takes a chunk of consecutive rows (about 2048 elements) and
// spreads the chunk's (row, element) pairs over its 256 threads, lanes running along the element index -- full lanes
// for every row length (17 elements of an N = 32 row as well as 257 of an N = 512 row), 32-bit index arithmetic, and
// no 64-bit division when the pass has a single batch dimension (row_bases).
__device__ __forceinline__ int chunk_rows(int count) { return count >= 2048 ? 1 : 2048 / count; }
#define PFFT_FOR_ROW_ELEMS(p, count, g, k)                                                                          \
  for (long long g0_ = (long long)blockIdx.x * chunk_rows(count); g0_ < p.batch_total;                                \
       g0_ += (long long)gridDim.x * chunk_rows(count))                                                               \
    for (unsigned idx_ = threadIdx.x,                                                                                 \
                  items_ = (unsigned)min((long long)chunk_rows(count), p.batch_total - g0_) * (unsigned)(count),      \
                  row_ = 0, k = 0;                                                                                    \
         idx_ < items_ && (row_ = idx_ / (unsigned)(count), k = idx_ - row_ * (unsigned)(count), true); idx_ += 256)  \
      for (long long g = g0_ + row_, once_ = 1; once_; once_ = 0)

// variant 0: out[m] = (x[2m], x[2m+1]), m < n/2;  variant 1: out[m] = (x[m], 0), m < n.  Input: REAL scalars.
template <typename T>
__global__ void __launch_bounds__(256) real_pack_kernel(const PassParams p, const int variant) {
  const int count = variant == 0 ? p.n / 2 : p.n;
  const T* x = reinterpret_cast<const T*>(p.in_re);
  cx<T>* out = reinterpret_cast<cx<T>*>(p.out_re);
  PFFT_FOR_ROW_ELEMS(p, count, g, m) {
    long long ib, ob;
    row_bases(p, g, ib, ob);
    {
      cx<T> v;
      if (variant == 0) {
        v.x = x[ib + (2LL * m) * p.is];
        v.y = x[ib + (2LL * m + 1) * p.is];
      } else {
        v.x = x[ib + (long long)m * p.is];
        v.y = T(0);
      }
      out[ob + m] = v;
    }
  }
}

// variant 0: out[2m] = Re z[m], out[2m+1] = Im z[m], m < n/2;  variant 1: out[m] = Re y[m], m < n.  Output: REAL scalars.
template <typename T>
__global__ void __launch_bounds__(256) real_unpack_kernel(const PassParams p, const int variant) {
  const int count = variant == 0 ? p.n / 2 : p.n;
  const cx<T>* in = reinterpret_cast<const cx<T>*>(p.in_re);
  T* x = reinterpret_cast<T*>(p.out_re);
  const T scale = p.apply_scale ? T(p.scale) : T(1);
  PFFT_FOR_ROW_ELEMS(p, count, g, m) {
    long long ib, ob;
    row_bases(p, g, ib, ob);
    {
      const cx<T> v = in[ib + m];
      if (variant == 0) {
        x[ob + (2LL * m) * p.os] = v.x * scale;
        x[ob + (2LL * m + 1) * p.os] = v.y * scale;
      } else {
        x[ob + (long long)m * p.os] = v.x * scale;
      }
    }
  }
}

// Input: packed interleaved rows (variant 0: Z of length n/2; variant 1: the full-length transform Y).  Output: the
// half spectrum X_k, k = 0..n/2, in the user's layout (element stride os, descriptor's storage).
// Variant 0 works on PAIRS (k, H - k), k = 0..H/2: with t = w^k O_k,  X_k = E_k + t  and  X_{H-k} = conj(E_k - t)
// (E_{H-k} = conj E_k, O_{H-k} = conj O_k, w^{H-k} = -conj w^k), so that every input element is read once and one
// twiddle serves two outputs (the element-wise form read each Z twice and ran at half the HBM rate on long rows).
template <typename T>
__global__ void __launch_bounds__(256) r2c_post_kernel(const PassParams p, const int variant, const bool il) {
  const int h = p.n / 2, count = variant == 0 ? h / 2 + 1 : h + 1;
  const cx<T>* in = reinterpret_cast<const cx<T>*>(p.in_re);
  const IoFlags fl{il, false};
  const T scale = p.apply_scale ? T(p.scale) : T(1);
  PFFT_FOR_ROW_ELEMS(p, count, g, ku) {
    const int k = (int)ku;
    long long ib, ob;
    row_bases(p, g, ib, ob);
    if (variant == 0) {
      const int m = h - k;  // partner (k = 0: Z_H = Z_0)
      const cx<T> a = in[ib + k];
      cx<T> b = in[ib + (k == 0 ? 0 : m)];
      b.y = -b.y;  // conj Z_{H-k}
      const cx<T> ev{(a.x + b.x) * T(0.5), (a.y + b.y) * T(0.5)};
      const cx<T> od{(a.y - b.y) * T(0.5), -(a.x - b.x) * T(0.5)};  // (a - b) / (2i)
      const cx<T> t = cmul(ldg_cx<T>(p.tw, k), od);
      gstore<T>(p, fl, ob + (long long)k * p.os, cscale(ev + t, scale));
      if (m != k) gstore<T>(p, fl, ob + (long long)m * p.os, cscale(cx<T>{ev.x - t.x, -(ev.y - t.y)}, scale));
    } else {
      gstore<T>(p, fl, ob + (long long)k * p.os, cscale(in[ib + k], scale));
    }
  }
}

// Input: the half spectrum in the user's layout (element stride is, descriptor's storage).  Output: packed
// interleaved rows, index-reversed (see the header): variant 0 length n/2, variant 1 the Hermitian extension, length n.
// Variant 0 works on pairs as above: with s = X_k + conj X_{H-k}, t = conj(w^k) (X_k - conj X_{H-k}):
// Z'_k = s + i t,  Z'_{H-k} = conj(s) + i conj(t).
template <typename T>
__global__ void __launch_bounds__(256) c2r_pre_kernel(const PassParams p, const int variant, const bool il) {
  const int n = p.n, h = n / 2;
  const int count = variant == 0 ? h / 2 + 1 : (n + 1) / 2;
  cx<T>* out = reinterpret_cast<cx<T>*>(p.out_re);
  const IoFlags fl{il, false};
  PFFT_FOR_ROW_ELEMS(p, count, g, ku) {
    const int k = (int)ku;
    long long ib, ob;
    row_bases(p, g, ib, ob);
    {
      cx<T> a = gload<T>(p, fl, ib + (long long)k * p.is);
      if (k == 0) a.y = T(0);  // the imaginary parts of X_0 (and X_{N/2}) do not enter a real inverse (numpy.fft.irfft)
      if (variant == 0) {
        const int m = h - k;
        cx<T> b = gload<T>(p, fl, ib + (long long)m * p.is);
        if (k == 0) b.y = T(0);
        b.y = -b.y;  // conj X_{H-k}
        const cx<T> s = a + b, d = a - b;
        cx<T> w = ldg_cx<T>(p.tw, k);
        w.y = -w.y;  // conj(w^k)
        const cx<T> t = cmul(w, d);
        out[ob + (k == 0 ? 0 : m)] = cx<T>{s.x - t.y, s.y + t.x};        // Z'_k = s + i t, stored at (H - k) mod H
        if (k != 0 && m != k) out[ob + k] = cx<T>{s.x + t.y, t.x - s.y};  // Z'_{H-k} = conj(s) + i conj(t), stored at k
      } else {
        // reversed Hermitian extension: row[k] = conj X_k, row[n - k] = X_k
        out[ob + k] = cx<T>{a.x, -a.y};
        if (k > 0) out[ob + n - k] = a;
      }
    }
  }
}
#undef PFFT_FOR_ROW_ELEMS

template <typename K, typename... Args>
cudaError_t launch_rows(K kern, int grid, cudaStream_t stream, Args... args) {
  kern<<<grid, 256, 0, stream>>>(args...);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_real_pack(const PassParams& p, bool is_double, int variant, int grid, cudaStream_t stream) {
  return is_double ? launch_rows(real_pack_kernel<double>, grid, stream, p, variant)
                   : launch_rows(real_pack_kernel<float>, grid, stream, p, variant);
}
cudaError_t launch_real_unpack(const PassParams& p, bool is_double, int variant, int grid, cudaStream_t stream) {
  return is_double ? launch_rows(real_unpack_kernel<double>, grid, stream, p, variant)
                   : launch_rows(real_unpack_kernel<float>, grid, stream, p, variant);
}
cudaError_t launch_r2c_post(const PassParams& p, bool is_double, bool il, int variant, int grid, cudaStream_t stream) {
  return is_double ? launch_rows(r2c_post_kernel<double>, grid, stream, p, variant, il)
                   : launch_rows(r2c_post_kernel<float>, grid, stream, p, variant, il);
}
cudaError_t launch_c2r_pre(const PassParams& p, bool is_double, bool il, int variant, int grid, cudaStream_t stream) {
  return is_double ? launch_rows(c2r_pre_kernel<double>, grid, stream, p, variant, il)
                   : launch_rows(c2r_pre_kernel<float>, grid, stream, p, variant, il);
}

}  // namespace pfft
