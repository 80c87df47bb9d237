// Element-wise pass: out[b] = modifiers(in[b]) over the batch multi-index of a PassParams with n == 1.
//
// Used by the Bluestein path for lengths whose padded convolution length M does not fit one CTA (plan.cpp): chirp
// multiply + zero padding (user layout -> packed scratch), the pointwise product with the transformed chirp between
// the two length-M transforms, and the final chirp multiply + truncation (scratch -> user layout).  The modifier index
// is the index along batch dimension 0 (the element index of the user's transform); lanes run along that dimension,
// so the packed side is coalesced and the user side is as coalesced as its element stride allows.
// The reference has no counterpart: it rejects lengths with large prime factors
// (/root/reference/src/portfft/committed_descriptor_impl.hpp:241, utils.hpp:102,126).
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "pass.h"

namespace pfft {

template <typename T>
__global__ void __launch_bounds__(256) ew_kernel(const PassParams p, const bool il_in, const bool il_out, const bool swap) {
  const IoFlags fl{il_in, swap && !(p.mod_flags & MOD_NO_USER_SWAP_IN)};
  const IoFlags flo{il_out, swap && !(p.mod_flags & MOD_NO_USER_SWAP_OUT)};
  const long long n0 = p.nb[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  // (row q, element j) of the flat index advance incrementally: one 64-bit division per thread, not per element
  const long long sq = stride / n0, sr = stride - sq * n0;
  const bool one_outer = p.nb[2] == 1 && p.nb[3] == 1;
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long q = g / n0, j = g - q * n0;
  for (; g < p.batch_total; g += stride, j += sr, q += sq) {
    if (j >= n0) {
      j -= n0;
      ++q;
    }
    long long ib = p.ioff + j * p.ibd[0], ob = p.ooff + j * p.obd[0];
    if (one_outer) {
      ib += q * p.ibd[1];
      ob += q * p.obd[1];
    } else {
      long long r = q;
#pragma unroll
      for (int d = 1; d < kMaxBatchDims; ++d) {
        const long long q2 = r / p.nb[d];
        const long long b = r - q2 * p.nb[d];
        r = q2;
        ib += b * p.ibd[d];
        ob += b * p.obd[d];
      }
    }
    if (p.valid_out > 0 && j >= p.valid_out) continue;
    cx<T> v{T(0), T(0)};
    if (p.valid_in == 0 || j < p.valid_in) {
      v = gload<T>(p, fl, ib);
      if (p.lmod != nullptr) v = cmul(v, ldg_cx<T>(p.lmod, j));
    }
    if (p.mod_flags & MOD_SWAP_PRE) v = cx<T>{v.y, v.x};
    if (p.smod != nullptr) v = cmul(v, ldg_cx<T>(p.smod, j));
    if (p.mod_flags & MOD_SWAP_POST) v = cx<T>{v.y, v.x};
    if (p.apply_scale) v = cscale(v, T(p.scale));
    gstore<T>(p, flo, ob, v);
  }
}

cudaError_t launch_ew(const PassParams& p, bool is_double, bool il_in, bool il_out, bool swap, int grid,
                      cudaStream_t stream) {
  if (is_double)
    ew_kernel<double><<<grid, 256, 0, stream>>>(p, il_in, il_out, swap);
  else
    ew_kernel<float><<<grid, 256, 0, stream>>>(p, il_in, il_out, swap);
  return cudaGetLastError();
}

}  // namespace pfft
