// Global-memory access helpers shared by the level kernels: interleaved / split storage, backward = (re <-> im)
// swap on interleaved data, and the batch multi-index -> base offset arithmetic of a pass (pass.h).
//
// Counterpart of the reference's views / strided copies (/root/reference/src/portfft/common/memory_views.hpp:44-288,
// transfers.hpp:52-195): there every access goes through offset_view / md_view / strided_view index math; here the
// per-transform base is computed once and the element stride is a single multiply.
#pragma once
#include "device_utils.cuh"
#include "pass.h"

namespace pfft {

struct IoFlags {
  bool il;    // interleaved storage
  bool swap;  // backward direction on interleaved storage (split storage swaps the pointers on the host)
};

template <typename T>
__device__ __forceinline__ cx<T> gload(const PassParams& p, IoFlags fl, long long idx) {
  cx<T> v;
  if (fl.il) {
    v = reinterpret_cast<const cx<T>*>(p.in_re)[idx];
    if (fl.swap) {
      T t = v.x;
      v.x = v.y;
      v.y = t;
    }
  } else {
    v.x = reinterpret_cast<const T*>(p.in_re)[idx];
    v.y = reinterpret_cast<const T*>(p.in_im)[idx];
  }
  return v;
}

template <typename T>
__device__ __forceinline__ void gstore(const PassParams& p, IoFlags fl, long long idx, cx<T> v) {
  if (fl.il) {
    if (fl.swap) {
      T t = v.x;
      v.x = v.y;
      v.y = t;
    }
    reinterpret_cast<cx<T>*>(p.out_re)[idx] = v;
  } else {
    reinterpret_cast<T*>(p.out_re)[idx] = v.x;
    reinterpret_cast<T*>(p.out_im)[idx] = v.y;
  }
}

__device__ __forceinline__ bool single_batch_dim(const PassParams& p) {
  return p.nb[1] == 1 && p.nb[2] == 1 && p.nb[3] == 1;
}

// input / output base offsets (complex elements) of batch entry g
__device__ __forceinline__ void batch_bases(const PassParams& p, bool one_dim, long long g, long long& ib,
                                            long long& ob) {
  if (one_dim) {
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

}  // namespace pfft
