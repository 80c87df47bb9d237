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

// `peer` < 0: the pass's own output buffer; otherwise entry `peer` of the peer table (memory of another GPU mapped
// into this address space: the store itself is the NVLink transfer)
template <typename T>
__device__ __forceinline__ void gstore(const PassParams& p, IoFlags fl, long long idx, cx<T> v, int peer = -1) {
  void* re = p.out_re;
  void* im = p.out_im;
  if (peer >= 0) {
    re = p.out_tab_re[peer];
    im = p.out_tab_im[peer];
  }
  if (fl.il) {
    if (fl.swap) {
      T t = v.x;
      v.x = v.y;
      v.y = t;
    }
    reinterpret_cast<cx<T>*>(re)[idx] = v;
  } else {
    reinterpret_cast<T*>(re)[idx] = v.x;
    reinterpret_cast<T*>(im)[idx] = v.y;
  }
}

__device__ __forceinline__ bool single_batch_dim(const PassParams& p) {
  return p.nb[1] == 1 && p.nb[2] == 1 && p.nb[3] == 1;
}

// input / output base offsets (complex elements) of batch entry g; `peer` = index along p.peer_dim (or -1)
__device__ __forceinline__ void batch_bases(const PassParams& p, bool one_dim, long long g, long long& ib,
                                            long long& ob, int& peer) {
  peer = -1;
  if (one_dim) {
    ib = p.ioff + g * p.ibd[0];
    ob = p.ooff + g * p.obd[0];
    if (p.peer_dim == 0) peer = (int)g;
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
    if (d == p.peer_dim) peer = (int)b;
  }
}

}  // namespace pfft
