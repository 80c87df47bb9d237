// WORKGROUP level, generic: F transforms per CTA, Stockham autosort passes through padded shared memory.
//
// This is the run-time-configurable block-level kernel (any length whose prime factors are <= 31 and that fits
// shared memory, any stride/distance/offset, interleaved or split storage, forward or backward, optional
// inter-factor twiddle on store).  It takes the place of the reference's `workgroup_impl` / `wg_dft` /
// `dimension_dft` (/root/reference/src/portfft/dispatcher/workgroup_dispatcher.hpp:94-281,
// /root/reference/src/portfft/common/workgroup.hpp:85-346), which runs a fixed two-factor Bailey split with two
// sub-groups per transform, scalar local-memory traffic and twiddles read from global memory.  Design here:
//   * radix passes up to 16 points per thread held in registers (dft.cuh), one block barrier per pass (ping-pong);
//   * first pass reads global memory directly and last pass writes it directly when the element stride is 1
//     (lanes -> consecutive complex elements), otherwise the tile is staged with lanes along the batch index;
//   * padding of one complex per 16 (fp32) / 8 (fp64) keeps the strided Stockham writes bank-conflict free;
//   * backward = swap(re, im) on load and store (no conjugation pass), scale fused on store.
// Hot sizes and layouts have specialised kernels (wg_cube.cu, wg_col.cu, wg_colg.cu, wg_r3.cu); this one guarantees
// coverage and carries the element-wise modifiers of the Bluestein passes.
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "launch_utils.h"
#include "pass.h"

namespace pfft {

template <typename T>
__device__ __forceinline__ int padidx(int i) {
  return i + (i >> (sizeof(T) == 4 ? 4 : 3));
}

// multiply by the inter-factor twiddle w_{gtw_n}^{c*k} and the scale (both optional), then store.
template <typename T>
__device__ __forceinline__ cx<T> finalize(const PassParams& p, cx<T> v, long long c, int k) {
  if (p.gtw_dim >= 0) {
    long long m = c * (long long)k;
    const cx<T> w = cmul(ldg_cx<T>(p.gtw_hi, m >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, m & ((1LL << p.gtw_bits) - 1)));
    v = cmul(v, w);
  }
  if (p.mod_flags & MOD_SWAP_PRE) v = cx<T>{v.y, v.x};
  if (p.smod != nullptr) v = cmul(v, ldg_cx<T>(p.smod, k));
  if (p.mod_flags & MOD_SWAP_POST) v = cx<T>{v.y, v.x};
  if (p.apply_scale) v = cscale(v, T(p.scale));
  return v;
}

// element j of the transform whose first element lives at `base`: zero beyond valid_in, times lmod[j] (both optional)
template <typename T>
__device__ __forceinline__ cx<T> load_in(const PassParams& p, IoFlags fl, long long base, int j) {
  if (p.valid_in > 0 && j >= p.valid_in) return cx<T>{T(0), T(0)};
  cx<T> v = gload<T>(p, fl, base + (long long)j * p.is);
  if (p.lmod != nullptr) v = cmul(v, ldg_cx<T>(p.lmod, j));
  return v;
}

template <typename T>
__device__ __forceinline__ void store_out(const PassParams& p, IoFlags fl, long long base, int k, cx<T> v, long long c,
                                          int peer) {
  if (p.valid_out > 0 && k >= p.valid_out) return;
  gstore<T>(p, fl, base + (long long)k * p.os, finalize<T>(p, v, c, k), peer);
}

template <typename T, int R>
__device__ __forceinline__ void stockham_pass(const PassParams& p, IoFlags fl, IoFlags flo, int ns, bool src_global,
                                              bool dst_global, const cx<T>* __restrict__ src, cx<T>* __restrict__ dst,
                                              long long ibase, long long obase, long long gtw_c, int peer, int tj) {
  const int n = p.n;
  const int nbf = n / R;
  for (int j = tj; j < nbf; j += p.threads_per_fft) {
    cx<T> v[R];
    if (src_global) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = load_in<T>(p, fl, ibase, j + r * nbf);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = src[padidx<T>(j + r * nbf)];
    }
    const int k = j % ns;
    if (ns > 1) {
      const int step = k * (n / (ns * R));
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, r * step));
    }
    DFT<R, T>::run(v);
    const int ob = (j - k) * R + k;
    if (dst_global) {
#pragma unroll
      for (int r = 0; r < R; ++r) store_out<T>(p, flo, obase, ob + r * ns, v[r], gtw_c, peer);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) dst[padidx<T>(ob + r * ns)] = v[r];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(sizeof(T) == 4 ? 512 : 256) wg_generic_kernel(const PassParams p, const bool il, const bool swap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = p.ffts_per_block;
  const int T_ = p.threads_per_fft;
  const int pitch = p.pitch;
  cx<T>* buf0 = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* buf1 = buf0 + (size_t)F * pitch;
  long long* s_ibase = reinterpret_cast<long long*>(buf1 + (size_t)F * pitch);
  long long* s_obase = s_ibase + F;
  long long* s_gtw = s_obase + F;
  int* s_peer = reinterpret_cast<int*>(s_gtw + F);
  const IoFlags fl{il, swap && !(p.mod_flags & MOD_NO_USER_SWAP_IN)};
  const IoFlags flo{il, swap && !(p.mod_flags & MOD_NO_USER_SWAP_OUT)};
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;
  const int f = tid / T_;
  const int tj = tid - f * T_;
  const int n = p.n;

  for (long long g0 = (long long)blockIdx.x * F; g0 < p.batch_total; g0 += (long long)gridDim.x * F) {
    const int nf = (int)min((long long)F, p.batch_total - g0);
    __syncthreads();
    if (tid < nf) {
      long long g = g0 + tid, ib = p.ioff, ob = p.ooff, c = 0;
      int peer = -1;
#pragma unroll
      for (int d = 0; d < kMaxBatchDims; ++d) {
        const long long q = g / p.nb[d];
        const long long b = g - q * p.nb[d];
        g = q;
        ib += b * p.ibd[d];
        ob += b * p.obd[d];
        if (d == p.gtw_dim) c = b;
        if (d == p.peer_dim) peer = (int)b;
      }
      s_ibase[tid] = ib;
      s_obase[tid] = ob;
      s_gtw[tid] = c;
      s_peer[tid] = peer;
    }
    __syncthreads();
    cx<T>* cur = buf0;
    cx<T>* nxt = buf1;
    if (p.in_mode == IO_STAGED_ELEM) {
      const unsigned total = (unsigned)nf * (unsigned)n;
      for (unsigned e = tid; e < total; e += nthreads) {
        const unsigned ff = e / (unsigned)n;
        const int i = (int)(e - ff * (unsigned)n);
        cur[ff * pitch + padidx<T>(i)] = load_in<T>(p, fl, s_ibase[ff], i);
      }
      __syncthreads();
    } else if (p.in_mode == IO_STAGED_BATCH) {
      const unsigned total = (unsigned)F * (unsigned)n;
      for (unsigned e = tid; e < total; e += nthreads) {
        const int ff = (int)(e & (unsigned)(F - 1));
        const int i = (int)(e / (unsigned)F);
        if (ff < nf) cur[ff * pitch + padidx<T>(i)] = load_in<T>(p, fl, s_ibase[ff], i);
      }
      __syncthreads();
    }
    const bool active = f < nf;
    const long long ibase = active ? s_ibase[f] : 0;
    const long long obase = active ? s_obase[f] : 0;
    const long long gtw_c = active ? s_gtw[f] : 0;
    const int peer = active ? s_peer[f] : -1;
    int ns = 1;
    for (int ps = 0; ps < p.num_radices; ++ps) {
      const bool src_global = (ps == 0) && (p.in_mode == IO_DIRECT);
      const bool dst_global = (ps == p.num_radices - 1) && (p.out_mode == IO_DIRECT);
      const int R = p.radix[ps];
      if (active) {
        const cx<T>* s = cur + (size_t)f * pitch;
        cx<T>* d = nxt + (size_t)f * pitch;
        switch (R) {
#define PFFT_CASE(RR)                                                                                         \
  case RR:                                                                                                    \
    stockham_pass<T, RR>(p, fl, flo, ns, src_global, dst_global, s, d, ibase, obase, gtw_c, peer, tj);             \
    break;
          PFFT_CASE(1)
          PFFT_CASE(2)
          PFFT_CASE(3)
          PFFT_CASE(4)
          PFFT_CASE(5)
          PFFT_CASE(6)
          PFFT_CASE(7)
          PFFT_CASE(8)
          PFFT_CASE(9)
          PFFT_CASE(10)
          PFFT_CASE(11)
          PFFT_CASE(12)
          PFFT_CASE(13)
          PFFT_CASE(16)
          PFFT_CASE(17)
          PFFT_CASE(19)
          PFFT_CASE(23)
          PFFT_CASE(29)
          PFFT_CASE(31)
#undef PFFT_CASE
          default:
            break;
        }
      }
      ns *= R;
      if (!dst_global) {
        __syncthreads();
        cx<T>* t = cur;
        cur = nxt;
        nxt = t;
      }
    }
    if (p.out_mode == IO_STAGED_ELEM) {
      const unsigned total = (unsigned)nf * (unsigned)n;
      for (unsigned e = tid; e < total; e += nthreads) {
        const unsigned ff = e / (unsigned)n;
        const int i = (int)(e - ff * (unsigned)n);
        store_out<T>(p, flo, s_obase[ff], i, cur[ff * pitch + padidx<T>(i)], s_gtw[ff], s_peer[ff]);
      }
    } else if (p.out_mode == IO_STAGED_BATCH) {
      const unsigned total = (unsigned)F * (unsigned)n;
      for (unsigned e = tid; e < total; e += nthreads) {
        const int ff = (int)(e & (unsigned)(F - 1));
        const int i = (int)(e / (unsigned)F);
        if (ff < nf) store_out<T>(p, flo, s_obase[ff], i, cur[ff * pitch + padidx<T>(i)], s_gtw[ff], s_peer[ff]);
      }
    }
  }
}

template <typename T>
static cudaError_t launch_wg_generic_t(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  const size_t smem = wg_generic_smem_bytes(p.ffts_per_block, p.pitch, sizeof(T));
  cudaError_t e = ensure_dynamic_smem(wg_generic_kernel<T>, smem);
  if (e != cudaSuccess) return e;
  wg_generic_kernel<T><<<grid, p.ffts_per_block * p.threads_per_fft, smem, stream>>>(p, il, swap);
  return cudaGetLastError();
}

cudaError_t launch_wg_generic(const PassParams& p, bool is_double, bool il, bool swap, int grid, cudaStream_t stream) {
  return is_double ? launch_wg_generic_t<double>(p, il, swap, grid, stream)
                   : launch_wg_generic_t<float>(p, il, swap, grid, stream);
}

}  // namespace pfft
