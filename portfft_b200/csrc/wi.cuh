// WORKITEM level: one CUDA thread computes one whole transform of compile-time length N <= 32 in registers.
//
// Reference counterpart: workitem_impl + wi_dft (/root/reference/src/portfft/dispatcher/workitem_dispatcher.hpp:
// 99-350, /root/reference/src/portfft/common/workitem.hpp:64-219): run-time-sized recursive Cooley-Tukey on a
// `priv[2*56]` array, sub-group cooperative global -> padded local -> private staging, store modifiers.  Here:
//   * N is a template parameter: the transform is one DFT<N,T>::run on N complex registers (dft.cuh), no loops over
//     run-time sizes, no private-memory indexing;
//   * DIRECT I/O: the thread reads / writes its own elements; lanes run along the batch index, so batch-interleaved
//     layouts (distance 1) and the outer dimensions of N-D transforms are perfectly coalesced and need no shared
//     memory at all (the reference's `interleaved_transforms_input` path, workitem_dispatcher.hpp:215-229);
//   * STAGED I/O (unit element stride): the CTA's tile of transforms is copied cooperatively with lanes along the
//     element index (coalesced) through shared memory with an odd pitch, so that the per-thread strided reads are
//     bank-conflict free (the reference pads with pad_local, memory_views.hpp:79-85);
//   * backward = (re <-> im) swap, scale fused on store; all N loads of a thread are in flight together.
#pragma once
#include <cstdint>

#include "io.cuh"
#include "launch_utils.h"
#include "kernels.h"

namespace pfft {

constexpr int kWiThreads = kWiBlock;

template <int N>
__host__ __device__ constexpr int wi_pitch() {
  return N | 1;
}

// one complex element global -> shared by cp.async: interleaved storage = one copy of sizeof(complex), split storage =
// one copy per component (the element size is the alignment both sides are guaranteed to have)
template <typename T>
__device__ __forceinline__ void cp_async_cx(cx<T>* dst, const PassParams& p, bool il, long long idx) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  if (il) {
    const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in_re) + idx;
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(2 * sizeof(T)) : "memory");
  } else {
    const T* re = reinterpret_cast<const T*>(p.in_re) + idx;
    const T* im = reinterpret_cast<const T*>(p.in_im) + idx;
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(re), "n"(sizeof(T)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d + (unsigned)sizeof(T)), "l"(im), "n"(sizeof(T))
                 : "memory");
  }
}

template <int N, typename T>
__global__ void __launch_bounds__(kWiThreads) wi_kernel(const PassParams p, const bool il, const bool swap,
                                                        const bool use_async) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int PITCH = wi_pitch<N>();
  constexpr int F = kWiThreads;
  cx<T>* buf = reinterpret_cast<cx<T>*>(smem_raw);
  long long* s_base = reinterpret_cast<long long*>(buf + (size_t)F * PITCH);
  int* s_peer = reinterpret_cast<int*>(s_base + F);
  const IoFlags fl{il, swap};
  const int tid = threadIdx.x;
  const bool one_dim = single_batch_dim(p);
  const bool in_staged = p.in_mode == IO_STAGED_ELEM, out_staged = p.out_mode == IO_STAGED_ELEM;
  const bool in_contig = one_dim && p.is == 1 && p.ibd[0] == N;
  const bool out_contig = one_dim && p.os == 1 && p.obd[0] == N && p.peer_dim < 0;
  const T scale = T(p.scale);

  for (long long g0 = (long long)blockIdx.x * F; g0 < p.batch_total; g0 += (long long)gridDim.x * F) {
    const int nf = (int)min((long long)F, p.batch_total - g0);
    const bool active = tid < nf;
    long long ib = 0, ob = 0;
    int peer = -1;
    if (active) batch_bases(p, one_dim, g0 + tid, ib, ob, peer);
    cx<T> v[N];
    if (in_staged) {
      __syncthreads();  // previous iteration's staged stores have left the buffer
      if (!in_contig) {
        if (active) s_base[tid] = ib;
        __syncthreads();
      }
      const long long tile0 = p.ioff + g0 * (long long)N;
      const int total = nf * N;
      if (use_async) {
        // cp.async (LDGSTS): global -> shared without staging registers, all N copies of a thread in flight at once
        // (rolled loop: the copies do not block, and unrolling only multiplies address registers)
#pragma unroll 2
        for (int i = 0; i < N; ++i) {
          const int e = tid + i * F;
          if (e < total) {
            const int ff = e / N, k = e - ff * N;
            cp_async_cx<T>(&buf[ff * PITCH + k], p, il, in_contig ? tile0 + e : s_base[ff] + k * p.is);
          }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      } else {
#pragma unroll 4
        for (int e = tid; e < total; e += F) {
          const int ff = e / N, k = e - ff * N;
          buf[ff * PITCH + k] = gload<T>(p, IoFlags{il, false}, in_contig ? tile0 + e : s_base[ff] + k * p.is);
        }
      }
      __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = buf[tid * PITCH + j];
        if (swap) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const T tmp = v[j].x;
            v[j].x = v[j].y;
            v[j].y = tmp;
          }
        }
      }
    } else if (active) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = gload<T>(p, fl, ib + (long long)j * p.is);
    }
    if (active) {
      DFT<N, T>::run(v);
      if (p.apply_scale) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = cscale(v[j], scale);
      }
    }
    if (out_staged) {
      __syncthreads();  // every thread has read its staged input
      if (active) {
#pragma unroll
        for (int j = 0; j < N; ++j) buf[tid * PITCH + j] = v[j];
        if (!out_contig) {
          s_base[tid] = ob;
          s_peer[tid] = peer;
        }
      }
      __syncthreads();
      const long long tile0 = p.ooff + g0 * (long long)N;
      const int total = nf * N;
#pragma unroll 4
      for (int i = 0; i < N; ++i) {
        const int e = tid + i * F;
        if (e < total) {
          const int ff = e / N, k = e - ff * N;
          if (out_contig)
            gstore<T>(p, fl, tile0 + e, buf[ff * PITCH + k]);
          else
            gstore<T>(p, fl, s_base[ff] + k * p.os, buf[ff * PITCH + k], s_peer[ff]);
        }
      }
    } else if (active) {
#pragma unroll
      for (int j = 0; j < N; ++j) gstore<T>(p, fl, ob + (long long)j * p.os, v[j], peer);
    }
  }
}

template <int N, typename T>
cudaError_t launch_wi_n(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  const bool staged = p.in_mode == IO_STAGED_ELEM || p.out_mode == IO_STAGED_ELEM;
  const size_t smem = staged ? wi_smem_bytes(N, sizeof(T)) : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = ensure_dynamic_smem(wi_kernel<N, T>, smem);
    if (e != cudaSuccess) return e;
  }
  // cp.async needs the source aligned to the copy size: element alignment of the user's pointers
  const size_t unit = il ? 2 * sizeof(T) : sizeof(T);
  const bool use_async = reinterpret_cast<uintptr_t>(p.in_re) % unit == 0 &&
                         (il || reinterpret_cast<uintptr_t>(p.in_im) % unit == 0);
  wi_kernel<N, T><<<grid, kWiThreads, smem, stream>>>(p, il, swap, use_async);
  return cudaGetLastError();
}

}  // namespace pfft
