// SUBGROUP level: L = 2..32 lanes of one warp compute one transform of length N = L * M, M points per lane.
//
// Reference counterpart: subgroup_impl + sg_dft / cross_sg_dft (/root/reference/src/portfft/dispatcher/
// subgroup_dispatcher.hpp:85-664, /root/reference/src/portfft/common/subgroup.hpp:80-291): f_sg lanes x f_wi
// registers, cross-lane DFT by recursive Cooley-Tukey over select_from_group / permute_group_by_xor with a lane
// transpose, twiddles staged in local memory, output left transposed and un-transposed through local memory.  Here:
//   * lane l loads x[brev(l) + L*r], r < M, straight from global memory (a transform's lanes read L consecutive
//     elements per instruction: coalesced) and runs the compile-time DFT<M> on its registers (dft.cuh);
//   * after the twiddle w_N^{brev(l)*k} the L-point cross-lane transform is log2(L) decimation-in-time radix-2
//     Stockham stages, each ONE __shfl_xor_sync exchange per register with the stage twiddle held in a register
//     (bit-reversed load order => natural output order, no lane transpose);
//   * lane l then holds X[k + M*l], k < M: the warp writes its 32*M results to a warp-private shared-memory tile with
//     odd pitch (conflict free) and streams them out with lanes along the element index (coalesced).  Only
//     __syncwarp() is used: warps never wait for each other;
//   * 32/L transforms per warp (as subgroup_dispatcher.hpp:130), persistent grid-stride over warp tiles;
//   * backward = (re <-> im) swap on load and store, scale fused on store.
#pragma once
#include "io.cuh"
#include "launch_utils.h"
#include "kernels.h"

namespace pfft {

template <typename T>
__device__ __forceinline__ cx<T> shfl_xor_cx(cx<T> a, int mask) {
  cx<T> b;
  b.x = __shfl_xor_sync(0xffffffffu, a.x, mask);
  b.y = __shfl_xor_sync(0xffffffffu, a.y, mask);
  return b;
}

template <int M, typename T>
__global__ void __launch_bounds__(kSgBlock) sg_kernel(const PassParams p, const bool il, const bool swap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int PM = M | 1;
  constexpr bool TW_REGS = M * sizeof(T) <= 64;  // per-lane twiddles stay in registers for the whole batch loop
  constexpr int NTW = TW_REGS ? (M > 1 ? M - 1 : 1) : 1;
  const IoFlags fl{il, swap};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = p.threads_per_fft;
  const int logL = __ffs(L) - 1;
  const int sub = lane >> logL;         // transform of this lane within the warp tile
  const int l = lane & (L - 1);         // lane within the transform
  const int lrev = (int)(__brev((unsigned)l) >> (32 - logL));
  const int fpw = 32 >> logL;           // transforms per warp
  const int N = L * M;
  cx<T>* wbuf = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)warp * 32 * PM;
  const bool one_dim = single_batch_dim(p);
  const T scale = T(p.scale);

  // stage twiddles: lanes in the upper half of a butterfly multiply by w_{2h}^{l mod h} = w_N^{(l mod h) * N/(2h)}
  cx<T> wst[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int h = 1 << s;
    wst[s] = {T(1), T(0)};
    if (s < logL && (l & h)) wst[s] = ldg_cx<T>(p.tw, (long long)(l & (h - 1)) * (N >> (s + 1)));
  }
  cx<T> wtw[NTW];
  if (TW_REGS) {
#pragma unroll
    for (int k = 1; k < M; ++k) wtw[k - 1] = ldg_cx<T>(p.tw, (long long)lrev * k);
  }

  const long long tiles = (p.batch_total + fpw - 1) / fpw;
  const long long tile_stride = (long long)gridDim.x * (kSgBlock / 32);
  for (long long tile = (long long)blockIdx.x * (kSgBlock / 32) + warp; tile < tiles; tile += tile_stride) {
    const long long g = tile * fpw + sub;
    const bool active = g < p.batch_total;
    long long ib = 0, ob = 0;
    int peer = -1;
    if (active) batch_bases(p, one_dim, g, ib, ob, peer);
    cx<T> v[M];
#pragma unroll
    for (int r = 0; r < M; ++r) v[r] = active ? gload<T>(p, fl, ib + lrev + L * r) : cx<T>{T(0), T(0)};
    DFT<M, T>::run(v);
#pragma unroll
    for (int k = 1; k < M; ++k) v[k] = cmul(v[k], TW_REGS ? wtw[k - 1] : ldg_cx<T>(p.tw, (long long)lrev * k));
    // cross-lane L-point transform: decimation in time, one shuffle exchange per register and stage
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      if (s < logL) {
        const int h = 1 << s;
        const bool upper = (l & h) != 0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
          cx<T> a = v[k];
          if (s > 0) a = cmul(a, wst[s]);  // stage 0: w_2^0 = 1
          const cx<T> b = shfl_xor_cx(a, h);
          v[k] = upper ? b - a : a + b;
        }
      }
    }
    if (p.apply_scale) {
#pragma unroll
      for (int k = 0; k < M; ++k) v[k] = cscale(v[k], scale);
    }
    // lane holds X[k + M*l]: transpose through the warp-private tile, store with lanes along the element index
    __syncwarp();
#pragma unroll
    for (int k = 0; k < M; ++k) wbuf[lane * PM + k] = v[k];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const int q = lane + 32 * i;  // element of the warp tile (fpw transforms x N)
      const int c = q / M;          // owning lane
      const int sub2 = c >> logL;
      const cx<T> val = wbuf[c * PM + (q - c * M)];
      long long ob2;
      int peer2 = -1;
      if (one_dim && p.peer_dim < 0) {
        ob2 = p.ooff + (tile * fpw + sub2) * p.obd[0];
      } else {
        ob2 = __shfl_sync(0xffffffffu, ob, sub2 << logL);
        peer2 = __shfl_sync(0xffffffffu, peer, sub2 << logL);
      }
      if (tile * fpw + sub2 < p.batch_total) gstore<T>(p, fl, ob2 + (q - sub2 * N), val, peer2);
    }
  }
}

template <int M, typename T>
cudaError_t launch_sg_m(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  const size_t smem = sg_smem_bytes(M, sizeof(T));
  if (smem > 48 * 1024) {
    cudaError_t e = ensure_dynamic_smem(sg_kernel<M, T>, smem);
    if (e != cudaSuccess) return e;
  }
  sg_kernel<M, T><<<grid, kSgBlock, smem, stream>>>(p, il, swap);
  return cudaGetLastError();
}

}  // namespace pfft
