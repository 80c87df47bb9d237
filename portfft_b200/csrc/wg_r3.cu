// WORKGROUP level, three compile-time radix passes: N = R0 * R1 * R2 (1000 = 10*10*10, 2048 = 16*16*8, ...), F
// transforms per CTA, any element stride / distance / offset, interleaved or split storage.
//
// This is the mixed-radix block-level kernel for layouts the TMA tile kernels cannot take (split storage, element
// strides != 1, odd offsets): BASELINE config C3 (N = 1000, split, stride 2, offsets) runs here.  Reference
// counterpart: workgroup_impl / wg_dft (/root/reference/src/portfft/dispatcher/workgroup_dispatcher.hpp:94-281,
// /root/reference/src/portfft/common/workgroup.hpp:85-346), which cannot run this configuration at all (UNPACKED
// layouts beyond one sub-group are rejected, descriptor_validation.hpp:67-80) and for PACKED data uses a two-factor
// split (25 x 40) with run-time sub-group DFTs.  Here:
//   * Stockham autosort with all radices, strides and twiddle indices compile-time; thread t of a transform owns
//     butterfly t of every pass, so its twiddles w^{k r} are loaded ONCE per CTA and stay in registers;
//   * first pass reads global memory directly, last pass writes it directly (lanes -> consecutive elements);
//   * two exchanges through ping-pong shared-memory buffers padded by one element per R0 (index i -> i + i/R0): with
//     the radix-R0 scatter of pass 0 this makes every access pattern of all three passes (nearly) conflict free for
//     radix 10 as well as for radix 16 -- the generic kernel's power-of-two padding gives 57% conflicts at N = 1000;
//   * the inputs of the CTA's NEXT group of transforms are loaded into registers before the current group is
//     transformed (one butterfly per thread in pass 0), so the global-load latency -- strided 4-byte loads for C3 --
//     is covered by three passes of arithmetic instead of stalling pass 0 (ncu: long_scoreboard was the top stall);
//   * storage (interleaved / split) and the backward swap are template parameters: no per-element branches;
//   * backward = (re <-> im) swap, scale fused into the store; two block barriers per transform.
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "launch_utils.h"

namespace pfft {

namespace r3 {
constexpr int cmax(int a, int b) { return a > b ? a : b; }
}

template <int R0, int R1, int R2>
struct R3Cfg {
  static constexpr int N = R0 * R1 * R2;
  static constexpr int RMAX = r3::cmax(R0, r3::cmax(R1, R2));
  static constexpr int TPF = N / RMAX;                      // threads per transform
  static constexpr bool ONE = (R0 == R1 && R1 == R2);       // one butterfly per thread and pass: register twiddles
  static constexpr int PITCH = (N + N / R0 + 1) | 1;
  __host__ __device__ static constexpr int pad(int i) { return i + i / R0; }
};

template <typename T, int R0, int R1, int R2, bool IL, bool SWAP>
__global__ void __launch_bounds__(512) wg_r3_kernel(const PassParams p) {
  using Cfg = R3Cfg<R0, R1, R2>;
  constexpr int N = Cfg::N, TPF = Cfg::TPF, PITCH = Cfg::PITCH;
  constexpr bool PREFETCH = N / R0 == TPF;  // one pass-0 butterfly per thread: its inputs can be fetched a group ahead
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int F = p.ffts_per_block;
  cx<T>* buf0 = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* buf1 = buf0 + (size_t)F * PITCH;
  const IoFlags fl{IL, SWAP};
  const int f = threadIdx.x / TPF, t = threadIdx.x - f * TPF;
  const bool one_dim = single_batch_dim(p);
  const T scale = T(p.scale);
  cx<T>* b0 = buf0 + (size_t)f * PITCH;
  cx<T>* b1 = buf1 + (size_t)f * PITCH;

  // register-resident twiddles (ONE): pass 1 uses w_{R0 R1}^{k1 r}, k1 = t % R0; pass 2 uses w_N^{k2 r}, k2 = t % (R0 R1)
  constexpr bool TWREG = Cfg::ONE && sizeof(T) == 4;
  cx<T> tw1[TWREG ? R1 : 1], tw2[TWREG ? R2 : 1];
  if (TWREG) {
#pragma unroll
    for (int r = 1; r < R1; ++r) tw1[r] = ldg_cx<T>(p.tw, (long long)(t % R0) * r * R2);
#pragma unroll
    for (int r = 1; r < R2; ++r) tw2[r] = ldg_cx<T>(p.tw, (long long)(t % (R0 * R1)) * r);
  }

  const long long gstride = (long long)gridDim.x * F;
  long long g0 = (long long)blockIdx.x * F;
  bool active = g0 + f < p.batch_total;
  long long ib = 0, ob = 0;
  int peer = -1;
  if (active) batch_bases(p, one_dim, g0 + f, ib, ob, peer);
  cx<T> nxt[PREFETCH ? R0 : 1];
  if (PREFETCH && active) {
#pragma unroll
    for (int r = 0; r < R0; ++r) nxt[r] = gload<T>(p, fl, ib + (long long)(t + r * (N / R0)) * p.is);
  }
  for (; g0 < p.batch_total; g0 += gstride) {
    const bool cur_active = active;
    const long long cur_ob = ob;
    const int cur_peer = peer;
    // ---- pass 0: radix R0, global -> buf0 ---------------------------------------------------------------------
    if (PREFETCH) {
      cx<T> v[R0];
#pragma unroll
      for (int r = 0; r < R0; ++r) v[r] = nxt[r];
      // inputs of the next group: in flight while this group runs its three passes
      active = g0 + gstride + f < p.batch_total;
      if (active) {
        batch_bases(p, one_dim, g0 + gstride + f, ib, ob, peer);
#pragma unroll
        for (int r = 0; r < R0; ++r) nxt[r] = gload<T>(p, fl, ib + (long long)(t + r * (N / R0)) * p.is);
      }
      if (cur_active) {
        DFT<R0, T>::run(v);
#pragma unroll
        for (int r = 0; r < R0; ++r) b0[Cfg::pad(t * R0 + r)] = v[r];
      }
    } else if (cur_active) {
#pragma unroll 1
      for (int j = t; j < N / R0; j += TPF) {
        cx<T> v[R0];
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = gload<T>(p, fl, ib + (long long)(j + r * (N / R0)) * p.is);
        DFT<R0, T>::run(v);
#pragma unroll
        for (int r = 0; r < R0; ++r) b0[Cfg::pad(j * R0 + r)] = v[r];
      }
    }
    __syncthreads();
    // ---- pass 1: radix R1, buf0 -> buf1 -----------------------------------------------------------------------
    if (cur_active) {
#pragma unroll 1
      for (int j = t; j < N / R1; j += TPF) {
        const int k = j % R0;
        cx<T> v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = b0[Cfg::pad(j + r * (N / R1))];
#pragma unroll
        for (int r = 1; r < R1; ++r) v[r] = cmul(v[r], TWREG ? tw1[r] : ldg_cx<T>(p.tw, (long long)k * r * R2));
        DFT<R1, T>::run(v);
        const int ob1 = (j - k) * R1 + k;
#pragma unroll
        for (int r = 0; r < R1; ++r) b1[Cfg::pad(ob1 + r * R0)] = v[r];
      }
    }
    __syncthreads();
    // ---- pass 2: radix R2, buf1 -> global ---------------------------------------------------------------------
    if (cur_active) {
#pragma unroll 1
      for (int j = t; j < N / R2; j += TPF) {
        const int k = j % (R0 * R1);
        cx<T> v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = b1[Cfg::pad(j + r * (N / R2))];
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], TWREG ? tw2[r] : ldg_cx<T>(p.tw, (long long)k * r));
        DFT<R2, T>::run(v);
        const int ob2 = (j - k) * R2 + k;
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          cx<T> o = v[r];
          if (p.apply_scale) o = cscale(o, scale);
          gstore<T>(p, fl, cur_ob + (long long)(ob2 + r * (R0 * R1)) * p.os, o, cur_peer);
        }
      }
    }
    if (!PREFETCH) {
      active = g0 + gstride + f < p.batch_total;
      if (active) batch_bases(p, one_dim, g0 + gstride + f, ib, ob, peer);
    }
  }
}

template <typename T, int R0, int R1, int R2, bool IL, bool SWAP>
static cudaError_t launch_r3_v(const PassParams& p, int grid, cudaStream_t stream) {
  using Cfg = R3Cfg<R0, R1, R2>;
  const size_t smem = (size_t)2 * p.ffts_per_block * Cfg::PITCH * sizeof(cx<T>);
  auto kern = wg_r3_kernel<T, R0, R1, R2, IL, SWAP>;
  cudaError_t e = ensure_dynamic_smem(kern, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, p.ffts_per_block * Cfg::TPF, smem, stream>>>(p);
  return cudaGetLastError();
}

template <typename T, int R0, int R1, int R2>
static cudaError_t launch_r3_t(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  if (!il) return launch_r3_v<T, R0, R1, R2, false, false>(p, grid, stream);
  return swap ? launch_r3_v<T, R0, R1, R2, true, true>(p, grid, stream)
              : launch_r3_v<T, R0, R1, R2, true, false>(p, grid, stream);
}

#define PFFT_R3_LIST(X) \
  X(1000, 10, 10, 10)   \
  X(1024, 16, 8, 8)     \
  X(1536, 16, 12, 8)    \
  X(2048, 16, 16, 8)    \
  X(3072, 16, 16, 12)   \
  X(4096, 16, 16, 16)

bool r3_supported(int n, bool is_double, int* threads_per_fft, int* pitch) {
  (void)is_double;
  switch (n) {
#define X(NN, A, B, C)                                        \
  case NN:                                                    \
    if (threads_per_fft) *threads_per_fft = R3Cfg<A, B, C>::TPF; \
    if (pitch) *pitch = R3Cfg<A, B, C>::PITCH;                \
    return true;
    PFFT_R3_LIST(X)
#undef X
    default:
      return false;
  }
}

cudaError_t launch_wg_r3(const PassParams& p, bool is_double, bool il, bool swap, int grid, cudaStream_t stream) {
  switch (p.n) {
#define X(NN, A, B, C)                                                                 \
  case NN:                                                                             \
    return is_double ? launch_r3_t<double, A, B, C>(p, il, swap, grid, stream)         \
                     : launch_r3_t<float, A, B, C>(p, il, swap, grid, stream);
    PFFT_R3_LIST(X)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pfft
