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
  // Second exchange buffer: pass 1 writes index R0 R1 a + k + R0 r' (k = lane % R0), i.e. groups of R0 lanes that are
  // G = R0 R1 + R1 padded elements apart.  A half-warp of 16 lanes spans two or three groups unless R0 is a multiple
  // of 16, and they collide in the banks unless G = R0 (mod 16): XPAD extra elements per R0 R1 make it so (radix 10:
  // G = 110 = 14 (mod 16) put lanes 10..15 onto the banks of lanes 0..3; 12 more per group remove the overlap).
  static constexpr int XPAD = R0 % 16 == 0 ? 0 : (((R0 - (R0 * R1 + R1)) % 16) + 16) % 16;
  static constexpr int PITCH = (N + N / R0 + XPAD * R2 + 1) | 1;
  __host__ __device__ static constexpr int pad(int i) { return i + i / R0; }
  __host__ __device__ static constexpr int pad1(int i) { return i + i / R0 + XPAD * (i / (R0 * R1)); }
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

  // Address arithmetic is kept out of the element loops: every shared-memory index of a pass is (per-butterfly base)
  // + r * (compile-time stride) -- the padding i + i/R0 is affine in r because every stride is a multiple of R0 or
  // smaller than R0 -- and every global address is (per-transform pointer) + r * (one 64-bit step).
  constexpr int S1 = N / R1 + N / (R1 * R0);  // pad(j + r N/R1) = pad(j) + r S1
  constexpr int S2 = N / R2 + N / (R2 * R0) + Cfg::XPAD;  // pad1(j + r N/R2) = pad1(j) + r S2  (N / R2 = R0 R1)
  const long long in_step = (long long)(N / R0) * p.is, out_step = (long long)(R0 * R1) * p.os;
  auto load_elem = [&](long long idx) -> cx<T> {
    cx<T> v;
    if (IL) {
      v = reinterpret_cast<const cx<T>*>(p.in_re)[idx];
      if (SWAP) v = cx<T>{v.y, v.x};
    } else {
      v.x = reinterpret_cast<const T*>(p.in_re)[idx];
      v.y = reinterpret_cast<const T*>(p.in_im)[idx];
    }
    return v;
  };

  const long long gstride = (long long)gridDim.x * F;
  long long g0 = (long long)blockIdx.x * F;
  bool active = g0 + f < p.batch_total;
  long long ib = 0, ob = 0;
  int peer = -1;
  if (active) batch_bases(p, one_dim, g0 + f, ib, ob, peer);
  cx<T> nxt[PREFETCH ? R0 : 1];
  if (PREFETCH && active) {
    const long long i0 = ib + (long long)t * p.is;
#pragma unroll
    for (int r = 0; r < R0; ++r) nxt[r] = load_elem(i0 + r * in_step);
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
        const long long i0 = ib + (long long)t * p.is;
#pragma unroll
        for (int r = 0; r < R0; ++r) nxt[r] = load_elem(i0 + r * in_step);
      }
      if (cur_active) {
        DFT<R0, T>::run(v);
        cx<T>* dst = b0 + t * (R0 + 1);  // pad(t R0 + r) = t (R0 + 1) + r
#pragma unroll
        for (int r = 0; r < R0; ++r) dst[r] = v[r];
      }
    } else if (cur_active) {
#pragma unroll 1
      for (int j = t; j < N / R0; j += TPF) {
        cx<T> v[R0];
        const long long i0 = ib + (long long)j * p.is;
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = load_elem(i0 + r * in_step);
        DFT<R0, T>::run(v);
        cx<T>* dst = b0 + j * (R0 + 1);
#pragma unroll
        for (int r = 0; r < R0; ++r) dst[r] = v[r];
      }
    }
    __syncthreads();
    // ---- pass 1: radix R1, buf0 -> buf1 -----------------------------------------------------------------------
    if (cur_active) {
#pragma unroll 1
      for (int j = t; j < N / R1; j += TPF) {
        const int k = j % R0;
        cx<T> v[R1];
        const cx<T>* src = b0 + Cfg::pad(j);
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = src[r * S1];
#pragma unroll
        for (int r = 1; r < R1; ++r) v[r] = cmul(v[r], TWREG ? tw1[r] : ldg_cx<T>(p.tw, (long long)k * r * R2));
        DFT<R1, T>::run(v);
        cx<T>* dst = b1 + Cfg::pad1((j - k) * R1 + k);  // (j - k) R1 is a multiple of R0 R1: pad1(. + r R0) = pad1(.) + r (R0 + 1)
#pragma unroll
        for (int r = 0; r < R1; ++r) dst[r * (R0 + 1)] = v[r];
      }
    }
    __syncthreads();
    // ---- pass 2: radix R2, buf1 -> global ---------------------------------------------------------------------
    if (cur_active) {
      // output pointers of this transform (peer table: memory of another GPU)
      T* ore = reinterpret_cast<T*>(cur_peer >= 0 ? p.out_tab_re[cur_peer] : p.out_re);
      T* oim = reinterpret_cast<T*>(cur_peer >= 0 ? p.out_tab_im[cur_peer] : p.out_im);
#pragma unroll 1
      for (int j = t; j < N / R2; j += TPF) {
        const int k = j % (R0 * R1);
        cx<T> v[R2];
        const cx<T>* src = b1 + Cfg::pad1(j);
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = src[r * S2];
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], TWREG ? tw2[r] : ldg_cx<T>(p.tw, (long long)k * r));
        DFT<R2, T>::run(v);
        const long long o0 = cur_ob + (long long)((j - k) * R2 + k) * p.os;
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          cx<T> o = v[r];
          if (p.apply_scale) o = cscale(o, scale);
          const long long idx = o0 + r * out_step;
          if (IL) {
            if (SWAP) o = cx<T>{o.y, o.x};
            reinterpret_cast<cx<T>*>(ore)[idx] = o;
          } else {
            ore[idx] = o.x;
            oim[idx] = o.y;
          }
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
