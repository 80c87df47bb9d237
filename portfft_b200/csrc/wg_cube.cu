// WORKGROUP level, specialised: N = R^3 (4096 = 16^3, 512 = 8^3), packed interleaved data, one transform per CTA
// iteration, persistent CTAs.  This is the hot kernel of BASELINE config C2 (fp32 N=4096, batch 65536, in place).
//
// Reference counterpart: workgroup_impl + wg_dft (/root/reference/src/portfft/dispatcher/workgroup_dispatcher.hpp:
// 94-281, /root/reference/src/portfft/common/workgroup.hpp:319-346): 64 work-items per 4096-point transform, 32
// sequential sub-FFT iterations per pass, scalar local-memory traffic, inter-step twiddles read from global memory
// and a strided 4-byte store.  Here:
//   * 256 threads x 16 points: three register-resident radix-16 passes (dft.cuh), all indices compile-time;
//   * the whole 32 KiB transform is fetched by ONE cp.async.bulk (TMA 1-D) into a 2-stage shared-memory ring
//     guarded by mbarriers, so the next transform streams from HBM while this one is computed;
//   * exchange 1 goes through a padded buffer (2 complex per 16: conflict-free 128-bit stores and 64-bit loads),
//     exchange 2 is written back into the consumed stage buffer (conflict-free unpadded) -> 2 block barriers per
//     transform;
//   * per-thread twiddles (w_256^{(t%16) r}, w_4096^{t r}) are loaded once per CTA and stay in registers;
//   * results leave with coalesced 64-bit stores straight from registers; backward = re/im swap; scale fused.
#include <cstdint>

#include "device_utils.cuh"
#include "kernels.h"

namespace pfft {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <int R>
__device__ __forceinline__ constexpr int epad(int i) {
  return i + 2 * (i / 16);  // 2 complex of padding per 16: keeps 16-byte alignment for vector stores
}

struct CubeArgs {
  const void* in;
  void* out;
  long long ioff, ooff, idist, odist, batch;
  const void* tw;  // w_N^k, k in [0, N)
  double scale;
  int apply_scale;
};

template <typename T, int R, bool SWAP, bool USE_TMA>
__global__ void __launch_bounds__(R* R, 2) wg_cube_kernel(const CubeArgs a) {
  constexpr int NT = R * R;
  constexpr int N = R * R * R;
  constexpr int EN = N + 2 * (N / 16);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T>* S0 = reinterpret_cast<cx<T>*>(smem_raw);  // TMA: stage 0 ; non-TMA: exchange-2 buffer
  cx<T>* S1 = S0 + N;                              // TMA: stage 1
  cx<T>* E = USE_TMA ? (S1 + N) : (S0 + N);        // padded exchange-1 buffer
  uint64_t* full = reinterpret_cast<uint64_t*>(E + EN);
  const int t = threadIdx.x;
  const int k2 = t % R;
  const cx<T>* gin = reinterpret_cast<const cx<T>*>(a.in) + a.ioff;
  cx<T>* gout = reinterpret_cast<cx<T>*>(a.out) + a.ooff;
  const long long stride = gridDim.x;

  // per-thread twiddles, resident in registers for the whole batch loop
  cx<T> tw2[R - 1], tw3[R - 1];
#pragma unroll
  for (int r = 1; r < R; ++r) {
    tw2[r - 1] = ldg_cx<T>(a.tw, (long long)k2 * r * R);  // w_{R^2}^{k2 r}
    tw3[r - 1] = ldg_cx<T>(a.tw, (long long)t * r);       // w_N^{t r}
  }
  const T scale = T(a.scale);

  if (USE_TMA) {
    if (t == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      fence_mbar_init();
      fence_proxy_async();
    }
    __syncthreads();
    if (t == 0) {
      long long k = blockIdx.x;
      if (k < a.batch) {
        mbar_expect_tx(&full[0], N * sizeof(cx<T>));
        bulk_g2s(S0, gin + k * a.idist, N * sizeof(cx<T>), &full[0]);
      }
      k += stride;
      if (k < a.batch) {
        mbar_expect_tx(&full[1], N * sizeof(cx<T>));
        bulk_g2s(S1, gin + k * a.idist, N * sizeof(cx<T>), &full[1]);
      }
    }
  }

  int it = 0;
  for (long long k = blockIdx.x; k < a.batch; k += stride, ++it) {
    cx<T>* S = USE_TMA ? ((it & 1) ? S1 : S0) : S0;
    cx<T> v[R];
    // ---- pass 1: x[t + NT r] -> radix R -> E[R t + r'] -------------------------------------------------------
    if (USE_TMA) {
      mbar_wait(&full[it & 1], (it >> 1) & 1);
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = S[t + NT * r];
    } else {
      const cx<T>* src = gin + k * a.idist;
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = src[t + NT * r];
    }
    if (SWAP) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T tmp = v[r].x;
        v[r].x = v[r].y;
        v[r].y = tmp;
      }
    }
    DFT<R, T>::run(v);
    {
      cx<T>* dst = E + (R * t + 2 * ((R * t) / 16));  // R*t..R*t+R-1 lie in whole 16-groups when R == 16
      if constexpr (R == 16 && sizeof(T) == 4) {
#pragma unroll
        for (int r = 0; r < R; r += 2)
          *reinterpret_cast<float4*>(dst + r) = make_float4(v[r].x, v[r].y, v[r + 1].x, v[r + 1].y);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) E[epad<R>(R * t + r)] = v[r];
      }
    }
    __syncthreads();
    if (USE_TMA && t == 0 && it >= 1) {
      // every thread has finished reading the stage used by iteration it-1 (its pass 3 precedes this barrier)
      const long long kn = k + stride;
      if (kn < a.batch) {
        cx<T>* Sn = (it & 1) ? S0 : S1;
        fence_proxy_async();
        mbar_expect_tx(&full[(it + 1) & 1], N * sizeof(cx<T>));
        bulk_g2s(Sn, gin + kn * a.idist, N * sizeof(cx<T>), &full[(it + 1) & 1]);
      }
    }
    // ---- pass 2: E[t + NT r] * w_{R^2}^{k2 r} -> radix R -> S[(t - k2) R + k2 + R r'] ---------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = E[epad<R>(t + NT * r)];
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw2[r - 1]);
    DFT<R, T>::run(v);
    {
      cx<T>* dst = S + (t - k2) * R + k2;
#pragma unroll
      for (int r = 0; r < R; ++r) dst[R * r] = v[r];
    }
    __syncthreads();
    // ---- pass 3: S[t + NT r] * w_N^{t r} -> radix R -> out[t + NT r'] ------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = S[t + NT * r];
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw3[r - 1]);
    DFT<R, T>::run(v);
    cx<T>* dst = gout + k * a.odist;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cx<T> o = v[r];
      if (a.apply_scale) o = cscale(o, scale);
      if (SWAP) {
        const T tmp = o.x;
        o.x = o.y;
        o.y = tmp;
      }
      dst[t + NT * r] = o;
    }
  }
}

template <typename T, int R>
size_t cube_smem_bytes(bool use_tma) {
  constexpr int N = R * R * R;
  constexpr int EN = N + 2 * (N / 16);
  return ((use_tma ? 2 : 1) * (size_t)N + EN) * sizeof(cx<T>) + 64;
}

template <typename T, int R, bool SWAP, bool USE_TMA>
static cudaError_t launch_cube_t(const CubeArgs& a, int grid, cudaStream_t stream) {
  const size_t smem = cube_smem_bytes<T, R>(USE_TMA);
  auto kern = wg_cube_kernel<T, R, SWAP, USE_TMA>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, R * R, smem, stream>>>(a);
  return cudaGetLastError();
}

// p: a single-pass plan entry with n == R^3, interleaved storage, unit strides, one batch dimension
cudaError_t launch_wg_cube(const PassParams& p, bool is_double, bool swap, int variant, int grid, cudaStream_t stream) {
  CubeArgs a;
  a.in = p.in_re;
  a.out = p.out_re;
  a.ioff = p.ioff;
  a.ooff = p.ooff;
  a.idist = p.ibd[0];
  a.odist = p.obd[0];
  a.batch = p.batch_total;
  a.tw = p.tw;
  a.scale = p.scale;
  a.apply_scale = p.apply_scale;
  const bool tma = variant == 0;
  if (!is_double && p.n == 4096) {
    if (tma) return swap ? launch_cube_t<float, 16, true, true>(a, grid, stream) : launch_cube_t<float, 16, false, true>(a, grid, stream);
    return swap ? launch_cube_t<float, 16, true, false>(a, grid, stream) : launch_cube_t<float, 16, false, false>(a, grid, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace pfft
