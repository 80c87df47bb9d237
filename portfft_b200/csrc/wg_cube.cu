// WORKGROUP level, specialised: N = R^3 (4096 = 16^3: one transform per CTA iteration; 512 = 8^3: a tile of F = 4
// consecutive transforms per iteration), packed interleaved data, persistent CTAs.  This is the hot kernel of BASELINE
// config C2 (fp32 N=4096, batch 65536, in place) and of the contiguous (z) pass of C5 (512^3).
//
// Reference counterpart: workgroup_impl + wg_dft (/root/reference/src/portfft/dispatcher/workgroup_dispatcher.hpp:
// 94-281, /root/reference/src/portfft/common/workgroup.hpp:319-346): 64 work-items per 4096-point transform, 32
// sequential sub-FFT iterations per pass, scalar local-memory traffic, inter-step twiddles read from global memory
// and a strided 4-byte store.  Here:
//   * 256 threads x 16 points: three register-resident radix-16 passes (dft.cuh), all indices compile-time;
//   * the whole 32 KiB transform is fetched by ONE cp.async.bulk (TMA 1-D) into a 2-stage shared-memory ring
//     guarded by mbarriers, so the next transform streams from HBM while this one is computed;
//   * exchange 1 goes through a padded buffer (2 complex per 16: conflict-free 128-bit stores and 64-bit loads),
//     exchange 2 is written back into the consumed stage buffer (R = 16: conflict-free as is; R = 8: bit 6 of the
//     index is folded into bit 3 so that the two index groups of a half-warp use different banks) -> 2 block
//     barriers per tile;
//   * per-thread twiddles (w_256^{(t%16) r}, w_4096^{t r}) are loaded once per CTA and stay in registers;
//   * results leave with coalesced 64-bit stores straight from registers; backward = re/im swap; scale fused.
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "device_utils.cuh"
#include "kernels.h"
#include "launch_utils.h"
#include "real_fuse.cuh"

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
  const void* tw2;  // REAL-domain fusion: w_{2N}^k, k in [0, 2N)
};

// ---------------------------------------------------------------------------------------------------------------
// REAL domain fused into the tile kernels (REAL = 1: real-to-complex, 2: complex-to-real; scheme and formulas:
// real.cu).  The complex transform of N = (real length) / 2 points runs on the real row read / written as interleaved
// pairs; what real.cu does in a separate pass over HBM happens here on the tile while it is in shared memory:
//   R2C: pass 3 puts Z back into the stage buffer, one more barrier, then every thread combines its own Z_k with the
//        partner Z_{N-k} and stores X_k (k = 0..N; rows of N + 1 outputs: plain 64-bit stores);
//   C2R: the stage receives the N + 1 input elements of a row (the bulk copy starts at the 16-byte boundary at or
//        below the row: `shift` = 0 or 1 elements), and pass 1 builds its inputs z'_j from X_{N-j} and X_j.
// The twiddle w_{2N}^k of element k = j + (N / R) r is (one per-thread register) x (the compile-time constant w_{2R}^r).
// ---------------------------------------------------------------------------------------------------------------
// One thread: fetch `rows` REAL-domain rows (row k + r of the batch into slot r of the stage) by bulk copies that start
// at the 16-byte boundary at or below each row.  REAL = 1: rows of N pairs (the real row); REAL = 2: rows of N + 1
// half-spectrum elements.  Rows are 8-byte aligned in general (in-place layouts: N + 1 elements apart), so a copy
// begins `shift` = 0 or 1 elements early -- inside the previous row, the buffer's base is 16-byte aligned -- and is
// rounded up to 16 bytes, which may take one element past the row's data: that is the row's own padding element for
// REAL = 1 rows that are N + 1 apart, the next row's first element otherwise, and past the buffer only for the last
// row of the batch, whose last element is fetched by an ordinary load instead (the arrive releases it to the waiters).
template <typename T, int N, int SN, int REAL>
__device__ __forceinline__ void real_issue_rows(cx<T>* Sd, const cx<T>* gin, long long k, int rows, long long idist,
                                                long long batch, uint64_t* bar) {
  constexpr int LEN = REAL == 2 ? N + 1 : N;  // elements of a row
  uint32_t total = 0;
  for (int r = 0; r < rows; ++r) {
    const cx<T>* src = gin + (k + r) * idist;
    const int shift = (int)((reinterpret_cast<uintptr_t>(src) & 15) / sizeof(cx<T>));
    uint32_t bytes = (uint32_t)(((shift + LEN) * sizeof(cx<T>) + 15) & ~(size_t)15);
    if (k + r == batch - 1 && bytes > (shift + LEN) * sizeof(cx<T>)) {
      bytes -= 16;
      Sd[r * SN + shift + LEN - 1] = src[LEN - 1];
    }
    total += bytes;
  }
  mbar_expect_tx(bar, total);
  for (int r = 0; r < rows; ++r) {
    const cx<T>* src = gin + (k + r) * idist;
    const int shift = (int)((reinterpret_cast<uintptr_t>(src) & 15) / sizeof(cx<T>));
    uint32_t bytes = (uint32_t)(((shift + LEN) * sizeof(cx<T>) + 15) & ~(size_t)15);
    if (k + r == batch - 1 && bytes > (shift + LEN) * sizeof(cx<T>)) bytes -= 16;
    if (bytes) bulk_g2s(Sd + r * SN, src - shift, bytes, bar);
  }
}

// position of element e of the exchange-2 layout inside a stage buffer (a permutation within aligned 16-groups, so
// that the pass-3 reads of 16 consecutive elements stay conflict free)
template <int R>
__device__ __forceinline__ constexpr int sw2(int e) {
  return R == 8 ? (e ^ ((e >> 3) & 8)) : e;
}

// EXE: the second exchange goes through the exchange buffer E (one more barrier between pass 2's reads and writes)
// instead of back into the consumed stage, which is then free right after pass 1: its refill is issued a whole tile
// earlier, i.e. both stages are always loaded or loading.  For the 64 KiB tiles of fp64 N = 4096 (one CTA of 8 warps
// per SM: a tile lasts 3.5 us, a refill issued two thirds of a tile ahead arrives late) -- see profiles/r2_ab_variants.txt.
template <typename T, int R, int F, bool SWAP, bool USE_TMA, int REAL = 0, bool EXE = false>
__global__ void __launch_bounds__(R* R* F, F == 1 ? (sizeof(T) == 8 ? 1 : 2) : 4) wg_cube_kernel(const CubeArgs a) {
  static_assert(REAL == 0 || (USE_TMA && !SWAP), "the REAL-domain forms are TMA fed and never swap");
  static_assert(!EXE || (USE_TMA && REAL == 0 && R == 16), "EXE: complex TMA form of N = 16^3");
  constexpr int NT = R * R;  // threads per transform
  constexpr int N = R * R * R;
  constexpr int EN = N + 2 * (N / 16);
  constexpr int SN = REAL != 0 ? N + 2 : N;  // elements of one transform's slot in a stage (REAL: shift + N (+ 1))
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T>* S0 = reinterpret_cast<cx<T>*>(smem_raw);  // TMA: stage 0 ; non-TMA: exchange-2 buffer
  cx<T>* S1 = S0 + F * SN;                         // TMA: stage 1
  cx<T>* E = USE_TMA ? (S1 + F * SN) : (S0 + F * N);  // padded exchange-1 buffer
  uint64_t* full = reinterpret_cast<uint64_t*>(E + F * EN);
  const int f = F == 1 ? 0 : threadIdx.x / NT;  // transform of the tile
  const int t = F == 1 ? threadIdx.x : threadIdx.x % NT;
  const int k2 = t % R;
  const cx<T>* gin = reinterpret_cast<const cx<T>*>(a.in) + a.ioff;
  cx<T>* gout = reinterpret_cast<cx<T>*>(a.out) + a.ooff;
  const long long stride = (long long)gridDim.x * F;
  const bool contig = a.idist == N;  // the F transforms of a tile are one contiguous run

  // per-thread twiddles, resident in registers for the whole batch loop
  cx<T> tw2[R - 1], tw3[R - 1];
#pragma unroll
  for (int r = 1; r < R; ++r) {
    tw2[r - 1] = ldg_cx<T>(a.tw, (long long)k2 * r * R);  // w_{R^2}^{k2 r}
    tw3[r - 1] = ldg_cx<T>(a.tw, (long long)t * r);       // w_N^{t r}
  }
  const T scale = T(a.scale);
  cx<T> wreal{T(1), T(0)};  // REAL: w_{2N}^t
  if (REAL != 0) wreal = ldg_cx<T>(a.tw2, t);

  // one thread: fetch the tile starting at transform k into stage buffer Sd
  auto issue = [&](long long k, cx<T>* Sd, uint64_t* bar) {
    const int rows = (int)min((long long)F, a.batch - k);
    if constexpr (REAL != 0) {
      real_issue_rows<T, N, SN, REAL>(Sd, gin, k, rows, a.idist, a.batch, bar);
      return;
    }
    mbar_expect_tx(bar, (uint32_t)(rows * N * sizeof(cx<T>)));
    if (F == 1 || contig) {
      bulk_g2s(Sd, gin + k * a.idist, (uint32_t)(rows * N * sizeof(cx<T>)), bar);
    } else {
      for (int r = 0; r < rows; ++r) bulk_g2s(Sd + r * N, gin + (k + r) * a.idist, (uint32_t)(N * sizeof(cx<T>)), bar);
    }
  };

  if (USE_TMA) {
    if (threadIdx.x == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      fence_mbar_init();
      fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long k = (long long)blockIdx.x * F;
      if (k < a.batch) issue(k, S0, &full[0]);
      k += stride;
      if (k < a.batch) issue(k, S1, &full[1]);
    }
  }

  int it = 0;
  for (long long k0 = (long long)blockIdx.x * F; k0 < a.batch; k0 += stride, ++it) {
    const long long k = k0 + f;
    const bool live = F == 1 || k < a.batch;  // ragged last tile: idle threads still take part in the barriers
    cx<T>* S = (USE_TMA ? ((it & 1) ? S1 : S0) : S0) + f * SN;
    cx<T>* Ef = E + f * EN;
    cx<T> v[R];
    // ---- pass 1: x[t + NT r] -> radix R -> E[R t + r'] -------------------------------------------------------
    if (USE_TMA) {
      mbar_wait(&full[it & 1], (it >> 1) & 1);
      if (live) {
        if constexpr (REAL == 2) {
          // input j = t + NT r of the transform is built from X_{N-j} and X_j of the row (see c2r_combine)
          const cx<T>* X = S + (int)((reinterpret_cast<uintptr_t>(gin + k * a.idist) & 15) / sizeof(cx<T>));
          static_for<0, R>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const int j = t + NT * r;
            const bool first = r == 0 && t == 0;
            v[r] = c2r_combine(X[first ? 0 : N - j], X[first ? N : j], mul_w<r, 2 * R, T>(wreal), first);
          });
        } else if constexpr (REAL == 1) {
          const cx<T>* X = S + (int)((reinterpret_cast<uintptr_t>(gin + k * a.idist) & 15) / sizeof(cx<T>));
#pragma unroll
          for (int r = 0; r < R; ++r) v[r] = X[t + NT * r];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) v[r] = S[t + NT * r];
        }
      }
    } else if (live) {
      const cx<T>* src = gin + k * a.idist;
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = src[t + NT * r];
    }
    if (live) {
      if (SWAP) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const T tmp = v[r].x;
          v[r].x = v[r].y;
          v[r].y = tmp;
        }
      }
      DFT<R, T>::run(v);
      cx<T>* dst = Ef + (R * t + 2 * ((R * t) / 16));  // R*t..R*t+R-1 lie inside one 16-group (R divides 16)
      if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int r = 0; r < R; r += 2)
          *reinterpret_cast<float4*>(dst + r) = make_float4(v[r].x, v[r].y, v[r + 1].x, v[r + 1].y);
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r] = v[r];
      }
    }
    __syncthreads();
    if (USE_TMA && threadIdx.x == 0 && (EXE || it >= 1)) {
      // every thread has finished reading the stage used by iteration it-1 (its pass 3 precedes this barrier);
      // EXE: ... and the stage of THIS iteration (nothing is written back into it): refill it with the tile after next
      const long long kn = k0 + (EXE ? 2 : 1) * stride;
      if (kn < a.batch) {
        fence_proxy_async();
        if (EXE)
          issue(kn, (it & 1) ? S1 : S0, &full[it & 1]);
        else
          issue(kn, (it & 1) ? S0 : S1, &full[(it + 1) & 1]);
      }
    }
    if (live) {
      // ---- pass 2: E[t + NT r] * w_{R^2}^{k2 r} -> radix R -> S[(t - k2) R + k2 + R r'] -------------------------
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = Ef[epad<R>(t + NT * r)];
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw2[r - 1]);
      DFT<R, T>::run(v);
    }
    if (EXE) __syncthreads();  // every thread holds its inputs: E takes the outputs
    if (live) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (EXE)
          Ef[epad<R>((t - k2) * R + k2 + R * r)] = v[r];
        else
          S[sw2<R>((t - k2) * R + k2 + R * r)] = v[r];
      }
    }
    __syncthreads();
    if (live) {
      // ---- pass 3: S[t + NT r] * w_N^{t r} -> radix R -> out[t + NT r'] ----------------------------------------
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = EXE ? Ef[epad<R>(t + NT * r)] : S[sw2<R>(t + NT * r)];
    }
    if (EXE) __syncthreads();  // E is read: the next tile's pass 1 may write it
    if (live) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw3[r - 1]);
      DFT<R, T>::run(v);
      if constexpr (REAL != 1) {
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
      } else {
        // Z back into the stage, in natural order (the positions this thread has just read)
#pragma unroll
        for (int r = 0; r < R; ++r) S[sw2<R>(t + NT * r)] = v[r];
      }
    }
    if constexpr (REAL == 1) {
      __syncthreads();
      if (live) {
        cx<T>* dst = gout + k * a.odist;
        static_for<0, R>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          const int kk = t + NT * r;
          const cx<T> b = S[sw2<R>(kk == 0 ? 0 : N - kk)];
          cx<T> o = r2c_combine(v[r], b, mul_w<r, 2 * R, T>(wreal));
          if (a.apply_scale) o = cscale(o, scale);
          dst[kk] = o;
        });
        if (t == 0) {
          cx<T> o{v[0].x - v[0].y, T(0)};  // X_N = Re Z_0 - Im Z_0
          if (a.apply_scale) o = cscale(o, scale);
          dst[N] = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Same scheme for the packed powers of two that are not cubes: N = R0 * R1 * R2 with R0 = 16 (1024 = 16*8*8, four
// transforms per tile; 2048 = 16*16*8, two per tile; 8192 = 16*16*32, one per tile).  N / 16 threads per transform:
// one radix-16 butterfly per thread in pass 1, N / (R1 NT) and N / (R2 NT) butterflies per thread in passes 2 and 3
// (twiddles from the L1-resident table instead of registers).  TMA ring, padded exchange 1, exchange 2 written back
// into the consumed stage, coalesced stores from registers -- as above.
// ---------------------------------------------------------------------------------------------------------------
//
// B1 = radix-16 butterflies per thread in pass 1 (threads per transform NT = N / 16 / B1).  8192 = 16*16*32 runs with
// B1 = 2: 256 threads, two radix-16 butterflies each in passes 1 and 2 and ONE radix-32 butterfly in pass 3 whose 31
// twiddles stay in registers (with B1 = 1 half of the 512 threads idle through pass 3 and every radix-32 butterfly
// fetches its twiddles from the table: 45 % of the HBM roofline).
template <int R0, int R1, int R2, int F, int B1, typename T>
constexpr int rows3_min_ctas() {
  constexpr size_t n = (size_t)R0 * R1 * R2;
  return (2 * n + n + 2 * (n / 16)) * F * 2 * sizeof(T) + 64 > 113 * 1024 ? 1 : 2;
}

// EXE: as in wg_cube_kernel -- the second exchange goes through E (all of pass 2's reads, a barrier, then its writes;
// another barrier after pass 3's reads) and the stage is refilled right after pass 1, a whole tile earlier.
template <typename T, int R0, int R1, int R2, int F, bool SWAP, int B1 = 1, int REAL = 0, bool EXE = false>
__global__ void __launch_bounds__((R1 * R2 / B1) * F, ((R1 * R2 / B1) * F > 256 ? 1 : rows3_min_ctas<R0, R1, R2, F, B1, T>()))
    wg_rows3_kernel(const CubeArgs a) {
  static_assert(R0 == 16, "pass 1 writes whole padded 16-groups");
  static_assert(!EXE || REAL == 0, "EXE: complex form");
  static_assert(REAL == 0 || !SWAP, "the REAL-domain forms never swap");
  constexpr int N = R0 * R1 * R2;
  constexpr int NT = N / R0 / B1;  // threads per transform
  constexpr int EN = N + 2 * (N / 16);
  constexpr int SN = REAL != 0 ? N + 2 : N;  // elements of one transform's slot in a stage (REAL: shift + N (+ 1))
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T>* S0 = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* S1 = S0 + F * SN;
  cx<T>* E = S1 + F * SN;
  uint64_t* full = reinterpret_cast<uint64_t*>(E + F * EN);
  const int f = F == 1 ? 0 : threadIdx.x / NT;
  const int t = F == 1 ? threadIdx.x : threadIdx.x % NT;
  const cx<T>* gin = reinterpret_cast<const cx<T>*>(a.in) + a.ioff;
  cx<T>* gout = reinterpret_cast<cx<T>*>(a.out) + a.ooff;
  const long long stride = (long long)gridDim.x * F;
  const bool contig = a.idist == N;
  const T scale = T(a.scale);
  // twiddles of the thread's own butterflies, resident in registers for the whole batch loop when they are few
  // (measured: table look-ups inside the passes cost 2048-point rows a third of their bandwidth)
  constexpr int B2 = (N / R1 + NT - 1) / NT, B3 = (N / R2 + NT - 1) / NT;  // butterflies per thread in passes 2, 3
  // pass-2 twiddles depend on j % R0 only: one set serves all of a thread's butterflies when R0 divides NT
  constexpr bool TW2SHARED = NT % R0 == 0;
  constexpr int S2 = TW2SHARED ? 0 : (R1 - 1);  // register offset between the twiddle sets of consecutive butterflies
  constexpr int NTW2 = (TW2SHARED ? 1 : B2) * (R1 - 1), NTW3 = B3 * (R2 - 1);
  constexpr int TWMAX = B1 > 1 ? 31 : 16;  // (B1 > 1: 256 threads on a whole SM, 255 registers each)
  constexpr bool TW2REG = NTW2 <= TWMAX, TW3REG = NTW3 <= TWMAX;
  cx<T> tw2[TW2REG ? NTW2 : 1], tw3[TW3REG ? NTW3 : 1];
  if (TW2REG) {
#pragma unroll
    for (int i = 0; i < (TW2SHARED ? 1 : B2); ++i)
#pragma unroll
      for (int r = 1; r < R1; ++r) tw2[i * (R1 - 1) + r - 1] = ldg_cx<T>(a.tw, ((t + i * NT) % R0) * r * R2);
  }
  if (TW3REG) {
#pragma unroll
    for (int i = 0; i < B3; ++i)
#pragma unroll
      for (int r = 1; r < R2; ++r) tw3[i * (R2 - 1) + r - 1] = ldg_cx<T>(a.tw, ((t + i * NT) % (N / R2)) * r);
  }

  // REAL: w_{2N}^{t + i NT}, the per-thread factor of the twiddles of passes 1 (C2R) and 3 (R2C)
  constexpr int BR = REAL == 0 ? 1 : (B1 > B3 ? B1 : B3);
  cx<T> wreal[BR];
  if (REAL != 0) {
#pragma unroll
    for (int i = 0; i < BR; ++i) wreal[i] = ldg_cx<T>(a.tw2, t + i * NT);
  }

  auto issue = [&](long long k, cx<T>* Sd, uint64_t* bar) {
    const int rows = (int)min((long long)F, a.batch - k);
    if constexpr (REAL != 0) {
      real_issue_rows<T, N, SN, REAL>(Sd, gin, k, rows, a.idist, a.batch, bar);
      return;
    }
    mbar_expect_tx(bar, (uint32_t)(rows * N * sizeof(cx<T>)));
    if (F == 1 || contig) {
      bulk_g2s(Sd, gin + k * a.idist, (uint32_t)(rows * N * sizeof(cx<T>)), bar);
    } else {
      for (int r = 0; r < rows; ++r) bulk_g2s(Sd + r * N, gin + (k + r) * a.idist, (uint32_t)(N * sizeof(cx<T>)), bar);
    }
  };

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long k = (long long)blockIdx.x * F;
    if (k < a.batch) issue(k, S0, &full[0]);
    k += stride;
    if (k < a.batch) issue(k, S1, &full[1]);
  }

  int it = 0;
  for (long long k0 = (long long)blockIdx.x * F; k0 < a.batch; k0 += stride, ++it) {
    const long long k = k0 + f;
    const bool live = F == 1 || k < a.batch;
    cx<T>* S = ((it & 1) ? S1 : S0) + f * SN;
    cx<T>* Ef = E + f * EN;
    // ---- pass 1: x[t + NT r] -> radix 16 -> E[16 t + r'] -------------------------------------------------------
    mbar_wait(&full[it & 1], (it >> 1) & 1);
    if (live) {
#pragma unroll
      for (int i = 0; i < B1; ++i) {
        const int j = t + i * NT;
        cx<T> v[R0];
        if constexpr (REAL == 2) {
          const cx<T>* X = S + (int)((reinterpret_cast<uintptr_t>(gin + k * a.idist) & 15) / sizeof(cx<T>));
          static_for<0, R0>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const int jj = j + (N / R0) * r;
            const bool first = r == 0 && j == 0;
            v[r] = c2r_combine(X[first ? 0 : N - jj], X[first ? N : jj], mul_w<r, 2 * R0, T>(wreal[i]), first);
          });
        } else if constexpr (REAL == 1) {
          const cx<T>* X = S + (int)((reinterpret_cast<uintptr_t>(gin + k * a.idist) & 15) / sizeof(cx<T>));
#pragma unroll
          for (int r = 0; r < R0; ++r) v[r] = X[j + (N / R0) * r];
        } else {
#pragma unroll
          for (int r = 0; r < R0; ++r) v[r] = S[j + (N / R0) * r];
        }
        if (SWAP) {
#pragma unroll
          for (int r = 0; r < R0; ++r) v[r] = cx<T>{v[r].y, v[r].x};
        }
        DFT<R0, T>::run(v);
        cx<T>* dst = Ef + 18 * j;  // epad(16 j + r) = 18 j + r
        if constexpr (sizeof(T) == 4) {
#pragma unroll
          for (int r = 0; r < R0; r += 2)
            *reinterpret_cast<float4*>(dst + r) = make_float4(v[r].x, v[r].y, v[r + 1].x, v[r + 1].y);
        } else {
#pragma unroll
          for (int r = 0; r < R0; ++r) dst[r] = v[r];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && (EXE || it >= 1)) {
      const long long kn = k0 + (EXE ? 2 : 1) * stride;
      if (kn < a.batch) {
        fence_proxy_async();
        if (EXE)
          issue(kn, (it & 1) ? S1 : S0, &full[it & 1]);
        else
          issue(kn, (it & 1) ? S0 : S1, &full[(it + 1) & 1]);
      }
    }
    if constexpr (EXE) {
      // ---- pass 2 through E: all butterflies of the thread read, then (barrier) written ---------------------------
      cx<T> vv[B2][R1];
      if (live) {
#pragma unroll
        for (int i = 0; i < B2; ++i) {
          const int j = t + i * NT;
          if (j >= N / R1) break;
          const int k1 = j % R0;
#pragma unroll
          for (int r = 0; r < R1; ++r) vv[i][r] = Ef[epad<R0>(j + (N / R1) * r)];
#pragma unroll
          for (int r = 1; r < R1; ++r)
            vv[i][r] = cmul(vv[i][r], TW2REG ? tw2[TW2REG ? i * S2 + r - 1 : 0] : ldg_cx<T>(a.tw, k1 * r * R2));
          DFT<R1, T>::run(vv[i]);
        }
      }
      __syncthreads();
      if (live) {
#pragma unroll
        for (int i = 0; i < B2; ++i) {
          const int j = t + i * NT;
          if (j >= N / R1) break;
          const int k1 = j % R0;
#pragma unroll
          for (int r = 0; r < R1; ++r) Ef[epad<R0>((j - k1) * R1 + k1 + R0 * r)] = vv[i][r];
        }
      }
    } else if (live) {
      // ---- pass 2: E[j + (N/R1) r] * w_{R0 R1}^{k1 r} -> radix R1 -> S[(j - k1) R1 + k1 + R0 r'] ------------------
#pragma unroll
      for (int i = 0; i < B2; ++i) {
        const int j = t + i * NT;
        if (j >= N / R1) break;
        const int k1 = j % R0;
        cx<T> v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = Ef[epad<R0>(j + (N / R1) * r)];
#pragma unroll
        for (int r = 1; r < R1; ++r)
          v[r] = cmul(v[r], TW2REG ? tw2[TW2REG ? i * S2 + r - 1 : 0] : ldg_cx<T>(a.tw, k1 * r * R2));
        DFT<R1, T>::run(v);
        cx<T>* dst = S + (j - k1) * R1 + k1;
#pragma unroll
        for (int r = 0; r < R1; ++r) dst[R0 * r] = v[r];
      }
    }
    __syncthreads();
    if constexpr (EXE) {
      // ---- pass 3 from E (one butterfly per thread): read, barrier (E is free for the next tile), compute, store ----
      static_assert(!EXE || B3 == 1, "EXE: one pass-3 butterfly per thread");
      cx<T> v[R2];
      const int j = t;
      if (live) {
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = Ef[epad<R0>(j + (N / R2) * r)];
      }
      __syncthreads();
      if (live) {
        cx<T>* dst = gout + k * a.odist;
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], TW3REG ? tw3[TW3REG ? r - 1 : 0] : ldg_cx<T>(a.tw, j * r));
        DFT<R2, T>::run(v);
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          cx<T> o = v[r];
          if (a.apply_scale) o = cscale(o, scale);
          if (SWAP) o = cx<T>{o.y, o.x};
          dst[j + (N / R2) * r] = o;
        }
      }
    } else if (live) {
      // ---- pass 3: S[j + (N/R2) r] * w_N^{j r} -> radix R2 -> out[j + (N/R2) r'] -----------------------------------
      cx<T>* dst = gout + k * a.odist;
#pragma unroll
      for (int i = 0; i < B3; ++i) {
        const int j = t + i * NT;
        if (j >= N / R2) break;
        cx<T> v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = S[j + (N / R2) * r];
#pragma unroll
        for (int r = 1; r < R2; ++r)
          v[r] = cmul(v[r], TW3REG ? tw3[TW3REG ? i * (R2 - 1) + r - 1 : 0] : ldg_cx<T>(a.tw, j * r));
        DFT<R2, T>::run(v);
        if constexpr (REAL == 1) {
          // Z back into the stage, in natural order (the positions this thread has just read)
#pragma unroll
          for (int r = 0; r < R2; ++r) S[j + (N / R2) * r] = v[r];
        } else {
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            cx<T> o = v[r];
            if (a.apply_scale) o = cscale(o, scale);
            if (SWAP) o = cx<T>{o.y, o.x};
            dst[j + (N / R2) * r] = o;
          }
        }
      }
    }
    if constexpr (REAL == 1) {
      __syncthreads();
      if (live) {
        cx<T>* dst = gout + k * a.odist;
#pragma unroll
        for (int i = 0; i < B3; ++i) {
          const int j = t + i * NT;
          if (j >= N / R2) break;
          static_for<0, R2>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const int kk = j + (N / R2) * r;
            const cx<T> b = S[kk == 0 ? 0 : N - kk];
            cx<T> o = r2c_combine(S[kk], b, mul_w<r, 2 * R2, T>(wreal[i]));
            if (a.apply_scale) o = cscale(o, scale);
            dst[kk] = o;
          });
          if (j == 0) {
            const cx<T> z0 = S[0];
            cx<T> o{z0.x - z0.y, T(0)};  // X_N = Re Z_0 - Im Z_0
            if (a.apply_scale) o = cscale(o, scale);
            dst[N] = o;
          }
        }
      }
    }
  }
}

template <typename T, int R0, int R1, int R2, int F, int B1, bool SWAP, int REAL, bool EXE = false>
static cudaError_t launch_rows3_k(const CubeArgs& a, int grid, cudaStream_t stream) {
  constexpr int N = R0 * R1 * R2;
  constexpr int SN = REAL != 0 ? N + 2 : N;
  constexpr size_t smem = (2 * (size_t)SN + (N + 2 * (N / 16))) * F * sizeof(cx<T>) + 64;
  auto kern = wg_rows3_kernel<T, R0, R1, R2, F, SWAP, B1, REAL, EXE>;
  const cudaError_t e = ensure_dynamic_smem(kern, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, (N / R0 / B1) * F, smem, stream>>>(a);
  return cudaGetLastError();
}

// real: 0 complex, 1 real-to-complex epilogue, 2 complex-to-real prologue (never with swap)
template <typename T, int R0, int R1, int R2, int F, int B1 = 1>
static cudaError_t launch_rows3(const CubeArgs& a, bool swap, int real, int grid, cudaStream_t stream) {
  if (real == 1) return launch_rows3_k<T, R0, R1, R2, F, B1, false, 1>(a, grid, stream);
  if (real == 2) return launch_rows3_k<T, R0, R1, R2, F, B1, false, 2>(a, grid, stream);
  return swap ? launch_rows3_k<T, R0, R1, R2, F, B1, true, 0>(a, grid, stream)
              : launch_rows3_k<T, R0, R1, R2, F, B1, false, 0>(a, grid, stream);
}

template <typename T, int R, int F>
size_t cube_smem_bytes_t(bool use_tma, int real = 0) {
  constexpr int N = R * R * R;
  constexpr int EN = N + 2 * (N / 16);
  return ((use_tma ? 2 : 1) * (size_t)(real != 0 ? N + 2 : N) + EN) * F * sizeof(cx<T>) + 64;
}

template <typename T, int R, int F, bool SWAP, bool USE_TMA, int REAL = 0, bool EXE = false>
static cudaError_t launch_cube_t(const CubeArgs& a, int grid, cudaStream_t stream) {
  const size_t smem = cube_smem_bytes_t<T, R, F>(USE_TMA, REAL);
  auto kern = wg_cube_kernel<T, R, F, SWAP, USE_TMA, REAL, EXE>;
  cudaError_t e = ensure_dynamic_smem(kern, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, R * R * F, smem, stream>>>(a);
  return cudaGetLastError();
}

template <typename T, int R, int F>
static cudaError_t launch_cube_v(const CubeArgs& a, bool swap, bool tma, int real, int grid, cudaStream_t stream) {
  if (real == 1) return launch_cube_t<T, R, F, false, true, 1>(a, grid, stream);
  if (real == 2) return launch_cube_t<T, R, F, false, true, 2>(a, grid, stream);
  if constexpr (R == 16 && sizeof(T) == 8) {
    static const bool exe = [] {  // PFFT_CUBE_EXE=0: exchange 2 back into the stage (the fp32 scheme)
      const char* e = std::getenv("PFFT_CUBE_EXE");
      return e ? std::atoi(e) != 0 : true;
    }();
    if (tma && exe)
      return swap ? launch_cube_t<T, R, F, true, true, 0, true>(a, grid, stream)
                  : launch_cube_t<T, R, F, false, true, 0, true>(a, grid, stream);
  }
  if (tma) return swap ? launch_cube_t<T, R, F, true, true>(a, grid, stream) : launch_cube_t<T, R, F, false, true>(a, grid, stream);
  return swap ? launch_cube_t<T, R, F, true, false>(a, grid, stream) : launch_cube_t<T, R, F, false, false>(a, grid, stream);
}

bool cube_supported(int n, bool is_double, int* transforms_per_tile, int* ctas_per_sm) {
  int f = 0, c = 0;
  if (is_double) {
    // same kernels, half the transforms per tile (the tile bytes stay the same); 4096 = 16^3 fills one SM with one
    // CTA (two 64 KiB stages + the padded exchange buffer)
    if (n == 4096) f = 1, c = 1;
    if (n == 512) f = kCube512Tile / 2, c = 4;
    if (n == 1024) f = 2, c = 2;
    if (n == 2048) f = 1, c = 2;
  } else {
    if (n == 4096) f = 1, c = 2;
    if (n == 512) f = kCube512Tile, c = 4;
    if (n == 1024) f = 4, c = 2;  // wg_rows3_kernel 16 x 8 x 8
    if (n == 2048) f = 2, c = 2;  // wg_rows3_kernel 16 x 16 x 8
    if (n == 8192) f = 1, c = 1;  // wg_rows3_kernel 16 x 16 x 32
  }
  if (f == 0) return false;
  if (transforms_per_tile) *transforms_per_tile = f;
  if (ctas_per_sm) *ctas_per_sm = c;
  return true;
}

// Half lengths that only the REAL-domain forms run here (their complex transforms belong to the tile kernel of
// wg_col.cu): 256 = 16 * 4 * 4, 128 = 16 * 4 * 2, 64 = 16 * 2 * 2; 16 KiB tiles of 8 / 16 / 32 (fp32) or 4 / 8 / 16
// (fp64) transforms -- the rows of real lengths 512, 256 and 128 (e.g. of 3-D real transforms of those edge lengths).
bool cube_real_supported(int n, bool is_double, int* transforms_per_tile, int* ctas_per_sm) {
  if (n != 256 && n != 128 && n != 64) return cube_supported(n, is_double, transforms_per_tile, ctas_per_sm);
  if (transforms_per_tile) *transforms_per_tile = (is_double ? 1024 : 2048) / n;
  if (ctas_per_sm) *ctas_per_sm = 4;
  return true;
}

// p: a single-pass plan entry with n == R^3, interleaved storage, unit strides, one batch dimension.
// real: 0 complex transform; 1 / 2: the REAL-domain forms (p.tw2 = w_{2n}^k; never with swap, TMA variant only)
cudaError_t launch_wg_cube(const PassParams& p, bool is_double, bool swap, int variant, int grid, cudaStream_t stream,
                           int real) {
  CubeArgs a;
  a.in = p.in_re;
  a.out = p.out_re;
  a.ioff = p.ioff;
  a.ooff = p.ooff;
  a.idist = p.ibd[0];
  a.odist = p.obd[0];
  a.batch = p.batch_total;
  a.tw = p.tw;
  a.tw2 = p.tw2;
  a.scale = p.scale;
  a.apply_scale = p.apply_scale;
  const bool tma = variant == 0;
  if (real != 0 && (!tma || swap || p.tw2 == nullptr)) return cudaErrorInvalidValue;
  if (real != 0 && p.n == 256)
    return is_double ? launch_rows3<double, 16, 4, 4, 4>(a, false, real, grid, stream)
                     : launch_rows3<float, 16, 4, 4, 8>(a, false, real, grid, stream);
  if (real != 0 && p.n == 128)
    return is_double ? launch_rows3<double, 16, 4, 2, 8>(a, false, real, grid, stream)
                     : launch_rows3<float, 16, 4, 2, 16>(a, false, real, grid, stream);
  if (real != 0 && p.n == 64)
    return is_double ? launch_rows3<double, 16, 2, 2, 16>(a, false, real, grid, stream)
                     : launch_rows3<float, 16, 2, 2, 32>(a, false, real, grid, stream);
  if (is_double) {
    if (!tma) return cudaErrorInvalidValue;
    if (p.n == 4096) return launch_cube_v<double, 16, 1>(a, swap, true, real, grid, stream);
    if (p.n == 512) return launch_cube_v<double, 8, kCube512Tile / 2>(a, swap, true, real, grid, stream);
    if (p.n == 1024) return launch_rows3<double, 16, 8, 8, 2>(a, swap, real, grid, stream);
    if (p.n == 2048) return launch_rows3<double, 16, 16, 8, 1>(a, swap, real, grid, stream);
    return cudaErrorInvalidValue;
  }
  if (p.n == 4096) return launch_cube_v<float, 16, 1>(a, swap, tma, real, grid, stream);
  if (p.n == 512) return launch_cube_v<float, 8, kCube512Tile>(a, swap, tma, real, grid, stream);
  if (tma && p.n == 1024) return launch_rows3<float, 16, 8, 8, 4>(a, swap, real, grid, stream);
  if (tma && p.n == 2048) return launch_rows3<float, 16, 16, 8, 2>(a, swap, real, grid, stream);
  if (tma && p.n == 8192) {
    static const bool wide = [] {  // A/B knob: the 512-thread form (one radix-16 butterfly per thread in pass 1)
      const char* e = std::getenv("PFFT_ROWS8192_WIDE");
      return e && std::atoi(e) != 0;
    }();
    static const bool exe = [] {  // PFFT_CUBE_EXE=0: exchange 2 back into the stage
      const char* e = std::getenv("PFFT_CUBE_EXE");
      return e ? std::atoi(e) != 0 : true;
    }();
    if (real == 0 && !wide && exe)
      return swap ? launch_rows3_k<float, 16, 16, 32, 1, 2, true, 0, true>(a, grid, stream)
                  : launch_rows3_k<float, 16, 16, 32, 1, 2, false, 0, true>(a, grid, stream);
    return wide && real == 0 ? launch_rows3<float, 16, 16, 32, 1>(a, swap, 0, grid, stream)
                             : launch_rows3<float, 16, 16, 32, 1, 2>(a, swap, real, grid, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace pfft
