// WORKITEM level, TMA in and out: packed, contiguous, interleaved power-of-two transforms of at most one 128-byte line
// (fp32 N = 2, 4, 8, 16 -- N = 16 is the reference's `small_1d` benchmark,
// /root/reference/test/bench/utils/reference_dft_set.hpp:91 -- and fp64 N = 2, 4, 8).  One thread owns one 128-byte
// line = 128 / (N * sizeof(complex)) whole transforms.
//
// The general thread-per-transform kernel (wi.cuh) stages its tile with 8-byte cp.async copies and per-thread
// global stores through a padded buffer; this variant lets the TMA engine move whole tiles in both directions:
//   * a tile of 128 lines (16 KiB) arrives by ONE cp.async.bulk.tensor.2d load with the 128-byte swizzle, so that
//     thread t finds 16-byte chunk c of its line at chunk position c ^ (t & 7): the 8 LDS.128 of a quarter-warp hit
//     8 different bank groups although every thread reads its own line (no padding, static register indices);
//   * the transforms run in registers (dft.cuh) and are written back over the line they came from;
//   * the tile leaves by ONE cp.async.bulk.tensor.2d store from the same stage (rows beyond the batch are clipped by
//     the tensor map), three stages per CTA: load(i+2) / compute(i) / store(i-1) overlap without any per-thread
//     global access.
#include <cuda.h>

#include <cstdint>
#include <cstring>

#include "device_utils.cuh"
#include "kernels.h"
#include "launch_utils.h"
#include "real_fuse.cuh"

namespace pfft {

namespace wt {
constexpr int kRows = 128, kStages = 3, kStageBytes = kRows * 128;
constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 + 64;  // + alignment slack + mbarriers

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int r0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int r0, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(r0), "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int Pending>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(Pending) : "memory");
}
}  // namespace wt

template <typename T, int N>
__global__ void __launch_bounds__(wt::kRows) wi_tma_kernel(const __grid_constant__ CUtensorMap in_map,
                                                           const __grid_constant__ CUtensorMap out_map,
                                                           const long long batch, const bool swap, const int apply_scale,
                                                           const T scale) {
  // `batch` counts 128-byte lines; every line holds PER whole transforms
  constexpr int PER = 128 / (N * 2 * (int)sizeof(T)), NV = 128 / (2 * (int)sizeof(T));
  static_assert(PER >= 1 && PER * N * 2 * sizeof(T) == 128, "whole transforms per 128-byte line");
  extern __shared__ unsigned char smem_dyn[];
  // the 128-byte swizzle pattern repeats every 1024 bytes: stages start on 1024-byte boundaries
  unsigned char* base = smem_dyn + ((1024 - (wt::smem_u32(smem_dyn) & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + wt::kStages * wt::kStageBytes);
  const int t = threadIdx.x;
  const long long tiles = (batch + wt::kRows - 1) / wt::kRows;

  if (t == 0) {
    for (int s = 0; s < wt::kStages; ++s) wt::mbar_init(&full[s], 1);
    wt::fence_mbar_init();
    wt::fence_proxy_async();
  }
  __syncthreads();
  if (t == 0) {
    long long tile = blockIdx.x;
    for (int s = 0; s < 2; ++s, tile += gridDim.x)
      if (tile < tiles) {
        wt::mbar_expect_tx(&full[s], wt::kStageBytes);
        wt::tma_load_2d(base + s * wt::kStageBytes, &in_map, 0, (int)(tile * wt::kRows), &full[s]);
      }
  }
  int it = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int st = it % wt::kStages;
    unsigned char* row = base + st * wt::kStageBytes + t * 128;
    wt::mbar_wait(&full[st], (uint32_t)((it / wt::kStages) & 1));
    cx<T> v[NV];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const T* src = reinterpret_cast<const T*>(row + ((c ^ (t & 7)) << 4));
      if constexpr (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(src);
        v[2 * c] = cx<T>{q.x, q.y};
        v[2 * c + 1] = cx<T>{q.z, q.w};
      } else {
        const double2 q = *reinterpret_cast<const double2*>(src);
        v[c] = cx<T>{q.x, q.y};
      }
    }
    if (swap) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = cx<T>{v[j].y, v[j].x};
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) DFT<N, T>::run(v + i * N);
    if (apply_scale) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = cscale(v[j], scale);
    }
    if (swap) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = cx<T>{v[j].y, v[j].x};
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      T* dst = reinterpret_cast<T*>(row + ((c ^ (t & 7)) << 4));
      if constexpr (sizeof(T) == 4)
        *reinterpret_cast<float4*>(dst) = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
      else
        *reinterpret_cast<double2*>(dst) = make_double2(v[c].x, v[c].y);
    }
    wt::fence_proxy_async();  // the rows written above must be visible to the TMA store
    __syncthreads();
    if (t == 0) {
      wt::tma_store_2d(&out_map, 0, (int)(tile * wt::kRows), base + st * wt::kStageBytes);
      wt::bulk_commit();
      // the store issued one iteration ago has finished reading its stage: refill it with the tile two ahead
      wt::bulk_wait_read<1>();
      const long long nxt = tile + 2LL * gridDim.x;
      if (nxt < tiles) {
        const int sn = (it + 2) % wt::kStages;
        wt::mbar_expect_tx(&full[sn], wt::kStageBytes);
        wt::tma_load_2d(base + sn * wt::kStageBytes, &in_map, 0, (int)(nxt * wt::kRows), &full[sn]);
      }
    }
  }
  if (t == 0) wt::bulk_wait_read<0>();  // shared memory must outlive the last store's reads
}

// ---------------------------------------------------------------------------------------------------------------
// REAL domain fused into the thread-level kernel: real rows of 2 N scalars = exactly one 128-byte line (fp32 real
// length 32 -- the reference's real `small_1d` benchmark, reference_dft_set.hpp -- and fp64 real length 16), dense half
// spectrum rows of N + 1 elements.  MODE 1 (real-to-complex): the line arrives by the swizzled tensor load as above,
// the thread transforms its N pairs and combines Z_k with Z_{N-k} in registers (all indices and twiddles compile
// time), writes the N + 1 outputs to a [row][N + 1] staging buffer (136-byte pitch: conflict free for 64-bit stores)
// and the whole tile -- contiguous in global memory when the half-spectrum rows are dense -- leaves by ONE
// cp.async.bulk (rows further apart: the threads copy the tile out, lanes along the row elements).  MODE 2 (complex-to-real): the
// mirror image: bulk load of the tile's half spectra, pre-processing and transform in registers, tensor store of the
// lines.  (An odd number of fp32 rows in the last tile is not a multiple of 16 bytes: its last row moves by ordinary
// loads / stores.)
// ---------------------------------------------------------------------------------------------------------------
namespace wt {
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <typename T, int N>
struct RealCfg {
  static constexpr int kSpecRow = (N + 1) * 2 * (int)sizeof(T);              // bytes of one half-spectrum row
  static constexpr int kSpecBytes = (kRows * kSpecRow + 127) / 128 * 128;    // staging buffer of one tile
  static constexpr size_t kSmem = 2 * (size_t)kStageBytes + 2 * (size_t)kSpecBytes + 1024 + 64;
};
}  // namespace wt

template <typename T, int N, int MODE>
__global__ void __launch_bounds__(wt::kRows) wi_tma_real_kernel(const __grid_constant__ CUtensorMap line_map,
                                                                cx<T>* spec, const long long spec_dist,
                                                                const long long batch, const int apply_scale, const T scale) {
  using Cfg = wt::RealCfg<T, N>;
  // dense half-spectrum rows (distance N + 1): a tile is contiguous in global memory and moves by one bulk copy;
  // otherwise the threads move it between global memory and the staging buffer, lanes along the row elements
  const bool dense = spec_dist == N + 1;
  static_assert(N * 2 * sizeof(T) == 128, "one transform per 128-byte line");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024 - (wt::smem_u32(smem_dyn) & 1023)) & 1023);  // line stages (swizzle period)
  unsigned char* sbase = base + 2 * wt::kStageBytes;                                     // half-spectrum staging
  uint64_t* full = reinterpret_cast<uint64_t*>(sbase + 2 * Cfg::kSpecBytes);
  const int t = threadIdx.x;
  const long long tiles = (batch + wt::kRows - 1) / wt::kRows;
  // rows of a tile and how many of them move by the bulk copy (a whole number of 16-byte units)
  auto tile_rows = [&](long long tile) { return (int)min((long long)wt::kRows, batch - tile * wt::kRows); };
  auto bulk_rows = [&](int rows) { return !dense || (rows * Cfg::kSpecRow) % 16 == 0 ? rows : rows - 1; };
  auto issue = [&](long long tile, int s) {
    if (MODE == 1) {
      wt::mbar_expect_tx(&full[s], wt::kStageBytes);
      wt::tma_load_2d(base + s * wt::kStageBytes, &line_map, 0, (int)(tile * wt::kRows), &full[s]);
    } else if (dense) {
      const uint32_t bytes = (uint32_t)(bulk_rows(tile_rows(tile)) * Cfg::kSpecRow);
      wt::mbar_expect_tx(&full[s], bytes);
      if (bytes) wt::bulk_g2s(sbase + s * Cfg::kSpecBytes, spec + tile * wt::kRows * (N + 1), bytes, &full[s]);
    }
  };

  if (t == 0) {
    wt::mbar_init(&full[0], 1);
    wt::mbar_init(&full[1], 1);
    wt::fence_mbar_init();
    wt::fence_proxy_async();
  }
  __syncthreads();
  if (t == 0) {
    long long tile = blockIdx.x;
    for (int s = 0; s < 2; ++s, tile += gridDim.x)
      if (tile < tiles) issue(tile, s);
  }
  int it = 0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int rows = tile_rows(tile), brows = bulk_rows(rows);
    unsigned char* line = base + s * wt::kStageBytes + t * 128;
    cx<T>* stile = reinterpret_cast<cx<T>*>(sbase + s * Cfg::kSpecBytes);
    cx<T>* srow = stile + t * (N + 1);
    cx<T>* grow = spec + (tile * wt::kRows + t) * spec_dist;
    if (MODE == 1 || dense) {
      wt::mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));
    } else {
      for (int idx = t; idx < rows * (N + 1); idx += wt::kRows) {
        const int r = idx / (N + 1);
        stile[idx] = spec[(tile * wt::kRows + r) * spec_dist + (idx - r * (N + 1))];
      }
      __syncthreads();
    }
    cx<T> v[N];
    if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const T* src = reinterpret_cast<const T*>(line + ((c ^ (t & 7)) << 4));
        if constexpr (sizeof(T) == 4) {
          const float4 q = *reinterpret_cast<const float4*>(src);
          v[2 * c] = cx<T>{q.x, q.y};
          v[2 * c + 1] = cx<T>{q.z, q.w};
        } else {
          const double2 q = *reinterpret_cast<const double2*>(src);
          v[c] = cx<T>{q.x, q.y};
        }
      }
      DFT<N, T>::run(v);
      cx<T> x[N + 1];
      static_for<0, N>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        x[k] = r2c_combine(v[k], v[k == 0 ? 0 : N - k], const_w<k, 2 * N, T>());
      });
      x[N] = cx<T>{v[0].x - v[0].y, T(0)};
      const bool direct = t >= brows && t < rows;  // the row the bulk copy cannot take
#pragma unroll
      for (int k = 0; k <= N; ++k) {
        const cx<T> o = apply_scale ? cscale(x[k], scale) : x[k];
        if (direct)
          grow[k] = o;
        else
          srow[k] = o;
      }
    } else {
      cx<T> x[N + 1];
      const bool direct = t >= brows && t < rows;
#pragma unroll
      for (int k = 0; k <= N; ++k) x[k] = direct ? grow[k] : srow[k];
      static_for<0, N>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        v[j] = c2r_combine(x[j == 0 ? 0 : N - j], x[j == 0 ? N : j], const_w<j, 2 * N, T>(), j == 0);
      });
      DFT<N, T>::run(v);
      if (apply_scale) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = cscale(v[j], scale);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        T* dst = reinterpret_cast<T*>(line + ((c ^ (t & 7)) << 4));
        if constexpr (sizeof(T) == 4)
          *reinterpret_cast<float4*>(dst) = make_float4(v[2 * c].x, v[2 * c].y, v[2 * c + 1].x, v[2 * c + 1].y);
        else
          *reinterpret_cast<double2*>(dst) = make_double2(v[c].x, v[c].y);
      }
    }
    wt::fence_proxy_async();  // what this thread wrote to shared memory must be visible to the bulk / tensor store
    // the store issued one iteration ago has finished reading the OTHER output buffer (it had a whole iteration)
    if (t == 0) wt::bulk_wait_read<0>();
    __syncthreads();
    if (MODE == 1 && !dense) {
      for (int idx = t; idx < rows * (N + 1); idx += wt::kRows) {
        const int r = idx / (N + 1);
        spec[(tile * wt::kRows + r) * spec_dist + (idx - r * (N + 1))] = stile[idx];
      }
    }
    if (t == 0) {
      if (MODE == 1) {
        if (dense && brows > 0)
          wt::bulk_s2g(spec + tile * wt::kRows * (N + 1), sbase + s * Cfg::kSpecBytes, (uint32_t)(brows * Cfg::kSpecRow));
      } else {
        wt::tma_store_2d(&line_map, 0, (int)(tile * wt::kRows), base + s * wt::kStageBytes);
      }
      wt::bulk_commit();
      // every thread has taken its inputs out of input buffer s: refill it with the tile two ahead
      const long long nxt = tile + 2LL * gridDim.x;
      if (nxt < tiles) issue(nxt, s);
    }
  }
  if (t == 0) wt::bulk_wait_read<0>();  // shared memory must outlive the last store's reads
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnW encode_fn_w() {
  static EncodeTiledFnW fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFnW>(f);
  }();
  return fn;
}

// rows of 128 bytes, `rows` of them `pitch_bytes` apart; box = 128 bytes x kRows, 128-byte swizzle
static bool make_row_map(const void* base, bool is_double, long long rows, long long pitch_bytes, CUtensorMap* map) {
  EncodeTiledFnW enc = encode_fn_w();
  if (enc == nullptr || reinterpret_cast<uintptr_t>(base) % 16 != 0 || pitch_bytes % 16 != 0 || rows <= 0 ||
      rows > (1LL << 31) - wt::kRows)
    return false;
  const cuuint32_t inner = is_double ? 16 : 32;
  cuuint64_t dims[2] = {inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {inner, (cuuint32_t)wt::kRows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base),
             dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool wi_tma_supported(int n, bool is_double) {
  return is_double ? (n == 2 || n == 4 || n == 8) : (n == 2 || n == 4 || n == 8 || n == 16);
}

template <typename T, int N>
static cudaError_t launch_wi_tma_n(const CUtensorMap& in_map, const CUtensorMap& out_map, long long lines, bool swap,
                                   const PassParams& p, int grid, cudaStream_t stream) {
  cudaError_t e = ensure_dynamic_smem(wi_tma_kernel<T, N>, wt::kSmem);
  if (e != cudaSuccess) return e;
  wi_tma_kernel<T, N><<<grid, wt::kRows, wt::kSmem, stream>>>(in_map, out_map, lines, swap, p.apply_scale, (T)p.scale);
  return cudaGetLastError();
}

// p: single batch dimension, contiguous transforms (distance == n), unit element strides, interleaved storage (the
// planner checks); *used == false with cudaSuccess when the buffers cannot be described as whole 128-byte lines ->
// the caller runs the general kernel
cudaError_t launch_wi_tma(const PassParams& p, bool is_double, bool swap, cudaStream_t stream, bool* used) {
  *used = false;
  const size_t esz = is_double ? 16 : 8;
  // one transform per line: rows may be padded (any 16-byte multiple pitch); several per line: contiguous buffers only
  const bool one_per_line = p.n * (long long)esz == 128;
  if (!one_per_line && (p.ibd[0] != p.n || p.obd[0] != p.n || (p.batch_total * p.n * (long long)esz) % 128 != 0))
    return cudaSuccess;
  const long long lines = one_per_line ? p.batch_total : p.batch_total * p.n * (long long)esz / 128;
  CUtensorMap in_map, out_map;
  memset(&in_map, 0, sizeof(in_map));
  memset(&out_map, 0, sizeof(out_map));
  const char* in = reinterpret_cast<const char*>(p.in_re) + (size_t)p.ioff * esz;
  char* out = reinterpret_cast<char*>(p.out_re) + (size_t)p.ooff * esz;
  if (!make_row_map(in, is_double, lines, one_per_line ? p.ibd[0] * (long long)esz : 128, &in_map)) return cudaSuccess;
  if (!make_row_map(out, is_double, lines, one_per_line ? p.obd[0] * (long long)esz : 128, &out_map)) return cudaSuccess;
  *used = true;
  const int sms = sm_count();
  if (sms <= 0) return cudaErrorLaunchOutOfResources;
  const long long tiles = (lines + wt::kRows - 1) / wt::kRows;
  const int grid = (int)(tiles < 4LL * sms ? tiles : 4LL * sms);
  if (is_double) {
    switch (p.n) {
      case 2: return launch_wi_tma_n<double, 2>(in_map, out_map, lines, swap, p, grid, stream);
      case 4: return launch_wi_tma_n<double, 4>(in_map, out_map, lines, swap, p, grid, stream);
      case 8: return launch_wi_tma_n<double, 8>(in_map, out_map, lines, swap, p, grid, stream);
    }
  } else {
    switch (p.n) {
      case 2: return launch_wi_tma_n<float, 2>(in_map, out_map, lines, swap, p, grid, stream);
      case 4: return launch_wi_tma_n<float, 4>(in_map, out_map, lines, swap, p, grid, stream);
      case 8: return launch_wi_tma_n<float, 8>(in_map, out_map, lines, swap, p, grid, stream);
      case 16: return launch_wi_tma_n<float, 16>(in_map, out_map, lines, swap, p, grid, stream);
    }
  }
  return cudaErrorInvalidValue;
}

// REAL domain fused (see wi_tma_real_kernel).  real: 1 real-to-complex (p reads the real rows as pairs, writes rows of
// n + 1 elements), 2 complex-to-real.  The planner guarantees one transform per line, dense half-spectrum rows and a
// single batch dimension; the runtime diverts unaligned buffers to the unfused plan.
bool wi_tma_real_supported(int n, bool is_double) { return is_double ? n == 8 : n == 16; }

template <typename T, int N>
static cudaError_t launch_wi_tma_real_n(const CUtensorMap& map, void* spec, long long spec_dist, int real,
                                        const PassParams& p, int grid, cudaStream_t stream) {
  using Cfg = wt::RealCfg<T, N>;
  if (real == 1) {
    cudaError_t e = ensure_dynamic_smem(wi_tma_real_kernel<T, N, 1>, Cfg::kSmem);
    if (e != cudaSuccess) return e;
    wi_tma_real_kernel<T, N, 1><<<grid, wt::kRows, Cfg::kSmem, stream>>>(map, reinterpret_cast<cx<T>*>(spec), spec_dist,
                                                                         p.batch_total, p.apply_scale, (T)p.scale);
  } else {
    cudaError_t e = ensure_dynamic_smem(wi_tma_real_kernel<T, N, 2>, Cfg::kSmem);
    if (e != cudaSuccess) return e;
    wi_tma_real_kernel<T, N, 2><<<grid, wt::kRows, Cfg::kSmem, stream>>>(map, reinterpret_cast<cx<T>*>(spec), spec_dist,
                                                                         p.batch_total, p.apply_scale, (T)p.scale);
  }
  return cudaGetLastError();
}

cudaError_t launch_wi_tma_real(const PassParams& p, bool is_double, int real, cudaStream_t stream) {
  const size_t esz = is_double ? 16 : 8;
  if (!wi_tma_real_supported(p.n, is_double) || (real != 1 && real != 2)) return cudaErrorInvalidValue;
  // the line side: the real rows (pairs); the other side: dense rows of n + 1 elements
  const char* lines = real == 1 ? reinterpret_cast<const char*>(p.in_re) + (size_t)p.ioff * esz
                                : reinterpret_cast<const char*>(p.out_re) + (size_t)p.ooff * esz;
  char* spec = real == 1 ? reinterpret_cast<char*>(p.out_re) + (size_t)p.ooff * esz
                         : const_cast<char*>(reinterpret_cast<const char*>(p.in_re)) + (size_t)p.ioff * esz;
  const long long line_pitch = (real == 1 ? p.ibd[0] : p.obd[0]) * (long long)esz;
  long long spec_dist = real == 1 ? p.obd[0] : p.ibd[0];
  if (p.batch_total == 1) spec_dist = p.n + 1;
  if (spec_dist < p.n + 1) return cudaErrorInvalidValue;
  // (dense rows move by bulk copies, which need the 16-byte alignment the runtime guarantees for fused passes)
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (!make_row_map(lines, is_double, p.batch_total, line_pitch, &map)) return cudaErrorInvalidValue;
  const int sms = sm_count();
  if (sms <= 0) return cudaErrorLaunchOutOfResources;
  const long long tiles = (p.batch_total + wt::kRows - 1) / wt::kRows;
  const int grid = (int)(tiles < 3LL * sms ? tiles : 3LL * sms);
  return is_double ? launch_wi_tma_real_n<double, 8>(map, spec, spec_dist, real, p, grid, stream)
                   : launch_wi_tma_real_n<float, 16>(map, spec, spec_dist, real, p, grid, stream);
}

}  // namespace pfft
