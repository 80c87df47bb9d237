// WORKGROUP level, column tiles of ANY 31-smooth length: C adjacent columns (transforms whose elements lie `is` apart
// and whose neighbours in the batch lie 1 apart) are transformed together, in place in shared memory.
//
// This is the batch-interleaved / outer-dimension counterpart of wg_col.cu for the lengths that kernel does not take
// (non powers of two such as 1000 = BASELINE config C3's batch-interleaved variant C3b, 96, 1536, 3072; powers of two
// beyond 512) and for split storage.  Reference counterparts: the BATCH_INTERLEAVED paths of workgroup_impl
// (/root/reference/src/portfft/dispatcher/workgroup_dispatcher.hpp:148-229: 32 transforms staged through local memory
// by transposing copies) and of the per-plane N-D loop (/root/reference/src/portfft/committed_descriptor_impl.hpp:
// 932-948).  Here:
//   * lanes run along the batch index in every phase: global loads and stores move full row segments (C complex
//     = 128 bytes for fp32 interleaved), and the [row][column] shared-memory tile is conflict free for any radix;
//   * the passes are decimation-in-frequency butterflies IN PLACE (radix R outputs go back to the R rows they came
//     from, twiddle w_{N'}^{j r} applied to the outputs), so one tile buffer suffices -- no ping-pong -- and lengths
//     up to 1792 (fp32, C = 16) fit one CTA; the digit-reversed order this leaves in the tile costs nothing because
//     the store picks the rows in output order while its lanes still run along the columns;
//   * backward = (re <-> im) swap on load and store, scale and the GLOBAL level's inter-factor twiddle fused into the
//     store (the column passes of non-power-of-two multi-pass lengths such as 68640 run here too).
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "launch_utils.h"

namespace pfft {

namespace {

// one in-place DIF pass of radix R on the blocks of length np: rows base + (np/R) r, r < R, of column c
template <typename T, int R>
__device__ __forceinline__ void dif_pass(const PassParams& p, cx<T>* S, int C, int c, int tb, int nthreads_b, int np) {
  const int n = p.n;
  const int sub = np / R;         // butterflies per block = distance between a butterfly's rows
  const int tw_step = n / np;     // w_np^{j r} = w_n^{j r tw_step}
  for (int b = tb; b < n / R; b += nthreads_b) {
    const int blk = b / sub, j = b - blk * sub;
    cx<T>* base = S + ((size_t)(blk * np + j)) * C + c;
    cx<T> v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = base[(size_t)r * sub * C];
    DFT<R, T>::run(v);
    if (sub > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, (long long)j * r * tw_step));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) base[(size_t)r * sub * C] = v[r];
  }
}

template <typename T>
__global__ void __launch_bounds__(512) wg_colg_kernel(const PassParams p, const bool il, const bool swap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem_raw);
  const int C = p.ffts_per_block;     // columns per tile (power of two)
  const int TB = p.threads_per_fft;   // butterfly threads per column
  const int n = p.n;
  const IoFlags fl{il, swap};
  const int tid = threadIdx.x;
  const int c = tid & (C - 1), tb = tid / C;
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2] * p.nb[3];
  const T scale = T(p.scale);

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    long long q = tile / tiles_c;
    const long long c0 = (tile - q * tiles_c) * C;
    long long ib = p.ioff + c0 * p.ibd[0], ob = p.ooff + c0 * p.obd[0];
#pragma unroll
    for (int d = 1; d < kMaxBatchDims; ++d) {
      const long long q2 = q / p.nb[d];
      const long long b = q - q2 * p.nb[d];
      q = q2;
      ib += b * p.ibd[d];
      ob += b * p.obd[d];
    }
    const bool live = c0 + c < p.nb[0];
    // ---- load: row by row, lanes along the columns -------------------------------------------------------------
    // (eight independent loads per thread in flight before the first shared-memory store: the tile is single
    // buffered, so memory-level parallelism inside the load phase is what hides the HBM latency)
    for (int row0 = tb; row0 < n; row0 += 8 * TB) {
      cx<T> v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = row0 + u * TB;
        v[u] = (live && row < n) ? gload<T>(p, fl, ib + (long long)c * p.ibd[0] + (long long)row * p.is) : cx<T>{T(0), T(0)};
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = row0 + u * TB;
        if (row < n) S[(size_t)row * C + c] = v[u];
      }
    }
    __syncthreads();
    // ---- in-place DIF passes ------------------------------------------------------------------------------------
    int np = n;
    for (int ps = 0; ps < p.num_radices; ++ps) {
      const int R = p.radix[ps];
      switch (R) {
#define PFFT_CASE(RR)                                  \
  case RR:                                             \
    dif_pass<T, RR>(p, S, C, c, tb, TB, np);           \
    break;
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
      np /= R;
      __syncthreads();
    }
    // ---- store: output index k -> the tile row that holds it (digit reversal over the radix list) ---------------
    for (int k = tb; k < n; k += TB) {
      int rem = k, row = 0, span = n;
      for (int ps = 0; ps < p.num_radices; ++ps) {
        const int R = p.radix[ps];
        span /= R;
        const int digit = rem % R;  // k = r_0 + R_0 (r_1 + R_1 (...)), row = r_0 n/R_0 + r_1 n/(R_0 R_1) + ...
        rem /= R;
        row += digit * span;
      }
      if (live) {
        cx<T> o = S[(size_t)row * C + c];
        if (p.gtw_dim == 0) {  // GLOBAL level: inter-factor twiddle w_M^{column * k} (two-level table)
          const long long m = (c0 + c) * (long long)k;
          o = cmul(o, cmul(ldg_cx<T>(p.gtw_hi, m >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, m & ((1LL << p.gtw_bits) - 1))));
        }
        if (p.apply_scale) o = cscale(o, scale);
        gstore<T>(p, fl, ob + (long long)c * p.obd[0] + (long long)k * p.os, o);
      }
    }
    __syncthreads();  // the tile buffer is reloaded by the next iteration
  }
}

}  // namespace

size_t colg_smem_bytes(int n, int columns, bool is_double) { return (size_t)n * columns * (is_double ? 16 : 8); }

cudaError_t launch_wg_colg(const PassParams& p, bool is_double, bool il, bool swap, int grid, cudaStream_t stream) {
  const size_t smem = colg_smem_bytes(p.n, p.ffts_per_block, is_double);
  const int threads = p.ffts_per_block * p.threads_per_fft;
  cudaError_t e;
  if (is_double) {
    e = ensure_dynamic_smem(wg_colg_kernel<double>, smem);
    if (e != cudaSuccess) return e;
    wg_colg_kernel<double><<<grid, threads, smem, stream>>>(p, il, swap);
  } else {
    e = ensure_dynamic_smem(wg_colg_kernel<float>, smem);
    if (e != cudaSuccess) return e;
    wg_colg_kernel<float><<<grid, threads, smem, stream>>>(p, il, swap);
  }
  return cudaGetLastError();
}

}  // namespace pfft
