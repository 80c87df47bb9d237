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
  const int rstride = sub * C;    // shared-memory distance between a butterfly's rows
  for (int b = tb; b < n / R; b += nthreads_b) {
    const int blk = b / sub, j = b - blk * sub;
    cx<T>* base = S + (blk * np + j) * C + c;
    cx<T> v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = base[r * rstride];
    DFT<R, T>::run(v);
    if (sub > 1) {
      const int tj = j * tw_step;
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, tj * r));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) base[r * rstride] = v[r];
  }
}

// IL: interleaved storage; BIG: the radix list may hold the primes 17..31 (their butterflies cost registers the common
// case should not pay for)
template <typename T, bool IL, bool BIG>
__global__ void __launch_bounds__(512) wg_colg_kernel(const PassParams p, const bool swap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* S = reinterpret_cast<cx<T>*>(smem_raw);
  const int C = p.ffts_per_block;     // columns per tile (power of two)
  const int TB = p.threads_per_fft;   // butterfly threads per column
  const int n = p.n;
  int* perm = reinterpret_cast<int*>(S + (size_t)n * C);  // output index k -> tile row holding it
  const int tid = threadIdx.x;
  const int c = tid & (C - 1), tb = tid / C;
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2] * p.nb[3];
  const T scale = T(p.scale);
  const bool sw = IL && swap;
  const long long in_step = (long long)TB * p.is, out_step = (long long)TB * p.os;

  // digit reversal over the radix list, once per CTA: k = r_0 + R_0 (r_1 + R_1 (...)) lives in row
  // r_0 n/R_0 + r_1 n/(R_0 R_1) + ...
  for (int k = tid; k < n; k += blockDim.x) {
    int rem = k, row = 0, span = n;
    for (int ps = 0; ps < p.num_radices; ++ps) {
      const int R = p.radix[ps];
      span /= R;
      row += (rem % R) * span;
      rem /= R;
    }
    perm[k] = row;
  }

  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    long long q = tile / tiles_c;
    const long long c0 = (tile - q * tiles_c) * C;
    long long ib = p.ioff + (c0 + c) * p.ibd[0], ob = p.ooff + (c0 + c) * p.obd[0];
#pragma unroll
    for (int d = 1; d < kMaxBatchDims; ++d) {
      const long long q2 = q / p.nb[d];
      const long long b = q - q2 * p.nb[d];
      q = q2;
      ib += b * p.ibd[d];
      ob += b * p.obd[d];
    }
    const bool live = c0 + c < p.nb[0];
    __syncthreads();  // the tile buffer of the previous iteration has been stored (and perm is complete)
    // ---- load: row by row, lanes along the columns; eight independent loads per thread in flight (the tile is
    // ---- single buffered, so memory-level parallelism inside the load phase is what hides the HBM latency)
    {
      long long idx = ib + (long long)tb * p.is;
      for (int row0 = tb; row0 < n; row0 += 8 * TB, idx += 8 * in_step) {
        cx<T> v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = cx<T>{T(0), T(0)};
          if (live && row0 + u * TB < n) {
            const long long e = idx + u * in_step;
            if (IL) {
              v[u] = reinterpret_cast<const cx<T>*>(p.in_re)[e];
            } else {
              v[u].x = reinterpret_cast<const T*>(p.in_re)[e];
              v[u].y = reinterpret_cast<const T*>(p.in_im)[e];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int row = row0 + u * TB;
          if (row < n) S[row * C + c] = sw ? cx<T>{v[u].y, v[u].x} : v[u];
        }
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
        default:
          if constexpr (BIG) {
            switch (R) {
              PFFT_CASE(17)
              PFFT_CASE(19)
              PFFT_CASE(23)
              PFFT_CASE(29)
              PFFT_CASE(31)
              default:
                break;
            }
          }
          break;
#undef PFFT_CASE
      }
      np /= R;
      __syncthreads();
    }
    // ---- store: output index k from the tile row that holds it --------------------------------------------------
    if (live) {
      long long idx = ob + (long long)tb * p.os;
      for (int k = tb; k < n; k += TB, idx += out_step) {
        cx<T> o = S[perm[k] * C + c];
        if (p.gtw_dim == 0) {  // GLOBAL level: inter-factor twiddle w_M^{column * k} (two-level table)
          const long long m = (c0 + c) * (long long)k;
          o = cmul(o, cmul(ldg_cx<T>(p.gtw_hi, m >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, m & ((1LL << p.gtw_bits) - 1))));
        }
        if (p.apply_scale) o = cscale(o, scale);
        if (IL) {
          reinterpret_cast<cx<T>*>(p.out_re)[idx] = sw ? cx<T>{o.y, o.x} : o;
        } else {
          reinterpret_cast<T*>(p.out_re)[idx] = o.x;
          reinterpret_cast<T*>(p.out_im)[idx] = o.y;
        }
      }
    }
  }
}

template <typename T, bool IL, bool BIG>
cudaError_t launch_colg_v(const PassParams& p, bool swap, size_t smem, int grid, cudaStream_t stream) {
  auto kern = wg_colg_kernel<T, IL, BIG>;
  cudaError_t e = ensure_dynamic_smem(kern, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, p.ffts_per_block * p.threads_per_fft, smem, stream>>>(p, swap);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_colg_t(const PassParams& p, bool il, bool swap, size_t smem, int grid, cudaStream_t stream) {
  bool big = false;
  for (int i = 0; i < p.num_radices; ++i) big = big || p.radix[i] > 16;
  if (il) return big ? launch_colg_v<T, true, true>(p, swap, smem, grid, stream) : launch_colg_v<T, true, false>(p, swap, smem, grid, stream);
  return big ? launch_colg_v<T, false, true>(p, swap, smem, grid, stream) : launch_colg_v<T, false, false>(p, swap, smem, grid, stream);
}

}  // namespace

// tile + the row permutation table
size_t colg_smem_bytes(int n, int columns, bool is_double) {
  return (size_t)n * columns * (is_double ? 16 : 8) + (size_t)n * sizeof(int);
}

cudaError_t launch_wg_colg(const PassParams& p, bool is_double, bool il, bool swap, int grid, cudaStream_t stream) {
  const size_t smem = colg_smem_bytes(p.n, p.ffts_per_block, is_double);
  return is_double ? launch_colg_t<double>(p, il, swap, smem, grid, stream)
                   : launch_colg_t<float>(p, il, swap, smem, grid, stream);
}

}  // namespace pfft
