// WORKGROUP level, column tiles with three compile-time radix passes: N = R0 * R1 * R2 (1000 = 10*10*10,
// 1024 = 16*8*8), C adjacent columns (transforms whose batch neighbours are adjacent in both domains) per tile, any
// storage.  The specialised form of wg_colg.cu for the lengths of its hot cases: BASELINE config C3's
// batch-interleaved variant (C3b: N = 1000 x 100000, split storage) runs here.
//
// Reference counterpart: the BATCH_INTERLEAVED path of workgroup_impl
// (/root/reference/src/portfft/dispatcher/workgroup_dispatcher.hpp:148-229: transposing copies of 32 transforms
// through local memory around wg_dft).  Here:
//   * lanes run along the columns in every phase (C columns = one 32- or 64-byte row segment per memory request), so
//     neither a transposition nor a staging copy exists;
//   * thread (t, c) owns butterfly t of column c in every pass: decimation-in-frequency, in place in ONE
//     [row][column] tile, all strides compile-time, row index padded by row / R2 so that the two rows a half-warp
//     touches fall into different bank halves in all three passes;
//   * a tile arrives by ONE tensor load per plane (cp.async.bulk.tensor.5d; the view (column, s, q) with row = R2 q + s
//     and a box of R2 + 1 rows per group drops the rows straight into the padded layout) and leaves by ONE tensor store
//     per plane (view (column, r, b1, a) with output row k = a + R0 b1 + R0 R1 r: the digit-reversed tile lands in
//     natural order); ring of three tile buffers: load in flight / three passes in place / store draining.  No thread
//     issues a global load or store except the twiddle look-ups;
//   * pointers, strides or batch shapes TMA cannot encode: the register form (inputs of the next tile loaded into
//     registers during the passes, per-thread stores in digit-reversed order);
//   * the generic kernel spends 127 instructions per point on this size (run-time radix list, look-ups, tile
//     load / store loops with bounds checks); this one about a third of that.
#include <cstdlib>
#include <cstring>

#include "col_common.cuh"
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "launch_utils.h"

namespace pfft {

namespace {

template <typename T, int R0_, int R1_, int R2_>
struct ColR3Cfg {
  static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int N = R0 * R1 * R2;
  static constexpr int TPC = R1 * R2;                 // threads per column = butterflies of pass 0
  static constexpr int C = 64 / (int)sizeof(cx<T>);   // columns per tile: 8 (fp32) / 4 (fp64)
  static constexpr int NT = TPC * C;
  static constexpr int ROWS = N + N / R2;             // padded rows: [R0][R1][R2 + 1]
  static constexpr size_t kTilePlane = (size_t)ROWS * C * sizeof(T);  // one scalar plane of the tile
  static constexpr size_t kTile = 2 * kTilePlane;
  static constexpr size_t kSmem = kTile;  // register-prefetch form: one tile
  static constexpr int NBUF = 3;          // TMA form: ring of three tiles (load in flight / compute / store draining)
  // TMA form: every thread owns butterfly t of W adjacent columns.  W = 2 (one 64- / 128-bit shared-memory access
  // serves both columns, shared twiddle look-ups and index arithmetic, 400 threads) was measured SLOWER on C3b than
  // W = 1 (800 threads): 0.477 ms against 0.431 ms -- with three barriers per tile the kernel needs the warps more
  // than it needs fewer instructions (profiles/r2_ab_variants.txt section 4).
#ifndef PFFT_COLR3_W
#define PFFT_COLR3_W 1
#endif
  static constexpr int W = PFFT_COLR3_W;
  static constexpr int NT_TMA = TPC * C / W;
  static constexpr size_t kSmemTma = NBUF * kTile + 64;
  static_assert(R0 >= R1 && R0 >= R2, "the first radix is the largest: one pass-0 butterfly per thread");
  static_assert(NT <= 1024, "block size");
  static_assert(N / R2 <= 256 && R2 + 1 <= 256, "TMA box dimensions");
  static_assert(kTilePlane % 128 == 0, "TMA sources / destinations are 128-byte aligned");
  __host__ __device__ static constexpr int pad(int row) { return row + row / R2; }
};

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(col::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct ColR3Maps {
  CUtensorMap in0, in1;    // input planes (interleaved storage: in0 only)
  CUtensorMap out0, out1;  // output planes
};

// The three passes on one tile buffer.  The tile keeps the storage of the data -- interleaved pairs for interleaved
// storage, two scalar planes for split storage -- so that tensor loads / stores move it plane by plane.
template <typename T, int R0, int R1, int R2, bool IL>
struct ColR3Tile {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  static constexpr int C = Cfg::C, TPC = Cfg::TPC, N = Cfg::N;
  cx<T>* Sx;
  T *Sre, *Sim;
  __device__ __forceinline__ ColR3Tile(unsigned char* raw, int c)
      : Sx(reinterpret_cast<cx<T>*>(raw) + c), Sre(reinterpret_cast<T*>(raw) + c),
        Sim(reinterpret_cast<T*>(raw + Cfg::kTilePlane) + c) {}
  // element at (padded row index * C)
  __device__ __forceinline__ cx<T> ld(int idx) const { return IL ? Sx[idx] : cx<T>{Sre[idx], Sim[idx]}; }
  __device__ __forceinline__ void st(int idx, cx<T> v) const {
    if (IL) {
      Sx[idx] = v;
    } else {
      Sre[idx] = v.x;
      Sim[idx] = v.y;
    }
  }
  // W adjacent columns at once (idx even for W = 2: one vector access per plane)
  template <int W>
  __device__ __forceinline__ void ldw(int idx, cx<T> (&o)[W]) const {
    if constexpr (W == 1) {
      o[0] = ld(idx);
    } else if constexpr (IL) {
      if constexpr (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(Sx + idx);
        o[0] = cx<T>{q.x, q.y};
        o[1] = cx<T>{q.z, q.w};
      } else {
        o[0] = Sx[idx];
        o[1] = Sx[idx + 1];
      }
    } else {
      using V = typename VecOf<T>::type;  // two scalars
      const V re = *reinterpret_cast<const V*>(Sre + idx), im = *reinterpret_cast<const V*>(Sim + idx);
      o[0] = cx<T>{re.x, im.x};
      o[1] = cx<T>{re.y, im.y};
    }
  }
  template <int W>
  __device__ __forceinline__ void stw(int idx, const cx<T> (&v)[W]) const {
    if constexpr (W == 1) {
      st(idx, v[0]);
    } else if constexpr (IL) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(Sx + idx) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
      } else {
        Sx[idx] = v[0];
        Sx[idx + 1] = v[1];
      }
    } else {
      using V = typename VecOf<T>::type;
      V re, im;
      re.x = v[0].x, re.y = v[1].x, im.x = v[0].y, im.y = v[1].y;
      *reinterpret_cast<V*>(Sre + idx) = re;
      *reinterpret_cast<V*>(Sim + idx) = im;
    }
  }
  // pass 0 of butterfly t: outputs (already transformed) times w_N^{t r} into rows t + TPC r
  // (pad(t + TPC r) = pad(t) + (TPC + R1) r)
  __device__ __forceinline__ static int row0(int t) { return Cfg::pad(t) * C; }
  static constexpr int kStep0 = (TPC + R1) * C;
  // pass 1: radix R1 inside each block of N / R0 rows, in place (W adjacent columns per thread)
  template <int W = 1>
  __device__ __forceinline__ void pass1(const PassParams& p, int t) const {
#pragma unroll 1
    for (int b = t; b < N / R1; b += TPC) {
      const int blk = b / R2, j = b - blk * R2;  // rows blk TPC + j + R2 r: pad = blk (TPC + R1) + j + (R2 + 1) r
      const int base = (blk * (TPC + R1) + j) * C;
      cx<T> v[W][R1];
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        cx<T> e[W];
        ldw<W>(base + r * (R2 + 1) * C, e);
#pragma unroll
        for (int w = 0; w < W; ++w) v[w][r] = e[w];
      }
#pragma unroll
      for (int w = 0; w < W; ++w) DFT<R1, T>::run(v[w]);
      if (R2 > 1) {
#pragma unroll
        for (int r = 1; r < R1; ++r) {
          const cx<T> tw = ldg_cx<T>(p.tw, (long long)j * r * R0);  // w_{N/R0}^{j r}
#pragma unroll
          for (int w = 0; w < W; ++w) v[w][r] = cmul(v[w][r], tw);
        }
      }
#pragma unroll
      for (int r = 0; r < R1; ++r) {
        cx<T> e[W];
#pragma unroll
        for (int w = 0; w < W; ++w) e[w] = v[w][r];
        stw<W>(base + r * (R2 + 1) * C, e);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// TMA form: tensor loads and tensor stores, ring of three tile buffers.  The input view (column, s, q, b1, b2) with
// row = R2 q + s and the box (C, R2 + 1, N / R2) puts the rows straight into the padded layout (the row s = R2 of
// every group lies outside the tensor and arrives as zeros); all three passes run in place; the output view
// (column, r, b1, a, batch) with row k = a + R0 b1 + R0 R1 r and the box (C, R2 + 1, R1, R0) writes the digit-reversed
// tile out in natural order (the padding rows, outside the tensor, are skipped).  Per tile: one tensor load and one
// tensor store per plane, three block barriers, no per-thread global access except the twiddle look-ups.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int R0, int R1, int R2, bool IL, bool SWAP>
__global__ void __launch_bounds__(ColR3Cfg<T, R0, R1, R2>::NT_TMA, 1)
    wg_colr3_tma_kernel(const PassParams p, const __grid_constant__ ColR3Maps maps) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  using Tile = ColR3Tile<T, R0, R1, R2, IL>;
  constexpr int N = Cfg::N, TPC = Cfg::TPC, C = Cfg::C, NBUF = Cfg::NBUF, W = Cfg::W;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NBUF * Cfg::kTile);
  const int c = (threadIdx.x % (C / W)) * W, t = threadIdx.x / (C / W);  // first of the thread's W columns
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2];
  const T scale = T(p.scale);

  constexpr bool TW0REG = sizeof(T) == 4 && Cfg::NT_TMA <= 512;  // (800 threads would leave 72 registers each)
  cx<T> tw0[TW0REG ? R0 : 1];
  if (TW0REG) {
#pragma unroll
    for (int r = 1; r < R0; ++r) tw0[r] = ldg_cx<T>(p.tw, (long long)t * r);
  }
  // tile -> first column, batch indices 1 and 2
  auto coords = [&](long long tl, int& c0, int& b1, int& b2) {
    long long q = tl / tiles_c;
    c0 = (int)((tl - q * tiles_c) * C);
    const long long q2 = q / p.nb[1];
    b1 = (int)(q - q2 * p.nb[1]);
    b2 = (int)q2;
  };
  auto issue_load = [&](long long tl, int buf) {
    int c0, b1, b2;
    coords(tl, c0, b1, b2);
    unsigned char* dst = smem_raw + (size_t)buf * Cfg::kTile;
    col::mbar_expect_tx(&full[buf], (uint32_t)Cfg::kTile);
    col::tma_load_5d(dst, &maps.in0, IL ? 2 * c0 : c0, 0, 0, b1, b2, &full[buf]);
    if (!IL) col::tma_load_5d(dst + Cfg::kTilePlane, &maps.in1, c0, 0, 0, b1, b2, &full[buf]);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NBUF; ++s) col::mbar_init(&full[s], 1);
    col::fence_mbar_init();
    col::fence_proxy_async();
  }
  __syncthreads();
  long long tile = blockIdx.x;
  if (threadIdx.x == 0) {
    if (tile < total_tiles) issue_load(tile, 0);
    if (tile + gridDim.x < total_tiles) issue_load(tile + gridDim.x, 1);
  }
  int buf = 0, phase = 0;  // buffer of the current tile; parity of its mbarrier phase (flips every NBUF tiles)
  for (; tile < total_tiles; tile += gridDim.x) {
    unsigned char* raw = smem_raw + (size_t)buf * Cfg::kTile;
    const Tile S(raw, c);
    col::mbar_wait(&full[buf], phase);
    // ---- pass 0: radix R0 on rows t + TPC r, in place -------------------------------------------------------------
    {
      cx<T> v[W][R0];
      const int base = Tile::row0(t);
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        cx<T> e[W];
        S.template ldw<W>(base + r * Tile::kStep0, e);
#pragma unroll
        for (int w = 0; w < W; ++w) v[w][r] = IL && SWAP ? cx<T>{e[w].y, e[w].x} : e[w];
      }
#pragma unroll
      for (int w = 0; w < W; ++w) DFT<R0, T>::run(v[w]);
#pragma unroll
      for (int r = 1; r < R0; ++r) {
        const cx<T> tw = TW0REG ? tw0[TW0REG ? r : 0] : ldg_cx<T>(p.tw, (long long)t * r);
#pragma unroll
        for (int w = 0; w < W; ++w) v[w][r] = cmul(v[w][r], tw);
      }
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        cx<T> e[W];
#pragma unroll
        for (int w = 0; w < W; ++w) e[w] = v[w][r];
        S.template stw<W>(base + r * Tile::kStep0, e);
      }
    }
    __syncthreads();
    S.template pass1<W>(p, t);
    __syncthreads();
    // ---- pass 2: radix R2 on rows b R2 + r, in place: row [a][b1][r] is row k = a + R0 b1 + R0 R1 r of the output ----
#pragma unroll 1
    for (int b = t; b < N / R2; b += TPC) {
      const int base = b * (R2 + 1) * C;  // pad(b R2 + r) = b (R2 + 1) + r
      cx<T> v[W][R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        cx<T> e[W];
        S.template ldw<W>(base + r * C, e);
#pragma unroll
        for (int w = 0; w < W; ++w) v[w][r] = e[w];
      }
#pragma unroll
      for (int w = 0; w < W; ++w) DFT<R2, T>::run(v[w]);
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        cx<T> e[W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
          cx<T> o = v[w][r];
          if (p.apply_scale) o = cscale(o, scale);
          e[w] = IL && SWAP ? cx<T>{o.y, o.x} : o;
        }
        S.template stw<W>(base + r * C, e);
      }
    }
    col::fence_proxy_async();  // this thread's tile writes become visible to the tensor store
    // The store of the previous tile has had a whole tile's time to read its buffer: once thread 0 has seen it
    // complete, that buffer takes the load of the tile after next.
    if (threadIdx.x == 0) bulk_wait_read_all();
    __syncthreads();
    if (threadIdx.x == 0) {
      const long long t2 = tile + 2 * (long long)gridDim.x;
      if (t2 < total_tiles) issue_load(t2, buf == 0 ? NBUF - 1 : buf - 1);
      int c0, b1, b2;
      coords(tile, c0, b1, b2);
      // (the store view has one batch dimension: b2 == 0 here, see colr3_make_store_map)
      tma_store_5d(&maps.out0, raw, IL ? 2 * c0 : c0, 0, 0, 0, b1);
      if (!IL) tma_store_5d(&maps.out1, raw + Cfg::kTilePlane, c0, 0, 0, 0, b1);
      bulk_commit();
    }
    if (++buf == NBUF) {
      buf = 0;
      phase ^= 1;
    }
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// Register form, any pointer alignment / stride / number of batch dimensions: the inputs of the next tile are loaded
// into registers while the current tile runs its three passes; per-thread stores.  (Measured on C3b: 0.95 ms -- every
// warp stalls in the burst of loads and stores, lg_throttle -- against 0.44 ms for the TMA form.)
// I: element-index type (int when every index of both buffers fits 31 bits).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int R0, int R1, int R2, bool IL, bool SWAP, typename I>
__global__ void __launch_bounds__(ColR3Cfg<T, R0, R1, R2>::NT, 1) wg_colr3_kernel(const PassParams p) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  using Tile = ColR3Tile<T, R0, R1, R2, IL>;
  constexpr int N = Cfg::N, TPC = Cfg::TPC, C = Cfg::C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int c = threadIdx.x % C, t = threadIdx.x / C;
  const Tile S(smem_raw, c);
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2] * p.nb[3];
  const T scale = T(p.scale);
  const I in_step = (I)(TPC * p.is), out_step = (I)((R0 * R1) * p.os);
  const I is = (I)p.is, os = (I)p.os;

  constexpr bool TW0REG = sizeof(T) == 4 && Cfg::NT <= 512;
  cx<T> tw0[TW0REG ? R0 : 1];
  if (TW0REG) {
#pragma unroll
    for (int r = 1; r < R0; ++r) tw0[r] = ldg_cx<T>(p.tw, (long long)t * r);
  }
  auto load_elem = [&](I idx) -> cx<T> {
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
  // input / output offsets of this thread's column in tile `tile`; false: the column lies beyond the batch
  auto bases = [&](long long tile, I& ib_out, I& ob_out) -> bool {
    long long q = tile / tiles_c;
    const long long col = (tile - q * tiles_c) * C + c;
    long long ib = p.ioff + col * p.ibd[0];
    long long ob = p.ooff + col * p.obd[0];
#pragma unroll
    for (int d = 1; d < kMaxBatchDims; ++d) {
      const long long q2 = q / p.nb[d];
      const long long b = q - q2 * p.nb[d];
      q = q2;
      ib += b * p.ibd[d];
      ob += b * p.obd[d];
    }
    ib_out = (I)ib;
    ob_out = (I)ob;
    return col < p.nb[0];
  };

  long long tile = blockIdx.x;
  I ib = 0, ob = 0;
  bool live = tile < total_tiles && bases(tile, ib, ob);
  cx<T> nxt[R0];
  if (live) {
    const I i0 = ib + (I)t * is;
#pragma unroll
    for (int r = 0; r < R0; ++r) nxt[r] = load_elem(i0 + r * in_step);
  }
  for (; tile < total_tiles; tile += gridDim.x) {
    const bool cur_live = live;
    const I cur_ob = ob;
    // ---- pass 0: radix R0 on rows t + TPC r, registers -> tile ------------------------------------------------------
    {
      cx<T> v[R0];
#pragma unroll
      for (int r = 0; r < R0; ++r) v[r] = nxt[r];
      // inputs of the next tile: in flight while this one runs its three passes
      const long long tn = tile + gridDim.x;
      live = tn < total_tiles && bases(tn, ib, ob);
      if (live) {
        const I i0 = ib + (I)t * is;
#pragma unroll
        for (int r = 0; r < R0; ++r) nxt[r] = load_elem(i0 + r * in_step);
      }
      if (cur_live) {
        DFT<R0, T>::run(v);
#pragma unroll
        for (int r = 1; r < R0; ++r) v[r] = cmul(v[r], TW0REG ? tw0[TW0REG ? r : 0] : ldg_cx<T>(p.tw, (long long)t * r));
        const int dst = Tile::row0(t);
#pragma unroll
        for (int r = 0; r < R0; ++r) S.st(dst + r * Tile::kStep0, v[r]);
      }
    }
    __syncthreads();
    if (cur_live) S.pass1(p, t);
    __syncthreads();
    // ---- pass 2: radix R2 on rows b R2 + r; tile -> global ----------------------------------------------------------
    if (cur_live) {
#pragma unroll 1
      for (int b = t; b < N / R2; b += TPC) {
        const int base = b * (R2 + 1) * C;  // pad(b R2 + r) = b (R2 + 1) + r
        cx<T> v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = S.ld(base + r * C);
        DFT<R2, T>::run(v);
        // row a N/R0 + b1 R2 + r holds output index k = a + R0 b1 + R0 R1 r (b = a R1 + b1)
        const int a = b / R1, b1 = b - a * R1;
        const I o0 = cur_ob + (I)(a + R0 * b1) * os;
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          cx<T> o = v[r];
          if (p.apply_scale) o = cscale(o, scale);
          const I idx = o0 + r * out_step;
          if (IL) {
            if (SWAP) o = cx<T>{o.y, o.x};
            reinterpret_cast<cx<T>*>(p.out_re)[idx] = o;
          } else {
            reinterpret_cast<T*>(p.out_re)[idx] = o.x;
            reinterpret_cast<T*>(p.out_im)[idx] = o.y;
          }
        }
      }
    }
    __syncthreads();  // the tile is free for the next pass 0
  }
}

// largest element index the pass touches on either side
inline long long colr3_max_index(const PassParams& p) {
  long long mi = p.ioff + (long long)(p.n - 1) * p.is, mo = p.ooff + (long long)(p.n - 1) * p.os;
  for (int d = 0; d < kMaxBatchDims; ++d) {
    mi += (p.nb[d] - 1) * p.ibd[d];
    mo += (p.nb[d] - 1) * p.obd[d];
  }
  return mi > mo ? mi : mo;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn colr3_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 5-D view of one plane: dimension 0 = columns (`scalars` per element), dimensions 1..4 = (size, stride in elements).
bool colr3_encode(const void* base, int scalars, bool is_double, long long columns, const long long (&size)[4],
                  const long long (&stride)[4], const int (&box)[5], int promo, CUtensorMap* map) {
  EncodeTiledFn enc = colr3_encode_fn();
  if (enc == nullptr || reinterpret_cast<uintptr_t>(base) % 16 != 0) return false;
  const size_t esz = (size_t)scalars * (is_double ? 8 : 4);
  cuuint64_t dims[5] = {(cuuint64_t)columns * scalars, 0, 0, 0, 0};
  cuuint64_t strides[4];
  for (int i = 0; i < 4; ++i) {
    const unsigned long long bytes = (unsigned long long)stride[i] * esz;
    if (size[i] <= 0 || stride[i] <= 0 || bytes % 16 != 0 || bytes >= (1ULL << 40)) return false;
    dims[i + 1] = (cuuint64_t)size[i];
    strides[i] = bytes;
  }
  for (int i = 0; i < 5; ++i)
    if (dims[i] == 0 || dims[i] > (1ULL << 32)) return false;
  cuuint32_t bx[5], estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) bx[i] = (cuuint32_t)box[i];
  const CUtensorMapL2promotion pr =
      promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                 : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                               : (promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
  return enc(map, is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base),
             dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// PFFT_COLR3_TMA: 0 = register form; 1..3 = TMA form with L2 promotion 64 / 128 / 256 B on the loads, 4 = none
int colr3_mode() {
  static const int mode = [] {
    const char* e = std::getenv("PFFT_COLR3_TMA");
    return e ? std::atoi(e) : 3;
  }();
  return mode;
}

// Tensor maps of the pass (one per scalar plane for split storage), from the plan's per-pass cache when the base
// addresses are the ones they were encoded for.  cache[0..1]: input planes, cache[2..3]: output planes.
// false: this geometry cannot be expressed (alignment, more than two / one batch dimensions beside the columns).
template <typename T, int R0, int R1, int R2, bool IL>
bool colr3_tensor_maps(const PassParams& p, ColMapCache* cache, ColR3Maps* m) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  const int mode = colr3_mode();
  memset(m, 0, sizeof(*m));
  if (mode == 0 || p.nb[2] != 1 || p.nb[3] != 1) return false;
  const int promo = mode == 4 ? 0 : mode;
  const size_t sc = sizeof(T);
  const bool dbl = sizeof(T) == 8;
  const int scalars = IL ? 2 : 1;
  const void* bases[4] = {reinterpret_cast<const char*>(p.in_re) + (size_t)p.ioff * (IL ? 2 * sc : sc),
                          IL ? nullptr : reinterpret_cast<const char*>(p.in_im) + (size_t)p.ioff * sc,
                          reinterpret_cast<const char*>(p.out_re) + (size_t)p.ooff * (IL ? 2 * sc : sc),
                          IL ? nullptr : reinterpret_cast<const char*>(p.out_im) + (size_t)p.ooff * sc};
  CUtensorMap* maps[4] = {&m->in0, &m->in1, &m->out0, &m->out1};
  const long long whole_in = p.is * p.n, whole_out = p.os * p.n;  // (stride of a dimension of size 1: any valid value)
  // input: row = R2 q + s                       output: row k = a + R0 b1 + R0 R1 r
  const long long in_size[4] = {R2, Cfg::N / R2, p.nb[1], 1}, in_stride[4] = {p.is, (long long)R2 * p.is,
                                                                              p.nb[1] > 1 ? p.ibd[1] : whole_in, whole_in};
  const long long out_size[4] = {R2, R1, R0, p.nb[1]},
                  out_stride[4] = {(long long)R0 * R1 * p.os, (long long)R0 * p.os, p.os, p.nb[1] > 1 ? p.obd[1] : whole_out};
  const int in_box[5] = {scalars * Cfg::C, R2 + 1, Cfg::N / R2, 1, 1}, out_box[5] = {scalars * Cfg::C, R2 + 1, R1, R0, 1};
  for (int i = 0; i < 4; ++i) {
    if (IL && (i & 1)) continue;
    ColMapCache* ce = cache ? cache + i : nullptr;
    if (ce != nullptr && ce->base == bases[i]) {
      memcpy(maps[i], ce->map, sizeof(CUtensorMap));
      continue;
    }
    const bool made = i < 2 ? colr3_encode(bases[i], scalars, dbl, p.nb[0], in_size, in_stride, in_box, promo, maps[i])
                            : colr3_encode(bases[i], scalars, dbl, p.nb[0], out_size, out_stride, out_box, 0, maps[i]);
    if (!made) return false;
    if (ce != nullptr) {
      memcpy(ce->map, maps[i], sizeof(CUtensorMap));
      ce->base = bases[i];
    }
  }
  return true;
}

template <typename T, int R0, int R1, int R2, bool IL, bool SWAP, typename I>
cudaError_t launch_colr3_regs(const PassParams& p, int grid, cudaStream_t stream) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  auto kern = wg_colr3_kernel<T, R0, R1, R2, IL, SWAP, I>;
  cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmem);
  if (e != cudaSuccess) return e;
  kern<<<grid, Cfg::NT, Cfg::kSmem, stream>>>(p);
  return cudaGetLastError();
}

template <typename T, int R0, int R1, int R2, bool IL, bool SWAP>
cudaError_t launch_colr3_v(const PassParams& p, int grid, cudaStream_t stream, ColMapCache* cache) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  ColR3Maps m;
  if (colr3_tensor_maps<T, R0, R1, R2, IL>(p, cache, &m)) {
    auto kern = wg_colr3_tma_kernel<T, R0, R1, R2, IL, SWAP>;
    cudaError_t e = ensure_dynamic_smem(kern, Cfg::kSmemTma);
    if (e != cudaSuccess) return e;
    kern<<<grid, Cfg::NT_TMA, Cfg::kSmemTma, stream>>>(p, m);
    return cudaGetLastError();
  }
  if (colr3_max_index(p) < (1LL << 31) - 1) return launch_colr3_regs<T, R0, R1, R2, IL, SWAP, int>(p, grid, stream);
  return launch_colr3_regs<T, R0, R1, R2, IL, SWAP, long long>(p, grid, stream);
}

template <typename T, int R0, int R1, int R2>
cudaError_t launch_colr3_t(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream, ColMapCache* cache) {
  if (!il) return launch_colr3_v<T, R0, R1, R2, false, false>(p, grid, stream, cache);
  return swap ? launch_colr3_v<T, R0, R1, R2, true, true>(p, grid, stream, cache)
              : launch_colr3_v<T, R0, R1, R2, true, false>(p, grid, stream, cache);
}

}  // namespace

#define PFFT_COLR3_LIST(X) \
  X(1000, 10, 10, 10)      \
  X(1024, 16, 8, 8)

bool colr3_supported(int n, bool is_double, int* columns, int* threads_per_column, size_t* smem) {
  switch (n) {
#define X(NN, A, B, CC)                                                                                       \
  case NN:                                                                                                    \
    if (columns) *columns = is_double ? ColR3Cfg<double, A, B, CC>::C : ColR3Cfg<float, A, B, CC>::C;         \
    if (threads_per_column) *threads_per_column = ColR3Cfg<float, A, B, CC>::TPC;                             \
    if (smem) *smem = is_double ? ColR3Cfg<double, A, B, CC>::kSmemTma : ColR3Cfg<float, A, B, CC>::kSmemTma; \
    return true;
    PFFT_COLR3_LIST(X)
#undef X
    default:
      return false;
  }
}

cudaError_t launch_wg_colr3(const PassParams& p, bool is_double, bool il, bool swap, int grid, cudaStream_t stream,
                            ColMapCache* cache) {
  switch (p.n) {
#define X(NN, A, B, CC)                                                                     \
  case NN:                                                                                  \
    return is_double ? launch_colr3_t<double, A, B, CC>(p, il, swap, grid, stream, cache)   \
                     : launch_colr3_t<float, A, B, CC>(p, il, swap, grid, stream, cache);
    PFFT_COLR3_LIST(X)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pfft
