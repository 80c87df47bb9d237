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
//   * the tile's inputs arrive by TMA (cp.async.bulk.tensor, boxes of 256 rows x one row segment; split storage: one
//     tensor map per plane) into a two-stage ring one and two tiles AHEAD, guarded by mbarriers: no thread issues a
//     global load, the LSU queues stay empty and the loads of the next tiles are in flight during all three passes.
//     Pointers or strides TMA cannot encode: inputs loaded into registers one tile ahead instead (measured on C3b:
//     0.95 ms -- every warp stalls in the load burst, lg_throttle -- against the TMA ring's figure in
//     profiles/r2_ab_variants.txt);
//   * pass 2 stores its outputs to global memory straight from registers in digit-reversed order -- the order costs
//     nothing because the lanes still run along the columns;
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
  static constexpr size_t kSmem = kTile;  // register-prefetch form
  // TMA form: two stages of whole 256-row boxes (rows beyond N arrive as zeros) + two mbarriers
  static constexpr int BOX = 256;
  static constexpr int NBOX = (N + BOX - 1) / BOX;
  static constexpr size_t kPlane = (size_t)NBOX * BOX * C * sizeof(T);  // one scalar plane of a stage
  static constexpr size_t kStage = 2 * kPlane;
  static constexpr size_t kSmemTma = kTile + 2 * kStage + 64;
  static_assert(R0 >= R1 && R0 >= R2, "the first radix is the largest: one pass-0 butterfly per thread");
  static_assert(NT <= 1024, "block size");
  static_assert(kTilePlane % 128 == 0 && kPlane % 128 == 0, "TMA sources / destinations are 128-byte aligned");
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

// I: element-index type of the per-thread global accesses (int when every index of both buffers fits 31 bits).
// TMA: inputs through the TMA ring; TMAST: outputs by TMA tensor stores from the tile (implies TMA).
// The tile keeps the storage of the data: interleaved pairs for interleaved storage, two scalar planes for split
// storage -- so that one tensor store per plane writes it out.
template <typename T, int R0, int R1, int R2, bool IL, bool SWAP, typename I, bool TMA, bool TMAST>
__global__ void __launch_bounds__(ColR3Cfg<T, R0, R1, R2>::NT, 1)
    wg_colr3_kernel(const PassParams p, const __grid_constant__ ColR3Maps maps) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  constexpr int N = Cfg::N, TPC = Cfg::TPC, C = Cfg::C;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage0 = smem_raw + Cfg::kTile;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + 2 * Cfg::kStage);
  const int c = threadIdx.x % C, t = threadIdx.x / C;
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2] * p.nb[3];
  const T scale = T(p.scale);
  const I in_step = (I)(TPC * p.is), out_step = (I)((R0 * R1) * p.os);
  const I is = (I)p.is, os = (I)p.os;

  // tile element (padded row index * C + c)
  cx<T>* Sx = reinterpret_cast<cx<T>*>(smem_raw) + c;
  T* Sre = reinterpret_cast<T*>(smem_raw) + c;
  T* Sim = reinterpret_cast<T*>(smem_raw + Cfg::kTilePlane) + c;
  auto tile_ld = [&](int idx) -> cx<T> { return IL ? Sx[idx] : cx<T>{Sre[idx], Sim[idx]}; };
  auto tile_st = [&](int idx, cx<T> v) {
    if (IL) {
      Sx[idx] = v;
    } else {
      Sre[idx] = v.x;
      Sim[idx] = v.y;
    }
  };

  // pass-0 twiddles w_N^{t r} stay in registers when the block leaves room; pass 1 needs only R2 * R1 distinct values:
  // table look-ups (L1 resident)
  constexpr bool TW0REG = sizeof(T) == 4 && Cfg::NT <= 512;  // (800 threads leave 72 registers each)
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
  // tile -> TMA coordinates: first column, batch indices 1..3
  auto coords = [&](long long tl, int& c0, int (&bc)[3]) {
    long long q = tl / tiles_c;
    c0 = (int)((tl - q * tiles_c) * C);
#pragma unroll
    for (int d = 1; d < kMaxBatchDims; ++d) {
      const long long q2 = q / p.nb[d];
      bc[d - 1] = (int)(q - q2 * p.nb[d]);
      q = q2;
    }
  };
  // one thread: TMA loads of tile `tl` into stage `s` (NBOX boxes per plane; the box rows beyond N are zero-filled)
  auto issue = [&](long long tl, int s) {
    int c0, bc[3];
    coords(tl, c0, bc);
    unsigned char* dst = stage0 + (size_t)s * Cfg::kStage;
    col::mbar_expect_tx(&full[s], (uint32_t)Cfg::kStage);
#pragma unroll
    for (int bx = 0; bx < Cfg::NBOX; ++bx) {
      if (IL) {  // one map of (re, im) pairs: rows of 2 C scalars
        col::tma_load_5d(dst + (size_t)bx * Cfg::BOX * C * sizeof(cx<T>), &maps.in0, 2 * c0, bx * Cfg::BOX, bc[0], bc[1],
                         bc[2], &full[s]);
      } else {
        col::tma_load_5d(dst + (size_t)bx * Cfg::BOX * C * sizeof(T), &maps.in0, c0, bx * Cfg::BOX, bc[0], bc[1], bc[2],
                         &full[s]);
        col::tma_load_5d(dst + Cfg::kPlane + (size_t)bx * Cfg::BOX * C * sizeof(T), &maps.in1, c0, bx * Cfg::BOX, bc[0],
                         bc[1], bc[2], &full[s]);
      }
    }
  };

  long long tile = blockIdx.x;
  I ib = 0, ob = 0;
  bool live = tile < total_tiles && bases(tile, ib, ob);
  cx<T> nxt[TMA ? 1 : R0];
  if (TMA) {
    if (threadIdx.x == 0) {
      col::mbar_init(&full[0], 1);
      col::mbar_init(&full[1], 1);
      col::fence_mbar_init();
      col::fence_proxy_async();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (tile < total_tiles) issue(tile, 0);
      if (tile + gridDim.x < total_tiles) issue(tile + gridDim.x, 1);
    }
  } else if (live) {
    const I i0 = ib + (I)t * is;
#pragma unroll
    for (int r = 0; r < R0; ++r) nxt[r] = load_elem(i0 + r * in_step);
  }
  int it = 0;
  for (; tile < total_tiles; tile += gridDim.x, ++it) {
    const bool cur_live = live;
    const I cur_ob = ob;
    // ---- pass 0: radix R0 on rows t + TPC r, stage (or registers) -> tile -------------------------------------------
    {
      cx<T> v[R0];
      const long long tn = tile + gridDim.x;
      if (TMA) {
        const unsigned char* st = stage0 + (size_t)(it & 1) * Cfg::kStage;
        col::mbar_wait(&full[it & 1], (it >> 1) & 1);
        if (IL) {
          const cx<T>* src = reinterpret_cast<const cx<T>*>(st) + t * C + c;
#pragma unroll
          for (int r = 0; r < R0; ++r) {
            v[r] = src[r * TPC * C];
            if (SWAP) v[r] = cx<T>{v[r].y, v[r].x};
          }
        } else {
          const T* sre = reinterpret_cast<const T*>(st) + t * C + c;
          const T* sim = reinterpret_cast<const T*>(st + Cfg::kPlane) + t * C + c;
#pragma unroll
          for (int r = 0; r < R0; ++r) v[r] = cx<T>{sre[r * TPC * C], sim[r * TPC * C]};
        }
        live = tn < total_tiles && bases(tn, ib, ob);
      } else {
#pragma unroll
        for (int r = 0; r < R0; ++r) v[r] = nxt[TMA ? 0 : r];
        // inputs of the next tile: in flight while this one runs its three passes
        live = tn < total_tiles && bases(tn, ib, ob);
        if (live) {
          const I i0 = ib + (I)t * is;
#pragma unroll
          for (int r = 0; r < R0; ++r) nxt[TMA ? 0 : r] = load_elem(i0 + r * in_step);
        }
      }
      if (cur_live) {
        DFT<R0, T>::run(v);
#pragma unroll
        for (int r = 1; r < R0; ++r) v[r] = cmul(v[r], TW0REG ? tw0[TW0REG ? r : 0] : ldg_cx<T>(p.tw, (long long)t * r));
      }
      if (TMAST && it > 0) {
        // the tensor store of the previous tile must have read the tile buffer before it is overwritten
        if (threadIdx.x == 0) bulk_wait_read_all();
        __syncthreads();
      }
      if (cur_live) {
        const int dst = Cfg::pad(t) * C;  // pad(t + TPC r) = pad(t) + (TPC + R1) r
#pragma unroll
        for (int r = 0; r < R0; ++r) tile_st(dst + r * (TPC + R1) * C, v[r]);
      }
    }
    __syncthreads();
    if (TMA && threadIdx.x == 0) {
      // every thread has taken its inputs out of this stage: refill it with the tile after next
      const long long t2 = tile + 2 * (long long)gridDim.x;
      if (t2 < total_tiles) {
        col::fence_proxy_async();
        issue(t2, it & 1);
      }
    }
    // ---- pass 1: radix R1 inside each block of N / R0 rows, in place ------------------------------------------------
    if (cur_live) {
#pragma unroll 1
      for (int b = t; b < N / R1; b += TPC) {
        const int blk = b / R2, j = b - blk * R2;  // rows blk TPC + j + R2 r: pad = blk (TPC + R1) + j + (R2 + 1) r
        const int base = (blk * (TPC + R1) + j) * C;
        cx<T> v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = tile_ld(base + r * (R2 + 1) * C);
        DFT<R1, T>::run(v);
        if (R2 > 1) {
#pragma unroll
          for (int r = 1; r < R1; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, (long long)j * r * R0));  // w_{N/R0}^{j r}
        }
#pragma unroll
        for (int r = 0; r < R1; ++r) tile_st(base + r * (R2 + 1) * C, v[r]);
      }
    }
    __syncthreads();
    // ---- pass 2: radix R2 on rows b R2 + r; tile -> global, or back into the tile for the tensor store --------------
    if (cur_live || TMAST) {
#pragma unroll 1
      for (int b = t; b < N / R2; b += TPC) {
        const int base = b * (R2 + 1) * C;  // pad(b R2 + r) = b (R2 + 1) + r
        cx<T> v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = tile_ld(base + r * C);
        DFT<R2, T>::run(v);
        if (TMAST) {
          // in place: row [a][b1][r] of the tile is row k = a + R0 b1 + R0 R1 r of the output (the store's tensor map)
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            cx<T> o = v[r];
            if (p.apply_scale) o = cscale(o, scale);
            if (IL && SWAP) o = cx<T>{o.y, o.x};
            tile_st(base + r * C, o);
          }
        } else {
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
    }
    if (TMAST) {
      col::fence_proxy_async();  // this thread's tile writes become visible to the tensor store
      __syncthreads();
      if (threadIdx.x == 0) {
        int c0, bc[3];
        coords(tile, c0, bc);
        // box (columns, R2 + 1, R1, R0, 1): the padding row r = R2 lies outside the tensor and is not written; neither
        // are the columns beyond the batch
        tma_store_5d(&maps.out0, smem_raw, IL ? 2 * c0 : c0, 0, 0, 0, bc[0]);
        if (!IL) tma_store_5d(&maps.out1, smem_raw + Cfg::kTilePlane, c0, 0, 0, 0, bc[0]);
        bulk_commit();
      }
    } else {
      __syncthreads();  // the tile is free for the next pass 0
    }
  }
  if (TMAST && threadIdx.x == 0) bulk_wait_all();
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

template <typename T, int R0, int R1, int R2, bool IL, bool SWAP, typename I, bool TMA, bool TMAST>
cudaError_t launch_colr3_i(const PassParams& p, const ColR3Maps& m, int grid, cudaStream_t stream) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  const size_t smem = TMA ? Cfg::kSmemTma : Cfg::kSmem;
  auto kern = wg_colr3_kernel<T, R0, R1, R2, IL, SWAP, I, TMA, TMAST>;
  cudaError_t e = ensure_dynamic_smem(kern, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, Cfg::NT, smem, stream>>>(p, m);
  return cudaGetLastError();
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

// Output view for the tensor store: (column, r, b1, a, batch dimension 1) with output row k = a + R0 b1 + R0 R1 r, so
// that the dense [a][b1][r (+ 1 padding row)][column] tile lands in natural order.  false: not encodable.
template <int R0, int R1, int R2>
bool colr3_make_store_map(const PassParams& p, const void* base_ptr, int scalars, bool is_double, int C, CUtensorMap* map) {
  EncodeTiledFn enc = colr3_encode_fn();
  if (enc == nullptr || p.nb[2] != 1 || p.nb[3] != 1) return false;
  const size_t sc = is_double ? 8 : 4, esz = scalars * sc;
  if (reinterpret_cast<uintptr_t>(base_ptr) % 16 != 0) return false;
  cuuint64_t dims[5] = {(cuuint64_t)p.nb[0] * scalars, (cuuint64_t)R2, (cuuint64_t)R1, (cuuint64_t)R0, (cuuint64_t)p.nb[1]};
  const long long st[4] = {(long long)R0 * R1 * p.os, (long long)R0 * p.os, p.os, p.nb[1] > 1 ? p.obd[1] : p.os * p.n};
  cuuint64_t strides[4];
  for (int i = 0; i < 4; ++i) {
    const unsigned long long bytes = (unsigned long long)st[i] * esz;
    if (st[i] <= 0 || bytes % 16 != 0 || bytes >= (1ULL << 40)) return false;
    strides[i] = bytes;
  }
  if (dims[0] > (1ULL << 32) || dims[4] > (1ULL << 32)) return false;
  cuuint32_t box[5] = {(cuuint32_t)(scalars * C), (cuuint32_t)(R2 + 1), (cuuint32_t)R1, (cuuint32_t)R0, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                         const_cast<void*>(base_ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// PFFT_COLR3_TMA: 0 = register prefetch and per-thread stores; 1..3 = TMA loads with L2 promotion 64 / 128 / 256 B,
// 4 = none; +8 = per-thread stores instead of tensor stores
int colr3_mode() {
  static const int mode = [] {
    const char* e = std::getenv("PFFT_COLR3_TMA");
    return e ? std::atoi(e) : 3;
  }();
  return mode;
}

// tensor maps of the pass (one per scalar plane for split storage), from the plan's per-pass cache when the base
// addresses are the ones they were encoded for.  cache[0..1]: input planes, cache[2..3]: output planes.
template <typename T, int R0, int R1, int R2, bool IL>
void colr3_tensor_maps(const PassParams& p, ColMapCache* cache, ColR3Maps* m, bool* loads, bool* stores) {
  using Cfg = ColR3Cfg<T, R0, R1, R2>;
  const int mode = colr3_mode();
  *loads = *stores = false;
  memset(m, 0, sizeof(*m));
  if ((mode & 7) == 0) return;
  const int promo = (mode & 7) == 4 ? 0 : (mode & 7);
  const size_t sc = sizeof(T);
  const bool dbl = sizeof(T) == 8;
  const void* bases[4] = {reinterpret_cast<const char*>(p.in_re) + (size_t)p.ioff * (IL ? 2 * sc : sc),
                          IL ? nullptr : reinterpret_cast<const char*>(p.in_im) + (size_t)p.ioff * sc,
                          reinterpret_cast<const char*>(p.out_re) + (size_t)p.ooff * (IL ? 2 * sc : sc),
                          IL ? nullptr : reinterpret_cast<const char*>(p.out_im) + (size_t)p.ooff * sc};
  CUtensorMap* maps[4] = {&m->in0, &m->in1, &m->out0, &m->out1};
  bool ok[2] = {true, (mode & 8) == 0};
  for (int i = 0; i < 4; ++i) {
    if ((IL && (i & 1)) || !ok[i / 2]) continue;
    ColMapCache* ce = cache ? cache + i : nullptr;
    if (ce != nullptr && ce->base == bases[i]) {
      memcpy(maps[i], ce->map, sizeof(CUtensorMap));
      continue;
    }
    const bool made = i < 2 ? col_make_tensor_map_plane(p, bases[i], IL ? 2 : 1, dbl, Cfg::C, Cfg::BOX, promo, maps[i])
                            : colr3_make_store_map<R0, R1, R2>(p, bases[i], IL ? 2 : 1, dbl, Cfg::C, maps[i]);
    if (!made) {
      ok[i / 2] = false;
      continue;
    }
    if (ce != nullptr) {
      memcpy(ce->map, maps[i], sizeof(CUtensorMap));
      ce->base = bases[i];
    }
  }
  *loads = ok[0];
  *stores = ok[0] && ok[1];
}

template <typename T, int R0, int R1, int R2, bool IL, bool SWAP>
cudaError_t launch_colr3_v(const PassParams& p, int grid, cudaStream_t stream, ColMapCache* cache) {
  ColR3Maps m;
  bool loads = false, stores = false;
  colr3_tensor_maps<T, R0, R1, R2, IL>(p, cache, &m, &loads, &stores);
  const bool small = colr3_max_index(p) < (1LL << 31) - 1;
  if (stores) return launch_colr3_i<T, R0, R1, R2, IL, SWAP, int, true, true>(p, m, grid, stream);
  if (loads)
    return small ? launch_colr3_i<T, R0, R1, R2, IL, SWAP, int, true, false>(p, m, grid, stream)
                 : launch_colr3_i<T, R0, R1, R2, IL, SWAP, long long, true, false>(p, m, grid, stream);
  return small ? launch_colr3_i<T, R0, R1, R2, IL, SWAP, int, false, false>(p, m, grid, stream)
               : launch_colr3_i<T, R0, R1, R2, IL, SWAP, long long, false, false>(p, m, grid, stream);
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
#define X(NN, A, B, CC)                                                                        \
  case NN:                                                                                     \
    if (columns) *columns = is_double ? ColR3Cfg<double, A, B, CC>::C : ColR3Cfg<float, A, B, CC>::C; \
    if (threads_per_column) *threads_per_column = ColR3Cfg<float, A, B, CC>::TPC;              \
    if (smem) *smem = ColR3Cfg<float, A, B, CC>::kSmemTma; /* same bytes for both precisions */ \
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
