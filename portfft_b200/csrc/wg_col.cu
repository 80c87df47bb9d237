// WORKGROUP level, column tiles: C adjacent columns (C * sizeof(complex) = 128 bytes) of a strided batch are
// transformed together, N = N1 * N2 points each, persistent CTAs.  This kernel runs
//   * the outer dimensions of N-D transforms (512^3: the y and x passes), and
//   * every pass of the GLOBAL level (four-step for large N: strided column transforms with the inter-factor
//     twiddle fused into the store, and the final row pass whose store performs the transposition).
//
// Reference counterparts: the per-plane launch loop of dispatch_dimensions
// (/root/reference/src/portfft/committed_descriptor_impl.hpp:932-948: batch*outer separate BATCH_INTERLEAVED
// launches), global_kernel / dispatch_level with store modifiers
// (/root/reference/src/portfft/common/global.hpp:135-170,303-401) and the 16x16 transpose kernels
// (/root/reference/src/portfft/common/transpose.hpp:44-100).  Here:
//   * IN_COLS: the N x C tile (rows `is` elements apart) is fetched by TMA tensor copies
//     (cp.async.bulk.tensor.5d, <= 256 rows per box) into a 2-stage shared-memory ring guarded by mbarriers, issued
//     by one thread; the next tile streams in while this one is computed.  Every row segment is one full 128-byte
//     line, so HBM sees only whole lines although the transform direction is strided;
//   * IN_ROWS (last pass of the GLOBAL level): each of the C transforms is a contiguous row, read directly with
//     lanes along the row; the stores below then write columns -> the transposition costs no extra pass;
//   * two register-resident passes (radix N1, then radix N2; dft.cuh) with ONE exchange through a padded
//     shared-memory buffer laid out [column][index] with odd pitch (conflict free for both access directions);
//   * results leave straight from registers with lanes along the column index: 128-byte row segments, full lines;
//     inter-factor twiddle w_M^{column * k} (two-level device table), scale and the backward (re <-> im) swap are
//     fused into that store.
#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "col_common.cuh"
#include "device_utils.cuh"
#include "io.cuh"
#include "kernels.h"
#include "launch_utils.h"

namespace pfft {


// IN: how a tile of C transforms reaches the CTA
enum : int {
  IN_COLS_TMA = 0,    // strided columns: TMA tensor tiles into the stage ring, stage layout [row j][column]
  IN_ROWS_DIRECT = 1, // contiguous rows: direct global loads (no staging, latency hidden by resident CTAs)
  IN_ROWS_BULK = 2    // contiguous rows, 16-byte aligned: cp.async.bulk (TMA 1-D) into the ring, layout [row][j]
};

// Tile coordinates (column tile, batch indices 1..3) of a persistent CTA, advanced by the grid size with carries
// instead of being re-derived from the linear tile index: the 64-bit divisions of that decode were ~300 of the ~1200
// instructions a thread spent per tile (SASS: 12 division subroutine calls per kernel).
struct TileCursor {
  int ct, b1, b2, b3;
  int s_ct, s_b1, s_b2, s_b3;
  int tiles_c, nb1, nb2;
  __device__ __forceinline__ void init(long long first, long long step, long long tiles_c_, long long nb1_, long long nb2_) {
    tiles_c = (int)tiles_c_;
    nb1 = (int)nb1_;
    nb2 = (int)nb2_;
    split(first, ct, b1, b2, b3);
    split(step, s_ct, s_b1, s_b2, s_b3);
  }
  __device__ __forceinline__ void split(long long t, int& c, int& x1, int& x2, int& x3) const {
    long long q = t / tiles_c;
    c = (int)(t - q * tiles_c);
    long long q2 = q / nb1;
    x1 = (int)(q - q2 * nb1);
    long long q3 = q2 / nb2;
    x2 = (int)(q2 - q3 * nb2);
    x3 = (int)q3;
  }
  __device__ __forceinline__ void set(long long tile) { split(tile, ct, b1, b2, b3); }
  __device__ __forceinline__ void advance() {
    ct += s_ct;
    if (ct >= tiles_c) {
      ct -= tiles_c;
      ++b1;
    }
    b1 += s_b1;
    if (b1 >= nb1) {
      b1 -= nb1;
      ++b2;
    }
    b2 += s_b2;
    if (b2 >= nb2) {
      b2 -= nb2;
      ++b3;
    }
    b3 += s_b3;
  }
};

// INPLACE (strided columns in and out, two passes, one last-pass butterfly per thread): the exchange between the two
// passes happens inside the stage buffer -- pass 1 writes its outputs back to the rows it read, pass 2 then finds its
// N2 inputs in N2 consecutive rows -- so no exchange buffer exists and more CTAs fit one SM.  The [row][column]
// stage layout is conflict free for both passes because lanes always run along the 128-byte row segment.
template <typename T, int N1, int N2, int N3, int IN, bool INPLACE = false>
struct ColCfg {
  static_assert(N3 == 1 || (N1 == N2 && N2 == N3), "three-pass variant: equal radices (one butterfly per thread)");
  static constexpr int N = N1 * N2 * N3;
  // transforms per tile: one 128-byte line per row segment (16 fp32 / 8 fp64).  Measured at N = 512: 64-byte segments
  // (two CTAs per SM) are 15-20 % slower than full lines with one CTA per SM.
  static constexpr int C = 128 / (2 * (int)sizeof(T));
  static constexpr int NL = N3 > 1 ? N3 : N2;           // radix of the last pass
  static constexpr int NS = N / NL;                     // butterflies of the last pass
  // threads per transform: two-pass variants loop over their butterflies, the three-pass variant has one per thread
  static constexpr int TPC = N3 > 1 ? N / N1 : col::cmin(col::cmin(N1, N2), 16);
  static constexpr int NT = C * TPC;
  static constexpr int PITCH = col::pitch<T>(N);
  // ring depth (a one-stage ring with three CTAs per SM was measured for fp64: 2.45 ms on C4 against 2.39 ms)
  static constexpr size_t kStageBytes = (size_t)N * C * 2 * sizeof(T);
  // in place: one stage when a tile is 64 KiB (the refill then overlaps the last pass's arithmetic and stores and the
  // other CTAs of the SM), two otherwise
  // fp64 in place (32 KiB stages, one CTA of 8 warps per SM, a tile lasts ~1.7 us): kRingF64 stages, so that a refill
  // -- issued only when its tile has left the stage -- has two tile times to arrive instead of one
#ifndef PFFT_COL_RING_F64
#define PFFT_COL_RING_F64 3
#endif
  static constexpr int RING =
      INPLACE && kStageBytes > 48 * 1024 ? 1 : (INPLACE && sizeof(T) == 8 && N3 == 1 ? PFFT_COL_RING_F64 : 2);
  static constexpr int STAGES = IN == IN_ROWS_DIRECT ? 0 : RING;
  static constexpr size_t kSmem = STAGES * kStageBytes + (INPLACE ? 0 : (size_t)C * PITCH * 2 * sizeof(T)) + 64;
  static_assert(!INPLACE || (IN == IN_COLS_TMA && N3 == 1 && N1 == TPC && (N / N1) % TPC == 0),
                "in-place exchange: TMA column tiles, two passes, one last-pass butterfly per thread");
  static constexpr int kBoxRows = N < 256 ? N : 256;
  // resident CTAs per SM the register allocation is bounded for (in place, fp32 N = 256: three 64 KiB rings per SM)
  // (the other fp32 variants spend 190-210 registers with the lean index arithmetic = one CTA per SM; bounding them
  // to two CTAs per SM was measured slower: 65536 x 2048 0.720 -> 0.772 ms)
  static constexpr int kMinBlocks =
      INPLACE && sizeof(T) == 4 && N == 256 ? 3 : 1;
};

template <typename T, int N1, int N2, int N3, int IN, bool OUT_ROWS, bool INPLACE = false>
__global__ void __launch_bounds__(ColCfg<T, N1, N2, N3, IN, INPLACE>::NT, ColCfg<T, N1, N2, N3, IN, INPLACE>::kMinBlocks)
    wg_col_kernel(const PassParams p, const __grid_constant__ CUtensorMap tmap, const bool swap) {
  using Cfg = ColCfg<T, N1, N2, N3, IN, INPLACE>;
  static_assert(!INPLACE || !OUT_ROWS, "in-place exchange: column output only");
  constexpr int N = Cfg::N, C = Cfg::C, TPC = Cfg::TPC, PITCH = Cfg::PITCH, NL = Cfg::NL, NS = Cfg::NS;
  constexpr int B1 = N / N1;  // butterflies of pass 1
  constexpr bool IN_ROWS = IN != IN_COLS_TMA;
  constexpr bool RING = IN != IN_ROWS_DIRECT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage0 = smem_raw;
  cx<T>* E = reinterpret_cast<cx<T>*>(smem_raw + Cfg::STAGES * Cfg::kStageBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(E + (INPLACE ? 0 : (size_t)C * PITCH));  // in place: no exchange buffer
  const int tid = threadIdx.x;
  // backward = (re <-> im) swap on the sides that are user data (plan-internal sides run the plain forward transform)
  const bool swap_in = swap && !(p.mod_flags & MOD_NO_USER_SWAP_IN), swap_out = swap && !(p.mod_flags & MOD_NO_USER_SWAP_OUT);
  // column mapping: lanes run along the transform index (used where memory is contiguous across transforms)
  const int cc = tid % C, tc = tid / C;
  // row mapping: lanes run along the element index (used where each transform is contiguous)
  const int cr = tid / TPC, tr = tid % TPC;
  const int c1 = IN_ROWS ? cr : cc, t1 = IN_ROWS ? tr : tc;     // pass 1
  const int c2 = OUT_ROWS ? cr : cc, t2 = OUT_ROWS ? tr : tc;   // pass 2
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2] * p.nb[3];
  const T scale = T(p.scale);
  const unsigned long long gmask = (1ULL << p.gtw_bits) - 1;
  const bool in_contig = p.ibd[0] == N;  // IN_ROWS_BULK: the C rows of a tile are one contiguous run

  // fp32: tile coordinates advance incrementally and results leave through a running pointer (-14 % / -6 % on the two
  // passes of 65536 x 2048).  fp64 keeps the per-tile decode and per-element address arithmetic: measured on C4, every
  // reduction of the instruction count made these DRAM-paced kernels (8 warps per SM) slower, 772 -> 802 us and
  // 741 -> 762 us per pass with the lean form, 835 / 809 us with the pass twiddles held in registers as well
  // (profiles/r2_ab_variants.txt).
  constexpr bool kLean = sizeof(T) == 4;
  TileCursor cur, icur;  // tile being transformed / next tile to load (thread 0)
  cur.init(blockIdx.x, gridDim.x, tiles_c, p.nb[1], p.nb[2]);
  icur = cur;
  long long icur_tile = blockIdx.x;
  auto issue = [&](int s) {  // loads the tile under `icur` into stage s and moves the cursor on
    const int c0 = icur.ct * C, b1 = icur.b1, b2 = icur.b2, b3 = icur.b3;
    if (kLean)
      icur.advance();
    else
      icur.set(icur_tile += gridDim.x);
    unsigned char* dst = stage0 + s * Cfg::kStageBytes;
    if (IN == IN_COLS_TMA) {
      col::mbar_expect_tx(&full[s], (uint32_t)Cfg::kStageBytes);
#pragma unroll
      for (int r0 = 0; r0 < N; r0 += Cfg::kBoxRows)
        col::tma_load_5d(dst + (size_t)r0 * C * sizeof(cx<T>), &tmap, 2 * c0, r0, b1, b2, b3,
                         &full[s]);  // innermost coordinate counts scalars
    } else {
      const int rows = (int)min((long long)C, p.nb[0] - c0);
      const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in_re) + p.ioff + (long long)c0 * p.ibd[0] +
                         (long long)b1 * p.ibd[1] + (long long)b2 * p.ibd[2] + (long long)b3 * p.ibd[3];
      col::mbar_expect_tx(&full[s], (uint32_t)(rows * N * sizeof(cx<T>)));
      if (in_contig) {
        col::bulk_g2s(dst, src, (uint32_t)(rows * N * sizeof(cx<T>)), &full[s]);
      } else {
        for (int r = 0; r < rows; ++r)
          col::bulk_g2s(dst + (size_t)r * N * sizeof(cx<T>), src + (long long)r * p.ibd[0], (uint32_t)(N * sizeof(cx<T>)),
                        &full[s]);
      }
    }
  };

  if (RING) {
    if (tid == 0) {
      for (int s = 0; s < Cfg::RING; ++s) col::mbar_init(&full[s], 1);
      col::fence_mbar_init();
      col::fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
      long long t0 = blockIdx.x;
      for (int s = 0; s < Cfg::RING; ++s, t0 += gridDim.x)
        if (t0 < total_tiles) issue(s);
    }
  }

  int it = 0;
  for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    if (kLean) {
      if (it) cur.advance();
    } else {
      cur.set(tile);
    }
    const int c0 = cur.ct * C, b1 = cur.b1, b2 = cur.b2, b3 = cur.b3;
    // ---- pass 1: radix N1 over x[j + N2*r], result to E[transform][pad(j*N1 + r)] -------------------------------
    {
      const int st = it % Cfg::RING;
      cx<T>* S = reinterpret_cast<cx<T>*>(stage0 + st * Cfg::kStageBytes);
      const bool live = c0 + c1 < p.nb[0];
      const long long ib = p.ioff + (long long)(c0 + c1) * p.ibd[0] + (long long)b1 * p.ibd[1] +
                           (long long)b2 * p.ibd[2] + (long long)b3 * p.ibd[3];
      if (RING) col::mbar_wait(&full[st], (it / Cfg::RING) & 1);
#pragma unroll 1
      for (int j = t1; j < B1; j += TPC) {
        cx<T> v[N1];
        if (IN == IN_COLS_TMA) {
#pragma unroll
          for (int r = 0; r < N1; ++r) v[r] = S[(j + B1 * r) * C + c1];
        } else if (IN == IN_ROWS_BULK) {
#pragma unroll
          for (int r = 0; r < N1; ++r) v[r] = S[c1 * N + j + B1 * r];
        } else {
#pragma unroll
          for (int r = 0; r < N1; ++r)
            v[r] = live ? reinterpret_cast<const cx<T>*>(p.in_re)[ib + (j + B1 * r)] : cx<T>{T(0), T(0)};
        }
        if (swap_in) {
#pragma unroll
          for (int r = 0; r < N1; ++r) {
            const T t = v[r].x;
            v[r].x = v[r].y;
            v[r].y = t;
          }
        }
        DFT<N1, T>::run(v);
        if (INPLACE) {
#pragma unroll
          for (int r = 0; r < N1; ++r) S[(j + B1 * r) * C + c1] = v[r];
        } else {
#pragma unroll
          for (int r = 0; r < N1; ++r) E[c1 * PITCH + col::pad<T>(j * N1 + r)] = v[r];
        }
      }
    }
    __syncthreads();
    if (RING && !INPLACE && tid == 0) {
      // the stage just consumed is free: refill it with the tile RING iterations ahead
      const long long nxt = tile + (long long)Cfg::RING * gridDim.x;
      if (nxt < total_tiles) {
        col::fence_proxy_async();
        issue(it % Cfg::RING);
      }
    }
    if (N3 > 1) {
      // ---- middle pass (three-pass variant), in place on E: read, barrier, write --------------------------------
      const int j = tc, k = j % N1;
      cx<T> v[N2];
#pragma unroll
      for (int r = 0; r < N2; ++r) v[r] = E[cc * PITCH + col::pad<T>(j + (N / N2) * r)];
#pragma unroll
      for (int r = 1; r < N2; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, k * r * N3));  // w_{N1 N2}^{k r}
      DFT<N2, T>::run(v);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < N2; ++r) E[cc * PITCH + col::pad<T>((j - k) * N2 + k + N1 * r)] = v[r];
      __syncthreads();
    }
    // ---- last pass: E[transform][pad(j + NS*r)] * w_N^{j r}, radix NL, store out[base + (j + NS*r)*os] -----------
    {
      const bool live = c0 + c2 < p.nb[0];
      const long long ob = p.ooff + (long long)(c0 + c2) * p.obd[0] + (long long)b1 * p.obd[1] +
                           (long long)b2 * p.obd[2] + (long long)b3 * p.obd[3];
      long long gidx = 0;
      if (p.gtw_dim >= 0) gidx = p.gtw_dim == 0 ? c0 + c2 : (p.gtw_dim == 1 ? b1 : (p.gtw_dim == 2 ? b2 : b3));
#pragma unroll 1
      for (int j = t2; j < NS; j += TPC) {
        cx<T> v[NL];
        if (INPLACE) {
          // element j + NS*r of the exchange is output j of pass-1 butterfly r: row r + (N/N1)*j of the stage
          const int st = it % Cfg::RING;
          const cx<T>* S = reinterpret_cast<const cx<T>*>(stage0 + st * Cfg::kStageBytes);
#pragma unroll
          for (int r = 0; r < NL; ++r) v[r] = S[(r + B1 * j) * C + c2];
          __syncthreads();  // (one iteration per thread) every thread has taken its inputs: the stage is free
          if (tid == 0) {
            const long long nxt = tile + (long long)Cfg::RING * gridDim.x;
            if (nxt < total_tiles) {
              col::fence_proxy_async();
              issue(st);
            }
          }
        } else {
#pragma unroll
          for (int r = 0; r < NL; ++r) v[r] = E[c2 * PITCH + col::pad<T>(j + NS * r)];
        }
#pragma unroll
        for (int r = 1; r < NL; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, j * r));
        DFT<NL, T>::run(v);
        if (!kLean) {
        if (live) {
          const IoFlags fl{true, swap_out};
          cx<T> tw_run{T(1), T(0)}, tw_step{T(1), T(0)};
          if (p.gtw_dim >= 0 && sizeof(T) == 8) {
            const long long mb = gidx * j, ms = gidx * NS;
            tw_run = cmul(ldg_cx<T>(p.gtw_hi, mb >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, mb & (long long)gmask));
            tw_step = cmul(ldg_cx<T>(p.gtw_hi, ms >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, ms & (long long)gmask));
          }
#pragma unroll
          for (int r = 0; r < NL; ++r) {
            const int k = j + NS * r;
            cx<T> o = v[r];
            if (p.gtw_dim >= 0) {
              if (sizeof(T) == 8) {
                o = cmul(o, tw_run);
                tw_run = cmul(tw_run, tw_step);
              } else {
                const long long m = gidx * k;
                o = cmul(o, cmul(ldg_cx<T>(p.gtw_hi, m >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, m & (long long)gmask)));
              }
            }
            bool keep = true;
            if (p.smod != nullptr) {  // table over the whole (multi-pass) transform: see PassParams::smod_mask / smod_n1
              const long long lin = p.smod_mask != 0 ? ((ob + (long long)k * p.os) & p.smod_mask)
                                                     : (long long)(c0 + c2) + (long long)p.smod_n1 * k;
              keep = p.smod_mask != 0 || lin < p.valid_out;
              if (p.mod_flags & MOD_SWAP_PRE) o = cx<T>{o.y, o.x};
              if (keep) o = cmul(o, ldg_cx<T>(p.smod, lin));
              if (p.mod_flags & MOD_SWAP_POST) o = cx<T>{o.y, o.x};
            }
            if (p.apply_scale) o = cscale(o, scale);
            if (keep) gstore<T>(p, fl, ob + (long long)k * p.os, o);
          }
        }
        } else
        if (live) {
          // results leave through a running pointer (base + k * os, k = j + NS * r)
          cx<T>* op = reinterpret_cast<cx<T>*>(p.out_re) + ob + (long long)j * p.os;
          const long long ostep = (long long)NS * p.os;
          if (p.gtw_dim >= 0) {
            // inter-factor twiddle w_M^{g*k}.  fp64: two table look-ups (base w^{g*j}, step w^{g*NS}) and a running
            // product (error ~ NL * 1.1e-16, far inside the fp64 bound) instead of 2*NL dependent L2 reads; fp32: the
            // tables are small enough to stay in L1, every element is looked up exactly
            cx<T> tw_run{T(1), T(0)}, tw_step{T(1), T(0)};
            if (sizeof(T) == 8) {
              tw_run = gtw_lookup<T>(p, (unsigned)gidx, (unsigned)j, gmask);
              tw_step = gtw_lookup<T>(p, (unsigned)gidx, (unsigned)NS, gmask);
            }
#pragma unroll
            for (int r = 0; r < NL; ++r) {
              cx<T> o;
              if (sizeof(T) == 8) {
                o = cmul(v[r], tw_run);
                tw_run = cmul(tw_run, tw_step);
              } else {
                o = cmul(v[r], gtw_lookup<T>(p, (unsigned)gidx, (unsigned)(j + NS * r), gmask));
              }
              if (p.apply_scale) o = cscale(o, scale);
              if (swap_out) o = cx<T>{o.y, o.x};
              *op = o;
              op += ostep;
            }
          } else if (p.smod != nullptr) {
            // last pass of a multi-pass transform with a table over the whole transform (Bluestein: times the
            // transformed chirp and a (re <-> im) swap; or swap, times chirp / M, truncation): see PassParams::smod_mask
            // (index = position inside the packed row) and PassParams::smod_n1 (index = linear output index)
            // (the table values are fetched as ONE batch of independent loads, clamped in range, before they are used:
            // interleaved with the stores their L2 latency was exposed load by load -- 1.7 -> 5.5 ms on the 512-point
            // pass of N = 65537 x 2048)
            const long long off = ob + (long long)j * p.os;
            cx<T> m[NL];
            if (p.smod_mask != 0) {
#pragma unroll
              for (int r = 0; r < NL; ++r) m[r] = ldg_cx<T>(p.smod, (off + r * ostep) & p.smod_mask);
            } else {
              const long long l0 = (long long)(c0 + c2) + (long long)p.smod_n1 * j, lstep = (long long)p.smod_n1 * NS;
#pragma unroll
              for (int r = 0; r < NL; ++r) m[r] = ldg_cx<T>(p.smod, min(l0 + r * lstep, (long long)p.valid_out - 1));
            }
            const long long keep_below = p.smod_mask != 0 ? (1LL << 62) : (long long)p.valid_out;  // linear index bound
            const long long l0 = (long long)(c0 + c2) + (long long)p.smod_n1 * j, lstep = (long long)p.smod_n1 * NS;
#pragma unroll
            for (int r = 0; r < NL; ++r) {
              cx<T> o = v[r];
              if (p.mod_flags & MOD_SWAP_PRE) o = cx<T>{o.y, o.x};
              o = cmul(o, m[r]);
              if (p.mod_flags & MOD_SWAP_POST) o = cx<T>{o.y, o.x};
              if (p.apply_scale) o = cscale(o, scale);
              if (swap_out) o = cx<T>{o.y, o.x};
              if (l0 + r * lstep < keep_below) *op = o;
              op += ostep;
            }
          } else {
#pragma unroll
            for (int r = 0; r < NL; ++r) {
              cx<T> o = v[r];
              if (p.apply_scale) o = cscale(o, scale);
              if (swap_out) o = cx<T>{o.y, o.x};
              *op = o;
              op += ostep;
            }
          }
        }
      }
    }
    if (!INPLACE) __syncthreads();  // E is rewritten by the next tile's pass 1
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 N = 512, columns in and out: two consumer groups over a three-stage ring.
//
// One tile (512 rows x 16 columns) is 64 KiB, so the generic variant above fits one CTA of 8 warps per SM and its
// arithmetic phases (4 issue slots, 2 warps each) cannot cover each other.  Here one CTA of 16 warps is two groups of
// 256 threads that work on alternate tiles, half a tile apart in time; each group runs the in-place two-pass scheme
// (radix 16, then radix 32 on 32 consecutive rows) on its stage with its own named barrier, and the third stage
// always holds the tile in flight: the thread that finishes reading a stage refills it with the tile three ahead.
// ---------------------------------------------------------------------------------------------------------------
namespace col {
__device__ __forceinline__ void group_sync(int group) {
  asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace col

struct Col512 {
  static constexpr int N = 512, N1 = 16, N2 = 32, C = 16, B1 = N / N1, RING = 3, GROUP = 256, NT = 2 * GROUP;
  static constexpr size_t kStageBytes = (size_t)N * C * sizeof(cx<float>);
  static constexpr size_t kSmem = RING * kStageBytes + 64;
};

// Stage s is used in turn by tiles s, s + 3, s + 6, ... of the CTA, alternately by the two groups.  `full[s]` completes
// one phase per load; `freed[s]` completes one phase per use, when the using group has taken its inputs and posted the
// next load.  A group waits for freed (previous use, by the other group) before it waits for full: it completed the
// use before that itself, so both one-bit phase parities are unambiguous however far the groups drift apart.
__global__ void __launch_bounds__(Col512::NT, 1)
    wg_col512_kernel(const PassParams p, const __grid_constant__ CUtensorMap tmap, const bool swap) {
  using T = float;
  constexpr int N = Col512::N, N1 = Col512::N1, N2 = Col512::N2, C = Col512::C, B1 = Col512::B1, RING = Col512::RING;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + RING * Col512::kStageBytes);
  uint64_t* freed = full + RING;
  const IoFlags fl{true, swap};
  const int group = threadIdx.x / Col512::GROUP, gt = threadIdx.x % Col512::GROUP;
  const int cc = gt % C, tc = gt / C;  // column of the tile, butterfly index (16 per column)
  const long long tiles_c = (p.nb[0] + C - 1) / C;
  const long long total_tiles = tiles_c * p.nb[1] * p.nb[2] * p.nb[3];
  const T scale = T(p.scale);
  const long long gmask = (1LL << p.gtw_bits) - 1;

  auto decode = [&](long long tile, int& c0, int& b1, int& b2, int& b3) {
    long long q = tile / tiles_c;
    c0 = (int)(tile - q * tiles_c) * C;
    long long q2 = q / p.nb[1];
    b1 = (int)(q - q2 * p.nb[1]);
    long long q3 = q2 / p.nb[2];
    b2 = (int)(q2 - q3 * p.nb[2]);
    b3 = (int)q3;
  };
  auto issue = [&](long long tile, int s) {
    int c0, b1, b2, b3;
    decode(tile, c0, b1, b2, b3);
    unsigned char* dst = smem_raw + s * Col512::kStageBytes;
    col::mbar_expect_tx(&full[s], (uint32_t)Col512::kStageBytes);
#pragma unroll
    for (int r0 = 0; r0 < N; r0 += 256)
      col::tma_load_5d(dst + (size_t)r0 * C * sizeof(cx<T>), &tmap, 2 * c0, r0, b1, b2, b3, &full[s]);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) {
      col::mbar_init(&full[s], 1);
      col::mbar_init(&freed[s], 1);
    }
    col::fence_mbar_init();
    col::fence_proxy_async();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = blockIdx.x;
    for (int s = 0; s < RING; ++s, t0 += gridDim.x)
      if (t0 < total_tiles) issue(t0, s);
  }

  // k-th tile of this CTA: blockIdx.x + k * gridDim.x, in stage k % 3, mbarrier phase (k / 3) & 1; group g takes k = g, g + 2, ...
  for (long long k = group;; k += 2) {
    const long long tile = blockIdx.x + k * gridDim.x;
    if (tile >= total_tiles) break;
    const int st = (int)(k % RING);
    cx<T>* S = reinterpret_cast<cx<T>*>(smem_raw + st * Col512::kStageBytes);
    int c0, b1, b2, b3;
    decode(tile, c0, b1, b2, b3);
    const long long use = k / RING;
    if (use > 0) col::mbar_wait(&freed[st], (uint32_t)((use - 1) & 1));
    col::mbar_wait(&full[st], (uint32_t)(use & 1));
    // ---- pass 1: radix 16 over rows j + 32 r, written back to the rows it read --------------------------------
#pragma unroll 1
    for (int j = tc; j < B1; j += 16) {
      cx<T> v[N1];
#pragma unroll
      for (int r = 0; r < N1; ++r) v[r] = S[(j + B1 * r) * C + cc];
      if (swap) {
#pragma unroll
        for (int r = 0; r < N1; ++r) v[r] = cx<T>{v[r].y, v[r].x};
      }
      DFT<N1, T>::run(v);
#pragma unroll
      for (int r = 0; r < N1; ++r) S[(j + B1 * r) * C + cc] = v[r];
    }
    col::group_sync(group);
    // ---- pass 2: butterfly tc takes rows 32 tc .. 32 tc + 31 (element tc + 16 r of the exchange), radix 32 -------
    cx<T> v[N2];
#pragma unroll
    for (int r = 0; r < N2; ++r) v[r] = S[(r + B1 * tc) * C + cc];
    col::group_sync(group);  // every thread of the group holds its inputs: the stage is free
    if (gt == 0) {
      const long long nxt = blockIdx.x + (k + RING) * gridDim.x;
      if (nxt < total_tiles) {
        col::fence_proxy_async();
        issue(nxt, st);
      }
      col::mbar_arrive(&freed[st]);
    }
#pragma unroll
    for (int r = 1; r < N2; ++r) v[r] = cmul(v[r], ldg_cx<T>(p.tw, tc * r));
    DFT<N2, T>::run(v);
    if (c0 + cc < p.nb[0]) {
      const long long ob = p.ooff + (long long)(c0 + cc) * p.obd[0] + (long long)b1 * p.obd[1] +
                           (long long)b2 * p.obd[2] + (long long)b3 * p.obd[3];
      long long gidx = 0;
      if (p.gtw_dim >= 0) gidx = p.gtw_dim == 0 ? c0 + cc : (p.gtw_dim == 1 ? b1 : (p.gtw_dim == 2 ? b2 : b3));
#pragma unroll
      for (int r = 0; r < N2; ++r) {
        const int kk = tc + N1 * r;
        cx<T> o = v[r];
        if (p.gtw_dim >= 0) {
          const long long m = gidx * kk;
          o = cmul(o, cmul(ldg_cx<T>(p.gtw_hi, m >> p.gtw_bits), ldg_cx<T>(p.gtw_lo, m & gmask)));
        }
        if (p.apply_scale) o = cscale(o, scale);
        gstore<T>(p, fl, ob + (long long)kk * p.os, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// 5-D view of the pass input: (column, row j, b1, b2, b3); element = one complex number described as 2 scalars
// folded into the innermost dimension (so that fp32 and fp64 both use a native TMA data type)
bool col_make_tensor_map(const PassParams& p, bool is_double, int C, int box_rows, CUtensorMap* map) {
  // L2 promotion of the strided 128-byte row segments (PFFT_COL_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B)
  static const int promo = [] {
    const char* e = std::getenv("PFFT_COL_L2PROMO");
    return e ? std::atoi(e) : 3;  // measured: 256 B is never slower, C4 1.7 % faster (profiles/r1_ab_variants.txt)
  }();
  const size_t esz = is_double ? 16 : 8;
  return col_make_tensor_map_plane(p, reinterpret_cast<const char*>(p.in_re) + (size_t)p.ioff * esz, 2, is_double, C,
                                   box_rows, promo, map);
}

// The same view over one scalar plane (`scalars` = 1: split storage, element = one scalar) or over interleaved pairs
// (`scalars` = 2); `base` = address of element ioff.  promo: 0 none, 1 64 B, 2 128 B, 3 256 B.
bool col_make_tensor_map_plane(const PassParams& p, const void* base_ptr, int scalars, bool is_double, int C,
                               int box_rows, int promo, CUtensorMap* map) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return false;
  const size_t sc = is_double ? 8 : 4, esz = scalars * sc;
  const char* base = reinterpret_cast<const char*>(base_ptr);
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return false;
  cuuint64_t dims[5] = {(cuuint64_t)p.nb[0] * scalars, (cuuint64_t)p.n, (cuuint64_t)p.nb[1], (cuuint64_t)p.nb[2],
                        (cuuint64_t)p.nb[3]};
  const long long st[4] = {p.is, p.nb[1] > 1 ? p.ibd[1] : p.is * p.n, p.nb[2] > 1 ? p.ibd[2] : p.is * p.n,
                           p.nb[3] > 1 ? p.ibd[3] : p.is * p.n};
  cuuint64_t strides[4];
  for (int i = 0; i < 4; ++i) {
    const unsigned long long bytes = (unsigned long long)st[i] * esz;
    if (st[i] <= 0 || bytes % 16 != 0 || bytes >= (1ULL << 40)) return false;
    strides[i] = bytes;
  }
  for (int i = 0; i < 5; ++i)
    if (dims[i] == 0 || dims[i] > (1ULL << 32)) return false;
  cuuint32_t box[5] = {(cuuint32_t)(scalars * C), (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapL2promotion pr =
      promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                 : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                               : (promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
  const CUresult r = enc(map, is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                         const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Persistent grid: one CTA per resident slot (occupancy of this instantiation x SMs), never more than there are tiles.
template <typename T, int N1, int N2, int N3, int IN, bool OUT_ROWS, bool INPLACE>
static cudaError_t launch_col_v(const PassParams& p, bool swap, const CUtensorMap& map, cudaStream_t stream) {
  using Cfg = ColCfg<T, N1, N2, N3, IN, INPLACE>;
  auto kern = wg_col_kernel<T, N1, N2, N3, IN, OUT_ROWS, INPLACE>;
  int slots = persistent_slots(kern, Cfg::NT, Cfg::kSmem);  // cached per (kernel, device)
  if (slots <= 0) return cudaErrorLaunchOutOfResources;
  static const int cap = [] {  // experiment knob: fewer resident CTAs per SM than the occupancy allows
    const char* e = std::getenv("PFFT_COL_CTAS_PER_SM");
    return e ? std::atoi(e) : 0;
  }();
  if (cap > 0) slots = std::min(slots, cap * sm_count());
  const long long tiles = ((p.nb[0] + Cfg::C - 1) / Cfg::C) * p.nb[1] * p.nb[2] * p.nb[3];
  const int grid = (int)(tiles < slots ? tiles : slots);
  kern<<<grid, Cfg::NT, Cfg::kSmem, stream>>>(p, map, swap);
  return cudaGetLastError();
}

static bool inplace_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("PFFT_COL_INPLACE");
    return e ? std::atoi(e) != 0 : true;
  }();
  return on;
}

// variant: bits 0-1 = input mode, bit 2 = OUT_ROWS
// tensor map of the pass input: from the plan's per-pass cache when the base address is the one it was encoded for
static bool col_tensor_map_cached(const PassParams& p, bool is_double, int C, int box_rows, ColMapCache* cache,
                                  CUtensorMap* map) {
  static_assert(sizeof(CUtensorMap) == sizeof(ColMapCache::map), "CUtensorMap is 128 bytes");
  const void* base = reinterpret_cast<const char*>(p.in_re) + (size_t)p.ioff * (is_double ? 16 : 8);
  if (cache != nullptr && cache->base == base) {
    memcpy(map, cache->map, sizeof(CUtensorMap));
    return true;
  }
  if (!col_make_tensor_map(p, is_double, C, box_rows, map)) return false;
  if (cache != nullptr) {
    memcpy(cache->map, map, sizeof(CUtensorMap));
    cache->base = base;
  }
  return true;
}

template <typename T, int N1, int N2, int N3>
static cudaError_t launch_col_t(const PassParams& p, bool swap, int variant, cudaStream_t stream, bool* used,
                                ColMapCache* cache) {
  *used = false;
  int in = variant & 3;
  const bool out_rows = (variant & 4) != 0;
  CUtensorMap map;
  if (in == IN_COLS_TMA) {
    using Cfg = ColCfg<T, N1, N2, N3, IN_COLS_TMA>;
    if (!col_tensor_map_cached(p, sizeof(T) == 8, Cfg::C, Cfg::kBoxRows, cache, &map)) return cudaSuccess;  // caller falls back
  } else {
    memset(&map, 0, sizeof(map));
  }
  if (in == IN_ROWS_BULK) {
    const uintptr_t base = reinterpret_cast<uintptr_t>(p.in_re) + (size_t)p.ioff * 2 * sizeof(T);
    if (base % 16 != 0) in = IN_ROWS_DIRECT;
  }
  *used = true;
  if (in == IN_COLS_TMA) {
    // columns in, columns out: exchange inside the stage buffer where the geometry allows (see ColCfg)
    constexpr int kTpc = col::cmin(col::cmin(N1, N2), 16);
    // measured on B200 (tools/ab_run.sh): fp32 N = 256 (L1D) 0.946 -> 0.823 ms with three in-place CTAs per SM, fp64
    // N = 256 (C4) 2.337 -> 2.316 ms; fp32 N = 512 (one 64 KiB stage, two CTAs per SM, no prefetch inside the CTA)
    // 1.178 -> 1.312 ms on 512^3, so that size keeps the exchange buffer and the two-stage ring
    constexpr bool kCanInplace = N3 == 1 && N1 == kTpc && (N2 % kTpc) == 0 && (size_t)N1 * N2 * 128 <= 48 * 1024;
    if constexpr (kCanInplace) {
      if (!out_rows && inplace_enabled()) return launch_col_v<T, N1, N2, N3, IN_COLS_TMA, false, true>(p, swap, map, stream);
    }
    return out_rows ? launch_col_v<T, N1, N2, N3, IN_COLS_TMA, true, false>(p, swap, map, stream)
                    : launch_col_v<T, N1, N2, N3, IN_COLS_TMA, false, false>(p, swap, map, stream);
  }
  if (in == IN_ROWS_BULK)
    return out_rows ? launch_col_v<T, N1, N2, N3, IN_ROWS_BULK, true, false>(p, swap, map, stream)
                    : launch_col_v<T, N1, N2, N3, IN_ROWS_BULK, false, false>(p, swap, map, stream);
  return out_rows ? launch_col_v<T, N1, N2, N3, IN_ROWS_DIRECT, true, false>(p, swap, map, stream)
                  : launch_col_v<T, N1, N2, N3, IN_ROWS_DIRECT, false, false>(p, swap, map, stream);
}

static bool col512_groups_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("PFFT_COL512_GROUPS");
    return e ? std::atoi(e) != 0 : true;
  }();
  return on;
}

static cudaError_t launch_col512(const PassParams& p, bool swap, cudaStream_t stream, bool* used, ColMapCache* cache) {
  *used = false;
  CUtensorMap map;
  if (!col_tensor_map_cached(p, false, Col512::C, 256, cache, &map)) return cudaSuccess;  // caller falls back
  *used = true;
  const int sms = sm_count();
  if (sms <= 0) return cudaErrorLaunchOutOfResources;
  cudaError_t e = ensure_dynamic_smem(wg_col512_kernel, Col512::kSmem);
  if (e != cudaSuccess) return e;
  const long long tiles = ((p.nb[0] + Col512::C - 1) / Col512::C) * p.nb[1] * p.nb[2] * p.nb[3];
  const int grid = (int)(tiles < sms ? tiles : sms);
  wg_col512_kernel<<<grid, Col512::NT, Col512::kSmem, stream>>>(p, map, swap);
  return cudaGetLastError();
}

bool col_supported(int n, bool is_double, int* n1, int* n2) {
  int a = 0, b = 0;
  switch (n) {
    case 64: a = 8; b = 8; break;
    case 128: a = 16; b = 8; break;
    case 256: a = 16; b = 16; break;
    case 512: a = 16; b = 32; break;
    default: return false;
  }
  if (n1) *n1 = a;
  if (n2) *n2 = b;
  return true;
}

int col_tile_columns(int n, bool is_double) {
  (void)n;
  return 128 / (is_double ? 16 : 8);
}

size_t col_smem_bytes(int n, bool is_double, bool ring) {
  const size_t esz = is_double ? 16 : 8;
  const int c = col_tile_columns(n, is_double);
  const int pitch = is_double ? col::pitch<double>(n) : col::pitch<float>(n);
  return (ring ? 2 : 0) * (size_t)n * c * esz + (size_t)c * pitch * esz + 64;
}

int col_threads(int n, bool is_double) {
  int a, b;
  if (!col_supported(n, is_double, &a, &b)) return 0;
  const int tpc = a < b ? a : b;  // (the three-pass N = 512 variant runs 64 threads per transform)
  return col_tile_columns(n, is_double) * (n == 512 && is_double ? 64 : (tpc < 16 ? tpc : 16));
}

// *used == false on return with cudaSuccess: the tensor map could not be built (alignment); run the generic kernel
cudaError_t launch_wg_col(const PassParams& p, bool is_double, bool swap, int variant, int grid, cudaStream_t stream,
                          bool* used, ColMapCache* cache) {
  (void)grid;  // the planner's estimate; the launch sizes the persistent grid from the kernel's actual occupancy
#define PFFT_COL(NN, A, B, CC)                                                                   \
  case NN:                                                                                       \
    return is_double ? launch_col_t<double, A, B, CC>(p, swap, variant, stream, used, cache) \
                     : launch_col_t<float, A, B, CC>(p, swap, variant, stream, used, cache);
  *used = false;
  switch (p.n) {
    PFFT_COL(64, 8, 8, 1)
    PFFT_COL(128, 16, 8, 1)
    PFFT_COL(256, 16, 16, 1)
    case 512: {
      // fp32: radix 16 x 32 (measured 1.41 ms on 512^3 against 1.45 ms for three radix-8 passes, PFFT_COL512=3);
      // fp64: three radix-8 passes (a radix-32 fp64 butterfly does not fit the register file)
      static const int three_pass = [] {
        const char* e = std::getenv("PFFT_COL512");
        return e ? std::atoi(e) == 3 : 0;
      }();
      if (!three_pass && !is_double && (variant & 3) == IN_COLS_TMA && (variant & 4) == 0 && col512_groups_enabled())
        return launch_col512(p, swap, stream, used, cache);
      if (!three_pass && !is_double) return launch_col_t<float, 16, 32, 1>(p, swap, variant, stream, used, cache);
      return is_double ? launch_col_t<double, 8, 8, 8>(p, swap, variant, stream, used, cache)
                       : launch_col_t<float, 8, 8, 8>(p, swap, variant, stream, used, cache);
    }
    default:
      return cudaErrorInvalidValue;
  }
#undef PFFT_COL
}

}  // namespace pfft
