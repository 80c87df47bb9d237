// GLOBAL level, two passes in ONE persistent kernel with the intermediate result L2 resident.
//
// The multi-pass (four-step) plans cross HBM once per factor: pass p writes the whole batch to the workspace, pass p+1
// reads it back.  But pass p+1 only needs a *chunk* of what pass p wrote -- for N = N_1 * N_2 one transform, for the
// last two factors of N = N_1 * N_2 * N_3 one N_2*N_3 sub-sequence -- and a chunk of a few MiB stays in the 126 MB L2.
// This kernel runs both passes as one work list, chunk by chunk:
//
//     A(0) .. A(lead-1) | A(lead) B(0) | A(lead+1) B(1) | ... | B(NC-lead) .. B(NC-1)
//
// A(c) = the column tiles of pass p on chunk c (TMA tensor tiles in, in-place two-radix exchange, inter-factor
// twiddle on store), written to slot c % NS of a small RING buffer instead of the full-size workspace; B(c) = the row
// tiles of pass p+1 on chunk c (cp.async.bulk rows in from the ring, transposing column store to the destination).
// The list is dealt round-robin to a persistent grid (item i -> CTA i % grid), every CTA works through its items in
// order, and two per-chunk arrival counters in global memory replace the kernel boundary:
//     B(c) tiles load only after done_a[c] has counted every A(c) tile (release/acquire + cross-proxy fence, since the
//     consumer reads through the async proxy), and A(c) tiles only after done_b[c - NS] freed their ring slot.
// Dependencies always point to earlier items, a CTA blocks only at the head of its own list (later items are issued
// ahead only when their dependency is already satisfied), and the grid is co-resident, so the earliest unfinished
// item can always run: no deadlock.  The ring slots are rewritten while still dirty in L2, so HBM sees one read of the
// source and one write of the destination: 1 round trip for N = 65536 (was 2), 2 for N = 2^24 (was 3).
//
// Reference counterpart: the `num_batches_in_l2` launch groups of the GLOBAL driver
// (/root/reference/src/portfft/dispatcher/global_dispatcher.hpp:343-408) -- there a host loop of kernel launches per
// batch group, here arrival counters inside one launch.
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

namespace fz {
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// L2 eviction priorities: the source and the destination stream through once (evict first), the ring must stay
// resident until its slot is rewritten (evict last) -- otherwise every dirty ring line is written back to HBM after its
// only use (measured without hints: 2.0 GB written for 1.07 GB of output on 65536 x 2048)
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_hint(cx<float>* a, cx<float> v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(a), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(cx<double>* a, cx<double> v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_5d_hint(void* dst, const CUtensorMap* map, int c0, int r0, int b1, int b2, int b3,
                                                 uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, "
      "%3, %4, %5, %6}], [%7], %8;" ::"r"(col::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(b1), "r"(b2), "r"(b3), "r"(col::smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          col::smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(col::smem_u32(bar)), "l"(pol)
      : "memory");
}
// generic-proxy writes (other CTAs' stores, made visible by the acquire above) -> async-proxy reads (TMA) and back
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
}  // namespace fz

template <typename T>
struct FusedCfg {
  static constexpr int N = 256, N1 = 16, N2 = 16, B1 = N / N1;
  static constexpr int C = 128 / (2 * (int)sizeof(T));  // transforms per tile: one 128-byte line per row segment
  static constexpr int TPC = 16;
  static constexpr int NC = C * TPC;      // consumer (butterfly) threads
  static constexpr int NT = NC + 32;      // + one service warp (loads, dependencies, completion signals)
  static constexpr int NW = NC / 32;      // consumer warps
  static constexpr int RING = 2;
  // Pass B keeps each contiguous row at a padded pitch and exchanges in place: 16-byte aligned rows for cp.async.bulk,
  // and a pitch of 4 banks (mod 32) so that the quarter-warp wavefronts of the 128-bit second-pass loads (lanes along
  // the rows) fall on distinct banks.  Pass A uses the same stage as [row j][column] without padding.
  static constexpr int PB = N + (sizeof(T) == 4 ? 2 : 1);
  static constexpr size_t kTileBytes = (size_t)N * C * 2 * sizeof(T);   // 32 KiB of payload for both precisions
  static constexpr size_t kStageBytes = (size_t)PB * C * 2 * sizeof(T);  // 33 024 / 32 896 bytes
  static constexpr size_t kSmem = RING * kStageBytes + 128;
  static constexpr int kBoxRows = 256;
  static constexpr int kMinBlocks = 3;
};

namespace fz {
__device__ __forceinline__ void consumer_sync(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_cta(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(col::smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_cta(unsigned* p) {
  asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(col::smem_u32(p)) : "memory");
}
}  // namespace fz

// Work of one CTA: items blockIdx.x + k * gridDim.x of the list.  The SERVICE warp (one lane) never blocks: it polls
// (1) completion counts of the consumer warps and turns them into release-increments of the per-chunk arrival
// counters, (2) free stages, posting the load of the next item as soon as its dependency is satisfied.  The CONSUMER
// warps wait only on the stage mbarriers.  Nothing in a CTA waits for another CTA except the service lane's polling,
// and that never holds back a completion signal: the earliest unfinished item of the whole list can always proceed.
template <typename T, int B_MODE>
__global__ void __launch_bounds__(FusedCfg<T>::NT, FusedCfg<T>::kMinBlocks)
    wg_fused2_kernel(const PassParams pa, const PassParams pb, const __grid_constant__ CUtensorMap tmap_a,
                     const FusedArgs fa, const bool swap_a, const bool swap_b) {
  using Cfg = FusedCfg<T>;
  constexpr int N = Cfg::N, N1 = Cfg::N1, N2 = Cfg::N2, B1 = Cfg::B1, C = Cfg::C, TPC = Cfg::TPC, PB = Cfg::PB;
  constexpr int RING = Cfg::RING;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage0 = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + RING * Cfg::kStageBytes);
  unsigned* rel_cnt = reinterpret_cast<unsigned*>(full + RING);  // consumer warps that handed a stage back, ever
  unsigned* done_cnt = rel_cnt + 1;                              // consumer warps that finished a tile, ever
  int4* info = reinterpret_cast<int4*>(rel_cnt + 4);             // per stage: phase, chunk, local tile
  const int tid = threadIdx.x;
  cx<T>* ring = reinterpret_cast<cx<T>*>(fa.ring);
  const unsigned per = (unsigned)(fa.tiles_a + fa.tiles_b), ta = (unsigned)fa.tiles_a, tb = (unsigned)fa.tiles_b;
  const unsigned total_items = (unsigned)fa.num_chunks * per;
  const unsigned my_items = total_items > blockIdx.x ? (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const unsigned tca = (unsigned)((pa.nb[0] + C - 1) / C);  // column tiles of pass A per chunk index q
  auto slot_base = [&](int chunk) { return (long long)(chunk % fa.slots) * fa.group * fa.unit; };

  if (tid == 0) {
    for (int s = 0; s < RING; ++s) col::mbar_init(&full[s], 1);
    *rel_cnt = 0;
    *done_cnt = 0;
    col::fence_mbar_init();
    col::fence_proxy_async();
  }
  __syncthreads();

  if (tid >= Cfg::NC) {
    // =========================== service warp ================================================================
    if (tid != Cfg::NC) return;
    const unsigned head = (unsigned)fa.lead * ta;            // prologue: A(0) .. A(lead-1)
    const unsigned body_chunks = (unsigned)(fa.num_chunks - fa.lead);  // then (A(m), B(m - lead)) pairs
    auto decode = [&](unsigned i, int& phase, int& chunk, int& local) {
      if (i < head) {
        phase = 0;
        chunk = (int)(i / ta);
        local = (int)(i - chunk * ta);
        return;
      }
      const unsigned r = i - head, m = r / per;
      if (m < body_chunks) {
        const unsigned rem = r - m * per;
        phase = rem < ta ? 0 : 1;
        chunk = (int)(rem < ta ? m + fa.lead : m);
        local = (int)(rem < ta ? rem : rem - ta);
        return;
      }
      const unsigned r2 = r - body_chunks * per;  // epilogue: B(NC - lead) .. B(NC - 1)
      phase = 1;
      chunk = (int)(body_chunks + r2 / tb);
      local = (int)(r2 % tb);
    };
    auto ready = [&](int phase, int chunk) -> bool {
      if (phase == 0) {
        if (chunk < fa.slots) return true;
        return fz::ld_acquire(fa.done_b + (chunk - fa.slots)) >= fa.epoch * (unsigned long long)tb;
      }
      return fz::ld_acquire(fa.done_a + chunk) >= fa.epoch * (unsigned long long)ta;
    };
    const uint64_t pol_stream = fz::policy_evict_first(), pol_ring = fz::policy_evict_last();
    auto issue = [&](int phase, int chunk, int local, int s) {
      unsigned char* dst = stage0 + s * Cfg::kStageBytes;
      info[s] = make_int4(phase, chunk, local, 0);
      fz::fence_proxy_async_all();
      col::mbar_expect_tx(&full[s], (uint32_t)Cfg::kTileBytes);
      if (phase == 0) {
        const int ct = (int)((unsigned)local % tca);
        const int q = chunk * fa.group + (int)((unsigned)local / tca);
        fz::tma_load_5d_hint(dst, &tmap_a, 2 * ct * C, 0, q, 0, 0, &full[s], pol_stream);  // innermost: scalars
      } else {
        // C rows of N contiguous elements each, to the padded row pitch of the stage
        const cx<T>* src;
        long long step;
        if (B_MODE == 0) {  // consecutive rows of one chunk index q
          const unsigned tcb = (unsigned)(pb.nb[0] / C);
          const unsigned ql = (unsigned)local / tcb;  // q - chunk * group
          const unsigned r0 = ((unsigned)local - ql * tcb) * C;
          src = ring + slot_base(chunk) + (long long)ql * fa.unit + (long long)r0 * N;
          step = N;
        } else {  // row r of C consecutive chunk indices q
          const unsigned gpc = (unsigned)fa.group / C;
          const unsigned r = (unsigned)local / gpc, g8 = (unsigned)local - r * gpc;
          src = ring + slot_base(chunk) + (long long)g8 * C * fa.unit + (long long)r * N;
          step = fa.unit;
        }
#pragma unroll 1
        for (int j = 0; j < C; ++j)
          fz::bulk_g2s_hint(dst + (size_t)j * PB * sizeof(cx<T>), src + (long long)j * step, (uint32_t)(N * sizeof(cx<T>)),
                            &full[s], pol_ring);
      }
    };
    unsigned next = 0, sig = 0;
    int nph = 0, nch = 0, nlo = 0;  // decoded item `next`
    if (my_items) decode(blockIdx.x, nph, nch, nlo);
    while (sig < my_items) {
      bool progress = false;
      if (fz::ld_acquire_cta(done_cnt) >= (sig + 1) * (unsigned)Cfg::NW) {
        // tile `sig` is complete: every consumer warp's stores precede its count (release.cta), the fence makes them
        // visible at gpu scope (and to the async proxy of the consumer CTA) before the arrival counter moves
        int ph, ch, lo;
        decode(blockIdx.x + sig * gridDim.x, ph, ch, lo);
        __threadfence();
        fz::fence_proxy_async_all();
        fz::red_release_add((ph == 0 ? fa.done_a : fa.done_b) + ch, 1ULL);
        ++sig;
        progress = true;
      }
      if (next < my_items && fz::ld_acquire_cta(rel_cnt) >= (next < RING ? 0u : (next - RING + 1) * (unsigned)Cfg::NW) &&
          ready(nph, nch)) {
        issue(nph, nch, nlo, (int)(next % RING));
        ++next;
        if (next < my_items) decode(blockIdx.x + next * gridDim.x, nph, nch, nlo);
        progress = true;
      }
      if (!progress) __nanosleep(32);
    }
    return;
  }

  // =========================== consumer warps ==================================================================
  const int cc = tid % C, tc = tid / C;      // lanes along the transform index (column accesses)
  const int cr = tid / TPC, tr = tid % TPC;  // lanes along the element index (row accesses)
  const int lane = tid & 31;
  const T scale_b = T(pb.scale);
  const unsigned long long gmask = (1ULL << pa.gtw_bits) - 1;
  const uint64_t pol_stream = fz::policy_evict_first(), pol_ring = fz::policy_evict_last();

  for (unsigned k = 0; k < my_items; ++k) {
    const int st = (int)(k % RING);
    cx<T>* S = reinterpret_cast<cx<T>*>(stage0 + st * Cfg::kStageBytes);
    col::mbar_wait(&full[st], (uint32_t)((k / RING) & 1));
    const int4 inf = info[st];
    const int phase = inf.x, chunk = inf.y, local = inf.z;
    if (phase == 0) {
      // ---- pass A: strided columns (stage [row][column]), radix 16 x 16 exchanged inside the stage ----------------
      const int ct = (int)((unsigned)local % tca);
      const int ql = (int)((unsigned)local / tca);
      {
        const int j = tc;  // B1 == TPC: one first-pass butterfly per thread
        cx<T> v[N1];
#pragma unroll
        for (int r = 0; r < N1; ++r) v[r] = S[(j + B1 * r) * C + cc];
        if (swap_a) {
#pragma unroll
          for (int r = 0; r < N1; ++r) v[r] = cx<T>{v[r].y, v[r].x};
        }
        DFT<N1, T>::run(v);
#pragma unroll
        for (int r = 0; r < N1; ++r) S[(j + B1 * r) * C + cc] = v[r];
      }
      fz::consumer_sync(Cfg::NC);
      cx<T> v[N2];
      const int j = tc;
#pragma unroll
      for (int r = 0; r < N2; ++r) v[r] = S[(r + B1 * j) * C + cc];
      __syncwarp();
      if (lane == 0) fz::red_release_cta(rel_cnt);  // this warp has taken its inputs out of the stage
#pragma unroll
      for (int r = 1; r < N2; ++r) v[r] = cmul(v[r], ldg_cx<T>(pa.tw, j * r));
      DFT<N2, T>::run(v);
      const int col_idx = ct * C + cc;
      if (col_idx < pa.nb[0]) {
        // inter-factor twiddle w_M^{column * k}, k = j + 16 r (fp64: running product from two look-ups, see wg_col.cu)
        cx<T>* out = ring + slot_base(chunk) + (long long)ql * fa.unit + col_idx;
        cx<T> tw_run{T(1), T(0)}, tw_step{T(1), T(0)};
        const unsigned gidx = (unsigned)col_idx;
        if (sizeof(T) == 8) {
          tw_run = gtw_lookup<T>(pa, gidx, (unsigned)j, gmask);
          tw_step = gtw_lookup<T>(pa, gidx, (unsigned)N1, gmask);
        }
#pragma unroll
        for (int r = 0; r < N2; ++r) {
          const int kk = j + N1 * r;
          cx<T> o = v[r];
          if (sizeof(T) == 8) {
            o = cmul(o, tw_run);
            tw_run = cmul(tw_run, tw_step);
          } else {
            o = cmul(o, gtw_lookup<T>(pa, gidx, (unsigned)kk, gmask));
          }
          fz::st_hint(out + (long long)kk * pa.os, o, pol_ring);  // plan-internal data: no swap, no scale
        }
      }
    } else {
      // ---- pass B: contiguous rows (stage [row][j] at pitch PB), radix 16 x 16 exchanged inside the stage ----------
      {
        const int j = tr;
        cx<T> v[N1];
#pragma unroll
        for (int r = 0; r < N1; ++r) v[r] = S[cr * PB + j + B1 * r];
        DFT<N1, T>::run(v);
#pragma unroll
        for (int r = 0; r < N1; ++r) S[cr * PB + j + B1 * r] = v[r];  // output r of butterfly j: element j + 16 r
      }
      fz::consumer_sync(Cfg::NC);
      long long ob;
      bool live;
      if (B_MODE == 0) {
        const unsigned tcb = (unsigned)(pb.nb[0] / C);
        const unsigned ql = (unsigned)local / tcb;
        const int r0 = (int)(((unsigned)local - ql * tcb) * C);
        const long long q = (long long)chunk * fa.group + ql;
        live = r0 + cc < pb.nb[0];
        ob = pb.ooff + (long long)(r0 + cc) * pb.obd[0] + q * pb.obd[1];
      } else {
        const unsigned gpc = (unsigned)fa.group / C;
        const unsigned r = (unsigned)local / gpc, g8 = (unsigned)local - r * gpc;
        const unsigned q0 = (unsigned)chunk * fa.group + g8 * C;  // first chunk index of the tile
        const unsigned qh = q0 / (unsigned)pb.nb[0], qlow = q0 - qh * (unsigned)pb.nb[0];
        live = true;
        ob = pb.ooff + (long long)(qlow + cc) * pb.obd[0] + (long long)r * pb.obd[1] + (long long)qh * pb.obd[2];
      }
      {
        // butterfly j of row cc takes output j of every first-pass butterfly: 16 consecutive elements from 16 j on
        const int j = tc;
        cx<T> v[N2];
        if (sizeof(T) == 4) {
          const float4* src = reinterpret_cast<const float4*>(S + cc * PB + B1 * j);
#pragma unroll
          for (int r = 0; r < N2 / 2; ++r) {
            const float4 t = src[r];
            v[2 * r] = cx<T>{(T)t.x, (T)t.y};
            v[2 * r + 1] = cx<T>{(T)t.z, (T)t.w};
          }
        } else {
#pragma unroll
          for (int r = 0; r < N2; ++r) v[r] = S[cc * PB + B1 * j + r];
        }
        __syncwarp();
        if (lane == 0) fz::red_release_cta(rel_cnt);  // this warp has taken its inputs out of the stage
#pragma unroll
        for (int r = 1; r < N2; ++r) v[r] = cmul(v[r], ldg_cx<T>(pb.tw, j * r));
        DFT<N2, T>::run(v);
        if (live) {
          cx<T>* out = reinterpret_cast<cx<T>*>(pb.out_re);
#pragma unroll
          for (int r = 0; r < N2; ++r) {
            cx<T> o = v[r];
            if (pb.apply_scale) o = cscale(o, scale_b);
            if (swap_b) o = cx<T>{o.y, o.x};
            fz::st_hint(out + ob + (long long)(j + N1 * r) * pb.os, o, pol_stream);
          }
        }
      }
    }
    // this warp's stores of the tile are issued (A: ring slot written; B: ring slot read, destination written)
    __syncwarp();
    if (lane == 0) fz::red_release_cta(done_cnt);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}

// Can passes (a, b) of a plan run fused?  a: 256-point TMA column pass with inter-factor twiddle writing packed
// columns of the workspace, b: 256-point row pass reading exactly those rows.  Fills the chunk geometry.
bool fused2_plan(const PassParams& a, int variant_a, const PassParams& b, int variant_b, bool is_double, int grid,
                 FusedGeom* g) {
  // Opt-in (PFFT_FUSE=1).  Measured on B200 (profiles/r2_ab_variants.txt): the fused kernel halves the DRAM traffic of
  // 65536 x 2048 fp32 (ncu: 1.11 GB read + 1.50 GB written against 2.15 + 2.06 GB for the two launches) but is not
  // faster, 0.848 ms against 0.720 ms: the two separate passes already run at 82 % / 98 % of the HBM roofline and the
  // work per tile is instruction-issue and latency bound (about one 256 x 16 tile per microsecond and SM either way),
  // so taking DRAM time away buys nothing until the per-tile cost drops.  fp64 2^24 x 8: 2.42 ms fused, 2.27 ms apart.
  if (!env_int("PFFT_FUSE", 0) || env_int("PFFT_NO_FUSE", 0)) return false;
  const int C = is_double ? 8 : 16;
  if (a.n != 256 || b.n != 256) return false;
  if ((variant_a & 7) != 0 || (variant_b & 7) != 2) return false;  // cols_tma -> cols ; rows_bulk -> cols
  if (a.gtw_dim != 0 || b.gtw_dim >= 0 || a.peer_dim >= 0 || b.peer_dim >= 0) return false;
  if (a.apply_scale || a.valid_in || a.valid_out || b.valid_in || b.valid_out || a.mod_flags || b.mod_flags) return false;
  if (a.nb[2] != 1 || a.nb[3] != 1 || a.nb[0] % C != 0) return false;
  const long long W = a.nb[0], Q = a.nb[1], unit = W * 256;
  if (a.obd[0] != 1 || a.os != W || a.ooff != 0 || (Q > 1 && a.obd[1] != unit)) return false;
  if (b.is != 1 || b.ioff != 0) return false;
  int mode;
  if (b.nb[0] == W && b.ibd[0] == 256 && b.nb[1] == Q && (Q == 1 || b.ibd[1] == unit) && b.nb[2] == 1 && b.nb[3] == 1) {
    mode = 0;  // rows of one chunk index in dimension 0
  } else if (b.ibd[0] == unit && b.nb[1] == W && b.ibd[1] == 256 && b.nb[0] * b.nb[2] == Q && b.nb[3] == 1 &&
             (b.nb[2] == 1 || b.ibd[2] == unit * b.nb[0]) && b.nb[0] % C == 0) {
    mode = 1;  // C consecutive chunk indices in dimension 0, rows in dimension 1
  } else {
    return false;
  }
  // chunk = `group` consecutive chunk indices q.  Sized so that one chunk's tiles of either pass cover the persistent
  // grid about once (the dependency of B(c) is then long satisfied when its turn comes) while `slots` chunks stay a
  // small part of the L2.
  const long long esz = is_double ? 16 : 8;
  const long long target = (long long)env_int("PFFT_FUSE_CHUNK_KB", 4096) * 1024;
  long long group = mode == 1 ? C : 1;  // (mode 1: a tile takes C consecutive chunk indices)
  const long long qdiv = mode == 1 ? b.nb[0] : Q;  // a chunk must not straddle dimension 2 of pass b
  while (group * 2 * unit * esz <= target && qdiv % (group * 2) == 0) group *= 2;
  if (qdiv % group != 0) return false;
  const long long chunks = Q / group;
  // pass a runs `lead` chunks ahead of pass b.  A CTA loads its items RING deals ahead; the load of a B(c) tile can be
  // posted that early only if A(c) is complete by then, i.e. if more than RING deals of the grid lie between the two
  // segments (measured with lead = 1 and segments shorter than the grid: every B tile loaded late, 1.21 ms against
  // 0.82 ms unfused on 65536 x 2048).
  const long long tiles_a = group * (W / C);
  const long long dist = (long long)env_int("PFFT_FUSE_LEAD_TENTHS", 20) * grid / 10;  // items between the segments
  int lead = env_int("PFFT_FUSE_LEAD", 0);
  if (lead <= 0) lead = (int)std::max<long long>(1, (dist + tiles_a - 1) / tiles_a);
  // ring slots: A(c) overwrites the slot of chunk c - slots, whose B tiles must be long finished as well
  const long long tiles_b = mode == 0 ? group * (b.nb[0] / C) : (group / C) * b.nb[1];
  const int slots = lead + 1 + (int)std::max<long long>(1, (dist + tiles_a + tiles_b - 1) / (tiles_a + tiles_b));
  if (chunks < 2 * slots) return false;  // too little work for the pipeline to matter
  g->mode = mode;
  g->group = (int)group;
  g->lead = lead;
  g->slots = slots;
  g->num_chunks = chunks;
  g->unit = unit;
  g->tiles_a = tiles_a;
  g->tiles_b = tiles_b;
  g->ring_bytes = (size_t)slots * group * unit * esz;
  return true;
}

template <typename T>
static int fused_grid_t() {
  using Cfg = FusedCfg<T>;
  int slots = persistent_slots(wg_fused2_kernel<T, 0>, Cfg::NT, Cfg::kSmem);
  const int cap = env_int("PFFT_FUSE_CTAS_PER_SM", 0);
  if (cap > 0) slots = std::min(slots, cap * sm_count());
  return slots;
}

// persistent grid of the fused kernel on the current device (both row modes have the same footprint)
int fused2_grid(bool is_double) { return is_double ? fused_grid_t<double>() : fused_grid_t<float>(); }

template <typename T, int B_MODE>
static cudaError_t launch_fused_t(const PassParams& a, const PassParams& b, const FusedArgs& fa, bool swap_a, bool swap_b,
                                  cudaStream_t stream, bool* used) {
  using Cfg = FusedCfg<T>;
  *used = false;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (!col_make_tensor_map(a, sizeof(T) == 8, Cfg::C, Cfg::kBoxRows, &map)) return cudaSuccess;  // caller runs the passes apart
  auto kern = wg_fused2_kernel<T, B_MODE>;
  const int slots = fused_grid_t<T>();
  if (slots <= 0) return cudaErrorLaunchOutOfResources;
  cudaError_t es = ensure_dynamic_smem(kern, Cfg::kSmem);
  if (es != cudaSuccess) return es;
  const long long items = fa.num_chunks * (fa.tiles_a + fa.tiles_b);
  const int grid = (int)(items < slots ? items : slots);
  *used = true;
  kern<<<grid, Cfg::NT, Cfg::kSmem, stream>>>(a, b, map, fa, swap_a, swap_b);
  return cudaGetLastError();
}

cudaError_t launch_wg_fused2(const PassParams& a, const PassParams& b, const FusedArgs& fa, int mode, bool is_double,
                             bool swap_a, bool swap_b, cudaStream_t stream, bool* used) {
  if (is_double)
    return mode == 0 ? launch_fused_t<double, 0>(a, b, fa, swap_a, swap_b, stream, used)
                     : launch_fused_t<double, 1>(a, b, fa, swap_a, swap_b, stream, used);
  return mode == 0 ? launch_fused_t<float, 0>(a, b, fa, swap_a, swap_b, stream, used)
                   : launch_fused_t<float, 1>(a, b, fa, swap_a, swap_b, stream, used);
}

}  // namespace pfft
