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
// generic-proxy writes (other CTAs' stores, made visible by the acquire above) -> async-proxy reads (TMA) and back
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
}  // namespace fz

template <typename T>
struct FusedCfg {
  static constexpr int N = 256, N1 = 16, N2 = 16, B1 = N / N1;
  static constexpr int C = 128 / (2 * (int)sizeof(T));  // transforms per tile: one 128-byte line per row segment
  static constexpr int TPC = 16, NT = C * TPC;
  static constexpr int PITCH = col::pitch<T>(N);
  static constexpr int RING = 2;
  static constexpr size_t kStageBytes = (size_t)N * C * 2 * sizeof(T);  // 32 KiB for both precisions
  static constexpr size_t kSmem = RING * kStageBytes + (size_t)C * PITCH * 2 * sizeof(T) + 64;
  static constexpr int kBoxRows = 256;
};

template <typename T, int B_MODE>
__global__ void __launch_bounds__(FusedCfg<T>::NT, 2)
    wg_fused2_kernel(const PassParams pa, const PassParams pb, const __grid_constant__ CUtensorMap tmap_a,
                     const FusedArgs fa, const bool swap_a, const bool swap_b) {
  using Cfg = FusedCfg<T>;
  constexpr int N = Cfg::N, N1 = Cfg::N1, N2 = Cfg::N2, B1 = Cfg::B1, C = Cfg::C, TPC = Cfg::TPC, PITCH = Cfg::PITCH;
  constexpr int RING = Cfg::RING;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage0 = smem_raw;
  cx<T>* E = reinterpret_cast<cx<T>*>(smem_raw + RING * Cfg::kStageBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(E + (size_t)C * PITCH);
  const int tid = threadIdx.x;
  const int cc = tid % C, tc = tid / C;    // lanes along the transform index (column accesses)
  const int cr = tid / TPC, tr = tid % TPC;  // lanes along the element index (row accesses)
  cx<T>* ring = reinterpret_cast<cx<T>*>(fa.ring);
  const long long per = fa.tiles_a + fa.tiles_b;
  const long long total_items = fa.num_chunks * per;
  const long long head = (long long)fa.lead * fa.tiles_a;  // prologue: A(0) .. A(lead-1)
  const long long body_chunks = fa.num_chunks - fa.lead;   // then (A(m), B(m - lead)) pairs
  const int tca = (int)((pa.nb[0] + C - 1) / C);           // column tiles of pass A per chunk index q

  // item -> (phase, chunk, local tile)
  auto decode = [&](long long i, int& phase, long long& chunk, long long& local) {
    if (i < head) {
      phase = 0;
      chunk = i / fa.tiles_a;
      local = i - chunk * fa.tiles_a;
      return;
    }
    const long long r = i - head;
    const long long m = r / per;
    if (m < body_chunks) {
      const long long rem = r - m * per;
      if (rem < fa.tiles_a) {
        phase = 0;
        chunk = m + fa.lead;
        local = rem;
      } else {
        phase = 1;
        chunk = m;
        local = rem - fa.tiles_a;
      }
      return;
    }
    const long long r2 = r - body_chunks * per;  // epilogue: B(NC - lead) .. B(NC - 1)
    phase = 1;
    chunk = body_chunks + r2 / fa.tiles_b;
    local = r2 % fa.tiles_b;
  };
  auto slot_base = [&](long long chunk) { return (chunk % fa.slots) * (long long)fa.group * fa.unit; };
  auto ready = [&](int phase, long long chunk) -> bool {
    if (phase == 0) {
      if (chunk < fa.slots) return true;
      return fz::ld_acquire(fa.done_b + (chunk - fa.slots)) >= fa.epoch * (unsigned long long)fa.tiles_b;
    }
    return fz::ld_acquire(fa.done_a + chunk) >= fa.epoch * (unsigned long long)fa.tiles_a;
  };
  // load of one item into stage s (thread 0; the item's dependency is satisfied)
  auto issue = [&](int phase, long long chunk, long long local, int s) {
    unsigned char* dst = stage0 + s * Cfg::kStageBytes;
    fz::fence_proxy_async_all();
    col::mbar_expect_tx(&full[s], (uint32_t)Cfg::kStageBytes);
    if (phase == 0) {
      const int ct = (int)(local % tca);
      const long long q = chunk * fa.group + local / tca;
      col::tma_load_5d(dst, &tmap_a, 2 * ct * C, 0, (int)q, 0, 0, &full[s]);  // innermost coordinate counts scalars
    } else if (B_MODE == 0) {
      // rows of one chunk index q: C consecutive rows = one contiguous run
      const int tcb = (int)(pb.nb[0] / C);
      const long long ql = local / tcb;  // q - chunk * group
      const int r0 = (int)(local - ql * tcb) * C;
      const cx<T>* src = ring + slot_base(chunk) + ql * fa.unit + (long long)r0 * N;
      col::bulk_g2s(dst, src, (uint32_t)Cfg::kStageBytes, &full[s]);
    } else {
      // row r of C consecutive chunk indices q
      const int gpc = fa.group / C;
      const long long r = local / gpc;
      const int g8 = (int)(local - r * gpc);
      const cx<T>* src = ring + slot_base(chunk) + (long long)g8 * C * fa.unit + r * N;
#pragma unroll 1
      for (int j = 0; j < C; ++j)
        col::bulk_g2s(dst + (size_t)j * N * sizeof(cx<T>), src + (long long)j * fa.unit, (uint32_t)(N * sizeof(cx<T>)),
                      &full[s]);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < RING; ++s) col::mbar_init(&full[s], 1);
    col::fence_mbar_init();
    col::fence_proxy_async();
  }
  __syncthreads();

  // number of items of this CTA, and the issue pump (thread 0 only): `next` = sequence number of the next item to
  // load, `released` = items whose stage has been handed back
  const long long my_items = total_items > blockIdx.x ? (total_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  long long next = 0, released = 0;
  auto pump = [&](bool block_for, long long k_block) {
    while (next < my_items && next < released + RING) {
      int ph;
      long long ch, lo;
      decode(blockIdx.x + next * gridDim.x, ph, ch, lo);
      if (!ready(ph, ch)) {
        if (!(block_for && next == k_block)) return;
        while (!ready(ph, ch)) __nanosleep(64);
      }
      issue(ph, ch, lo, (int)(next % RING));
      ++next;
    }
  };
  if (tid == 0) pump(true, 0);

  const T scale_b = T(pb.scale);
  const IoFlags fl_b{true, swap_b};
  const long long gmask = (1LL << pa.gtw_bits) - 1;

  for (long long k = 0; k < my_items; ++k) {
    int phase;
    long long chunk, local;
    decode(blockIdx.x + k * gridDim.x, phase, chunk, local);
    const int st = (int)(k % RING);
    cx<T>* S = reinterpret_cast<cx<T>*>(stage0 + st * Cfg::kStageBytes);
    if (tid == 0 && next <= k) pump(true, k);  // head of the list and not yet loaded: wait for its dependency
    col::mbar_wait(&full[st], (uint32_t)((k / RING) & 1));
    if (phase == 0) {
      // ---- pass A: strided columns (stage [row][column]), radix 16 x 16 exchanged inside the stage ----------------
      const int ct = (int)(local % tca);
      const long long ql = local / tca;
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
      __syncthreads();
      cx<T> v[N2];
      const int j = tc;
#pragma unroll
      for (int r = 0; r < N2; ++r) v[r] = S[(r + B1 * j) * C + cc];
      __syncthreads();  // every thread has taken its inputs: the stage is free
      if (tid == 0) {
        ++released;
        pump(false, 0);
      }
#pragma unroll
      for (int r = 1; r < N2; ++r) v[r] = cmul(v[r], ldg_cx<T>(pa.tw, j * r));
      DFT<N2, T>::run(v);
      const int col_idx = ct * C + cc;
      if (col_idx < pa.nb[0]) {
        // inter-factor twiddle w_M^{column * k}, k = j + 16 r (fp64: running product from two look-ups, see wg_col.cu)
        cx<T>* out = ring + slot_base(chunk) + ql * fa.unit + col_idx;
        cx<T> tw_run{T(1), T(0)}, tw_step{T(1), T(0)};
        const long long gidx = col_idx;
        if (sizeof(T) == 8) {
          const long long mb = gidx * j, ms = gidx * N1;
          tw_run = cmul(ldg_cx<T>(pa.gtw_hi, mb >> pa.gtw_bits), ldg_cx<T>(pa.gtw_lo, mb & gmask));
          tw_step = cmul(ldg_cx<T>(pa.gtw_hi, ms >> pa.gtw_bits), ldg_cx<T>(pa.gtw_lo, ms & gmask));
        }
#pragma unroll
        for (int r = 0; r < N2; ++r) {
          const int kk = j + N1 * r;
          cx<T> o = v[r];
          if (sizeof(T) == 8) {
            o = cmul(o, tw_run);
            tw_run = cmul(tw_run, tw_step);
          } else {
            const long long m = gidx * kk;
            o = cmul(o, cmul(ldg_cx<T>(pa.gtw_hi, m >> pa.gtw_bits), ldg_cx<T>(pa.gtw_lo, m & gmask)));
          }
          out[(long long)kk * pa.os] = o;  // plan-internal data: no swap, no scale (both belong to the last pass)
        }
      }
      __syncthreads();  // all stores of the tile are issued
      if (tid == 0) {
        __threadfence();
        fz::fence_proxy_async_all();
        fz::red_release_add(fa.done_a + chunk, 1ULL);
      }
    } else {
      // ---- pass B: contiguous rows (stage [row][j]), radix 16, exchange through E, radix 16, column store ---------
      {
        const int j = tr;
        cx<T> v[N1];
#pragma unroll
        for (int r = 0; r < N1; ++r) v[r] = S[cr * N + j + B1 * r];
        DFT<N1, T>::run(v);
#pragma unroll
        for (int r = 0; r < N1; ++r) E[cr * PITCH + col::pad<T>(j * N1 + r)] = v[r];
      }
      __syncthreads();  // the stage is consumed, E is complete
      if (tid == 0) {
        ++released;
        pump(false, 0);
      }
      long long ob;
      bool live;
      if (B_MODE == 0) {
        const int tcb = (int)(pb.nb[0] / C);
        const long long ql = local / tcb;
        const int r0 = (int)(local - ql * tcb) * C;
        const long long q = chunk * fa.group + ql;
        live = r0 + cc < pb.nb[0];
        ob = pb.ooff + (long long)(r0 + cc) * pb.obd[0] + q * pb.obd[1];
      } else {
        const int gpc = fa.group / C;
        const long long r = local / gpc;
        const int g8 = (int)(local - r * gpc);
        const long long q0 = chunk * fa.group + (long long)g8 * C;  // first chunk index of the tile
        const long long qh = q0 / pb.nb[0], qlow = q0 - qh * pb.nb[0];
        live = true;
        ob = pb.ooff + (qlow + cc) * pb.obd[0] + r * pb.obd[1] + qh * pb.obd[2];
      }
      {
        const int j = tc;
        cx<T> v[N2];
#pragma unroll
        for (int r = 0; r < N2; ++r) v[r] = E[cc * PITCH + col::pad<T>(j + B1 * r)];
#pragma unroll
        for (int r = 1; r < N2; ++r) v[r] = cmul(v[r], ldg_cx<T>(pb.tw, j * r));
        DFT<N2, T>::run(v);
        if (live) {
#pragma unroll
          for (int r = 0; r < N2; ++r) {
            cx<T> o = v[r];
            if (pb.apply_scale) o = cscale(o, scale_b);
            gstore<T>(pb, fl_b, ob + (long long)(j + N1 * r) * pb.os, o);
          }
        }
      }
      __syncthreads();  // E is rewritten by the next B tile; the ring slot has been read completely
      if (tid == 0) fz::red_release_add(fa.done_b + chunk, 1ULL);
    }
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
bool fused2_plan(const PassParams& a, int variant_a, const PassParams& b, int variant_b, bool is_double,
                 FusedGeom* g) {
  if (env_int("PFFT_NO_FUSE", 0)) return false;
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
  const long long target = (long long)env_int("PFFT_FUSE_CHUNK_KB", 8192) * 1024;
  long long group = mode == 1 ? C : 1;  // (mode 1: a tile takes C consecutive chunk indices)
  const long long qdiv = mode == 1 ? b.nb[0] : Q;  // a chunk must not straddle dimension 2 of pass b
  while (group * 2 * unit * esz <= target && qdiv % (group * 2) == 0) group *= 2;
  if (qdiv % group != 0) return false;
  const long long chunks = Q / group;
  const int lead = env_int("PFFT_FUSE_LEAD", 1), slots = lead + 2;
  if (chunks < 2 * slots) return false;  // too little work for the pipeline to matter
  g->mode = mode;
  g->group = (int)group;
  g->lead = lead;
  g->slots = slots;
  g->num_chunks = chunks;
  g->unit = unit;
  g->tiles_a = group * (W / C);
  g->tiles_b = mode == 0 ? group * (b.nb[0] / C) : (group / C) * b.nb[1];
  g->ring_bytes = (size_t)slots * group * unit * esz;
  return true;
}

template <typename T, int B_MODE>
static cudaError_t launch_fused_t(const PassParams& a, const PassParams& b, const FusedArgs& fa, bool swap_a, bool swap_b,
                                  cudaStream_t stream, bool* used) {
  using Cfg = FusedCfg<T>;
  *used = false;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (!col_make_tensor_map(a, sizeof(T) == 8, Cfg::C, Cfg::kBoxRows, &map)) return cudaSuccess;  // caller runs the passes apart
  auto kern = wg_fused2_kernel<T, B_MODE>;
  int slots = persistent_slots(kern, Cfg::NT, Cfg::kSmem);
  if (slots <= 0) return cudaErrorLaunchOutOfResources;
  const int cap = env_int("PFFT_FUSE_CTAS_PER_SM", 0);
  if (cap > 0) slots = std::min(slots, cap * sm_count());
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
