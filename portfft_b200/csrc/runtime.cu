// Runtime of a committed plan + the C ABI (include/pfft.h).
//
// Owns what the reference's committed_descriptor_impl owns (twiddles, scratch, kernel parameters;
// /root/reference/src/portfft/committed_descriptor_impl.hpp:716-768,579-708) and runs the pass list on a CUDA
// stream.  Twiddle tables are built once per plan in extended precision on the host, rounded once to the plan's
// scalar type and kept device resident (north_star: "device-resident twiddle tables" instead of
// scripts/generate_twiddles.py + per-level device kernels, subgroup_dispatcher.hpp:666-693,
// workgroup_dispatcher.hpp:382-443, global_dispatcher.hpp:107-256).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pfft.h"
#include "kernels.h"
#include "plan.h"
#include "tables.h"

namespace pfft {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }  // for the other translation units of the C ABI
static std::atomic<unsigned long long> g_total_launches{0};

#define PFFT_CUDA_CHECK(expr)                                                                         \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      throw PlanError(PFFT_CUDA_ERROR, std::string(#expr) + " failed: " + cudaGetErrorString(_e));    \
  } while (0)

struct GtwTable {
  void* lo = nullptr;
  void* hi = nullptr;
  int bits = 0;
};

}  // namespace pfft

using namespace pfft;

// Device-resident tables of a plan: immutable after commit, shared (ref-counted) between a plan and its copies
// (pfft_clone) exactly as the reference's copy constructor shares the twiddles and re-allocates only the scratch
// (/root/reference/src/portfft/committed_descriptor_impl.hpp:774-803).
struct PlanTables {
  int device = 0;
  std::map<long long, void*> tw;
  std::map<long long, GtwTable> gtw;
  std::map<std::pair<int, std::pair<long long, long long>>, void*> mod;  // (ModTable, (L, M)) -> device table
  std::vector<void*> owned;
  ~PlanTables() {
    if (owned.empty()) return;
    int prev = -1;
    const bool sw = cudaGetDevice(&prev) == cudaSuccess && prev != device && cudaSetDevice(device) == cudaSuccess;
    for (void* p : owned) cudaFree(p);
    if (sw) cudaSetDevice(prev);
  }
};

struct pfft_plan {
  PlanHost host;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::shared_ptr<PlanTables> tables;
  void* scratch = nullptr;
  void* scratch2 = nullptr;
  void* scratch3 = nullptr;
  size_t scratch_bytes = 0, scratch2_bytes = 0, scratch3_bytes = 0;
  // device staging for pfft_compute_host
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes[2] = {0, 0};
  // pairs of consecutive GLOBAL-level passes that run as ONE persistent kernel with their intermediate result in an
  // L2-resident ring (wg_fused.cu): index of the first pass, chunk geometry, ring + arrival counters, launch count
  struct FusedPair {
    size_t first = 0;
    FusedGeom geom;
    void* ring = nullptr;
    unsigned long long* counters = nullptr;  // [2 * num_chunks]: done_a, done_b
    unsigned long long epoch = 0;
  };
  std::vector<FusedPair> fused[2];
  // encoded TMA tensor maps of the column-tile passes, [direction][pass]: valid while the caller keeps passing the same
  // buffers (the usual case), re-encoded when an address changes.  A different box geometry (col512 vs the two-pass
  // tile kernel) is a different pass, so the address alone identifies the map.
  std::vector<ColMapCache> col_maps[2];  // four entries per pass: input planes 0 / 1, output planes 0 / 1
  ColMapCache* map_cache(int dir, size_t pass) {
    std::vector<ColMapCache>& m = col_maps[dir];
    if (m.size() != 4 * host.passes[dir].size()) m.assign(4 * host.passes[dir].size(), ColMapCache());
    return &m[4 * pass];
  }
  // pfft_compute_host pipeline: copy streams, per-chunk events, sub-batch plans (number_of_transforms -> plan)
  cudaStream_t copy_stream[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> chunk_up, chunk_done;
  std::map<size_t, pfft_plan*> child;
  // L2-resident execution of multi-pass plans: the batch is processed in chunks of `l2_chunk` transforms (0: off),
  // each chunk by a sub-batch plan whose small workspace stays in the 126 MB L2 between its passes
  size_t l2_chunk = 0;
  bool allow_l2_chunk = true;
  // N-D variant (packed layouts): the passes along dimensions 1..D-1 run L2 resident on chunks of `nd_chunk`
  // dimension-0 planes (sub-plans `child`), then ONE pass along dimension 0 over everything (`nd_outer`)
  size_t nd_chunk = 0;
  std::map<size_t, pfft_plan*> nd_child;     // planes per chunk -> sub-plan over dimensions 1..D-1
  pfft_plan* nd_outer[2] = {nullptr, nullptr};  // [direction]: the pass along dimension 0, in place on the output
  pfft_plan* unfused = nullptr;  // REAL domain: the plan without kernel fusion, for buffers the fused kernels cannot take

  ~pfft_plan() {
    delete nd_outer[0];
    delete nd_outer[1];
    delete unfused;
    for (auto& kv : nd_child) delete kv.second;
    for (auto& kv : child) delete kv.second;
    for (cudaEvent_t e : chunk_up) cudaEventDestroy(e);
    for (cudaEvent_t e : chunk_done) cudaEventDestroy(e);
    for (cudaStream_t s : copy_stream)
      if (s) cudaStreamDestroy(s);
    for (auto& v : fused)
      for (FusedPair& f : v) {
        if (f.ring) cudaFree(f.ring);
        if (f.counters) cudaFree(f.counters);
      }
    if (scratch) cudaFree(scratch);
    if (scratch2) cudaFree(scratch2);
    if (scratch3) cudaFree(scratch3);
    for (void* p : stage)
      if (p) cudaFree(p);
  }
};

namespace pfft {

static void* upload(pfft_plan* plan, const void* host, size_t bytes) {
  void* d = nullptr;
  PFFT_CUDA_CHECK(cudaMalloc(&d, bytes));
  plan->tables->owned.push_back(d);
  PFFT_CUDA_CHECK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, plan->stream));
  PFFT_CUDA_CHECK(cudaStreamSynchronize(plan->stream));
  return d;
}

template <typename T>
static void build_tables(pfft_plan* plan) {
  PlanTables* tb = plan->tables.get();
  for (int dir = 0; dir < 2; ++dir) {
    for (PassHost& ps : plan->host.passes[dir]) {
      for (long long n : {ps.tw_n, ps.tw2_n}) {
        if (n > 0 && !tb->tw.count(n)) {
          std::vector<T> t = make_twiddles<T>(n, n, 1);
          tb->tw[n] = upload(plan, t.data(), t.size() * sizeof(T));
        }
      }
      if (ps.pp.gtw_dim >= 0 && !tb->gtw.count(ps.pp.gtw_n)) {
        const long long n = ps.pp.gtw_n;
        int total_bits = 0;
        while ((1LL << total_bits) < n) ++total_bits;
        GtwTable g;
        g.bits = (total_bits + 1) / 2;
        const long long lo_count = 1LL << g.bits;
        const long long hi_count = ((n - 1) >> g.bits) + 1;
        std::vector<T> lo = make_twiddles<T>(n, lo_count, 1);
        std::vector<T> hi = make_twiddles<T>(n, hi_count, lo_count);
        g.lo = upload(plan, lo.data(), lo.size() * sizeof(T));
        g.hi = upload(plan, hi.data(), hi.size() * sizeof(T));
        tb->gtw[n] = g;
      }
      for (int kind : {ps.lmod_kind, ps.smod_kind}) {
        const auto key = std::make_pair(kind, std::make_pair(ps.mod_l, ps.mod_m));
        if (kind != MODT_NONE && !tb->mod.count(key)) {
          std::vector<T> t = make_mod_table<T>(kind, ps.mod_l, ps.mod_m);
          tb->mod[key] = upload(plan, t.data(), t.size() * sizeof(T));
        }
      }
    }
  }
}

// workspaces of one plan instance (a copy of a plan owns its own: committed_descriptor_impl.hpp:774-803) and the
// table / workspace addresses patched into its pass list
static void attach_device_state(pfft_plan* plan) {
  PlanTables* tb = plan->tables.get();
  const size_t scalar = plan->host.desc.is_double ? 8 : 4;
  if (plan->l2_chunk == 0 && plan->nd_chunk == 0) {  // (chunked plans run through their sub-plans and own no workspace)
    plan->scratch_bytes = plan->host.scratch_elems * 2 * scalar;
    if (plan->scratch_bytes) PFFT_CUDA_CHECK(cudaMalloc(&plan->scratch, plan->scratch_bytes));
    plan->scratch2_bytes = plan->host.scratch2_elems * 2 * scalar;
    if (plan->scratch2_bytes) PFFT_CUDA_CHECK(cudaMalloc(&plan->scratch2, plan->scratch2_bytes));
    plan->scratch3_bytes = plan->host.scratch3_elems * 2 * scalar;
    if (plan->scratch3_bytes) PFFT_CUDA_CHECK(cudaMalloc(&plan->scratch3, plan->scratch3_bytes));
  }
  for (int dir = 0; dir < 2; ++dir) {
    for (PassHost& ps : plan->host.passes[dir]) {
      ps.pp.tw = ps.tw_n > 0 ? tb->tw[ps.tw_n] : nullptr;
      ps.pp.tw2 = ps.tw2_n > 0 ? tb->tw[ps.tw2_n] : nullptr;
      if (ps.pp.gtw_dim >= 0) {
        const GtwTable& g = tb->gtw[ps.pp.gtw_n];
        ps.pp.gtw_lo = g.lo;
        ps.pp.gtw_hi = g.hi;
        ps.pp.gtw_bits = g.bits;
      }
      if (ps.lmod_kind != MODT_NONE) ps.pp.lmod = tb->mod[std::make_pair(ps.lmod_kind, std::make_pair(ps.mod_l, ps.mod_m))];
      if (ps.smod_kind != MODT_NONE) ps.pp.smod = tb->mod[std::make_pair(ps.smod_kind, std::make_pair(ps.mod_l, ps.mod_m))];
    }
    // fusable pairs: pass i writes a workspace that only pass i + 1 reads
    std::vector<PassHost>& ps = plan->host.passes[dir];
    plan->fused[dir].clear();
    for (size_t i = 0; i + 1 < ps.size(); ++i) {
      const PassHost &a = ps[i], &b = ps[i + 1];
      if (a.kernel != KERNEL_WG_COL || b.kernel != KERNEL_WG_COL || a.dst != b.src) continue;
      if (a.dst != BUF_SCRATCH && a.dst != BUF_SCRATCH2) continue;
      if (a.internal_storage != b.internal_storage || a.real_view || b.real_view) continue;
      if (a.pp.smod_mask != 0 || b.pp.smod_mask != 0) continue;  // (store modifier over the whole transform: wg_col only)
      bool read_later = false;
      for (size_t j = i + 2; j < ps.size() && !read_later; ++j) {
        if (ps[j].src == a.dst) read_later = true;
        if (ps[j].dst == a.dst) break;
      }
      if (read_later) continue;
      pfft_plan::FusedPair f;
      f.first = i;
      if (a.pp.n != 256 || b.pp.n != 256) continue;
      const int fgrid = fused2_grid(plan->host.desc.is_double);
      if (fgrid <= 0 || !fused2_plan(a.pp, a.variant, b.pp, b.variant, plan->host.desc.is_double, fgrid, &f.geom)) continue;
      PFFT_CUDA_CHECK(cudaMalloc(&f.ring, f.geom.ring_bytes));
      const size_t cbytes = (size_t)2 * f.geom.num_chunks * sizeof(unsigned long long);
      PFFT_CUDA_CHECK(cudaMalloc((void**)&f.counters, cbytes));
      PFFT_CUDA_CHECK(cudaMemsetAsync(f.counters, 0, cbytes, plan->stream));
      PFFT_CUDA_CHECK(cudaStreamSynchronize(plan->stream));
      plan->fused[dir].push_back(f);
      ++i;  // a pass belongs to one pair at most
    }
  }
}

static void commit_device(pfft_plan* plan) {
  plan->tables = std::make_shared<PlanTables>();
  plan->tables->device = plan->device;
  if (plan->host.desc.is_double)
    build_tables<double>(plan);
  else
    build_tables<float>(plan);
  attach_device_state(plan);
}

// The kernels of a plan must be launched with the plan's device current (its streams, tables and workspaces live
// there); a caller that drives several GPUs from one thread gets its own current device back afterwards.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      ok = false;
    } else if (prev != device) {
      switched = cudaSetDevice(device) == cudaSuccess;
      ok = switched;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

struct PeerTable {
  size_t n = 0;
  void* const* re = nullptr;
  void* const* im = nullptr;
};

static pfft_plan* make_plan(const DescHost& d, int device, cudaStream_t stream, bool allow_l2_chunk = true);
static void execute(pfft_plan* plan, int dir, const void* in, const void* in_imag, void* out, void* out_imag,
                    cudaStream_t stream, const PeerTable* peers = nullptr);

// L2-resident execution (see choose_l2_chunk): sub-batch plans over consecutive batch ranges
static void execute_chunked(pfft_plan* plan, int dir, const void* in, const void* in_imag, void* out, void* out_imag,
                            cudaStream_t stream) {
  const DescHost& d = plan->host.desc;
  const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  const int odir = dir == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD;
  const size_t scalar = d.is_double ? 8 : 4;
  // bytes between consecutive transforms, per plane, on each side (REAL: the forward domain counts real scalars)
  const size_t unit_in = (d.is_real() && dir == PFFT_FORWARD) ? scalar : (il ? 2 * scalar : scalar);
  const size_t unit_out = (d.is_real() && dir == PFFT_BACKWARD) ? scalar : (il ? 2 * scalar : scalar);
  const size_t step_in = d.distance(dir) * unit_in, step_out = d.distance(odir) * unit_out;
  const size_t batch = d.number_of_transforms;
  for (size_t b0 = 0; b0 < batch; b0 += plan->l2_chunk) {
    const size_t nb = std::min(plan->l2_chunk, batch - b0);
    pfft_plan*& sub = plan->child[nb];
    if (sub == nullptr) {
      DescHost dc = d;
      dc.number_of_transforms = nb;
      sub = make_plan(dc, plan->device, plan->stream, false);
    }
    const char* i0 = (const char*)in + b0 * step_in;
    const char* i1 = in_imag ? (const char*)in_imag + b0 * step_in : nullptr;
    char* o0 = (char*)out + b0 * step_out;
    char* o1 = out_imag ? (char*)out_imag + b0 * step_out : nullptr;
    execute(sub, dir, i0, i1, o0, o1, stream, nullptr);
  }
}

// N-D, packed layouts: planes of dimension 0 (and of the batch) are independent for every pass except the one along
// dimension 0.  Chunks of planes run those passes back to back (the first reads the user's input, the others work in
// place on the output planes, which stay in L2 meanwhile); one strided pass along dimension 0 finishes in place.
// 512^3: HBM sees 2 round trips instead of 3.
static void execute_nd_chunked(pfft_plan* plan, int dir, const void* in, const void* in_imag, void* out, void* out_imag,
                               cudaStream_t stream) {
  const DescHost& d = plan->host.desc;
  const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  const size_t scalar = d.is_double ? 8 : 4;
  const size_t unit = il ? 2 * scalar : scalar;
  size_t plane = 1;
  for (size_t i = 1; i < d.lengths.size(); ++i) plane *= d.lengths[i];
  const size_t planes = d.number_of_transforms * d.lengths[0];
  // (the sub-plans carry the descriptor's offsets themselves)
  for (size_t p0 = 0; p0 < planes; p0 += plan->nd_chunk) {
    const size_t np = std::min(plan->nd_chunk, planes - p0);
    pfft_plan*& sub = plan->nd_child[np];
    if (sub == nullptr) {
      DescHost dc = d;
      dc.lengths.assign(d.lengths.begin() + 1, d.lengths.end());
      dc.forward_strides.assign(d.forward_strides.begin() + 1, d.forward_strides.end());
      dc.backward_strides.assign(d.backward_strides.begin() + 1, d.backward_strides.end());
      dc.number_of_transforms = np;
      dc.forward_distance = dc.backward_distance = plane;
      dc.forward_scale = dc.backward_scale = 1.0;  // the scale belongs to the last pass (nd_outer)
      sub = make_plan(dc, plan->device, plan->stream, false);
    }
    const size_t step = p0 * plane * unit;
    execute(sub, dir, (const char*)in + step, in_imag ? (const char*)in_imag + step : nullptr, (char*)out + step,
            out_imag ? (char*)out_imag + step : nullptr, stream, nullptr);
  }
  pfft_plan*& outer = plan->nd_outer[dir == PFFT_FORWARD ? 0 : 1];
  if (outer == nullptr) {
    DescHost dc = d;
    dc.lengths.assign(1, d.lengths[0]);
    dc.forward_strides.assign(1, plane);
    dc.backward_strides.assign(1, plane);
    dc.number_of_transforms = plane;
    dc.forward_distance = dc.backward_distance = 1;
    dc.placement = PFFT_IN_PLACE;
    // in place on the OUTPUT buffer in either direction: both domains of this sub-plan carry the output's offset
    dc.forward_offset = dc.backward_offset = d.offset(dir == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD);
    outer = make_plan(dc, plan->device, plan->stream, false);
  }
  const size_t total = d.lengths[0] * plane * unit;
  for (size_t b = 0; b < d.number_of_transforms; ++b)
    execute(outer, dir, (const char*)out + b * total, out_imag ? (const char*)out_imag + b * total : nullptr,
            (char*)out + b * total, out_imag ? (char*)out_imag + b * total : nullptr, stream, nullptr);
}

static void execute(pfft_plan* plan, int dir, const void* in, const void* in_imag, void* out, void* out_imag,
                    cudaStream_t stream, const PeerTable* peers) {
  const DescHost& d = plan->host.desc;
  const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  const bool real = d.is_real();
  const bool bwd = dir == PFFT_BACKWARD;
  // committed_descriptor_impl.hpp:862-871 (for REAL descriptors the check applies to the complex side of the call;
  // the real side is one scalar array: committed_descriptor.hpp:201-206,273-278)
  const bool in_cplx = !real || bwd, out_cplx = !real || !bwd;
  if ((in_cplx && il && in_imag != nullptr) || (out_cplx && il && out_imag != nullptr))
    throw PlanError(PFFT_INVALID_CONFIGURATION,
                    "To use interleaved data layout, please set the storage in the descriptor to INTERLEAVED_COMPLEX");
  if ((in_cplx && !il && in_imag == nullptr) || (out_cplx && !il && out_imag == nullptr))
    throw PlanError(PFFT_INVALID_CONFIGURATION,
                    "To use split data layout, please set the storage in the descriptor to SPLIT_COMPLEX");
  if ((!in_cplx && in_imag != nullptr) || (!out_cplx && out_imag != nullptr))
    throw PlanError(PFFT_INVALID_CONFIGURATION, "the real side of a REAL-domain transform is a single scalar array");
  if (in == nullptr || out == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null data pointer");
  // REAL-domain passes fused into the TMA kernels need 16-byte aligned buffers (their rows may start 8 bytes off such a
  // boundary: the bulk copies then begin one element early, inside the buffer); other pointers run the unfused plan
  // (made on first use)
  for (const PassHost& ps : plan->host.passes[dir]) {
    if (ps.fuse_real == 0) continue;
    if ((uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0) continue;
    if (plan->unfused == nullptr) {
      DescHost du = d;
      du.no_real_fuse = true;
      plan->unfused = make_plan(du, plan->device, plan->stream, plan->allow_l2_chunk);
    }
    execute(plan->unfused, dir, in, in_imag, out, out_imag, stream, peers);
    return;
  }
  if (plan->l2_chunk != 0 && peers == nullptr) {
    execute_chunked(plan, dir, in, in_imag, out, out_imag, stream);
    return;
  }
  if (plan->nd_chunk != 0 && peers == nullptr) {
    execute_nd_chunked(plan, dir, in, in_imag, out, out_imag, stream);
    return;
  }
  const size_t scalar = d.is_double ? 8 : 4;
  // backward = forward transform of the (re <-> im)-swapped data, swapped back: free for split storage
  // (REAL plans never swap: c2r_pre writes its rows index-reversed instead, real.cu)
  const void* uin_re = in;
  const void* uin_im = in_imag;
  void* uout_re = out;
  void* uout_im = out_imag;
  if (!il && bwd && !real) {
    std::swap(uin_re, uin_im);
    std::swap(uout_re, uout_im);
  }
  void* s_re = plan->scratch;
  void* s_im = il ? nullptr : (void*)((char*)plan->scratch + plan->host.scratch_elems * scalar);
  void* s2_re = plan->scratch2;
  void* s2_im = il ? nullptr : (void*)((char*)plan->scratch2 + plan->host.scratch2_elems * scalar);
  struct Prepared {
    PassParams p;
    bool swap, il_in, il_out, pil;
  };
  auto prepare = [&](const PassHost& ps) {
    Prepared r;
    PassParams& p = r.p;
    p = ps.pp;
    switch (ps.src) {
      case BUF_IN: p.in_re = uin_re; p.in_im = uin_im; break;
      case BUF_OUT: p.in_re = uout_re; p.in_im = uout_im; break;
      case BUF_SCRATCH2: p.in_re = s2_re; p.in_im = s2_im; break;
      case BUF_SCRATCH3: p.in_re = plan->scratch3; p.in_im = nullptr; break;  // always interleaved
      default: p.in_re = s_re; p.in_im = s_im; break;
    }
    switch (ps.dst) {
      case BUF_OUT: p.out_re = uout_re; p.out_im = uout_im; break;
      case BUF_SCRATCH2: p.out_re = s2_re; p.out_im = s2_im; break;
      case BUF_SCRATCH3: p.out_re = plan->scratch3; p.out_im = nullptr; break;
      default: p.out_re = s_re; p.out_im = s_im; break;
    }
    // backward on interleaved data = (re <-> im) swap on load and store; passes between plan-internal buffers
    // (Bluestein's inner transforms) always run the plain forward transform
    const int internal = MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT;
    // storage of each side of this pass: the descriptor's, or interleaved for plan-internal rows of a REAL plan
    r.il_in = il || (ps.internal_storage & 1);
    r.il_out = il || (ps.internal_storage & 2);
    r.pil = r.il_in && r.il_out;  // the transform kernels take one storage for both sides (the planner pairs them)
    // (force_swap: inverse complex passes of a REAL N-D backward plan.)  A pass whose both sides are plan-internal never
    // swaps, in any kernel family.
    r.swap = (ps.force_swap || (r.pil && bwd && !real)) && (p.mod_flags & internal) != internal;
    if (ps.real_view) {
      // the user's real rows addressed as interleaved complex pairs: needs complex alignment of the first element
      const uintptr_t a = (ps.real_view & 1) ? (uintptr_t)p.in_re + (size_t)p.ioff * 2 * scalar
                                            : (uintptr_t)p.out_re + (size_t)p.ooff * 2 * scalar;
      if (a % (2 * scalar) != 0)
        throw PlanError(PFFT_UNSUPPORTED_CONFIGURATION, "REAL domain: the real buffer must be aligned to a complex element");
    }
    if (p.peer_dim >= 0) {
      if (peers == nullptr || peers->n != (size_t)p.nb[p.peer_dim])
        throw PlanError(PFFT_INVALID_CONFIGURATION,
                        "plan committed with PFFT_GURU_PEER_LAST_DIM: call pfft_compute_peer with one output buffer "
                        "per index of the last extra batch dimension");
      for (size_t i = 0; i < peers->n; ++i) {
        void* re = peers->re[i];
        void* im = il ? nullptr : peers->im[i];
        if (!il && bwd) std::swap(re, im);
        p.out_tab_re[i] = re;
        p.out_tab_im[i] = im;
      }
    }
    return r;
  };
  std::vector<PassHost>& passes = plan->host.passes[dir];
  std::vector<pfft_plan::FusedPair>& fused = plan->fused[dir];
  size_t next_fused = 0;
  for (size_t pi = 0; pi < passes.size(); ++pi) {
    const PassHost& ps = passes[pi];
    const Prepared pr = prepare(ps);
    const PassParams& p = pr.p;
    const bool swap = pr.swap, il_in = pr.il_in, il_out = pr.il_out, pil = pr.pil;
    cudaError_t e = cudaSuccess;
    while (next_fused < fused.size() && fused[next_fused].first < pi) ++next_fused;
    if (next_fused < fused.size() && fused[next_fused].first == pi) {
      // this pass and the next one as ONE persistent kernel, the data between them in an L2-resident ring.  The source
      // of the first must not be the destination of the second (an in-place transform through the workspace): the
      // second pass of a chunk would overwrite what the first pass of a later chunk still has to read.
      pfft_plan::FusedPair& f = fused[next_fused];
      const Prepared pb = prepare(passes[pi + 1]);
      if (pil && pb.pil && p.in_re != pb.p.out_re && pr.swap == pb.swap) {
        FusedArgs fa;
        fa.ring = f.ring;
        fa.done_a = f.counters;
        fa.done_b = f.counters + f.geom.num_chunks;
        fa.epoch = f.epoch + 1;
        fa.group = f.geom.group;
        fa.lead = f.geom.lead;
        fa.slots = f.geom.slots;
        fa.num_chunks = f.geom.num_chunks;
        fa.unit = f.geom.unit;
        fa.tiles_a = f.geom.tiles_a;
        fa.tiles_b = f.geom.tiles_b;
        bool used = false;
        // swap of the pair: on the load of the first pass and the store of the second (what lies between is internal)
        e = launch_wg_fused2(p, pb.p, fa, f.geom.mode, d.is_double, pr.swap, pb.swap, stream, &used);
        if (e != cudaSuccess)
          throw PlanError(PFFT_CUDA_ERROR, std::string("kernel launch failed: ") + cudaGetErrorString(e));
        if (used) {
          ++f.epoch;
          g_total_launches.fetch_add(1, std::memory_order_relaxed);
          ++pi;
          continue;
        }
      }
    }
    switch (ps.kernel) {
      case KERNEL_WG_GENERIC:
        e = launch_wg_generic(p, d.is_double, pil, swap, ps.grid, stream);
        break;
      case KERNEL_WI: {
        bool used = false;
        if (ps.fuse_real != 0) {
          e = launch_wi_tma_real(p, d.is_double, ps.fuse_real, stream);
          break;
        }
        if (ps.variant == 1) e = launch_wi_tma(p, d.is_double, swap, stream, &used);
        if (e == cudaSuccess && !used)
          e = d.is_double ? launch_wi_f64(p, pil, swap, ps.grid, stream) : launch_wi_f32(p, pil, swap, ps.grid, stream);
        break;
      }
      case KERNEL_SG:
        e = d.is_double ? launch_sg_f64(p, pil, swap, ps.grid, stream) : launch_sg_f32(p, pil, swap, ps.grid, stream);
        break;
      case KERNEL_WG_R3:
        e = launch_wg_r3(p, d.is_double, pil, swap, ps.grid, stream);
        break;
      case KERNEL_WG_COL: {
        bool used = false;
        e = launch_wg_col(p, d.is_double, swap, ps.variant, ps.alt_grid, stream, &used, plan->map_cache(dir, pi));
        if (e == cudaSuccess && !used) e = launch_wg_generic(p, d.is_double, pil, swap, ps.grid, stream);
        break;
      }
      case KERNEL_WG_CUBE:
        // cp.async.bulk needs 16-byte aligned global addresses; otherwise run the generic kernel (fused REAL-domain
        // passes were diverted to the unfused plan above)
        if (ps.fuse_real != 0 || (((uintptr_t)p.in_re % 16) == 0 && ((uintptr_t)p.out_re % 16) == 0))
          e = launch_wg_cube(p, d.is_double, swap, ps.variant, ps.alt_grid, stream, ps.fuse_real);
        else
          e = launch_wg_generic(p, d.is_double, pil, swap, ps.grid, stream);
        break;
      case KERNEL_WG_COLG:
        e = ps.variant == 1 ? launch_wg_colr3(p, d.is_double, pil, swap, ps.grid, stream, plan->map_cache(dir, pi))
                            : launch_wg_colg(p, d.is_double, pil, swap, ps.grid, stream);
        break;
      case KERNEL_EW:
        e = launch_ew(p, d.is_double, il_in, il_out, swap, ps.grid, stream);
        break;
      case KERNEL_REAL_PACK:
        e = launch_real_pack(p, d.is_double, ps.variant, ps.grid, stream);
        break;
      case KERNEL_REAL_UNPACK:
        e = launch_real_unpack(p, d.is_double, ps.variant, ps.grid, stream);
        break;
      case KERNEL_R2C_POST:
        e = launch_r2c_post(p, d.is_double, il_out, ps.variant, ps.grid, stream);
        break;
      case KERNEL_C2R_PRE:
        e = launch_c2r_pre(p, d.is_double, il_in, ps.variant, ps.grid, stream);
        break;
      default:
        throw PlanError(PFFT_INTERNAL_ERROR, "unknown kernel kind");
    }
    if (e != cudaSuccess) throw PlanError(PFFT_CUDA_ERROR, std::string("kernel launch failed: ") + cudaGetErrorString(e));
    g_total_launches.fetch_add(1, std::memory_order_relaxed);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// L2-resident chunking.  A plan with k > 1 passes moves its intermediate results through a workspace; run over the
// whole batch at once, every pass streams the full data set through HBM (k round trips).  When the per-transform
// workspace is small against the 126 MB L2, the batch is instead cut into chunks whose workspace (a few tens of MB,
// the SAME addresses for every chunk) stays L2 resident: each chunk runs all its passes back to back, the
// intermediate writes are absorbed and re-read by L2, and HBM sees one read of the input and one write of the output.
// Chunks are sub-batch plans on the base pointers advanced by b0 * distance (the descriptor's own addressing:
// /root/reference/src/portfft/descriptor.hpp:91-92), exactly like the host pipeline below.
// ---------------------------------------------------------------------------------------------------------------
static void choose_l2_chunk(pfft_plan* plan) {
  const PlanHost& h = plan->host;
  const DescHost& d = h.desc;
  plan->l2_chunk = 0;
  if (!plan->allow_l2_chunk || !d.extra.empty() || d.number_of_transforms < 2) return;
  const size_t ws = h.scratch_elems + h.scratch2_elems + h.scratch3_elems;
  if (ws == 0) return;
  // Opt-in (PFFT_L2_CHUNK_BYTES): measured on B200, 65536-point x 2048 fp32 (2 passes): 0.82 ms in one piece, 1.09 /
  // 1.27 / 2.23 ms with 64 / 32 / 8 MB chunks -- every chunk pass is a ~10 us launch whose ramp-up and tail are not
  // amortised; the idea needs the passes of a chunk fused into one persistent kernel (profiles/r1_ab_variants.txt).
  const char* env = std::getenv("PFFT_L2_CHUNK_BYTES");
  const size_t budget = env ? (size_t)std::atoll(env) : 0;
  if (budget == 0) return;
  const size_t per = ws * 2 * (d.is_double ? 8 : 4) / d.number_of_transforms;  // workspace bytes per transform
  if (per == 0 || per > budget) return;                                        // one transform alone outgrows L2
  const size_t chunk = budget / per;
  if (chunk >= d.number_of_transforms) return;  // the whole batch fits: nothing to cut
  plan->l2_chunk = chunk;
}

static void choose_nd_chunk(pfft_plan* plan) {
  const DescHost& d = plan->host.desc;
  plan->nd_chunk = 0;
  if (!plan->allow_l2_chunk || plan->l2_chunk != 0 || !d.extra.empty() || d.is_real() || d.lengths.size() < 2 ||
      d.lengths[0] < 2)
    return;
  if (get_layout(d, PFFT_FORWARD) != PFFT_LAYOUT_PACKED || get_layout(d, PFFT_BACKWARD) != PFFT_LAYOUT_PACKED) return;
  if (d.forward_offset != d.backward_offset && d.placement == PFFT_IN_PLACE) return;
  const char* env = std::getenv("PFFT_L2_CHUNK_BYTES");  // opt-in, see choose_l2_chunk
  const size_t budget = env ? (size_t)std::atoll(env) : 0;
  if (budget == 0) return;
  size_t plane = 1;
  for (size_t i = 1; i < d.lengths.size(); ++i) plane *= d.lengths[i];
  const size_t plane_bytes = plane * 2 * (d.is_double ? 8 : 4);
  const size_t planes = d.number_of_transforms * d.lengths[0];
  const size_t chunk = budget / plane_bytes;
  // worth it only when the whole array does not fit L2 anyway and a chunk holds at least two planes
  if (chunk < 2 || chunk >= planes || planes * plane_bytes <= 3 * budget) return;
  plan->nd_chunk = chunk;
}

// SM count and opt-in shared memory of a device, queried once per device (cudaGetDeviceProperties costs ~1 ms)
static DeviceLimits device_limits(int device) {
  static std::mutex mu;
  static std::map<int, DeviceLimits> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(device);
  if (it != cache.end()) return it->second;
  DeviceLimits lim;
  int v = 0;
  PFFT_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
  lim.num_sms = v;
  PFFT_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  lim.max_smem_per_block = (size_t)v;
  cache[device] = lim;
  return lim;
}

static pfft_plan* make_plan(const DescHost& d, int device, cudaStream_t stream, bool allow_l2_chunk) {
  DeviceGuard guard(device);  // the caller's current device is restored on return
  if (!guard.ok) throw PlanError(PFFT_CUDA_ERROR, "cudaSetDevice failed for the plan's device");
  const DeviceLimits lim = device_limits(device);
  std::unique_ptr<pfft_plan> plan(new pfft_plan);
  plan->host = build_plan(d, lim);
  plan->device = device;
  plan->stream = stream;
  plan->allow_l2_chunk = allow_l2_chunk;
  choose_l2_chunk(plan.get());
  choose_nd_chunk(plan.get());
  commit_device(plan.get());
  return plan.release();
}

// ---------------------------------------------------------------------------------------------------------------
// pfft_compute_host: batch-chunked pipeline.  Chunk c is copied to the device on copy stream 0, transformed on the
// plan's stream and copied back on copy stream 1, so that H2D of chunk c+1, the kernels of chunk c and D2H of
// chunk c-1 overlap (PCIe is full duplex; the kernels are ~50x shorter than either copy).
// ---------------------------------------------------------------------------------------------------------------
struct HostDomain {
  size_t dist, off, extent, count;
};

static HostDomain host_domain(const DescHost& d, int dir) {
  HostDomain h;
  h.dist = d.distance(dir);
  h.off = d.offset(dir);
  h.extent = 1;
  const std::vector<size_t> len = d.domain_lengths(dir);  // (REAL: n / 2 + 1 along the last dimension of the backward domain)
  for (size_t i = 0; i < len.size(); ++i) h.extent += (len[i] - 1) * d.strides(dir)[i];
  h.count = d.buffer_count(dir);
  return h;
}

// every element of the buffer on side `dir` is addressed by the descriptor (nothing to preserve around the results)
static bool dense_domain(const DescHost& d, int dir) {
  const std::vector<size_t> len = d.domain_lengths(dir);
  size_t total = 1;
  for (size_t l : len) total *= l;
  return d.offset(dir) == 0 && d.strides(dir) == default_strides(len) && d.distance(dir) == total;
}

static void compute_host_monolithic(pfft_plan* plan, int direction, const void* in, const void* in_imag, void* out,
                                    void* out_imag, char* din, char* dout, size_t plane_in, size_t plane_out) {
  const DescHost& d = plan->host.desc;
  const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  const int odir = direction == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD;
  const bool inplace = in == out;
  // REAL descriptors: the real side of the call is a single plane of scalars
  const bool in2 = !il && !(d.is_real() && direction == PFFT_FORWARD);
  const bool out2 = !il && !(d.is_real() && direction == PFFT_BACKWARD);
  cudaStream_t s = plan->stream;
  PFFT_CUDA_CHECK(cudaMemcpyAsync(din, in, plane_in, cudaMemcpyHostToDevice, s));
  if (in2) PFFT_CUDA_CHECK(cudaMemcpyAsync(din + plane_in, in_imag, plane_in, cudaMemcpyHostToDevice, s));
  // elements of the output buffer that the descriptor does not address must survive the round trip
  const bool out_dense = dense_domain(d, odir);
  if (!inplace && !out_dense) {
    PFFT_CUDA_CHECK(cudaMemcpyAsync(dout, out, plane_out, cudaMemcpyHostToDevice, s));
    if (out2) PFFT_CUDA_CHECK(cudaMemcpyAsync(dout + plane_out, out_imag, plane_out, cudaMemcpyHostToDevice, s));
  }
  execute(plan, direction, din, in2 ? din + plane_in : nullptr, dout, out2 ? dout + plane_out : nullptr, s);
  PFFT_CUDA_CHECK(cudaMemcpyAsync(out, dout, plane_out, cudaMemcpyDeviceToHost, s));
  if (out2) PFFT_CUDA_CHECK(cudaMemcpyAsync(out_imag, dout + plane_out, plane_out, cudaMemcpyDeviceToHost, s));
  PFFT_CUDA_CHECK(cudaStreamSynchronize(s));
}

static void compute_host_pipelined(pfft_plan* plan, int direction, const void* in, const void* in_imag, void* out,
                                   void* out_imag, char* din, char* dout, size_t plane_in, size_t plane_out,
                                   size_t chunks) {
  const DescHost& d = plan->host.desc;
  const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  const int odir = direction == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD;
  const bool inplace = in == out;
  // bytes of one addressed unit of a plane, per side (REAL descriptors: the forward domain counts real scalars in ONE
  // plane whatever the complex storage)
  const size_t sc = d.is_double ? 8 : 4;
  const bool in_real = d.is_real() && direction == PFFT_FORWARD, out_real = d.is_real() && direction == PFFT_BACKWARD;
  const size_t esz_in = in_real ? sc : sc * (il ? 2 : 1), esz_out = out_real ? sc : sc * (il ? 2 : 1);
  const bool in2 = !il && !in_real, out2 = !il && !out_real;  // a second (imaginary) plane on that side
  const HostDomain hi = host_domain(d, direction), ho = host_domain(d, odir);
  const bool out_dense = dense_domain(d, odir);
  const size_t batch = d.number_of_transforms;
  const size_t per = (batch + chunks - 1) / chunks;
  chunks = (batch + per - 1) / per;
  for (int i = 0; i < 2; ++i)
    if (!plan->copy_stream[i]) PFFT_CUDA_CHECK(cudaStreamCreateWithFlags(&plan->copy_stream[i], cudaStreamNonBlocking));
  while (plan->chunk_up.size() < chunks) {
    cudaEvent_t a, b;
    PFFT_CUDA_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    PFFT_CUDA_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    plan->chunk_up.push_back(a);
    plan->chunk_done.push_back(b);
  }
  cudaStream_t s_up = plan->copy_stream[0], s_dn = plan->copy_stream[1], s = plan->stream;
  for (size_t c = 0; c < chunks; ++c) {
    const size_t b0 = c * per, b1 = std::min(batch, b0 + per), nb = b1 - b0;
    pfft_plan*& sub = plan->child[nb];
    if (sub == nullptr) {
      DescHost dc = d;
      dc.number_of_transforms = nb;
      sub = make_plan(dc, plan->device, s);
    }
    // byte ranges of this chunk in the input and output planes
    const size_t i0 = c == 0 ? 0 : (hi.off + b0 * hi.dist) * esz_in, i1 = b1 == batch ? plane_in : (hi.off + b1 * hi.dist) * esz_in;
    const size_t o0 = c == 0 ? 0 : (ho.off + b0 * ho.dist) * esz_out, o1 = b1 == batch ? plane_out : (ho.off + b1 * ho.dist) * esz_out;
    PFFT_CUDA_CHECK(cudaMemcpyAsync(din + i0, (const char*)in + i0, i1 - i0, cudaMemcpyHostToDevice, s_up));
    if (in2)
      PFFT_CUDA_CHECK(cudaMemcpyAsync(din + plane_in + i0, (const char*)in_imag + i0, i1 - i0, cudaMemcpyHostToDevice, s_up));
    if (!inplace && !out_dense) {
      PFFT_CUDA_CHECK(cudaMemcpyAsync(dout + o0, (const char*)out + o0, o1 - o0, cudaMemcpyHostToDevice, s_up));
      if (out2)
        PFFT_CUDA_CHECK(cudaMemcpyAsync(dout + plane_out + o0, (const char*)out_imag + o0, o1 - o0, cudaMemcpyHostToDevice, s_up));
    }
    PFFT_CUDA_CHECK(cudaEventRecord(plan->chunk_up[c], s_up));
    PFFT_CUDA_CHECK(cudaStreamWaitEvent(s, plan->chunk_up[c], 0));
    const size_t pi = b0 * hi.dist * esz_in, po = b0 * ho.dist * esz_out;  // the sub-plan keeps the descriptor's offsets
    execute(sub, direction, din + pi, in2 ? din + plane_in + pi : nullptr, dout + po, out2 ? dout + plane_out + po : nullptr, s);
    PFFT_CUDA_CHECK(cudaEventRecord(plan->chunk_done[c], s));
    PFFT_CUDA_CHECK(cudaStreamWaitEvent(s_dn, plan->chunk_done[c], 0));
    PFFT_CUDA_CHECK(cudaMemcpyAsync((char*)out + o0, dout + o0, o1 - o0, cudaMemcpyDeviceToHost, s_dn));
    if (out2)
      PFFT_CUDA_CHECK(cudaMemcpyAsync((char*)out_imag + o0, dout + plane_out + o0, o1 - o0, cudaMemcpyDeviceToHost, s_dn));
  }
  PFFT_CUDA_CHECK(cudaStreamSynchronize(s_dn));
  PFFT_CUDA_CHECK(cudaStreamSynchronize(s));
}

template <typename F>
static pfft_status guarded(F&& f) {
  try {
    f();
    return PFFT_OK;
  } catch (const PlanError& e) {
    g_last_error = e.what();
    return e.status;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return PFFT_INTERNAL_ERROR;
  }
}

}  // namespace pfft

extern "C" {

pfft_status pfft_validate(const pfft_desc* desc) {
  return guarded([&] { validate_descriptor(desc_from_c(desc)); });
}

size_t pfft_get_flattened_length(const pfft_desc* desc) {
  size_t r = 0;
  guarded([&] { r = desc_from_c(desc).flattened_length(); });
  return r;
}

size_t pfft_get_buffer_count(const pfft_desc* desc, int direction) {
  size_t r = 0;
  guarded([&] { r = desc_from_c(desc).buffer_count(direction); });
  return r;
}

int pfft_get_layout(const pfft_desc* desc, int direction) {
  int r = -1;
  guarded([&] { r = get_layout(desc_from_c(desc), direction); });
  return r;
}

pfft_status pfft_plan_describe(const pfft_desc* desc, int direction, char* buf, size_t buf_len, size_t* needed) {
  return guarded([&] {
    PlanHost plan = build_plan(desc_from_c(desc), DeviceLimits{});
    std::string s = describe_plan(plan, direction);
    if (needed) *needed = s.size() + 1;
    if (buf && buf_len) {
      const size_t n = std::min(buf_len - 1, s.size());
      std::memcpy(buf, s.data(), n);
      buf[n] = 0;
    }
  });
}

pfft_status pfft_plan_export(const pfft_desc* desc, int direction, char* buf, size_t buf_len, size_t* needed) {
  return guarded([&] {
    PlanHost plan = build_plan(desc_from_c(desc), DeviceLimits{});
    std::string s = export_plan_json(plan, direction);
    if (needed) *needed = s.size() + 1;
    if (buf && buf_len) {
      const size_t n = std::min(buf_len - 1, s.size());
      std::memcpy(buf, s.data(), n);
      buf[n] = 0;
    }
  });
}

pfft_status pfft_table_host(int precision, int kind, size_t transform_length, size_t convolution_length, void* out) {
  return guarded([&] {
    if (out == nullptr || kind < MODT_CHIRP || kind > MODT_CONV || transform_length == 0)
      throw PlanError(PFFT_INVALID_CONFIGURATION, "pfft_table_host: bad arguments");
    if (precision == PFFT_DOUBLE) {
      std::vector<double> t = make_mod_table<double>(kind, (long long)transform_length, (long long)convolution_length);
      std::memcpy(out, t.data(), t.size() * sizeof(double));
    } else {
      std::vector<float> t = make_mod_table<float>(kind, (long long)transform_length, (long long)convolution_length);
      std::memcpy(out, t.data(), t.size() * sizeof(float));
    }
  });
}

pfft_status pfft_commit(const pfft_desc* desc, int device, void* stream, pfft_plan** plan_out) {
  return guarded([&] {
    if (plan_out == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null plan_out");
    *plan_out = nullptr;
    DescHost d = desc_from_c(desc);
    validate_descriptor(d);  // host-only errors first, before any CUDA call
    *plan_out = make_plan(d, device, (cudaStream_t)stream);
  });
}

pfft_status pfft_clone(const pfft_plan* plan, pfft_plan** plan_out) {
  return guarded([&] {
    if (plan == nullptr || plan_out == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null plan");
    *plan_out = nullptr;
    DeviceGuard guard(plan->device);
    std::unique_ptr<pfft_plan> copy(new pfft_plan);
    copy->host = plan->host;
    copy->device = plan->device;
    copy->stream = plan->stream;
    copy->tables = plan->tables;  // shared, immutable
    copy->allow_l2_chunk = plan->allow_l2_chunk;
    copy->l2_chunk = plan->l2_chunk;
    copy->nd_chunk = plan->nd_chunk;
    attach_device_state(copy.get());  // own workspaces; staging buffers, copy streams and sub-plans are made on demand
    *plan_out = copy.release();
  });
}

pfft_status pfft_compute(pfft_plan* plan, int direction, const void* in, const void* in_imag, void* out,
                         void* out_imag, void* stream) {
  return guarded([&] {
    if (plan == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null plan");
    if (direction != PFFT_FORWARD && direction != PFFT_BACKWARD)
      throw PlanError(PFFT_INVALID_CONFIGURATION, "invalid direction");
    cudaStream_t s = stream ? (cudaStream_t)stream : plan->stream;
    DeviceGuard guard(plan->device);
    execute(plan, direction, in, in_imag, out, out_imag, s);
  });
}

pfft_status pfft_commit_guru(const pfft_desc* desc, size_t n_extra, const pfft_batch_dim* extra, int flags, int device,
                             void* stream, pfft_plan** plan_out) {
  return guarded([&] {
    if (plan_out == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null plan_out");
    *plan_out = nullptr;
    DescHost d = desc_from_c(desc);
    if (n_extra > 0 && extra == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null extra batch dimensions");
    for (size_t i = 0; i < n_extra; ++i)
      d.extra.push_back({extra[i].count, extra[i].forward_distance, extra[i].backward_distance});
    d.peer_last = (flags & PFFT_GURU_PEER_LAST_DIM) != 0;
    if (d.peer_last && n_extra == 0) throw PlanError(PFFT_INVALID_CONFIGURATION, "PFFT_GURU_PEER_LAST_DIM needs an extra dimension");
    validate_descriptor(d);
    *plan_out = make_plan(d, device, (cudaStream_t)stream);
  });
}

pfft_status pfft_compute_peer(pfft_plan* plan, int direction, const void* in, const void* in_imag, size_t n_peers,
                              void* const* out, void* const* out_imag, void* stream) {
  return guarded([&] {
    if (plan == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null plan");
    if (direction != PFFT_FORWARD) throw PlanError(PFFT_UNSUPPORTED_CONFIGURATION, "peer output: forward direction only");
    if (out == nullptr || n_peers == 0) throw PlanError(PFFT_INVALID_CONFIGURATION, "null peer table");
    const bool il = plan->host.desc.complex_storage == PFFT_INTERLEAVED_COMPLEX;
    if (!il && out_imag == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null imaginary peer table");
    PeerTable t{n_peers, out, out_imag};
    cudaStream_t s = stream ? (cudaStream_t)stream : plan->stream;
    DeviceGuard guard(plan->device);
    // the "out" argument of execute only feeds its null checks and passes that do not write the final output
    execute(plan, direction, in, in_imag, out[0], il ? nullptr : out_imag[0], s, &t);
  });
}

pfft_status pfft_compute_host(pfft_plan* plan, int direction, const void* in, const void* in_imag, void* out,
                              void* out_imag) {
  return guarded([&] {
    if (plan == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null plan");
    const DescHost& d = plan->host.desc;
    const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
    const int odir = direction == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD;
    const size_t scalar = d.is_double ? 8 : 4;
    const size_t in_elems = d.buffer_count(direction), out_elems = d.buffer_count(odir);
    // bytes of one plane per side; REAL descriptors: the real side is one plane of scalars
    const bool real = d.is_real();
    const bool in_real = real && direction == PFFT_FORWARD, out_real = real && direction == PFFT_BACKWARD;
    const size_t plane_in = in_elems * scalar * (in_real ? 1 : (il ? 2 : 1));
    const size_t plane_out = out_elems * scalar * (out_real ? 1 : (il ? 2 : 1));
    const size_t planes_in = (il || in_real) ? 1 : 2, planes_out = (il || out_real) ? 1 : 2;
    if (real && in == out)
      throw PlanError(PFFT_UNSUPPORTED_CONFIGURATION, "pfft_compute_host: REAL-domain transforms take distinct host buffers");
    const bool inplace = in == out;
    // in place on the host = in place on the device: both domains then live in ONE staged buffer, which is only
    // meaningful when they address it alike (same offset, same plane size)
    if (inplace && (plane_in != plane_out || d.forward_offset != d.backward_offset))
      throw PlanError(PFFT_UNSUPPORTED_CONFIGURATION,
                      "pfft_compute_host: an in-place call needs equal forward / backward offsets and buffer sizes");
    DeviceGuard guard(plan->device);
    const size_t need[2] = {plane_in * planes_in, inplace ? 0 : plane_out * planes_out};
    for (int i = 0; i < 2; ++i) {
      if (need[i] > plan->stage_bytes[i]) {
        if (plan->stage[i]) PFFT_CUDA_CHECK(cudaFree(plan->stage[i]));
        plan->stage[i] = nullptr;
        PFFT_CUDA_CHECK(cudaMalloc(&plan->stage[i], need[i]));
        plan->stage_bytes[i] = need[i];
      }
    }
    char* din = (char*)plan->stage[0];
    char* dout = inplace ? din : (char*)plan->stage[1];
    // Chunk along the batch when both domains are batch-major (every transform lives inside its own
    // [offset + b*distance, offset + (b+1)*distance) window); otherwise one H2D / compute / D2H sequence.
    const HostDomain hi = host_domain(d, direction), ho = host_domain(d, odir);
    const size_t batch = d.number_of_transforms;
    size_t chunks = 1;
    if (batch >= 2 && hi.extent <= hi.dist && ho.extent <= ho.dist) {
      const char* env = std::getenv("PFFT_HOST_CHUNK_BYTES");
      const size_t target = env ? (size_t)std::atoll(env) : ((size_t)32 << 20);
      chunks = std::min<size_t>(std::min<size_t>(batch, 64), std::max<size_t>(1, (plane_in * planes_in) / std::max<size_t>(1, target)));
    }
    if (chunks <= 1)
      compute_host_monolithic(plan, direction, in, in_imag, out, out_imag, din, dout, plane_in, plane_out);
    else
      compute_host_pipelined(plan, direction, in, in_imag, out, out_imag, din, dout, plane_in, plane_out, chunks);
  });
}

pfft_status pfft_destroy(pfft_plan* plan) {
  return guarded([&] {
    if (plan == nullptr) return;
    DeviceGuard guard(plan->device);
    cudaStreamSynchronize(plan->stream);  // committed_descriptor_impl.hpp:825-828
    delete plan;
  });
}

size_t pfft_workspace_bytes(const pfft_plan* plan) {
  if (!plan) return 0;
  size_t total = plan->scratch_bytes + plan->scratch2_bytes + plan->scratch3_bytes;
  for (const auto& kv : plan->child) total += pfft_workspace_bytes(kv.second);
  for (const auto& kv : plan->nd_child) total += pfft_workspace_bytes(kv.second);
  for (const pfft_plan* o : plan->nd_outer)
    if (o) total += pfft_workspace_bytes(o);
  if (plan->unfused) total += pfft_workspace_bytes(plan->unfused);
  return total;
}

size_t pfft_plan_chunk_transforms(const pfft_plan* plan) { return plan ? (plan->l2_chunk ? plan->l2_chunk : plan->nd_chunk) : 0; }

int pfft_plan_level(const pfft_plan* plan, size_t dimension) {
  if (!plan || dimension >= plan->host.dim_level.size()) return -1;
  return plan->host.dim_level[dimension];
}

size_t pfft_plan_num_launches(const pfft_plan* plan, int direction) {
  if (!plan || direction < 0 || direction > 1) return 0;
  return plan->host.passes[direction].size();
}

unsigned long long pfft_total_launches(void) { return g_total_launches.load(); }

const char* pfft_last_error(void) { return g_last_error.c_str(); }

const char* pfft_version(void) { return "pfft_b200 0.1 (sm_100a)"; }

}  // extern "C"
