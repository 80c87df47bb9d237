// Multi-GPU layer of the C ABI (include/pfft.h, "Multi-GPU"): batch sharding and the slab-decomposed 3-D transform.
//
// The reference is single-device (one sycl::queue, /root/reference/src/portfft/committed_descriptor_impl.hpp:109; no
// collective call site anywhere), so nothing here restates reference code; the descriptor vocabulary is the
// reference's.  Everything in this file sits ON TOP of the single-GPU entry points (pfft_commit_guru, pfft_compute,
// pfft_compute_peer, pfft_compute_host): a local pass of the slab transform is an ordinary plan.
//
// Slab transform of lengths (n0, n1, n2) over W ranks, XL = n0 / W, YB = n1 / W, rank r:
//   in  [XL][n1][n2]   --y pass-->  A [XL][n1][n2]
//   A                  --z pass-->  block d = rows with y in [d YB, (d+1) YB), stored STRAIGHT into rank d's window at
//                                   x = r XL ..: the kernel's stores are the all-to-all (NVLink peer memory), tile by tile
//   flag barrier (every block has landed everywhere)
//   B [n0][YB][n2]     --x pass, in place-->  the y-slab of the spectrum
// The barrier is a one-warp kernel: lane d release-stores this rank's epoch into rank d's flag array and acquire-spins
// on its own flag d.  No host round trip, no NCCL call; works between processes (windows mapped through CUDA IPC)
// and inside one process alike.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pfft.h"
#include "pass.h"
#include "plan.h"

namespace pfft {
namespace {

struct MultiError {
  pfft_status status;
  std::string msg;
};

[[noreturn]] void fail(pfft_status s, const std::string& m) { throw MultiError{s, m}; }

void check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) fail(PFFT_CUDA_ERROR, std::string(what) + " failed: " + cudaGetErrorString(e));
}
#define MULTI_CUDA(expr) check_cuda((expr), #expr)

// status of a nested C-ABI call: its message is already the thread's last error
void check(pfft_status s) {
  if (s != PFFT_OK) throw MultiError{s, std::string(pfft_last_error())};
}

template <typename F>
pfft_status guarded(F&& f) {
  try {
    f();
    return PFFT_OK;
  } catch (const MultiError& e) {
    set_last_error(e.msg);
    return e.status;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return PFFT_INTERNAL_ERROR;
  }
}

struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int device) {
    MULTI_CUDA(cudaGetDevice(&prev));
    if (prev != device) MULTI_CUDA(cudaSetDevice(device));
  }
  ~DeviceScope() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

void partition(size_t n, int world, int rank, size_t* first, size_t* count) {
  if (world <= 0 || rank < 0 || rank >= world) fail(PFFT_INVALID_CONFIGURATION, "bad rank / world size");
  const size_t base = n / (size_t)world, rem = n % (size_t)world;
  *count = base + ((size_t)rank < rem ? 1 : 0);
  *first = (size_t)rank * base + std::min<size_t>((size_t)rank, rem);
}

size_t element_bytes(const pfft_desc& d, int direction) {
  const size_t sc = d.precision == PFFT_DOUBLE ? 8 : 4;
  const bool real_side = d.domain == PFFT_DOMAIN_REAL && direction == PFFT_FORWARD;
  return (real_side || d.complex_storage == PFFT_SPLIT_COMPLEX) ? sc : 2 * sc;
}

// Descriptor of the rank-local shard.  `strides` backs the two one-entry stride arrays of a re-laid batch-interleaved
// shard; the other pointers keep referring to the caller's arrays.
struct LocalDesc {
  pfft_desc d;
  size_t strides[2];
  pfft_shard_info info;
};

void make_local(const pfft_desc* desc, int world, int rank, LocalDesc* l) {
  if (desc == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null descriptor");
  l->d = *desc;
  partition(desc->number_of_transforms, world, rank, &l->info.first, &l->info.count);
  const size_t local = std::max<size_t>(l->info.count, 1);
  l->d.number_of_transforms = local;
  const size_t nt = desc->number_of_transforms;
  auto interleaved = [&](size_t n_strides, const size_t* strides, size_t distance) {
    return desc->rank == 1 && nt > 1 && distance == 1 && n_strides == 1 && strides != nullptr && strides[0] == nt;
  };
  if (interleaved(desc->n_forward_strides, desc->forward_strides, desc->forward_distance)) {
    l->strides[0] = local;
    l->d.forward_strides = &l->strides[0];
  }
  if (interleaved(desc->n_backward_strides, desc->backward_strides, desc->backward_distance)) {
    l->strides[1] = local;
    l->d.backward_strides = &l->strides[1];
  }
  l->info.forward_start = l->info.first * desc->forward_distance;
  l->info.backward_start = l->info.first * desc->backward_distance;
}

}  // namespace
}  // namespace pfft

using namespace pfft;

// ---------------------------------------------------------------------------------------------------------------
// batch sharding
// ---------------------------------------------------------------------------------------------------------------
struct pfft_multi {
  pfft_desc desc;  // the un-sharded descriptor (array pointers re-targeted at the copies below)
  std::vector<size_t> lengths, fstrides, bstrides;
  std::vector<int> device;
  std::vector<cudaStream_t> stream;
  std::vector<bool> own_stream;
  std::vector<pfft_plan*> plan;
  std::vector<pfft_shard_info> info;
  // pfft_multi_compute_host on a descriptor with offsets: shards r > 0 start AT their first transform (offset folded
  // into the base pointer, zero offsets in the plan), so that the host ranges of two shards never overlap
  std::vector<pfft_plan*> host_plan;

  ~pfft_multi() {
    for (size_t r = 0; r < host_plan.size(); ++r)
      if (host_plan[r]) pfft_destroy(host_plan[r]);
    for (size_t r = 0; r < plan.size(); ++r)
      if (plan[r]) pfft_destroy(plan[r]);
    for (size_t r = 0; r < stream.size(); ++r)
      if (own_stream[r] && stream[r]) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device[r]);
        cudaStreamDestroy(stream[r]);
        if (prev >= 0) cudaSetDevice(prev);
      }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// slab decomposition
// ---------------------------------------------------------------------------------------------------------------
namespace pfft {
namespace {

struct BarrierArgs {
  unsigned long long* peer_flags[kMaxPeers];  // flag array inside every rank's window
  unsigned long long* my_flags;
  unsigned long long epoch;
  unsigned long long timeout_ns;
  unsigned int* status;  // set to 1 when a peer did not arrive in time
  int world, rank;
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Lane d: tell rank d that this rank has arrived (everything this rank's stream issued before -- including stores into
// peer memory -- is complete: the kernel boundary orders it, the release makes it visible system-wide), then wait for
// rank d's arrival here.  Epochs only grow, so a peer that is already one barrier ahead does no harm.
__global__ void slab_barrier_kernel(const BarrierArgs a) {
  const int d = threadIdx.x;
  if (d >= a.world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_flags[d] + a.rank), "l"(a.epoch) : "memory");
  const unsigned long long t0 = global_timer_ns();
  unsigned long long seen = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.my_flags + d) : "memory");
    if (seen >= a.epoch) break;
    if (global_timer_ns() - t0 > a.timeout_ns) {
      atomicExch(a.status, 1u);
      break;
    }
  }
  __threadfence_system();
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace pfft

struct pfft_slab {
  int world = 1, rank = 0, device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool is_double = false;
  size_t n0 = 0, n1 = 0, n2 = 0, xl = 0, yb = 0, slab_elems = 0, block_elems = 0, esz = 8;
  double backward_scale = 1.0;
  pfft_plan* py = nullptr;        // along y: in -> A (backward: A -> out, carries backward_scale)
  pfft_plan* pz_peer = nullptr;   // along z: A -> the peers' windows
  pfft_plan* pz_local = nullptr;  // along z: A <-> S (exchange through a caller collective; backward)
  pfft_plan* px = nullptr;        // along x: in place on B
  char* A = nullptr;
  char* S = nullptr;       // exchange buffer of the backward transform and of the caller-collective path
  char* window = nullptr;  // [B: slab_elems complex][flags: kMaxPeers u64]
  bool own_window = true;
  size_t window_bytes = 0, flags_offset = 0;
  unsigned int* status = nullptr;
  char* peer_window[kMaxPeers] = {};
  bool imported[kMaxPeers] = {};
  unsigned long long epoch = 0;
  unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
  pfft_alltoall_fn a2a = nullptr;
  void* a2a_user = nullptr;

  char* B() const { return window; }
  unsigned long long* flags(char* win) const { return reinterpret_cast<unsigned long long*>(win + flags_offset); }

  ~pfft_slab() {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (pfft_plan* p : {py, pz_peer, pz_local, px})
      if (p) pfft_destroy(p);
    for (int i = 0; i < kMaxPeers; ++i)
      if (imported[i] && peer_window[i]) cudaIpcCloseMemHandle(peer_window[i]);
    if (A) cudaFree(A);
    if (S) cudaFree(S);
    if (window && own_window) cudaFree(window);
    if (status) cudaFree(status);
    if (own_stream && stream) cudaStreamDestroy(stream);
    if (prev >= 0) cudaSetDevice(prev);
  }
};

namespace pfft {
namespace {

pfft_plan* commit_pass(const pfft_slab& s, size_t length, size_t transforms, size_t stride, size_t distance, bool in_place,
                       double backward_scale, std::vector<pfft_batch_dim> extra, bool peer_last) {
  pfft_desc d;
  std::memset(&d, 0, sizeof(d));
  d.precision = s.is_double ? PFFT_DOUBLE : PFFT_FLOAT;
  d.domain = PFFT_DOMAIN_COMPLEX;
  d.rank = 1;
  d.lengths = &length;
  d.forward_scale = 1.0;
  d.backward_scale = backward_scale;
  d.number_of_transforms = transforms;
  d.complex_storage = PFFT_INTERLEAVED_COMPLEX;
  d.placement = in_place ? PFFT_IN_PLACE : PFFT_OUT_OF_PLACE;
  d.n_forward_strides = d.n_backward_strides = 1;
  d.forward_strides = d.backward_strides = &stride;
  d.forward_distance = d.backward_distance = distance;
  pfft_plan* plan = nullptr;
  check(pfft_commit_guru(&d, extra.size(), extra.empty() ? nullptr : extra.data(), peer_last ? PFFT_GURU_PEER_LAST_DIM : 0,
                         s.device, s.stream, &plan));
  return plan;
}

pfft_plan* commit_z(const pfft_slab& s, bool peer) {
  // rows along z; batch = local y row, extra = (x plane, destination rank); the destination selects a peer window
  // (peer) or a block of the send buffer
  std::vector<pfft_batch_dim> extra = {{s.xl, s.n1 * s.n2, s.yb * s.n2},
                                       {(size_t)s.world, s.yb * s.n2, peer ? 0 : s.xl * s.yb * s.n2}};
  return commit_pass(s, s.n2, s.yb, 1, s.n2, false, 1.0, extra, peer);
}

void launch_barrier(pfft_slab* s) {
  BarrierArgs a;
  std::memset(&a, 0, sizeof(a));
  for (int d = 0; d < s->world; ++d) {
    if (s->peer_window[d] == nullptr)
      fail(PFFT_INVALID_CONFIGURATION, "slab: the window of rank " + std::to_string(d) + " has not been attached");
    a.peer_flags[d] = s->flags(s->peer_window[d]);
  }
  a.my_flags = s->flags(s->window);
  a.epoch = ++s->epoch;
  a.timeout_ns = s->timeout_ns;
  a.status = s->status;
  a.world = s->world;
  a.rank = s->rank;
  slab_barrier_kernel<<<1, 32, 0, s->stream>>>(a);
  MULTI_CUDA(cudaGetLastError());
}

// First launches load their modules (lazy loading) and raise shared-memory limits, both of which may synchronise the
// device: with several ranks in one process that would deadlock against a peer spinning in the barrier.  Run every
// kernel of the plan once, on the rank's own buffers, before any barrier can be pending (for the same reason every
// buffer is allocated at commit, none on first use).
void warm_up(pfft_slab* s) {
  std::vector<void*> self(s->world, s->B());
  for (int dir : {PFFT_FORWARD, PFFT_BACKWARD}) {
    check(pfft_compute(s->py, dir, s->B(), nullptr, s->A, nullptr, s->stream));
    check(pfft_compute(s->pz_local, dir, s->A, nullptr, s->S, nullptr, s->stream));
    check(pfft_compute(s->px, dir, s->B(), nullptr, s->B(), nullptr, s->stream));
  }
  check(pfft_compute_peer(s->pz_peer, PFFT_FORWARD, s->A, nullptr, (size_t)s->world, self.data(), nullptr, s->stream));
  BarrierArgs a;
  std::memset(&a, 0, sizeof(a));
  a.peer_flags[0] = a.my_flags = s->flags(s->window);
  a.world = 1;  // epoch 0: passes at once
  a.timeout_ns = s->timeout_ns;
  a.status = s->status;
  slab_barrier_kernel<<<1, 32, 0, s->stream>>>(a);
  MULTI_CUDA(cudaGetLastError());
  MULTI_CUDA(cudaMemsetAsync(s->window, 0, s->window_bytes, s->stream));
  MULTI_CUDA(cudaStreamSynchronize(s->stream));
}

// stream == nullptr: the legacy default stream, as in pfft_commit -- unless `own_stream` asks for a new non-blocking one
// (several ranks in one process must not share a stream: their barriers wait for each other)
pfft_slab* slab_commit(const pfft_desc* desc, int world, int rank, int device, cudaStream_t stream, bool own_stream) {
  if (desc == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null descriptor");
  if (world <= 0 || world > kMaxPeers || rank < 0 || rank >= world)
    fail(PFFT_INVALID_CONFIGURATION, "slab: world size must be 1.." + std::to_string(kMaxPeers) + " and 0 <= rank < world");
  check(pfft_validate(desc));
  if (desc->domain != PFFT_DOMAIN_COMPLEX || desc->complex_storage != PFFT_INTERLEAVED_COMPLEX || desc->rank != 3 ||
      desc->number_of_transforms != 1 || desc->forward_offset != 0 || desc->backward_offset != 0)
    fail(PFFT_UNSUPPORTED_CONFIGURATION,
         "slab: needs one 3-D COMPLEX transform with INTERLEAVED_COMPLEX storage and zero offsets");
  const size_t n0 = desc->lengths[0], n1 = desc->lengths[1], n2 = desc->lengths[2];
  const size_t def[3] = {n1 * n2, n2, 1};
  for (int i = 0; i < 3; ++i)
    if ((desc->n_forward_strides == 3 && desc->forward_strides[i] != def[i]) ||
        (desc->n_backward_strides == 3 && desc->backward_strides[i] != def[i]))
      fail(PFFT_UNSUPPORTED_CONFIGURATION, "slab: default strides only");
  if (n0 % (size_t)world || n1 % (size_t)world)
    fail(PFFT_INVALID_CONFIGURATION, "slab decomposition needs lengths[0] and lengths[1] divisible by the world size");
  if (desc->forward_scale != 1.0) fail(PFFT_UNSUPPORTED_CONFIGURATION, "slab: forward_scale must be 1");
  DeviceScope scope(device);
  std::unique_ptr<pfft_slab> s(new pfft_slab);
  s->world = world;
  s->rank = rank;
  s->device = device;
  s->stream = stream;
  if (s->stream == nullptr && own_stream) {
    MULTI_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->own_stream = true;
  }
  s->is_double = desc->precision == PFFT_DOUBLE;
  s->esz = s->is_double ? 16 : 8;
  s->n0 = n0, s->n1 = n1, s->n2 = n2;
  s->xl = n0 / world, s->yb = n1 / world;
  s->slab_elems = s->xl * n1 * n2;
  s->block_elems = s->xl * s->yb * n2;
  s->backward_scale = desc->backward_scale;
  if (const char* e = std::getenv("PFFT_SLAB_TIMEOUT_MS")) s->timeout_ns = (unsigned long long)std::atoll(e) * 1000000ull;
  s->flags_offset = align_up(s->slab_elems * s->esz, 256);
  s->window_bytes = s->flags_offset + kMaxPeers * sizeof(unsigned long long);
  MULTI_CUDA(cudaMalloc((void**)&s->window, s->window_bytes));
  MULTI_CUDA(cudaMalloc((void**)&s->A, s->slab_elems * s->esz));
  MULTI_CUDA(cudaMalloc((void**)&s->S, s->slab_elems * s->esz));
  MULTI_CUDA(cudaMalloc((void**)&s->status, sizeof(unsigned int)));
  MULTI_CUDA(cudaMemsetAsync(s->status, 0, sizeof(unsigned int), s->stream));
  s->peer_window[rank] = s->window;
  // y: element stride n2, batch = z (distance 1), extra = x plane
  s->py = commit_pass(*s, n1, n2, n2, 1, false, s->backward_scale, {{s->xl, n1 * n2, n1 * n2}}, false);
  s->pz_peer = commit_z(*s, true);
  s->pz_local = commit_z(*s, false);
  // x: element stride YB n2, batch-interleaved over (y row, z), in place on the window
  s->px = commit_pass(*s, n0, s->yb * n2, s->yb * n2, 1, true, 1.0, {}, false);
  warm_up(s.get());
  return s.release();
}

void slab_forward(pfft_slab* s, const void* in, void** out) {
  if (in == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null data pointer");
  DeviceScope scope(s->device);
  check(pfft_compute(s->py, PFFT_FORWARD, in, nullptr, s->A, nullptr, s->stream));
  if (s->a2a == nullptr) {
    launch_barrier(s);  // nobody still reads its window (the x pass of the previous call) when remote stores begin
    void* dst[kMaxPeers];
    for (int d = 0; d < s->world; ++d) dst[d] = s->peer_window[d] + (size_t)s->rank * s->block_elems * s->esz;
    check(pfft_compute_peer(s->pz_peer, PFFT_FORWARD, s->A, nullptr, (size_t)s->world, dst, nullptr, s->stream));
    launch_barrier(s);  // every block has landed everywhere
  } else {
    check(pfft_compute(s->pz_local, PFFT_FORWARD, s->A, nullptr, s->S, nullptr, s->stream));
    if (s->a2a(s->a2a_user, s->S, s->B(), s->block_elems * s->esz, s->stream) != 0)
      fail(PFFT_NCCL_ERROR, "slab: the caller's all-to-all reported an error");
  }
  check(pfft_compute(s->px, PFFT_FORWARD, s->B(), nullptr, s->B(), nullptr, s->stream));
  if (out) *out = s->B();
}

void slab_backward(pfft_slab* s, const void* in, void* out) {
  if (in == nullptr || out == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null data pointer");
  DeviceScope scope(s->device);
  const size_t bytes = s->slab_elems * s->esz, block = s->block_elems * s->esz;
  if (in != s->B()) MULTI_CUDA(cudaMemcpyAsync(s->B(), in, bytes, cudaMemcpyDeviceToDevice, s->stream));
  check(pfft_compute(s->px, PFFT_BACKWARD, s->B(), nullptr, s->B(), nullptr, s->stream));
  if (s->a2a == nullptr) {
    launch_barrier(s);  // every rank's x pass is complete
    for (int d = 0; d < s->world; ++d)  // pull block `rank` of every window over NVLink
      MULTI_CUDA(cudaMemcpyAsync(s->S + (size_t)d * block, s->peer_window[d] + (size_t)s->rank * block, block,
                                 cudaMemcpyDefault, s->stream));
    launch_barrier(s);  // nobody overwrites its window while peers still read it
  } else {
    if (s->a2a(s->a2a_user, s->B(), s->S, block, s->stream) != 0)
      fail(PFFT_NCCL_ERROR, "slab: the caller's all-to-all reported an error");
  }
  check(pfft_compute(s->pz_local, PFFT_BACKWARD, s->S, nullptr, s->A, nullptr, s->stream));
  check(pfft_compute(s->py, PFFT_BACKWARD, s->A, nullptr, out, nullptr, s->stream));
}

}  // namespace
}  // namespace pfft

extern "C" {

pfft_status pfft_partition(size_t n_items, int world, int rank, size_t* first, size_t* count) {
  return guarded([&] {
    if (first == nullptr || count == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null output pointer");
    partition(n_items, world, rank, first, count);
  });
}

pfft_status pfft_commit_shard(const pfft_desc* desc, int world, int rank, int device, void* stream, pfft_plan** plan_out,
                              pfft_shard_info* info) {
  return guarded([&] {
    if (plan_out == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null plan_out");
    *plan_out = nullptr;
    if (desc != nullptr) check(pfft_validate(desc));  // the un-sharded descriptor must itself be valid
    LocalDesc l;
    make_local(desc, world, rank, &l);
    check(pfft_commit(&l.d, device, stream, plan_out));
    if (info) *info = l.info;
  });
}

pfft_status pfft_commit_multi(const pfft_desc* desc, int n_dev, const int* devices, void* const* streams,
                              pfft_multi** multi_out) {
  return guarded([&] {
    if (multi_out == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null multi_out");
    *multi_out = nullptr;
    if (desc == nullptr || devices == nullptr || n_dev <= 0) fail(PFFT_INVALID_CONFIGURATION, "pfft_commit_multi: bad arguments");
    check(pfft_validate(desc));
    std::unique_ptr<pfft_multi> m(new pfft_multi);
    m->desc = *desc;
    m->lengths.assign(desc->lengths, desc->lengths + desc->rank);
    if (desc->forward_strides) m->fstrides.assign(desc->forward_strides, desc->forward_strides + desc->n_forward_strides);
    if (desc->backward_strides) m->bstrides.assign(desc->backward_strides, desc->backward_strides + desc->n_backward_strides);
    m->desc.lengths = m->lengths.data();
    m->desc.forward_strides = m->fstrides.data();
    m->desc.backward_strides = m->bstrides.data();
    for (int r = 0; r < n_dev; ++r) {
      m->device.push_back(devices[r]);
      m->stream.push_back(nullptr);
      m->own_stream.push_back(false);
      m->plan.push_back(nullptr);
      m->info.push_back(pfft_shard_info{});
      DeviceScope scope(devices[r]);
      if (streams != nullptr && streams[r] != nullptr) {
        m->stream[r] = (cudaStream_t)streams[r];
      } else {
        MULTI_CUDA(cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking));
        m->own_stream[r] = true;
      }
      check(pfft_commit_shard(&m->desc, n_dev, r, devices[r], m->stream[r], &m->plan[r], &m->info[r]));
    }
    *multi_out = m.release();
  });
}

int pfft_multi_size(const pfft_multi* multi) { return multi ? (int)multi->plan.size() : 0; }

pfft_status pfft_multi_shard(const pfft_multi* multi, int r, pfft_shard_info* info, pfft_plan** plan) {
  return guarded([&] {
    if (multi == nullptr || r < 0 || r >= (int)multi->plan.size()) fail(PFFT_INVALID_CONFIGURATION, "pfft_multi_shard: bad rank");
    if (info) *info = multi->info[r];
    if (plan) *plan = multi->plan[r];
  });
}

pfft_status pfft_multi_compute(pfft_multi* multi, int direction, const void* const* in, const void* const* in_imag,
                               void* const* out, void* const* out_imag) {
  return guarded([&] {
    if (multi == nullptr || in == nullptr || out == nullptr) fail(PFFT_INVALID_CONFIGURATION, "pfft_multi_compute: null argument");
    for (size_t r = 0; r < multi->plan.size(); ++r) {
      if (multi->info[r].count == 0) continue;
      check(pfft_compute(multi->plan[r], direction, in[r], in_imag ? in_imag[r] : nullptr, out[r],
                         out_imag ? out_imag[r] : nullptr, multi->stream[r]));
    }
  });
}

pfft_status pfft_multi_compute_host(pfft_multi* multi, int direction, const void* in, const void* in_imag, void* out,
                                    void* out_imag) {
  return guarded([&] {
    if (multi == nullptr || in == nullptr || out == nullptr) fail(PFFT_INVALID_CONFIGURATION, "pfft_multi_compute_host: null argument");
    if (direction != PFFT_FORWARD && direction != PFFT_BACKWARD) fail(PFFT_INVALID_CONFIGURATION, "invalid direction");
    const pfft_desc& d = multi->desc;
    const size_t n = multi->plan.size();
    if (n > 1 && (pfft_get_layout(&d, PFFT_FORWARD) == PFFT_LAYOUT_BATCH_INTERLEAVED ||
                  pfft_get_layout(&d, PFFT_BACKWARD) == PFFT_LAYOUT_BATCH_INTERLEAVED))
      fail(PFFT_UNSUPPORTED_CONFIGURATION,
           "pfft_multi_compute_host: batch-interleaved shards are not contiguous in the host buffers; scatter them and "
           "use pfft_multi_compute");
    const int odir = direction == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD;
    const size_t ein = element_bytes(d, direction), eout = element_bytes(d, odir);
    // A shard's host range runs from its base pointer over offset + count * distance elements; with a non-zero offset
    // the range of shard r would reach into the unaddressed prefix of shard r + 1, which that shard's thread uploads
    // and writes back concurrently.  Shards r > 0 therefore run zero-offset plans on pointers that include the offset.
    const bool offsets = d.forward_offset != 0 || d.backward_offset != 0;
    if (offsets && multi->host_plan.empty()) multi->host_plan.assign(n, nullptr);
    for (size_t r = 1; offsets && r < n; ++r) {
      if (multi->host_plan[r] != nullptr || multi->info[r].count == 0) continue;
      LocalDesc l;
      make_local(&d, (int)n, (int)r, &l);
      l.d.forward_offset = l.d.backward_offset = 0;
      check(pfft_commit(&l.d, multi->device[r], multi->stream[r], &multi->host_plan[r]));
    }
    const size_t off_in = direction == PFFT_FORWARD ? d.forward_offset : d.backward_offset;
    const size_t off_out = direction == PFFT_FORWARD ? d.backward_offset : d.forward_offset;
    std::vector<pfft_status> st(n, PFFT_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> workers;
    for (size_t r = 0; r < n; ++r) {
      if (multi->info[r].count == 0) continue;
      workers.emplace_back([&, r] {
        const pfft_shard_info& si = multi->info[r];
        const bool folded = offsets && r > 0;
        const size_t oi = ((direction == PFFT_FORWARD ? si.forward_start : si.backward_start) + (folded ? off_in : 0)) * ein;
        const size_t oo = ((direction == PFFT_FORWARD ? si.backward_start : si.forward_start) + (folded ? off_out : 0)) * eout;
        st[r] = pfft_compute_host(folded ? multi->host_plan[r] : multi->plan[r], direction, (const char*)in + oi, in_imag ? (const char*)in_imag + oi : nullptr,
                                  (char*)out + oo, out_imag ? (char*)out_imag + oo : nullptr);
        if (st[r] != PFFT_OK) msg[r] = pfft_last_error();
      });
    }
    for (std::thread& w : workers) w.join();
    for (size_t r = 0; r < n; ++r)
      if (st[r] != PFFT_OK) fail(st[r], "GPU " + std::to_string(multi->device[r]) + ": " + msg[r]);
  });
}

pfft_status pfft_multi_sync(pfft_multi* multi) {
  return guarded([&] {
    if (multi == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null multi");
    for (size_t r = 0; r < multi->plan.size(); ++r) {
      DeviceScope scope(multi->device[r]);
      MULTI_CUDA(cudaStreamSynchronize(multi->stream[r]));
    }
  });
}

pfft_status pfft_multi_destroy(pfft_multi* multi) {
  return guarded([&] { delete multi; });
}

pfft_status pfft_slab_commit(const pfft_desc* desc, int world, int rank, int device, void* stream, pfft_slab** slab_out) {
  return guarded([&] {
    if (slab_out == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null slab_out");
    *slab_out = nullptr;
    *slab_out = slab_commit(desc, world, rank, device, (cudaStream_t)stream, false);
  });
}

size_t pfft_slab_elems(const pfft_slab* slab) { return slab ? slab->slab_elems : 0; }

pfft_status pfft_slab_window(pfft_slab* slab, void** base, size_t* bytes) {
  return guarded([&] {
    if (slab == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null slab");
    if (base) *base = slab->window;
    if (bytes) *bytes = slab->window_bytes;
  });
}

pfft_status pfft_slab_export(pfft_slab* slab, void* ipc_handle) {
  return guarded([&] {
    static_assert(sizeof(cudaIpcMemHandle_t) == PFFT_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    if (slab == nullptr || ipc_handle == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null argument");
    DeviceScope scope(slab->device);
    cudaIpcMemHandle_t h;
    MULTI_CUDA(cudaIpcGetMemHandle(&h, slab->window));
    std::memcpy(ipc_handle, &h, sizeof(h));
  });
}

pfft_status pfft_slab_import(pfft_slab* slab, int peer_rank, const void* ipc_handle) {
  return guarded([&] {
    if (slab == nullptr || ipc_handle == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null argument");
    if (peer_rank < 0 || peer_rank >= slab->world) fail(PFFT_INVALID_CONFIGURATION, "slab: bad peer rank");
    if (peer_rank == slab->rank) return;  // a rank's own window is attached at commit
    DeviceScope scope(slab->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    MULTI_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (slab->imported[peer_rank] && slab->peer_window[peer_rank]) cudaIpcCloseMemHandle(slab->peer_window[peer_rank]);
    slab->peer_window[peer_rank] = (char*)p;
    slab->imported[peer_rank] = true;
  });
}

pfft_status pfft_slab_attach(pfft_slab* slab, int peer_rank, void* peer_window) {
  return guarded([&] {
    if (slab == nullptr || peer_window == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null argument");
    if (peer_rank < 0 || peer_rank >= slab->world) fail(PFFT_INVALID_CONFIGURATION, "slab: bad peer rank");
    if (peer_rank == slab->rank) return;
    if (slab->imported[peer_rank] && slab->peer_window[peer_rank]) cudaIpcCloseMemHandle(slab->peer_window[peer_rank]);
    slab->imported[peer_rank] = false;
    slab->peer_window[peer_rank] = (char*)peer_window;
  });
}

pfft_status pfft_slab_use_window(pfft_slab* slab, void* base, size_t bytes) {
  return guarded([&] {
    if (slab == nullptr || base == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null argument");
    if (bytes < slab->window_bytes) fail(PFFT_INVALID_CONFIGURATION, "slab: the window is too small");
    if (slab->epoch != 0) fail(PFFT_INVALID_CONFIGURATION, "slab: the window can only be replaced before the first transform");
    DeviceScope scope(slab->device);
    MULTI_CUDA(cudaMemsetAsync(base, 0, slab->window_bytes, slab->stream));
    MULTI_CUDA(cudaStreamSynchronize(slab->stream));
    if (slab->own_window && slab->window) MULTI_CUDA(cudaFree(slab->window));
    slab->window = (char*)base;
    slab->own_window = false;
    slab->peer_window[slab->rank] = slab->window;
  });
}

pfft_status pfft_slab_commit_local(const pfft_desc* desc, int n_dev, const int* devices, void* const* streams,
                                   pfft_slab** slabs_out) {
  return guarded([&] {
    if (slabs_out == nullptr || devices == nullptr || n_dev <= 0) fail(PFFT_INVALID_CONFIGURATION, "pfft_slab_commit_local: bad arguments");
    std::vector<std::unique_ptr<pfft_slab>> s;
    for (int r = 0; r < n_dev; ++r) slabs_out[r] = nullptr;
    for (int r = 0; r < n_dev; ++r)
      s.emplace_back(slab_commit(desc, n_dev, r, devices[r], streams ? (cudaStream_t)streams[r] : nullptr, true));
    for (int a = 0; a < n_dev; ++a) {
      DeviceScope scope(devices[a]);
      for (int b = 0; b < n_dev; ++b) {
        if (devices[a] == devices[b]) continue;
        int can = 0;
        MULTI_CUDA(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
        if (!can)
          fail(PFFT_UNSUPPORTED_CONFIGURATION, "slab: GPU " + std::to_string(devices[a]) + " cannot map the memory of GPU " +
                                                   std::to_string(devices[b]) + " (no peer access)");
        const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled)
          cudaGetLastError();
        else
          MULTI_CUDA(e);
      }
      for (int b = 0; b < n_dev; ++b) s[a]->peer_window[b] = s[b]->window;
    }
    for (int r = 0; r < n_dev; ++r) slabs_out[r] = s[r].release();
  });
}

pfft_status pfft_slab_set_alltoall(pfft_slab* slab, pfft_alltoall_fn fn, void* user) {
  return guarded([&] {
    if (slab == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null slab");
    slab->a2a = fn;
    slab->a2a_user = user;
  });
}

pfft_status pfft_slab_forward(pfft_slab* slab, const void* in_xslab, void** out_yslab) {
  return guarded([&] {
    if (slab == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null slab");
    slab_forward(slab, in_xslab, out_yslab);
  });
}

pfft_status pfft_slab_backward(pfft_slab* slab, const void* in_yslab, void* out_xslab) {
  return guarded([&] {
    if (slab == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null slab");
    slab_backward(slab, in_yslab, out_xslab);
  });
}

pfft_status pfft_slab_sync(pfft_slab* slab) {
  return guarded([&] {
    if (slab == nullptr) fail(PFFT_INVALID_CONFIGURATION, "null slab");
    DeviceScope scope(slab->device);
    MULTI_CUDA(cudaStreamSynchronize(slab->stream));
    unsigned int st = 0;
    MULTI_CUDA(cudaMemcpy(&st, slab->status, sizeof(st), cudaMemcpyDeviceToHost));
    if (st != 0) {
      MULTI_CUDA(cudaMemset(slab->status, 0, sizeof(st)));
      fail(PFFT_CUDA_ERROR, "slab: a peer did not reach the exchange barrier in time (results are invalid)");
    }
  });
}

pfft_status pfft_slab_destroy(pfft_slab* slab) {
  return guarded([&] { delete slab; });
}

}  // extern "C"
