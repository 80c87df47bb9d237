// Host planner. See plan.h for the reference counterparts.
#include "plan.h"

#include "tables.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <numeric>
#include <sstream>

namespace pfft {

// ---------------------------------------------------------------------------------------------------------------
// descriptor helpers (src/portfft/descriptor.hpp:161-183,262-270; src/portfft/utils.hpp:190-246)
// ---------------------------------------------------------------------------------------------------------------
size_t DescHost::flattened_length() const {
  size_t t = 1;
  for (size_t l : lengths) t *= l;
  return t;
}

std::vector<size_t> DescHost::domain_lengths(int dir) const {
  std::vector<size_t> l = lengths;
  if (is_real() && dir == PFFT_BACKWARD && !l.empty()) l.back() = l.back() / 2 + 1;
  return l;
}

size_t DescHost::buffer_count(int dir) const {
  const auto& s = strides(dir);
  const std::vector<size_t> len = domain_lengths(dir);
  size_t last = (number_of_transforms - 1) * distance(dir);
  for (size_t i = 0; i < len.size() && i < s.size(); ++i) last += (len[i] - 1) * s[i];
  for (size_t i = 0; i < extra.size(); ++i) {
    if (peer_last && i + 1 == extra.size() && dir == PFFT_BACKWARD) continue;  // selects a buffer, not an address
    last += (extra[i].count - 1) * (dir == PFFT_FORWARD ? extra[i].forward_distance : extra[i].backward_distance);
  }
  return offset(dir) + last + 1;
}

std::vector<size_t> default_strides(const std::vector<size_t>& lengths) {
  std::vector<size_t> s(lengths.size());
  size_t total = 1;
  for (size_t i = lengths.size(); i > 0; --i) {
    s[i - 1] = total;
    total *= lengths[i - 1];
  }
  return s;
}

int get_layout(const DescHost& d, int dir) {
  if (d.lengths.empty() || d.strides(dir).size() != d.lengths.size()) return PFFT_LAYOUT_UNPACKED;  // (unvalidated input)
  if (d.strides(dir) == default_strides(d.lengths) && d.distance(dir) == d.flattened_length())
    return PFFT_LAYOUT_PACKED;
  if (d.lengths.size() == 1 && d.distance(dir) == 1 && d.strides(dir).back() == d.number_of_transforms)
    return PFFT_LAYOUT_BATCH_INTERLEAVED;
  return PFFT_LAYOUT_UNPACKED;
}

DescHost desc_from_c(const pfft_desc* c) {
  if (c == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null descriptor");
  DescHost d;
  d.is_double = c->precision == PFFT_DOUBLE;
  d.domain = c->domain;
  if (c->rank > 0 && c->lengths == nullptr) throw PlanError(PFFT_INVALID_CONFIGURATION, "null lengths");
  d.lengths.assign(c->lengths, c->lengths + c->rank);
  d.forward_scale = c->forward_scale;
  d.backward_scale = c->backward_scale;
  d.number_of_transforms = c->number_of_transforms;
  d.complex_storage = c->complex_storage;
  d.placement = c->placement;
  if (c->forward_strides) d.forward_strides.assign(c->forward_strides, c->forward_strides + c->n_forward_strides);
  if (c->backward_strides) d.backward_strides.assign(c->backward_strides, c->backward_strides + c->n_backward_strides);
  d.forward_distance = c->forward_distance;
  d.backward_distance = c->backward_distance;
  d.forward_offset = c->forward_offset;
  d.backward_offset = c->backward_offset;
  return d;
}

// ---------------------------------------------------------------------------------------------------------------
// validation: same accept / reject set for *invalid* configurations as descriptor_validation.hpp
// ---------------------------------------------------------------------------------------------------------------
namespace {

template <typename... Ts>
[[noreturn]] void invalid(const Ts&... args) {
  std::stringstream ss;
  (ss << ... << args);
  throw PlanError(PFFT_INVALID_CONFIGURATION, ss.str());
}
template <typename... Ts>
[[noreturn]] void unsupported(const Ts&... args) {
  std::stringstream ss;
  (ss << ... << args);
  throw PlanError(PFFT_UNSUPPORTED_CONFIGURATION, ss.str());
}

// descriptor_validation.hpp:92-111
void check_basic(const std::vector<size_t>& lengths, size_t batch, const std::vector<size_t>& strides, size_t distance,
                 const char* dom) {
  if (strides.size() != lengths.size())
    invalid("Mismatching ", dom, " strides length got ", strides.size(), " expected ", lengths.size());
  for (size_t i = 0; i < strides.size(); ++i)
    if (strides[i] == 0) invalid("Invalid ", dom, " stride[", i, "]=", strides[i], ", must be positive");
  if (batch > 1 && distance == 0) invalid("Invalid ", dom, " distance ", distance, ", must be positive for batched FFTs");
}

// descriptor_validation.hpp:123-151
void check_multidim(const std::vector<size_t>& lengths, size_t batch, const std::vector<size_t>& strides,
                    size_t distance, const char* dom) {
  std::vector<size_t> gs = strides, gn = lengths;
  if (batch > 1) {
    gs.push_back(distance);
    gn.push_back(batch);
  }
  std::vector<size_t> idx(gn.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return gs[a] < gs[b]; });
  for (size_t i = 1; i < idx.size(); ++i)
    if (!(gs[idx[i - 1]] * gn[idx[i - 1]] <= gs[idx[i]]))
      invalid("Domain ", dom, ": multi-dimension strides are not large enough to avoid overlap");
}

// descriptor_validation.hpp:162-204 -- sub-linear in the batch count
void check_1d(const std::vector<size_t>& lengths, size_t batch, const std::vector<size_t>& strides, size_t distance,
              const char* dom) {
  const size_t fft_size = lengths[0], stride = strides[0];
  const size_t first_batch_limit = stride * fft_size;
  const size_t first_length_limit = distance * batch;
  if ((stride <= distance && first_batch_limit <= distance) || (distance <= stride && first_length_limit <= stride))
    return;
  for (size_t b = 1; b < batch;) {
    const size_t first = b * distance;
    const size_t column = first % stride;
    if (column == 0) {
      if (first >= first_batch_limit) return;
      invalid("Domain ", dom, ": batch ", b, " collides with first batch at index ", first);
    }
    size_t skip = (stride - column) / distance;
    if ((stride - column) % distance != 0) skip += 1;
    b += skip;
  }
}

void check_strides_distance(const std::vector<size_t>& lengths, size_t batch, const std::vector<size_t>& strides,
                            size_t distance, const char* dom) {
  check_basic(lengths, batch, strides, distance, dom);
  if (lengths.size() > 1)
    check_multidim(lengths, batch, strides, distance, dom);
  else
    check_1d(lengths, batch, strides, distance, dom);
}

}  // namespace

void validate_descriptor(const DescHost& d) {
  if (d.number_of_transforms == 0) invalid("Invalid number of transform ", d.number_of_transforms, ", must be positive");
  // :38-47
  if (d.lengths.empty()) invalid("Invalid lengths, must have at least 1 dimension");
  for (size_t i = 0; i < d.lengths.size(); ++i)
    if (d.lengths[i] == 0) invalid("Invalid lengths[", i, "]=", d.lengths[i], ", must be positive");
  if (d.is_real()) {
    // The reference stops here (descriptor_validation.hpp:268-270: "REAL domain is unsupported").  Forward domain:
    // real scalars of `lengths`; backward domain: lengths[last] / 2 + 1 complex elements along the last dimension.
    // Every pass that reads the input finishes before the first pass writes the output, so IN_PLACE needs no
    // relation between the two layouts beyond each being overlap-free.
    if (!d.extra.empty()) unsupported("REAL domain: extra batch dimensions are not supported");
    check_strides_distance(d.lengths, d.number_of_transforms, d.forward_strides, d.forward_distance, "forward");
    check_strides_distance(d.domain_lengths(PFFT_BACKWARD), d.number_of_transforms, d.backward_strides,
                           d.backward_distance, "backward");
    return;
  }
  // :237-253
  if (d.placement == PFFT_IN_PLACE) {
    if (d.forward_strides != d.backward_strides)
      invalid("Invalid forward and backward strides must match for in-place configurations");
    if (d.forward_distance != d.backward_distance)
      invalid("Invalid forward and backward distances must match for in-place configurations");
    check_strides_distance(d.lengths, d.number_of_transforms, d.forward_strides, d.forward_distance, "forward");
  } else {
    check_strides_distance(d.lengths, d.number_of_transforms, d.forward_strides, d.forward_distance, "forward");
    check_strides_distance(d.lengths, d.number_of_transforms, d.backward_strides, d.backward_distance, "backward");
  }
  // guru extension (no reference counterpart): plain sanity only, overlap is the caller's responsibility
  if (!d.extra.empty()) {
    if (d.lengths.size() != 1 || d.extra.size() > 2)
      unsupported("extra batch dimensions need a 1-D descriptor and at most 2 extra dimensions");
    for (const auto& e : d.extra)
      if (e.count == 0) invalid("Invalid extra batch dimension count 0");
    if (d.peer_last && d.extra.back().count > (size_t)kMaxPeers)
      unsupported("at most ", kMaxPeers, " peer output buffers");
  }
  // validate_layout (:57-81) only raises *unsupported*; those restrictions (N-D non-default strides, arbitrary
  // strides beyond one sub-group) do not exist in this implementation.
}

// ---------------------------------------------------------------------------------------------------------------
// factorisation
// ---------------------------------------------------------------------------------------------------------------
namespace {

const int kRadixSet[] = {16, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 17, 19, 23, 29, 31};

bool smooth31(size_t n) {
  for (int p : {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31})
    while (n % p == 0) n /= p;
  return n == 1;
}

struct RadixChoice {
  long cost;
  int radix;
};

RadixChoice best_radix(size_t n, std::map<size_t, RadixChoice>& memo) {
  if (n == 1) return {0, 1};
  auto it = memo.find(n);
  if (it != memo.end()) return it->second;
  RadixChoice best{1L << 60, 0};
  for (int r : kRadixSet) {
    if (n % r != 0) continue;
    RadixChoice sub = best_radix(n / r, memo);
    if (sub.cost >= (1L << 60)) continue;
    long c = 1000 + r + sub.cost;
    if (c < best.cost) best = {c, r};
  }
  memo[n] = best;
  return best;
}

int pad_index(int i, bool dbl) { return i + (i >> (dbl ? 3 : 4)); }
int pitch_for(int n, bool dbl) { return (pad_index(n - 1, dbl) + 1) | 1; }

int pow2_floor(long long v) {
  int r = 1;
  while ((long long)r * 2 <= v) r *= 2;
  return r;
}

constexpr size_t kSoftSmem = 72 * 1024;  // three CTAs per SM

// PFFT_FORCE_LEVEL=0/1/2 restricts the single-pass kernel choice (testing / benchmarking); unset: automatic
int force_level() {
  const char* env = std::getenv("PFFT_FORCE_LEVEL");
  return env ? std::atoi(env) : -1;
}

struct BDim {
  long long n, in, out;
  bool peer = false;  // selects an output buffer (never merged, may have n == 1)
};

// merge adjacent batch dimensions that are contiguous in both domains; drop unit dimensions
std::vector<BDim> merge_dims(std::vector<BDim> dims, size_t keep_front) {
  std::vector<BDim> head(dims.begin(), dims.begin() + keep_front), rest;
  for (size_t i = keep_front; i < dims.size(); ++i)
    if (dims[i].n > 1 || dims[i].peer) rest.push_back(dims[i]);
  std::stable_sort(rest.begin(), rest.end(),
                   [](const BDim& a, const BDim& b) { return std::min(a.in, a.out) < std::min(b.in, b.out); });
  std::vector<BDim> merged;
  for (const BDim& d : rest) {
    if (!merged.empty() && !d.peer && !merged.back().peer && d.in == merged.back().n * merged.back().in &&
        d.out == merged.back().n * merged.back().out)
      merged.back().n *= d.n;
    else
      merged.push_back(d);
  }
  head.insert(head.end(), merged.begin(), merged.end());
  return head;
}

}  // namespace

std::vector<int> choose_radices(size_t n) {
  if (n == 1) return {1};
  if (!smooth31(n)) return {};
  std::map<size_t, RadixChoice> memo;
  std::vector<int> out;
  while (n > 1) {
    RadixChoice c = best_radix(n, memo);
    out.push_back(c.radix);
    n /= c.radix;
  }
  std::sort(out.begin(), out.end(), std::greater<int>());
  return out;
}

static size_t wg_smem(int F, int pitch, bool dbl) { return (size_t)2 * F * pitch * (dbl ? 16 : 8) + (size_t)32 * F; }

size_t max_workgroup_length(bool is_double, const DeviceLimits& lim) {
  // largest n whose ping-pong buffers fit one CTA (F = 1)
  size_t n = 1;
  while (wg_smem(1, pitch_for((int)(n * 2), is_double), is_double) <= lim.max_smem_per_block - 1024) n *= 2;
  return n;  // 8192 (fp32) / 4096 (fp64) with 227 KB
}

namespace {

void set_batch_dims(PassParams& p, const std::vector<BDim>& dims) {
  if (dims.size() > (size_t)kMaxBatchDims)
    unsupported("transform needs ", dims.size(), " independent batch dimensions in one pass; at most ", kMaxBatchDims,
                " are supported");
  p.batch_total = 1;
  p.peer_dim = -1;
  for (int d = 0; d < kMaxBatchDims; ++d) {
    if ((size_t)d < dims.size()) {
      if (dims[d].peer) p.peer_dim = d;
      p.nb[d] = dims[d].n;
      p.ibd[d] = dims[d].in;
      p.obd[d] = dims[d].out;
      p.batch_total *= dims[d].n;
    } else {
      p.nb[d] = 1;
      p.ibd[d] = 0;
      p.obd[d] = 0;
    }
  }
}

// launch geometry, shared-memory pitch and I/O modes of the generic block-level kernel
void configure_wg_generic(PassHost& ps, bool dbl, const DeviceLimits& lim, bool force_stage) {
  PassParams& p = ps.pp;
  const int n = p.n;
  const int elem = dbl ? 16 : 8;
  const int max_threads = dbl ? 256 : 512;
  int rmax = 1;
  for (int i = 0; i < p.num_radices; ++i) rmax = std::max(rmax, p.radix[i]);
  const int t_nat = std::max(1, n / rmax);
  p.pitch = pitch_for(n, dbl);

  const bool in_batch = (p.is != 1 && p.nb[0] > 1 && p.ibd[0] < p.is) || (force_stage && p.is != 1);
  const bool out_batch = (p.os != 1 && p.nb[0] > 1 && p.obd[0] < p.os) || (force_stage && p.os != 1);
  int F, T;
  if (in_batch || out_batch) {
    F = 128 / elem;
    while (F > 32 / elem && wg_smem(F, p.pitch, dbl) > kSoftSmem) F /= 2;
    while (F > 1 && wg_smem(F, p.pitch, dbl) > lim.max_smem_per_block - 1024) F /= 2;
    while (F > 1 && F / 2 >= p.nb[0] && F / 2 >= 1) F /= 2;
    T = std::max(1, std::min(t_nat, max_threads / F));
  } else {
    T = std::min(t_nat, max_threads);
    F = std::min(64, pow2_floor(std::max(1, 256 / T)));
    while (F > 1 && wg_smem(F, p.pitch, dbl) > kSoftSmem) F /= 2;
    while (F > 1 && F / 2 >= p.batch_total) F /= 2;
  }
  if (wg_smem(F, p.pitch, dbl) > lim.max_smem_per_block - 1024)
    throw PlanError(PFFT_OUT_OF_LOCAL_MEMORY, "transform does not fit shared memory");
  p.threads_per_fft = T;
  p.ffts_per_block = F;
  p.in_mode = in_batch ? IO_STAGED_BATCH : IO_DIRECT;
  p.out_mode = out_batch ? IO_STAGED_BATCH : IO_DIRECT;
  // tiny transforms: one thread owns a whole transform; stage the contiguous tile so that global access coalesces
  if (p.in_mode == IO_DIRECT && p.is == 1 && T * elem < 32 && F > 1 && p.ibd[0] == n) p.in_mode = IO_STAGED_ELEM;
  if (p.out_mode == IO_DIRECT && p.os == 1 && T * elem < 32 && F > 1 && p.obd[0] == n) p.out_mode = IO_STAGED_ELEM;
  ps.block = F * T;
  ps.smem = wg_smem(F, p.pitch, dbl);
  const long long blocks = (p.batch_total + F - 1) / F;
  ps.grid = (int)std::min<long long>(blocks, (long long)lim.num_sms * 16);
  ps.kernel = KERNEL_WG_GENERIC;
  ps.tw_n = n;
}

void set_radices(PassParams& p, size_t n) {
  std::vector<int> r = choose_radices(n);
  if (r.empty()) unsupported("FFT size ", n, " has a prime factor larger than 31, which is not supported");
  if (r.size() > (size_t)kMaxRadices) unsupported("FFT size ", n, " needs too many radix passes");
  p.n = (int)n;
  p.num_radices = (int)r.size();
  for (size_t i = 0; i < r.size(); ++i) p.radix[i] = r[i];
}

// WORKITEM level (wi.cuh): one thread per transform.  Staged (coalesced tile copy through shared memory) on a side
// whose elements lie closer together than its batches, direct (lanes along the batch) otherwise.
bool configure_wi(PassHost& ps, bool dbl, bool interleaved, const DeviceLimits& lim) {
  PassParams& p = ps.pp;
  if (p.n > (dbl ? kWiMaxNDouble : kWiMaxNFloat) || p.gtw_dim >= 0) return false;
  p.threads_per_fft = 1;
  p.ffts_per_block = kWiBlock;
  p.pitch = p.n | 1;
  p.num_radices = 1;
  p.radix[0] = p.n;
  p.in_mode = (p.n > 1 && p.is < p.ibd[0]) ? IO_STAGED_ELEM : IO_DIRECT;
  p.out_mode = (p.n > 1 && p.os < p.obd[0]) ? IO_STAGED_ELEM : IO_DIRECT;
  if (p.nb[0] == 1) {  // a single transform per row: no batch neighbour to coalesce with
    if (p.n > 1) p.in_mode = p.out_mode = IO_STAGED_ELEM;
  }
  ps.block = kWiBlock;
  const bool staged = p.in_mode == IO_STAGED_ELEM || p.out_mode == IO_STAGED_ELEM;
  ps.smem = staged ? (size_t)kWiBlock * p.pitch * (dbl ? 16 : 8) + (size_t)kWiBlock * 16 : 0;
  const long long blocks = (p.batch_total + kWiBlock - 1) / kWiBlock;
  ps.grid = (int)std::min<long long>(blocks, (long long)lim.num_sms * 16);
  ps.kernel = KERNEL_WI;
  ps.level = LEVEL_WORKITEM;
  ps.tw_n = 0;
  // packed contiguous power-of-two rows of at most one 128-byte line, large batches: TMA tiles in and out (wi_tma.cu);
  // the geometry above stays valid as the fallback for buffers a tensor map of whole lines cannot describe
  ps.variant = 0;
  const char* env = std::getenv("PFFT_NO_WI_TMA");
  bool one_dim = true;
  for (int i = 1; i < kMaxBatchDims; ++i) one_dim = one_dim && p.nb[i] == 1;
  if (!(env && std::atoi(env) != 0) && wi_tma_supported(p.n, dbl) && interleaved && one_dim && p.is == 1 && p.os == 1 &&
      ((p.ibd[0] == p.n && p.obd[0] == p.n) ||
       (p.n * (dbl ? 16 : 8) == 128 && p.ibd[0] >= p.n && p.obd[0] >= p.n && (p.ibd[0] * (dbl ? 16 : 8)) % 16 == 0 &&
        (p.obd[0] * (dbl ? 16 : 8)) % 16 == 0)) &&
      (p.ioff * (dbl ? 16 : 8)) % 16 == 0 && (p.ooff * (dbl ? 16 : 8)) % 16 == 0 && p.peer_dim < 0 &&
      p.batch_total * p.n >= 65536)
    ps.variant = 1;
  return true;
}

// SUBGROUP level (sg.cuh): n = lanes * m with lanes a power of two <= 32 and m points per lane; unit element strides.
bool configure_sg(PassHost& ps, bool dbl, const DeviceLimits& lim) {
  PassParams& p = ps.pp;
  if (p.is != 1 || p.os != 1 || p.gtw_dim >= 0) return false;
  const int max_m = dbl ? kSgMaxMDouble : kSgMaxMFloat;
  int best_l = 0, best_m = 0;
  for (int l = 2; l <= 32; l *= 2) {
    if (p.n % l != 0) continue;
    const int m = p.n / l;
    if (m < 2 || m > max_m || !sg_supports_m(m, dbl)) continue;
    // prefer the largest m <= 16 (fewest shuffle stages per element at moderate register use), else the smallest m
    const bool better = best_m == 0 || (m <= 16 && (best_m > 16 || m > best_m)) || (m > 16 && best_m > 16 && m < best_m);
    if (better) {
      best_l = l;
      best_m = m;
    }
  }
  if (best_m == 0) return false;
  p.threads_per_fft = best_l;
  p.ffts_per_block = kSgBlock / best_l;
  p.pitch = best_m | 1;
  p.num_radices = 2;
  p.radix[0] = best_m;
  p.radix[1] = best_l;
  p.in_mode = IO_DIRECT;
  p.out_mode = IO_STAGED_ELEM;
  ps.block = kSgBlock;
  ps.smem = (size_t)kSgBlock * p.pitch * (dbl ? 16 : 8);
  const long long per_block = (long long)(kSgBlock / 32) * (32 / best_l);
  const long long blocks = (p.batch_total + per_block - 1) / per_block;
  ps.grid = (int)std::min<long long>(blocks, (long long)lim.num_sms * 4);
  ps.kernel = KERNEL_SG;
  ps.level = LEVEL_SUBGROUP;
  ps.tw_n = p.n;
  return true;
}

// Hot sizes: hand-specialised kernels (same numerics, compile-time geometry). The generic configuration stays in
// the pass as the fallback for unaligned pointers.
void select_specialised(PassHost& ps, const DescHost& d, const DeviceLimits& lim) {
  const PassParams& p = ps.pp;
  const char* env = std::getenv("PFFT_CUBE_VARIANT");
  const int variant = env ? std::atoi(env) : 0;
  if (variant < 0) return;
  const bool il = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  bool single_batch_dim = true;
  for (int i = 1; i < kMaxBatchDims; ++i) single_batch_dim = single_batch_dim && p.nb[i] == 1;
  int tile = 0, per_sm = 0;
  const char* env512 = std::getenv("PFFT_NO_CUBE512");
  if (p.n == 512 && env512 && std::atoi(env512) != 0) return;
  const char* envr3 = std::getenv("PFFT_NO_ROWS3");
  if ((p.n == 1024 || p.n == 2048 || p.n == 8192) && ((envr3 && std::atoi(envr3) != 0) || variant != 0)) return;
  if (d.is_double && variant != 0) return;  // fp64: the TMA-fed variants only
  const char* envd = std::getenv("PFFT_NO_CUBE_F64");
  if (d.is_double && envd && std::atoi(envd) != 0) return;
  if (cube_supported(p.n, d.is_double, &tile, &per_sm) && il && p.is == 1 && p.os == 1 && single_batch_dim &&
      p.gtw_dim < 0 && p.peer_dim < 0 && p.valid_in == 0 && p.valid_out == 0 &&
      (d.is_double || (p.ioff % 2 == 0 && p.ooff % 2 == 0 && p.ibd[0] % 2 == 0 && p.obd[0] % 2 == 0))) {
    ps.kernel = KERNEL_WG_CUBE;
    ps.variant = variant;
    const long long tiles = (p.batch_total + tile - 1) / tile;
    ps.alt_grid = (int)std::min<long long>(tiles, (long long)per_sm * lim.num_sms);
  }
}

// Three compile-time radix passes (wg_r3.cu) for layouts the tile kernels cannot take: replaces the generic kernel
// when both sides are accessed directly by the butterfly threads.  Rewrites the pass geometry (no fallback needed:
// the kernel takes every pointer alignment).
void select_r3(PassHost& ps, const DescHost& d, const DeviceLimits& lim) {
  PassParams& p = ps.pp;
  const char* env = std::getenv("PFFT_NO_R3");
  if (env && std::atoi(env) != 0) return;
  int tpf = 0, pitch = 0;
  if (p.gtw_dim >= 0 || p.in_mode != IO_DIRECT || p.out_mode != IO_DIRECT) return;
  if (!r3_supported(p.n, d.is_double, &tpf, &pitch)) return;
  const size_t esz = d.is_double ? 16 : 8;
  int F = std::max(1, 256 / tpf);
  while (F > 1 && (size_t)2 * F * pitch * esz > kSoftSmem) F /= 2;
  while (F > 1 && F / 2 >= p.batch_total) F /= 2;
  if (F * tpf > 512 || (size_t)2 * F * pitch * esz > lim.max_smem_per_block) return;
  p.threads_per_fft = tpf;
  p.ffts_per_block = F;
  p.pitch = pitch;
  ps.block = F * tpf;
  ps.smem = (size_t)2 * F * pitch * esz;
  const long long blocks = (p.batch_total + F - 1) / F;
  const int per_sm = std::max<int>(1, std::min<int>(2048 / ps.block, (int)((lim.max_smem_per_block + 1024) / (ps.smem + 1024))));
  ps.grid = (int)std::min<long long>(blocks, (long long)lim.num_sms * per_sm);
  ps.kernel = KERNEL_WG_R3;
}

// Generic column-tile kernel (wg_colg.cu): any 31-smooth length, any storage, transforms whose batch neighbours are
// adjacent in both domains (batch-interleaved layouts, outer dimensions of N-D transforms), in place in shared memory.
// Rewrites the pass geometry (no fallback needed).
void select_colg(PassHost& ps, const DescHost& d, const DeviceLimits& lim) {
  PassParams& p = ps.pp;
  const char* env = std::getenv("PFFT_NO_COLG");
  if (env && std::atoi(env) != 0) return;
  if (p.gtw_dim > 0 || p.peer_dim >= 0 || p.valid_in || p.valid_out || p.n < 2) return;
  if (p.ibd[0] != 1 || p.obd[0] != 1 || p.nb[0] < 2) return;
  // three compile-time radices, inputs prefetched a tile ahead (wg_colr3.cu): C3b 0.745 -> see profiles/r2_ab_variants.txt
  {
    int cols = 0, tpc = 0;
    size_t smem = 0;
    const char* e3 = std::getenv("PFFT_NO_COLR3");
    if (!(e3 && std::atoi(e3) != 0) && p.gtw_dim < 0 && colr3_supported(p.n, d.is_double, &cols, &tpc, &smem) &&
        smem <= lim.max_smem_per_block && p.nb[0] >= cols) {
      p.ffts_per_block = cols;
      p.threads_per_fft = tpc;
      p.in_mode = p.out_mode = IO_DIRECT;
      ps.block = cols * tpc;
      ps.smem = smem;
      ps.variant = 1;
      const long long tiles = ((p.nb[0] + cols - 1) / cols) * p.nb[1] * p.nb[2] * p.nb[3];
      ps.grid = (int)std::min<long long>(tiles, (long long)lim.num_sms);  // one CTA per SM (registers)
      ps.kernel = KERNEL_WG_COLG;
      return;
    }
  }
  // columns per tile: as wide as shared memory allows, up to one 128-byte row segment.  Measured on C3b (N = 1000,
  // split fp32): 16 columns (128 KB tile, one CTA per SM) 1.61 ms, 8 columns (three CTAs per SM) 1.73 ms -- the
  // kernel is bound by the number of distinct lines per memory instruction, which narrower tiles make worse.
  int c = d.is_double ? 8 : 16;
  while (c > 1 && c / 2 >= p.nb[0]) c /= 2;
  while (c > 1 && colg_smem_bytes(p.n, c, d.is_double) > lim.max_smem_per_block) c /= 2;
  if (c * (d.is_double ? 16 : 8) < 32 || colg_smem_bytes(p.n, c, d.is_double) > lim.max_smem_per_block) return;
  int rmax = 1;
  for (int i = 0; i < p.num_radices; ++i) rmax = std::max(rmax, p.radix[i]);
  const int tb = std::max(1, std::min(512 / c, p.n / rmax));
  p.ffts_per_block = c;
  p.threads_per_fft = tb;
  p.in_mode = p.out_mode = IO_DIRECT;
  ps.block = c * tb;
  ps.smem = colg_smem_bytes(p.n, c, d.is_double);
  const long long tiles = ((p.nb[0] + c - 1) / c) * p.nb[1] * p.nb[2] * p.nb[3];
  const int per_sm = std::max<int>(1, std::min<size_t>(2048 / ps.block, (lim.max_smem_per_block + 1024) / (ps.smem + 1024)));
  ps.grid = (int)std::min<long long>(tiles, (long long)lim.num_sms * per_sm);
  ps.kernel = KERNEL_WG_COLG;
}

// Tile kernel (wg_col.cu): 16 (fp32) / 8 (fp64) transforms per CTA iteration, fed by TMA.  Input side: strided columns
// whose fastest batch dimension is contiguous (TMA tensor tiles) or contiguous rows (cp.async.bulk / direct loads);
// output side: columns (fastest batch dimension contiguous) or contiguous rows.  Covers the packed 1-D sizes
// 64..512, the outer dimensions of N-D transforms and every pass of the GLOBAL level.  The generic configuration
// stays valid as the fallback (tensor map not encodable for the actual pointers).
void select_col(PassHost& ps, const DescHost& d, const DeviceLimits& lim) {
  const PassParams& p = ps.pp;
  const char* env = std::getenv("PFFT_NO_COL");
  if (env && std::atoi(env) != 0) return;
  if (d.complex_storage != PFFT_INTERLEAVED_COMPLEX || p.peer_dim >= 0) return;
  if (!col_supported(p.n, d.is_double, nullptr, nullptr)) return;
  const long long esz = d.is_double ? 16 : 8;
  auto aligned = [&](bool need_is) {
    if ((p.ioff * esz) % 16 != 0 || (need_is && (p.is * esz) % 16 != 0)) return false;
    for (int i = need_is ? 1 : 0; i < kMaxBatchDims; ++i)
      if (p.nb[i] > 1 && (p.ibd[i] * esz) % 16 != 0) return false;
    return true;
  };
  int in_mode;
  if (p.is == 1) {
    in_mode = aligned(false) ? 2 : 1;
  } else if (p.ibd[0] == 1 && p.nb[0] >= 2 && aligned(true) && p.nb[0] * 2 < (1LL << 31) && p.nb[1] < (1LL << 31) &&
             p.nb[2] < (1LL << 31) && p.nb[3] < (1LL << 31)) {
    in_mode = 0;
  } else {
    return;
  }
  bool out_rows;
  if (p.obd[0] == 1 && p.nb[0] >= 2 && p.os != 1)
    out_rows = false;
  else if (p.os == 1)
    out_rows = true;
  else
    return;
  const int c = col_tile_columns(p.n, d.is_double);
  const size_t smem = col_smem_bytes(p.n, d.is_double, in_mode != 1);
  if (smem > lim.max_smem_per_block) return;
  const long long tiles = ((p.nb[0] + c - 1) / c) * p.nb[1] * p.nb[2] * p.nb[3];
  const int per_sm = std::max<int>(1, std::min<size_t>(8, (lim.max_smem_per_block + 1024) / (smem + 1024)));
  ps.kernel = KERNEL_WG_COL;
  ps.variant = in_mode | (out_rows ? 4 : 0);
  ps.alt_grid = (int)std::min<long long>(tiles, (long long)lim.num_sms * per_sm);
}

// split n into k factors, each <= fmax, as balanced as possible
bool split_factors(size_t n, int k, size_t fmax, std::vector<size_t>& out) {
  if (k == 1) {
    if (n > fmax) return false;
    out.push_back(n);
    return true;
  }
  const double target = std::pow((double)n, 1.0 / k);
  std::vector<size_t> cands;
  for (size_t a = 2; a <= fmax && a <= n; ++a)
    if (n % a == 0) cands.push_back(a);
  std::sort(cands.begin(), cands.end(), [&](size_t a, size_t b) {
    return std::fabs(std::log((double)a) - std::log(target)) < std::fabs(std::log((double)b) - std::log(target));
  });
  for (size_t a : cands) {
    std::vector<size_t> sub;
    if (split_factors(n / a, k - 1, fmax, sub)) {
      out.push_back(a);
      out.insert(out.end(), sub.begin(), sub.end());
      return true;
    }
  }
  return false;
}

// One launch for a whole transform of length L <= max_workgroup_length: picks the level / kernel (thread, warp or
// block level; specialised block-level kernels where the layout allows).
PassHost single_pass(const DescHost& d, const DeviceLimits& lim, size_t L, long long es_in, long long es_out,
                     long long off_in, long long off_out, const std::vector<BDim>& dims, int src, int dst) {
  const bool dbl = d.is_double;
  PassHost ps;
  set_radices(ps.pp, L);
  ps.pp.is = es_in;
  ps.pp.os = es_out;
  ps.pp.ioff = off_in;
  ps.pp.ooff = off_out;
  ps.pp.gtw_dim = -1;
  set_batch_dims(ps.pp, dims);
  ps.src = src;
  ps.dst = dst;
  ps.level = LEVEL_WORKGROUP;
  const int force = force_level();
  if ((force < 0 || force == LEVEL_WORKITEM) && configure_wi(ps, dbl, d.complex_storage == PFFT_INTERLEAVED_COMPLEX, lim)) return ps;
  // block-level tile kernel first where it applies (it out-runs the warp-level kernel on B200: TMA-fed,
  // persistent); the warp-level kernel takes the unit-stride sizes it does not cover
  PassHost wg = ps;
  configure_wg_generic(wg, dbl, lim, false);
  select_specialised(wg, d, lim);
  if (wg.kernel == KERNEL_WG_GENERIC) select_col(wg, d, lim);
  if (wg.kernel == KERNEL_WG_GENERIC) select_r3(wg, d, lim);
  if (wg.kernel == KERNEL_WG_GENERIC) select_colg(wg, d, lim);
  if (force == LEVEL_WORKGROUP || (force < 0 && wg.kernel != KERNEL_WG_GENERIC)) return wg;
  if ((force < 0 || force == LEVEL_SUBGROUP) && configure_sg(ps, dbl, lim)) return ps;
  return wg;
}

// One side (source or destination) of a multi-pass transform: buffer, element stride, offset and the distance of
// every (merged) outer batch dimension.
struct View {
  int buf;
  long long es, off;
  std::vector<long long> dist;
};

// GLOBAL level for one smooth length L = N_1 * ... * N_k (k = 2..4): one pass per factor.  Pass p < k transforms the
// stride-M_p columns (first pass: src -> work, then in place on `work`) and multiplies by the inter-factor twiddle
// w_{M_{p-1}}^{c*k} on store; the last pass reads contiguous rows of `work` and writes the digit-reversed positions
// of dst, so the transposition is folded into the store addresses.  `work` holds packed rows of L elements per outer
// batch element (it may be the source buffer itself when that is packed: every pass before the last is in place).
void emit_multipass(std::vector<PassHost>& passes, const DescHost& d, const DeviceLimits& lim, size_t L,
                    const std::vector<long long>& outer_n, const View& src, int work, const View& dst) {
  const bool dbl = d.is_double;
  std::vector<size_t> factors;
  for (int k = 2; k <= 4 && factors.empty(); ++k) {
    std::vector<size_t> f;
    const bool pow2 = (L & (L - 1)) == 0 && d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
    const size_t col_max = pow2 ? (dbl ? 256 : 512) : 1024;
    if (split_factors(L, k, col_max, f)) factors = f;
  }
  if (factors.empty()) unsupported("FFT size ", L, " is too large");
  std::sort(factors.begin(), factors.end(), std::greater<size_t>());
  const size_t k = factors.size();
  // work distance of every outer batch dimension (packed rows)
  std::vector<long long> sdist(outer_n.size());
  {
    long long acc = (long long)L;
    for (size_t i = 0; i < outer_n.size(); ++i) {
      sdist[i] = acc;
      acc *= outer_n[i];
    }
  }
  long long M = (long long)L;  // M_{p-1}
  long long done = 1;          // N_1 * ... * N_{p-1}
  for (size_t pi = 0; pi < k; ++pi) {
    const long long Np = (long long)factors[pi];
    const long long Mp = M / Np;
    const bool pfirst = pi == 0, plast = pi + 1 == k;
    PassHost ps;
    set_radices(ps.pp, (size_t)Np);
    ps.level = LEVEL_GLOBAL;
    std::vector<BDim> dims;
    if (!plast) {
      // columns c' (dim 0, carries the twiddle index), combined earlier digits K, outer batches
      const long long in_unit = pfirst ? src.es : 1;
      dims.push_back({Mp, in_unit, 1});
      if (done > 1) dims.push_back({done, M, M});  // only for pi >= 1, work -> work
      for (size_t i = 0; i < outer_n.size(); ++i) dims.push_back({outer_n[i], pfirst ? src.dist[i] : sdist[i], sdist[i]});
      ps.pp.is = Mp * in_unit;
      ps.pp.os = Mp;
      ps.pp.ioff = pfirst ? src.off : 0;
      ps.pp.ooff = 0;
      ps.pp.gtw_dim = 0;
      ps.pp.gtw_n = M;
      ps.src = pfirst ? src.buf : work;
      ps.dst = work;
      set_batch_dims(ps.pp, merge_dims(dims, 1));
    } else {
      // last factor: contiguous rows of the work buffer, output digit-reversed into the destination layout
      // rows are indexed by digits k_1..k_{k-1}: work distance M_q, output distance N_1..N_{q-1}
      long long mq = (long long)L, prod = 1;
      for (size_t q = 0; q + 1 < k; ++q) {
        mq /= (long long)factors[q];
        dims.push_back({(long long)factors[q], mq, prod * dst.es});
        prod *= (long long)factors[q];
      }
      for (size_t i = 0; i < outer_n.size(); ++i) dims.push_back({outer_n[i], sdist[i], dst.dist[i]});
      ps.pp.is = 1;
      ps.pp.os = done * dst.es;
      ps.pp.ioff = 0;
      ps.pp.ooff = dst.off;
      ps.pp.gtw_dim = -1;
      ps.src = work;
      ps.dst = dst.buf;
      // dim 0 must be the unit-output-distance digit (k_1) so that the staged store coalesces along it
      set_batch_dims(ps.pp, merge_dims(dims, 1));
    }
    configure_wg_generic(ps, dbl, lim, true);
    select_col(ps, d, lim);
    if (ps.kernel == KERNEL_WG_GENERIC) select_colg(ps, d, lim);
    passes.push_back(ps);
    M = Mp;
    done *= Np;
  }
}

struct Domain {
  std::vector<size_t> strides;
  size_t distance, offset;
};

// All passes of ONE dimension of a complex transform: length L along element strides (es_in, es_out), `outer` = the
// batch dimensions this dimension sees (fastest first), reading src0 and writing dst_buf.  Single pass when L fits
// one CTA, one pass per factor (GLOBAL level) when it is smooth, Bluestein otherwise.  Returns the level.
int emit_dim(PlanHost& plan, std::vector<PassHost>& passes, const DescHost& d, const DeviceLimits& lim, size_t L,
             const std::vector<BDim>& outer, long long es_in, long long es_out, long long off_in, long long off_out,
             int src0, int dst_buf) {
  const bool dbl = d.is_double;
  const size_t wg_max = max_workgroup_length(dbl, lim);
  if (L <= wg_max && !choose_radices(L).empty()) {
    PassHost ps = single_pass(d, lim, L, es_in, es_out, off_in, off_out, merge_dims(outer, 0), src0, dst_buf);
    passes.push_back(ps);
    return ps.level;
  }

  if (d.peer_last) unsupported("peer output buffers are supported for single-pass transform lengths only");
  std::vector<BDim> outer_m = merge_dims(outer, 0);
  std::vector<long long> outer_n, outer_in, outer_out;
  long long outer_total = 1;
  for (const BDim& b : outer_m) {
    outer_n.push_back(b.n);
    outer_in.push_back(b.in);
    outer_out.push_back(b.out);
    outer_total *= b.n;
  }
  const View vin{src0, es_in, off_in, outer_in}, vout{dst_buf, es_out, off_out, outer_out};
  if (smooth31(L)) {
    // GLOBAL level: L = N_1 * ... * N_k, one pass per factor, twiddle + transposition fused into the stores
    plan.scratch_elems = std::max(plan.scratch_elems, (size_t)outer_total * L);
    emit_multipass(passes, d, lim, L, outer_n, vin, BUF_SCRATCH, vout);
    return PFFT_LEVEL_GLOBAL;
  }
  // Bluestein: a length with a prime factor > 31 becomes a circular convolution of power-of-two length M >= 2L - 1
  //   X_k = w_k * sum_j (x_j w_j) conj(w)_{k-j},  w_j = exp(-i pi j^2 / L)
  // = two length-M transforms with the chirp multiplies, the zero padding, the product with FFT_M(conj w) and the
  // truncation fused into their loads and stores.  (The reference rejects these lengths:
  // committed_descriptor_impl.hpp:241, utils.hpp:102,126.)
  size_t M = 1;
  while (M < 2 * L - 1) M *= 2;
  if (M > ((size_t)1 << 25)) unsupported("FFT size ", L, " (large prime factor) is too large");
  if (outer_m.size() + 1 > (size_t)kMaxBatchDims) unsupported("too many independent batch dimensions");
  std::vector<long long> sdist(outer_m.size());
  {
    long long acc = (long long)M;
    for (size_t i = 0; i < outer_m.size(); ++i) {
      sdist[i] = acc;
      acc *= outer_n[i];
    }
  }
  plan.scratch_elems = std::max(plan.scratch_elems, (size_t)outer_total * M);
  auto bdims = [&](const std::vector<long long>& in_d, const std::vector<long long>& out_d) {
    std::vector<BDim> v;
    for (size_t i = 0; i < outer_m.size(); ++i) v.push_back({outer_n[i], in_d[i], out_d[i]});
    return v;
  };
  auto mods = [&](PassHost& ps, int lkind, int skind, int valid_in, int valid_out, int flags) {
    ps.lmod_kind = lkind;
    ps.smod_kind = skind;
    ps.mod_l = (long long)L;
    ps.mod_m = (long long)M;
    ps.pp.valid_in = valid_in;
    ps.pp.valid_out = valid_out;
    ps.pp.mod_flags = flags;
  };
  if (M <= wg_max) {
    PassHost a;
    set_radices(a.pp, M);
    a.pp.is = es_in;
    a.pp.os = 1;
    a.pp.ioff = off_in;
    a.pp.ooff = 0;
    a.pp.gtw_dim = -1;
    set_batch_dims(a.pp, merge_dims(bdims(outer_in, sdist), 0));
    a.src = src0;
    a.dst = BUF_SCRATCH;
    configure_wg_generic(a, dbl, lim, false);
    mods(a, MODT_CHIRP, MODT_CONV, (int)L, 0, MOD_SWAP_POST | MOD_NO_USER_SWAP_OUT);
    passes.push_back(a);
    PassHost b;
    set_radices(b.pp, M);
    b.pp.is = 1;
    b.pp.os = es_out;
    b.pp.ioff = 0;
    b.pp.ooff = off_out;
    b.pp.gtw_dim = -1;
    set_batch_dims(b.pp, merge_dims(bdims(sdist, outer_out), 0));
    b.src = BUF_SCRATCH;
    b.dst = dst_buf;
    configure_wg_generic(b, dbl, lim, false);
    mods(b, MODT_NONE, MODT_CHIRP_OVER_M, 0, (int)L, MOD_SWAP_PRE | MOD_NO_USER_SWAP_IN);
    passes.push_back(b);
    return PFFT_LEVEL_WORKGROUP;
  }
  // convolution length beyond one CTA: element-wise passes around two multi-pass transforms, ping-pong between two
  // packed scratch buffers (rows of M elements)
  plan.scratch2_elems = std::max(plan.scratch2_elems, (size_t)outer_total * M);
  auto ew = [&](int src, int dst, long long n0, long long e_in, long long o_in, const std::vector<long long>& d_in,
                long long e_out, long long o_out, const std::vector<long long>& d_out) {
    PassHost ps;
    ps.pp.n = 1;
    ps.pp.num_radices = 1;
    ps.pp.radix[0] = 1;
    ps.pp.threads_per_fft = 1;
    ps.pp.ffts_per_block = 256;
    ps.pp.is = ps.pp.os = 1;
    ps.pp.ioff = o_in;
    ps.pp.ooff = o_out;
    ps.pp.gtw_dim = -1;
    std::vector<BDim> dims{{n0, e_in, e_out}};
    for (size_t i = 0; i < outer_m.size(); ++i) dims.push_back({outer_n[i], d_in[i], d_out[i]});
    set_batch_dims(ps.pp, dims);  // dimension 0 (the element index) is never merged
    ps.src = src;
    ps.dst = dst;
    ps.kernel = KERNEL_EW;
    ps.level = LEVEL_GLOBAL;
    ps.block = 256;
    ps.grid = (int)std::min<long long>((ps.pp.batch_total + 255) / 256, (long long)lim.num_sms * 16);
    ps.tw_n = 0;
    return ps;
  };
  const std::vector<long long> row_n{outer_total}, row_d{(long long)M};
  const View s1{BUF_SCRATCH, 1, 0, row_d}, s2{BUF_SCRATCH2, 1, 0, row_d};
  {
    PassHost ps = ew(src0, BUF_SCRATCH, (long long)M, es_in, off_in, outer_in, 1, 0, sdist);
    mods(ps, MODT_CHIRP, MODT_NONE, (int)L, 0, MOD_NO_USER_SWAP_OUT);
    passes.push_back(ps);
  }
  // the two length-M transforms work on plan-internal data: plain forward transforms in either direction
  const size_t inner0 = passes.size();
  emit_multipass(passes, d, lim, M, row_n, s1, BUF_SCRATCH, s2);
  {
    // times the transformed chirp, (re <-> im) swap: folded into the store of the transform's last pass when that is the
    // tile kernel reading rows (its store can index a table by the position inside the packed row); else its own pass
    PassHost& last = passes.back();
    static const bool fold_off = [] {  // PFFT_NO_BLUESTEIN_FOLD: 1 = no fold, 2 = keep this one only, 3 = the other one only
      const char* e = std::getenv("PFFT_NO_BLUESTEIN_FOLD");
      return e && (std::atoi(e) == 1 || std::atoi(e) == 3);
    }();
    if (!fold_off && last.kernel == KERNEL_WG_COL && (last.variant & 3) != 0 && (last.variant & 4) == 0 &&
        last.pp.gtw_dim < 0 && (M & (M - 1)) == 0) {
      last.smod_kind = MODT_CONV;
      last.mod_l = (long long)L;
      last.mod_m = (long long)M;
      last.pp.smod_mask = (long long)M - 1;
      last.pp.mod_flags |= MOD_SWAP_POST;
    } else {
      PassHost ps = ew(BUF_SCRATCH2, BUF_SCRATCH2, (long long)M, 1, 0, sdist, 1, 0, sdist);
      mods(ps, MODT_NONE, MODT_CONV, 0, 0, MOD_SWAP_POST | MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT);
      passes.push_back(ps);
    }
  }
  // Second transform.  Its last pass can write the user's layout itself -- (re <-> im) swap, times chirp / M,
  // truncation to the L outputs, scale, backward swap -- when it is the tile kernel reading rows of a two-factor
  // transform and the user's data is interleaved; otherwise it returns to the workspace and an element-wise pass
  // finishes.
  bool folded_out = false;
  const size_t second0 = passes.size();
  static const bool fold_off2 = [] {
    const char* e = std::getenv("PFFT_NO_BLUESTEIN_FOLD");
    return e && (std::atoi(e) == 1 || std::atoi(e) == 2);
  }();
  if (!fold_off2 && d.complex_storage == PFFT_INTERLEAVED_COMPLEX) {
    const View s2v{BUF_SCRATCH2, 1, 0, sdist};
    emit_multipass(passes, d, lim, M, outer_n, s2v, BUF_SCRATCH2, vout);
    PassHost& last = passes.back();
    if (passes.size() == second0 + 2 && last.kernel == KERNEL_WG_COL && (last.variant & 3) != 0 &&
        (last.variant & 4) == 0 && last.pp.gtw_dim < 0 && last.pp.nb[0] * (long long)last.pp.n == (long long)M) {
      last.smod_kind = MODT_CHIRP_OVER_M;
      last.mod_l = (long long)L;
      last.mod_m = (long long)M;
      last.pp.smod_n1 = (int)last.pp.nb[0];
      last.pp.valid_out = (int)L;
      last.pp.mod_flags |= MOD_SWAP_PRE;
      folded_out = true;
    } else {
      passes.resize(second0);
    }
  }
  if (!folded_out) emit_multipass(passes, d, lim, M, row_n, s2, BUF_SCRATCH2, s1);
  for (size_t i = inner0; i < passes.size(); ++i) passes[i].pp.mod_flags |= MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT;
  if (folded_out) {
    passes.back().pp.mod_flags &= ~MOD_NO_USER_SWAP_OUT;  // its output is the user's data
  } else {
    PassHost ps = ew(BUF_SCRATCH, dst_buf, (long long)L, 1, 0, sdist, es_out, off_out, outer_out);
    mods(ps, MODT_NONE, MODT_CHIRP_OVER_M, 0, 0, MOD_SWAP_PRE | MOD_NO_USER_SWAP_IN);
    passes.push_back(ps);
  }
  return PFFT_LEVEL_GLOBAL;
}

// REAL domain (see real.cu for the scheme).  Along the last dimension, forward: [pack] -> complex transform of length
// L (N/2 for even N, N for odd N) -> r2c_post; backward: c2r_pre -> complex transform -> [unpack].  The pack / unpack
// passes disappear when the real rows can be addressed as interleaved complex pairs (even N, unit stride, even offset
// and distances).  Plan-internal rows live packed and interleaved in the workspaces, whatever the descriptor's
// complex storage.  N-D: forward = the real-to-complex rows first, then complex passes along the other dimensions in
// place on the half-spectrum output; backward = the half spectrum copied to a packed workspace, inverse complex
// passes along the other dimensions there, the complex-to-real rows last (the Hermitian symmetry that c2r_pre relies
// on holds along the last dimension only after the other dimensions are transformed).
void build_real_direction(PlanHost& plan, int dir, const DeviceLimits& lim) {
  const DescHost& d = plan.desc;
  DescHost dint = d;  // the descriptor as the inner complex passes see it
  dint.complex_storage = PFFT_INTERLEAVED_COMPLEX;
  const bool dbl = d.is_double;
  const bool fwd = dir == PFFT_FORWARD;
  const size_t D = d.lengths.size();
  const std::vector<size_t> clen = d.domain_lengths(PFFT_BACKWARD);
  const long long N = (long long)d.lengths[D - 1], H = N / 2;
  const bool even = N % 2 == 0;
  const int variant = even ? 0 : 1;
  const long long L = even ? H : N;
  const long long rs = (long long)d.forward_strides[D - 1], roff = (long long)d.forward_offset;
  const long long cs = (long long)d.backward_strides[D - 1], coff = (long long)d.backward_offset;
  const size_t wg_max = max_workgroup_length(dbl, lim);
  std::vector<PassHost>& passes = plan.passes[dir];
  if (!smooth31((size_t)L))
    unsupported("REAL domain: length ", N, " needs a complex transform of length ", L,
                " with a prime factor larger than 31, which is not supported");
  const bool single = (size_t)L <= wg_max;
  plan.dim_level.assign(D, single ? PFFT_LEVEL_WORKGROUP : PFFT_LEVEL_GLOBAL);
  // rows of the last dimension: (i_{D-2}, ..., i_0, batch), fastest first; real / complex / workspace distances.
  // The workspace W (backward, N-D) holds the half spectrum packed: [batch][d_0]..[d_{D-2}][H + 1].
  std::vector<long long> row_n, row_r, row_c, row_w, row_s;
  {
    long long wacc = H + 1;
    for (size_t e = D - 1; e > 0; --e) {
      row_n.push_back((long long)d.lengths[e - 1]);
      row_r.push_back((long long)d.forward_strides[e - 1]);
      row_c.push_back((long long)d.backward_strides[e - 1]);
      row_w.push_back(wacc);
      wacc *= (long long)d.lengths[e - 1];
    }
    row_n.push_back((long long)d.number_of_transforms);
    row_r.push_back((long long)d.forward_distance);
    row_c.push_back((long long)d.backward_distance);
    row_w.push_back(wacc);
    long long sacc = L;
    for (size_t i = 0; i < row_n.size(); ++i) {
      row_s.push_back(sacc);  // packed rows of L elements in the scratch buffers
      sacc *= row_n[i];
    }
  }
  long long rows = 1;
  for (long long n : row_n) rows *= n;
  bool pairs = even && rs == 1 && roff % 2 == 0;
  std::vector<long long> pair_d;
  for (size_t i = 0; i < row_n.size(); ++i) {
    if (row_n[i] > 1 && row_r[i] % 2 != 0) pairs = false;
    pair_d.push_back(row_r[i] / 2);
  }
  const View s1{BUF_SCRATCH, 1, 0, row_s}, s2{BUF_SCRATCH2, 1, 0, row_s};
  const View rview{fwd ? BUF_IN : BUF_OUT, 1, roff / 2, pair_d};  // the real rows as complex pairs

  auto row_dims = [&](const std::vector<long long>& d_in, const std::vector<long long>& d_out) {
    std::vector<BDim> v;
    for (size_t i = 0; i < row_n.size(); ++i) v.push_back({row_n[i], d_in[i], d_out[i]});
    return merge_dims(v, 0);
  };
  auto rows_pass = [&](int kernel, int src, int dst, long long es_in, long long o_in, const std::vector<long long>& d_in,
                       long long es_out, long long o_out, const std::vector<long long>& d_out, long long work_per_row) {
    PassHost ps;
    ps.pp.n = (int)N;
    ps.pp.num_radices = 1;
    ps.pp.radix[0] = 1;
    ps.pp.threads_per_fft = 1;
    ps.pp.ffts_per_block = 256;
    ps.pp.is = es_in;
    ps.pp.os = es_out;
    ps.pp.ioff = o_in;
    ps.pp.ooff = o_out;
    ps.pp.gtw_dim = -1;
    set_batch_dims(ps.pp, row_dims(d_in, d_out));
    ps.src = src;
    ps.dst = dst;
    ps.kernel = kernel;
    ps.variant = variant;
    ps.level = single ? LEVEL_WORKGROUP : LEVEL_GLOBAL;
    ps.block = 256;
    {
      // real.cu: a CTA takes chunks of about 2048 elements = 2048 / (elements per row) consecutive rows
      const long long chunk = work_per_row >= 2048 ? 1 : 2048 / std::max<long long>(1, work_per_row);
      ps.grid = (int)std::min<long long>((rows + chunk - 1) / chunk, (long long)lim.num_sms * 16);
    }
    ps.tw_n = (even && (kernel == KERNEL_R2C_POST || kernel == KERNEL_C2R_PRE)) ? N : 0;
    // the scratch side of these passes is interleaved whatever the descriptor says
    ps.internal_storage = (src != BUF_IN ? 1 : 0) | (dst != BUF_OUT ? 2 : 0);
    return ps;
  };
  // complex transform of length L between two views; returns the buffer holding the result
  auto transform = [&](const View& src, const View* final_dst) {
    const size_t first = passes.size();
    int result;
    if (single) {
      const View dst = final_dst ? *final_dst : View{src.buf == BUF_IN ? BUF_SCRATCH : src.buf, 1, 0, row_s};
      passes.push_back(single_pass(dint, lim, (size_t)L, src.es, dst.es, src.off, dst.off, row_dims(src.dist, dst.dist),
                                   src.buf, dst.buf));
      result = dst.buf;
    } else {
      // multi-pass: in place on a packed workspace, the last pass moves to the other one (or to the final view)
      const int work = src.buf == BUF_IN ? BUF_SCRATCH2 : src.buf;
      const View other = work == BUF_SCRATCH ? s2 : s1;
      const View dst = final_dst ? *final_dst : other;
      if (!final_dst || src.buf == BUF_IN) plan.scratch2_elems = std::max(plan.scratch2_elems, (size_t)(rows * L));
      emit_multipass(passes, dint, lim, (size_t)L, row_n, src, work, dst);
      result = dst.buf;
    }
    for (size_t i = first; i < passes.size(); ++i) {
      passes[i].pp.mod_flags |= MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT;  // plain forward transforms
      passes[i].internal_storage = 3;
    }
    return result;
  };
  // batch dimensions seen by a complex transform along dimension `dim` of the half-spectrum array (fastest first)
  auto outer_of = [&](size_t dim, const std::vector<long long>& strides, long long distance) {
    std::vector<BDim> outer;
    for (size_t e = D; e > 0; --e)
      if (e - 1 != dim) outer.push_back({(long long)clen[e - 1], strides[e - 1], strides[e - 1]});
    outer.push_back({(long long)d.number_of_transforms, distance, distance});
    return outer;
  };

  // Pre / post-processing fused into the TMA-fed tile kernel of the rows (wg_cube.cu): the transform pass itself reads
  // / writes the user's half spectrum, one HBM round trip instead of two.  Needs the packed-row tile kernel, interleaved
  // unit-stride complex rows and a single (merged) batch dimension on both sides.
  const bool il_user = d.complex_storage == PFFT_INTERLEAVED_COMPLEX;
  static const bool fuse_off = [] {
    const char* e = std::getenv("PFFT_NO_REAL_FUSE");
    return e && std::atoi(e) != 0;
  }();
  auto fusable = [&](const PassHost& ps, const std::vector<BDim>& dims) {
    if (fuse_off || d.no_real_fuse || !even || !pairs || !single || !il_user || dims.size() != 1) return false;
    if (ps.kernel == KERNEL_WG_CUBE) return ps.variant == 0;
    if (ps.kernel == KERNEL_WI) {
      // thread-level TMA kernel, one real row per 128-byte line (wi_tma.cu): half-spectrum rows that do not overlap,
      // 16-byte aligned start (dense rows move by bulk copies)
      const long long spec_dist = fwd ? dims[0].out : dims[0].in, spec_off = fwd ? coff : (D > 1 ? 0 : coff);
      return ps.variant == 1 && wi_tma_real_supported(ps.pp.n, dbl) && (dims[0].n == 1 || spec_dist >= H + 1) &&
             (spec_off * (dbl ? 16 : 8)) % 16 == 0;
    }
    // The pass was planned for another kernel: a half length whose complex transform runs elsewhere (256), or rows
    // that are only 8-byte aligned (in-place layouts: rows of n + 2 reals = n/2 + 1 pairs apart) -- the REAL forms
    // of the tile kernel take both (their bulk copies start at the 16-byte boundary below the row, wg_cube.cu).
    const PassParams& p = ps.pp;
    bool one_dim = true;
    for (int i = 1; i < kMaxBatchDims; ++i) one_dim = one_dim && p.nb[i] == 1;
    return cube_real_supported(p.n, dbl, nullptr, nullptr) && p.is == 1 && p.os == 1 && one_dim && p.gtw_dim < 0 &&
           p.peer_dim < 0 && p.valid_in == 0 && p.valid_out == 0 && dims[0].in >= p.n + (fwd ? 0 : 1) &&
           dims[0].out >= p.n + (fwd ? 1 : 0);
  };
  auto set_fused = [&](PassHost& ps, int mode, const std::vector<BDim>& dims) {
    if (ps.kernel != KERNEL_WG_CUBE && ps.kernel != KERNEL_WI) {
      int tile = 1, per_sm = 1;
      cube_real_supported(ps.pp.n, dbl, &tile, &per_sm);
      ps.kernel = KERNEL_WG_CUBE;
      ps.variant = 0;
      ps.alt_grid = (int)std::min<long long>((ps.pp.batch_total + tile - 1) / tile, (long long)per_sm * lim.num_sms);
    }
    set_batch_dims(ps.pp, dims);
    ps.fuse_real = mode;
    ps.tw2_n = ps.kernel == KERNEL_WG_CUBE ? N : 0;  // (the thread-level kernel's twiddles are compile-time constants)
  };

  if (fwd) {
    int zbuf;
    bool fused = false;
    if (pairs) {
      zbuf = transform(rview, nullptr);
      const std::vector<BDim> dims = row_dims(rview.dist, row_c);
      if (passes.size() == 1 && cs == 1 && fusable(passes.back(), dims)) {
        PassHost& ps = passes.back();
        set_fused(ps, 1, dims);
        ps.dst = BUF_OUT;
        ps.pp.ooff = coff;
        ps.internal_storage = 0;
        fused = true;
      }
    } else {
      passes.push_back(rows_pass(KERNEL_REAL_PACK, BUF_IN, BUF_SCRATCH, rs, roff, row_r, 1, 0, row_s, L));
      zbuf = transform(s1, nullptr);
    }
    if (!fused) passes.push_back(rows_pass(KERNEL_R2C_POST, zbuf, BUF_OUT, 1, 0, row_s, cs, coff, row_c, H + 1));
    // the other dimensions: complex passes in place on the half-spectrum output
    std::vector<long long> bst(d.backward_strides.begin(), d.backward_strides.end());
    for (size_t e = D - 1; e > 0; --e) {
      const size_t dim = e - 1;
      plan.dim_level[dim] = emit_dim(plan, passes, d, lim, d.lengths[dim], outer_of(dim, bst, (long long)d.backward_distance),
                                     bst[dim], bst[dim], coff, coff, BUF_OUT, BUF_OUT);
    }
  } else {
    int xbuf = BUF_IN;
    long long xs = cs, xoff = coff;
    std::vector<long long> xd = row_c;
    if (D > 1) {
      // half spectrum -> packed interleaved workspace W, then inverse complex passes along the other dimensions on W
      std::vector<long long> wst(D);
      {
        long long acc = 1;
        for (size_t e = D; e > 0; --e) {
          wst[e - 1] = acc;
          acc *= (long long)clen[e - 1];
        }
        plan.scratch3_elems = std::max(plan.scratch3_elems, (size_t)(acc * (long long)d.number_of_transforms));
      }
      const long long wdist = wst[0] * (long long)clen[0];
      {
        PassHost ps;
        ps.pp.n = 1;
        ps.pp.num_radices = 1;
        ps.pp.radix[0] = 1;
        ps.pp.threads_per_fft = 1;
        ps.pp.ffts_per_block = 256;
        ps.pp.is = ps.pp.os = 1;
        ps.pp.ioff = coff;
        ps.pp.ooff = 0;
        ps.pp.gtw_dim = -1;
        std::vector<BDim> dims{{H + 1, cs, 1}};
        std::vector<BDim> rest;
        for (size_t i = 0; i < row_n.size(); ++i) rest.push_back({row_n[i], row_c[i], row_w[i]});
        rest = merge_dims(rest, 0);
        dims.insert(dims.end(), rest.begin(), rest.end());
        set_batch_dims(ps.pp, dims);  // dimension 0 (the element index) is never merged
        ps.src = BUF_IN;
        ps.dst = BUF_SCRATCH3;
        ps.kernel = KERNEL_EW;
        ps.level = LEVEL_WORKGROUP;
        ps.block = 256;
        ps.grid = (int)std::min<long long>((ps.pp.batch_total + 255) / 256, (long long)lim.num_sms * 16);
        ps.internal_storage = 2;
        ps.pp.mod_flags = MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT;
        passes.push_back(ps);
      }
      for (size_t e = D - 1; e > 0; --e) {
        const size_t dim = e - 1;
        const size_t first = passes.size();
        plan.dim_level[dim] = emit_dim(plan, passes, dint, lim, d.lengths[dim], outer_of(dim, wst, wdist), wst[dim], wst[dim],
                                       0, 0, BUF_SCRATCH3, BUF_SCRATCH3);
        for (size_t i = first; i < passes.size(); ++i) {
          passes[i].internal_storage = 3;
          passes[i].force_swap = 1;  // unnormalised inverse = forward transform of the (re <-> im)-swapped data
        }
      }
      xbuf = BUF_SCRATCH3;
      xs = 1;
      xoff = 0;
      xd = row_w;
    }
    bool fused = false;
    if (pairs && xs == 1) {
      // try the fused form: plan the transform on packed rows (what the unfused plan runs), then let it read the half
      // spectrum itself (rows of H + 1 elements at any 8-byte alignment: wg_cube.cu handles the shift)
      const size_t first = passes.size();
      transform(s1, &rview);
      const std::vector<BDim> dims = row_dims(xd, rview.dist);
      if (passes.size() == first + 1 && fusable(passes.back(), dims)) {
        PassHost& ps = passes.back();
        set_fused(ps, 2, dims);
        ps.src = xbuf;
        ps.pp.ioff = xoff;
        ps.internal_storage = xbuf == BUF_IN ? 0 : 1;
        fused = true;
      } else {
        passes.resize(first);
      }
    }
    if (fused) {
      // (nothing else: the pass reads xbuf and writes the real rows)
    } else {
    passes.push_back(rows_pass(KERNEL_C2R_PRE, xbuf, BUF_SCRATCH, xs, xoff, xd, 1, 0, row_s, even ? H / 2 + 1 : (N + 1) / 2));
    if (pairs) {
      transform(s1, &rview);
    } else {
      const int ybuf = transform(s1, nullptr);
      passes.push_back(rows_pass(KERNEL_REAL_UNPACK, ybuf, BUF_OUT, 1, 0, row_s, rs, roff, row_r, L));
    }
    }
  }
  // the packed-row workspace, unless every pass of this direction works on the user's buffers (fused forms)
  for (const PassHost& ps : passes)
    if (ps.src == BUF_SCRATCH || ps.dst == BUF_SCRATCH) plan.scratch_elems = std::max(plan.scratch_elems, (size_t)(rows * L));
  // passes that address the user's real buffer as complex pairs
  for (PassHost& ps : passes) {
    if (ps.kernel >= KERNEL_EW && ps.kernel <= KERNEL_REAL_UNPACK) continue;
    if (fwd && ps.src == BUF_IN) ps.real_view |= 1;
    if (!fwd && ps.dst == BUF_OUT) ps.real_view |= 2;
  }
  const double scale = d.scale(dir);
  PassHost& last = passes.back();
  last.pp.scale = d.is_double ? scale : (double)(float)scale;
  last.pp.apply_scale = (d.is_double ? scale : (double)(float)scale) != 1.0;
}

void build_direction(PlanHost& plan, int dir, const DeviceLimits& lim) {
  const DescHost& d = plan.desc;
  if (d.is_real()) {
    build_real_direction(plan, dir, lim);
    return;
  }
  const Domain in{d.strides(dir), d.distance(dir), d.offset(dir)};
  const int odir = dir == PFFT_FORWARD ? PFFT_BACKWARD : PFFT_FORWARD;
  const Domain out{d.strides(odir), d.distance(odir), d.offset(odir)};
  const size_t D = d.lengths.size();
  std::vector<PassHost>& passes = plan.passes[dir];
  if (plan.dim_level.size() != D) plan.dim_level.assign(D, PFFT_LEVEL_WORKGROUP);

  for (size_t step = 0; step < D; ++step) {
    const size_t dim = D - 1 - step;
    const bool first = step == 0;
    const size_t L = d.lengths[dim];
    // batch dimensions seen by this dimension's transform (fastest first)
    std::vector<BDim> outer;
    for (size_t e = D; e > 0; --e) {
      if (e - 1 == dim) continue;
      outer.push_back({(long long)d.lengths[e - 1], (long long)(first ? in.strides[e - 1] : out.strides[e - 1]),
                       (long long)out.strides[e - 1]});
    }
    outer.push_back({(long long)d.number_of_transforms, (long long)(first ? in.distance : out.distance),
                     (long long)out.distance});
    for (size_t e = 0; e < d.extra.size(); ++e) {
      const long long fd = (long long)d.extra[e].forward_distance, bd = (long long)d.extra[e].backward_distance;
      BDim b{(long long)d.extra[e].count, dir == PFFT_FORWARD ? fd : bd, dir == PFFT_FORWARD ? bd : fd};
      // the peer dimension addresses the forward domain normally and selects a buffer in the backward domain
      if (d.peer_last && e + 1 == d.extra.size() && dir == PFFT_FORWARD) {
        b.peer = true;
        b.out = 0;
      }
      outer.push_back(b);
    }
    const long long es_in = (long long)(first ? in.strides[dim] : out.strides[dim]);
    const long long es_out = (long long)out.strides[dim];
    const long long off_in = (long long)(first ? in.offset : out.offset);
    const long long off_out = (long long)out.offset;
    const int src0 = first ? BUF_IN : BUF_OUT;

    plan.dim_level[dim] = emit_dim(plan, passes, d, lim, L, outer, es_in, es_out, off_in, off_out, src0, BUF_OUT);
  }
  // scale on the last pass executed (committed_descriptor_impl.hpp:473-474)
  const double scale = d.scale(dir);
  PassHost& last = passes.back();
  last.pp.scale = d.is_double ? scale : (double)(float)scale;
  last.pp.apply_scale = (d.is_double ? scale : (double)(float)scale) != 1.0;
}

}  // namespace

PlanHost build_plan(const DescHost& d, const DeviceLimits& lim) {
  validate_descriptor(d);
  PlanHost plan;
  plan.desc = d;
  build_direction(plan, PFFT_FORWARD, lim);
  build_direction(plan, PFFT_BACKWARD, lim);
  return plan;
}

std::string describe_plan(const PlanHost& plan, int direction) {
  static const char* level_names[] = {"WORKITEM", "SUBGROUP", "WORKGROUP", "GLOBAL"};
  static const char* kernel_names[] = {"wg_generic", "wi", "sg", "wg_cube", "wg_col", "wg_r3", "ew", "real_pack", "r2c_post", "c2r_pre", "real_unpack", "wg_colg"};
  static const char* mode_names[] = {"direct", "staged_elem", "staged_batch"};
  static const char* buf_names[] = {"in", "out", "scratch", "scratch2", "scratch3"};
  static const char* mod_names[] = {"none", "chirp", "chirp/M", "conv"};
  std::stringstream ss;
  ss << "levels:";
  for (int l : plan.dim_level) ss << " " << level_names[l];
  ss << "; scratch_elems=" << plan.scratch_elems;
  if (plan.scratch2_elems) ss << " scratch2_elems=" << plan.scratch2_elems;
  if (plan.scratch3_elems) ss << " scratch3_elems=" << plan.scratch3_elems;
  ss << "\n";
  for (const PassHost& ps : plan.passes[direction]) {
    const PassParams& p = ps.pp;
    ss << "pass kernel=" << kernel_names[ps.kernel] << " level=" << level_names[ps.level] << " n=" << p.n << " radices=";
    for (int i = 0; i < p.num_radices; ++i) ss << (i ? "x" : "") << p.radix[i];
    ss << " T=" << p.threads_per_fft << " F=" << p.ffts_per_block << " block=" << ps.block << " grid=" << ps.grid
       << " smem=" << ps.smem << " " << buf_names[ps.src] << "->" << buf_names[ps.dst] << " in=" << mode_names[p.in_mode]
       << " out=" << mode_names[p.out_mode] << " is=" << p.is << " os=" << p.os << " batch=[";
    for (int dd = 0; dd < kMaxBatchDims; ++dd)
      if (p.nb[dd] > 1 || dd == 0) ss << (dd ? " " : "") << p.nb[dd] << ":" << p.ibd[dd] << ":" << p.obd[dd];
    ss << "] gtw_dim=" << p.gtw_dim;
    if (p.peer_dim >= 0) ss << " peer_dim=" << p.peer_dim;
    if (ps.kernel == KERNEL_WG_COL) {
      static const char* in_names[] = {"cols_tma", "rows_direct", "rows_bulk", "?"};
      ss << " tile_in=" << in_names[ps.variant & 3] << " tile_out=" << ((ps.variant & 4) ? "rows" : "cols")
         << " tile_grid=" << ps.alt_grid << " (generic geometry above is the fallback)";
    }
    if (p.gtw_dim >= 0) ss << " gtw_n=" << p.gtw_n;
    if (ps.lmod_kind || ps.smod_kind || p.valid_in || p.valid_out || p.mod_flags)
      ss << " lmod=" << mod_names[ps.lmod_kind] << " smod=" << mod_names[ps.smod_kind] << " valid_in=" << p.valid_in
         << " valid_out=" << p.valid_out << " mod_flags=" << p.mod_flags << " mod_l=" << ps.mod_l << " mod_m=" << ps.mod_m;
    if (p.apply_scale) ss << " scale=" << p.scale;
    if (ps.real_view) ss << " real_view=" << ps.real_view;
    if (ps.fuse_real) ss << " fuse_real=" << ps.fuse_real;
    if (ps.force_swap) ss << " force_swap";
    if (ps.kernel >= KERNEL_REAL_PACK && ps.kernel <= KERNEL_REAL_UNPACK) ss << " variant=" << ps.variant;
    ss << "\n";
  }
  return ss.str();
}

std::string export_plan_json(const PlanHost& plan, int direction) {
  std::stringstream ss;
  ss.precision(17);
  auto arr = [&](const char* name, const long long* v, int n) {
    ss << "\"" << name << "\": [";
    for (int i = 0; i < n; ++i) ss << (i ? ", " : "") << v[i];
    ss << "]";
  };
  ss << "{\"scratch_elems\": " << plan.scratch_elems << ", \"scratch2_elems\": " << plan.scratch2_elems
     << ", \"scratch3_elems\": " << plan.scratch3_elems
     << ", \"is_double\": " << (plan.desc.is_double ? 1 : 0) << ", \"is_real\": " << (plan.desc.is_real() ? 1 : 0)
     << ", \"passes\": [";
  bool firstp = true;
  for (const PassHost& ps : plan.passes[direction]) {
    const PassParams& p = ps.pp;
    ss << (firstp ? "" : ", ") << "{\"kernel\": " << ps.kernel << ", \"level\": " << ps.level << ", \"src\": " << ps.src
       << ", \"dst\": " << ps.dst << ", \"n\": " << p.n << ", \"is\": " << p.is << ", \"os\": " << p.os
       << ", \"ioff\": " << p.ioff << ", \"ooff\": " << p.ooff << ", ";
    arr("nb", p.nb, kMaxBatchDims);
    ss << ", ";
    arr("ibd", p.ibd, kMaxBatchDims);
    ss << ", ";
    arr("obd", p.obd, kMaxBatchDims);
    ss << ", \"gtw_dim\": " << p.gtw_dim << ", \"gtw_n\": " << (p.gtw_dim >= 0 ? p.gtw_n : 0) << ", \"peer_dim\": " << p.peer_dim
       << ", \"valid_in\": " << p.valid_in << ", \"valid_out\": " << p.valid_out << ", \"mod_flags\": " << p.mod_flags
       << ", \"lmod\": " << ps.lmod_kind << ", \"smod\": " << ps.smod_kind << ", \"mod_l\": " << ps.mod_l
       << ", \"mod_m\": " << ps.mod_m << ", \"apply_scale\": " << p.apply_scale << ", \"scale\": " << p.scale
       << ", \"variant\": " << ps.variant << ", \"real_view\": " << ps.real_view << ", \"fuse_real\": " << ps.fuse_real
       << ", \"smod_mask\": " << ps.pp.smod_mask << ", \"smod_n1\": " << ps.pp.smod_n1
       << ", \"force_swap\": " << ps.force_swap << "}";
    firstp = false;
  }
  ss << "]}";
  return ss.str();
}

}  // namespace pfft
