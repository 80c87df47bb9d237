// Builds oracle/_ref/libportfft_ref.so: the REFERENCE's own code, compiled from /root/reference where it lies
// (nothing is copied), behind a C interface so that the tests can pin the oracle against it.
//   wi_dft                 /root/reference/src/portfft/common/workitem.hpp:200-219
//   sg_dft                 /root/reference/src/portfft/common/subgroup.hpp:271-291 (32 lock-step host threads)
//   factorize, wi_temps, fits_in_wi, factorize_sg, fits_in_sg   workitem.hpp:135-185, subgroup.hpp:226-253
//   wg_dft                 /root/reference/src/portfft/common/workgroup.hpp:319-346 (dimension_dft :85-286) on an
//                          emulated work-group of 2 sub-groups x 32 lock-step host threads, twiddles laid out as
//                          dispatcher/workgroup_dispatcher.hpp:382-443 does
//   validate_descriptor    /root/reference/src/portfft/descriptor_validation.hpp:264-281 (whole file)
//   get_layout             /root/reference/src/portfft/utils.hpp:210-246
// The SYCL runtime is replaced by oracle/ref_shim/sycl/sycl.hpp; the build defines are the reference's CMake
// defaults (CMakeLists.txt:38-59).  TEST INFRASTRUCTURE ONLY.
#define PORTFFT_REGISTERS_PER_WI 128
#define PORTFFT_SUBGROUP_SIZES 32
#define PORTFFT_VEC_LOAD_BYTES 16
#define PORTFFT_SGS_IN_WG 2
#define PORTFFT_MAX_CONCURRENT_KERNELS 16
#define PORTFFT_SLOW_SG_SHUFFLES 0
#define PORTFFT_UNROLL

#include <portfft/common/subgroup.hpp>
#include <portfft/common/workitem.hpp>
#include <portfft/common/memory_views.hpp>
#include <portfft/common/transfers.hpp>
#include <portfft/common/workgroup.hpp>
#include <portfft/descriptor_validation.hpp>
#include <portfft/utils.hpp>

#include <thread>
#include <vector>

namespace {

template <typename T>
void run_wi(const T* in, T* out, int n) {
  std::vector<T> priv(in, in + 2 * n), scratch(4 * 2 * 64 + 8 * n);
  portfft::wi_dft<0>(priv.data(), priv.data(), n, 1, 1, scratch.data());
  for (int i = 0; i < 2 * n; ++i) out[i] = priv[i];
}

// one transform of size factor_sg * factor_wi held by lanes 0..factor_sg-1 (lane l, slot k <- x[l*factor_wi + k])
template <typename T>
void run_sg(const T* in, T* out, int factor_wi, int factor_sg) {
  const int n = factor_wi * factor_sg;
  std::vector<T> tw(2 * n);
  for (int l = 0; l < factor_sg; ++l)
    for (int k = 0; k < factor_wi; ++k) portfft::sg_calc_twiddles<T>(factor_sg, factor_wi, l, k, tw.data());
  sycl::sg_shared_state state;
  std::vector<std::vector<T>> priv(32, std::vector<T>(2 * factor_wi, T(0)));
  for (int l = 0; l < factor_sg; ++l)
    for (int k = 0; k < factor_wi; ++k) {
      priv[l][2 * k] = in[2 * (l * factor_wi + k)];
      priv[l][2 * k + 1] = in[2 * (l * factor_wi + k) + 1];
    }
  std::vector<std::thread> lanes;
  for (int l = 0; l < 32; ++l)
    lanes.emplace_back([&, l] {
      sycl::sub_group sg(&state, static_cast<sycl::sub_group::linear_id_type>(l));
      std::vector<T> scratch(4 * 2 * 64 + 8 * factor_wi);
      portfft::sg_dft<32>(priv[l].data(), sg, factor_wi, factor_sg, tw.data(), scratch.data());
    });
  for (auto& t : lanes) t.join();
  // output is transposed: lane l slot j holds X[j*factor_sg + l] (subgroup.hpp:262-263)
  for (int l = 0; l < factor_sg; ++l)
    for (int j = 0; j < factor_wi; ++j) {
      out[2 * (j * factor_sg + l)] = priv[l][2 * j];
      out[2 * (j * factor_sg + l) + 1] = priv[l][2 * j + 1];
    }
}

// One transform of the WORKGROUP level: what workgroup_impl does for a PACKED interleaved transform
// (dispatcher/workgroup_dispatcher.hpp:232-251) -- input to padded local memory, wg_dft, transposing copy out.
template <typename T>
void run_wg(const T* in, T* out, int fft_size) {
  using namespace portfft;
  using namespace portfft::detail;
  const Idx n = factorize(static_cast<Idx>(fft_size)), m = fft_size / n;
  const Idx fsg_n = factorize_sg(n, 32), fwi_n = n / fsg_n, fsg_m = factorize_sg(m, 32), fwi_m = m / fsg_m;
  std::vector<T> tw(static_cast<std::size_t>(2 * (m + n + fft_size)));
  for (Idx a = 0; a < fsg_n; ++a)
    for (Idx k = 0; k < fwi_n; ++k) sg_calc_twiddles<T>(fsg_n, fwi_n, a, k, tw.data() + 2 * m);
  for (Idx a = 0; a < fsg_m; ++a)
    for (Idx k = 0; k < fwi_m; ++k) sg_calc_twiddles<T>(fsg_m, fwi_m, a, k, tw.data());
  for (Idx i = 0; i < n; ++i)
    for (Idx j_wi = 0; j_wi < fwi_m; ++j_wi)
      for (Idx j_sg = 0; j_sg < fsg_m; ++j_sg) {
        const Idx j = j_wi + j_sg * fwi_m, j_loc = j_wi * fsg_m + j_sg;
        const std::complex<T> w = calculate_twiddle<T>(i * j, static_cast<Idx>(fft_size));
        tw[static_cast<std::size_t>(2 * (n + m + i * m + j_loc))] = w.real();
        tw[static_cast<std::size_t>(2 * (n + m + i * m + j_loc) + 1)] = w.imag();
      }
  const Idx blp = bank_lines_per_pad_wg(2 * static_cast<Idx>(sizeof(T)) * m);
  std::vector<T> loc(static_cast<std::size_t>(pad_local(2 * fft_size, blp)) + 64, T(0));
  auto loc_view = padded_view(loc.data(), blp);
  for (Idx i = 0; i < 2 * fft_size; ++i) loc_view[i] = in[i];
  std::vector<T> loc_tw(tw.begin(), tw.begin() + 2 * (m + n));
  const T* wg_tw = tw.data() + 2 * (m + n);
  constexpr int kWg = 32 * PORTFFT_SGS_IN_WG;
  sycl::wg_shared_state wg;
  wg.size = kWg;
  std::vector<sycl::sg_shared_state> sgs(PORTFFT_SGS_IN_WG);
  std::vector<std::thread> items;
  for (int lid = 0; lid < kWg; ++lid)
    items.emplace_back([&, lid] {
      sycl::sub_group sg(&sgs[lid / 32], static_cast<sycl::sub_group::linear_id_type>(lid % 32),
                         static_cast<sycl::sub_group::linear_id_type>(lid / 32));
      sycl::nd_item<1> it(&wg, sg, static_cast<std::size_t>(lid), kWg);
      global_data_struct<1> gd(it);
      std::vector<T> wi_scratch(2 * 4 * 64 + 16 * 64), priv(2 * 64 + 64);
      wg_dft<32>(loc_view, loc_tw.data(), wg_tw, T(1), 1, 0, IdxGlobal(0), static_cast<const T*>(nullptr),
                 static_cast<const T*>(nullptr), static_cast<Idx>(fft_size), n, m, complex_storage::INTERLEAVED_COMPLEX,
                 layout::PACKED, elementwise_multiply::NOT_APPLIED, elementwise_multiply::NOT_APPLIED,
                 apply_scale_factor::NOT_APPLIED, complex_conjugate::NOT_APPLIED, complex_conjugate::NOT_APPLIED, gd,
                 wi_scratch.data(), priv.data());
    });
  for (auto& t : items) t.join();
  // local (c, b) of the n x m matrix -> output (b, c): workgroup_dispatcher.hpp:247-251
  for (Idx c = 0; c < n; ++c)
    for (Idx b = 0; b < m; ++b) {
      out[2 * (b * n + c)] = loc_view[2 * (c * m + b)];
      out[2 * (b * n + c) + 1] = loc_view[2 * (c * m + b) + 1];
    }
}

// The public fields and getters validate_descriptor / get_layout read from a portfft::descriptor
// (descriptor.hpp:59-129,161-251); portfft::descriptor itself would drag the SYCL kernels in through
// committed_descriptor.hpp.
template <typename S, portfft::domain D>
struct plain_descriptor {
  using Scalar = S;
  static constexpr portfft::domain Domain = D;
  std::vector<std::size_t> lengths, forward_strides, backward_strides;
  std::size_t number_of_transforms = 1, forward_distance = 1, backward_distance = 1;
  portfft::placement placement = portfft::placement::OUT_OF_PLACE;
  const std::vector<std::size_t>& get_strides(portfft::direction d) const {
    return d == portfft::direction::FORWARD ? forward_strides : backward_strides;
  }
  std::size_t get_distance(portfft::direction d) const {
    return d == portfft::direction::FORWARD ? forward_distance : backward_distance;
  }
  std::size_t get_flattened_length() const {
    std::size_t t = 1;
    for (std::size_t l : lengths) t *= l;
    return t;
  }
};

template <typename S>
plain_descriptor<S, portfft::domain::COMPLEX> make_desc(int placement, std::size_t rank, const std::size_t* lengths,
                                                        std::size_t nfs, const std::size_t* fs, std::size_t nbs,
                                                        const std::size_t* bs, std::size_t fd, std::size_t bd,
                                                        std::size_t batch) {
  plain_descriptor<S, portfft::domain::COMPLEX> d;
  d.lengths.assign(lengths, lengths + rank);
  d.forward_strides.assign(fs, fs + nfs);
  d.backward_strides.assign(bs, bs + nbs);
  d.forward_distance = fd;
  d.backward_distance = bd;
  d.number_of_transforms = batch;
  d.placement = placement == 0 ? portfft::placement::IN_PLACE : portfft::placement::OUT_OF_PLACE;
  return d;
}

}  // namespace

extern "C" {
// the reference's validate_descriptor: 0 accepted, 1 invalid_configuration, 2 unsupported_configuration
int refshim_validate(int is_double, int placement, std::size_t rank, const std::size_t* lengths, std::size_t nfs,
                     const std::size_t* fs, std::size_t nbs, const std::size_t* bs, std::size_t fd, std::size_t bd,
                     std::size_t batch) {
  try {
    if (is_double)
      portfft::detail::validate::validate_descriptor(make_desc<double>(placement, rank, lengths, nfs, fs, nbs, bs, fd, bd, batch));
    else
      portfft::detail::validate::validate_descriptor(make_desc<float>(placement, rank, lengths, nfs, fs, nbs, bs, fd, bd, batch));
  } catch (const portfft::invalid_configuration&) {
    return 1;
  } catch (const portfft::unsupported_configuration&) {
    return 2;
  }
  return 0;
}
// the reference's get_layout: 0 PACKED, 1 UNPACKED, 2 BATCH_INTERLEAVED (enums.hpp:46)
int refshim_get_layout(int direction, std::size_t rank, const std::size_t* lengths, const std::size_t* fs,
                       const std::size_t* bs, std::size_t fd, std::size_t bd, std::size_t batch) {
  auto d = make_desc<float>(1, rank, lengths, rank, fs, rank, bs, fd, bd, batch);
  return static_cast<int>(portfft::detail::get_layout(
      d, direction == 0 ? portfft::direction::FORWARD : portfft::direction::BACKWARD));
}
void ref_wi_dft_f32(const float* in, float* out, int n) { run_wi<float>(in, out, n); }
void ref_wi_dft_f64(const double* in, double* out, int n) { run_wi<double>(in, out, n); }
void ref_sg_dft_f32(const float* in, float* out, int factor_wi, int factor_sg) { run_sg<float>(in, out, factor_wi, factor_sg); }
void ref_sg_dft_f64(const double* in, double* out, int factor_wi, int factor_sg) { run_sg<double>(in, out, factor_wi, factor_sg); }
void ref_wg_dft_f32(const float* in, float* out, int n) { run_wg<float>(in, out, n); }
void ref_wg_dft_f64(const double* in, double* out, int n) { run_wg<double>(in, out, n); }
long long refshim_factorize(long long n) { return portfft::detail::factorize<long long>(n); }
long long refshim_wi_temps(long long n) { return portfft::detail::wi_temps<long long>(n); }
int refshim_fits_in_wi(long long n, int is_double) {
  return is_double ? portfft::detail::fits_in_wi<double>(n) : portfft::detail::fits_in_wi<float>(n);
}
long long refshim_factorize_sg(long long n, int sg) { return portfft::detail::factorize_sg<long long>(n, sg); }
int refshim_fits_in_sg(long long n, int sg, int is_double) {
  return is_double ? portfft::detail::fits_in_sg<double>(n, sg) : portfft::detail::fits_in_sg<float>(n, sg);
}
}
