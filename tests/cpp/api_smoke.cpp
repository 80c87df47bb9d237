// C++ smoke / parity test of the drop-in header: reads like a reference unit test
// (/root/reference/test/unit_test/fft_test_utils.hpp:276-347) with sycl::queue -> portfft::queue.
// Checks a forward/backward round trip and a naive DFT on the host, in-place interleaved and out-of-place split,
// plus the error convention (invalid_configuration from commit, storage mismatch from compute).
#include <portfft/portfft.hpp>

#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

using namespace portfft;

template <typename T>
static double check(std::size_t n, std::size_t batch) {
  descriptor<T, domain::COMPLEX> desc({n});
  desc.number_of_transforms = batch;
  desc.placement = placement::IN_PLACE;
  desc.backward_scale = T(1) / T(n);
  cudaStream_t s;
  cudaStreamCreate(&s);
  double fwd_err = 0, err = 0, nrm = 0;
  std::complex<T>* dev = nullptr;
  {  // the committed descriptor must be gone before its stream is destroyed (its destructor synchronises the stream)
  queue q(s);
  auto committed = desc.commit(q);
  std::vector<std::complex<T>> host(n * batch), ref(n * batch), back(n * batch);
  for (std::size_t i = 0; i < host.size(); ++i) host[i] = {T(std::sin(0.37 * i)), T(std::cos(1.1 * i + 0.3))};
  for (std::size_t b = 0; b < batch; ++b)
    for (std::size_t k = 0; k < n; ++k) {
      std::complex<double> acc = 0;
      for (std::size_t j = 0; j < n; ++j)
        acc += std::complex<double>(host[b * n + j]) * std::polar(1.0, -2.0 * M_PI * double(j * k % n) / double(n));
      ref[b * n + k] = std::complex<T>(acc);
    }
  cudaMalloc(&dev, sizeof(std::complex<T>) * host.size());
  cudaMemcpyAsync(dev, host.data(), sizeof(std::complex<T>) * host.size(), cudaMemcpyHostToDevice, s);
  event e1 = committed.compute_forward(dev);
  e1.wait();
  cudaMemcpy(back.data(), dev, sizeof(std::complex<T>) * host.size(), cudaMemcpyDeviceToHost);
  for (std::size_t i = 0; i < host.size(); ++i) {
    err += std::norm(std::complex<double>(back[i]) - std::complex<double>(ref[i]));
    nrm += std::norm(std::complex<double>(ref[i]));
  }
  fwd_err = std::sqrt(err / nrm);
  event e2 = committed.compute_backward(dev, {e1});
  e2.wait();
  cudaMemcpy(back.data(), dev, sizeof(std::complex<T>) * host.size(), cudaMemcpyDeviceToHost);
  err = nrm = 0;
  for (std::size_t i = 0; i < host.size(); ++i) {
    err += std::norm(std::complex<double>(back[i]) - std::complex<double>(host[i]));
    nrm += std::norm(std::complex<double>(host[i]));
  }
  }
  cudaFree(dev);
  cudaStreamDestroy(s);
  return std::max(fwd_err, std::sqrt(err / nrm));
}

// REAL domain: real-to-complex forward against a naive DFT on the host, then complex-to-real backward (round trip)
template <typename T>
static double check_real(std::size_t n, std::size_t batch) {
  descriptor<T, domain::REAL> desc({n});
  const std::size_t h = n / 2 + 1;
  desc.number_of_transforms = batch;
  desc.backward_distance = h;
  desc.backward_scale = T(1) / T(n);
  queue q;
  auto committed = desc.commit(q);
  std::vector<T> host(n * batch), back(n * batch);
  std::vector<std::complex<T>> spec(h * batch);
  for (std::size_t i = 0; i < host.size(); ++i) host[i] = T(std::sin(0.37 * i) + 0.25 * std::cos(1.3 * i));
  T* din = nullptr;
  std::complex<T>* dspec = nullptr;
  cudaMalloc(&din, sizeof(T) * host.size());
  cudaMalloc(&dspec, sizeof(std::complex<T>) * spec.size());
  cudaMemcpy(din, host.data(), sizeof(T) * host.size(), cudaMemcpyHostToDevice);
  committed.compute_forward(static_cast<const T*>(din), dspec).wait();
  cudaMemcpy(spec.data(), dspec, sizeof(std::complex<T>) * spec.size(), cudaMemcpyDeviceToHost);
  double err = 0, nrm = 0;
  for (std::size_t b = 0; b < batch; ++b)
    for (std::size_t k = 0; k < h; ++k) {
      std::complex<double> acc = 0;
      for (std::size_t j = 0; j < n; ++j)
        acc += double(host[b * n + j]) * std::polar(1.0, -2.0 * M_PI * double(j * k % n) / double(n));
      err += std::norm(std::complex<double>(spec[b * h + k]) - acc);
      nrm += std::norm(acc);
    }
  const double fwd_err = std::sqrt(err / nrm);
  cudaMemset(din, 0, sizeof(T) * host.size());
  committed.compute_backward(static_cast<const std::complex<T>*>(dspec), din).wait();
  cudaMemcpy(back.data(), din, sizeof(T) * host.size(), cudaMemcpyDeviceToHost);
  err = nrm = 0;
  for (std::size_t i = 0; i < host.size(); ++i) {
    err += (double(back[i]) - double(host[i])) * (double(back[i]) - double(host[i]));
    nrm += double(host[i]) * double(host[i]);
  }
  cudaFree(din);
  cudaFree(dspec);
  return std::max(fwd_err, std::sqrt(err / nrm));
}

int main() {
  int fails = 0;
  for (std::size_t n : {16, 1000, 81}) {
    double ef = check_real<float>(n, 3), ed = check_real<double>(n, 3);
    double bf = 1e-5 * std::log2(double(n)), bd = 1e-13 * std::log2(double(n));
    std::printf("real n=%zu float relL2=%.2e (bound %.1e) double relL2=%.2e (bound %.1e)\n", n, ef, bf, ed, bd);
    if (!(ef <= bf) || !(ed <= bd)) ++fails;
  }
  for (std::size_t n : {8, 64, 1000, 4096}) {
    double ef = check<float>(n, 5), ed = check<double>(n, 5);
    double bf = 1e-5 * std::log2(double(n)), bd = 1e-13 * std::log2(double(n));
    std::printf("n=%zu float relL2=%.2e (bound %.1e) double relL2=%.2e (bound %.1e)\n", n, ef, bf, ed, bd);
    if (!(ef <= bf) || !(ed <= bd)) ++fails;
  }
  // error convention
  {
    descriptor<float, domain::COMPLEX> bad({8});
    bad.number_of_transforms = 2;
    bad.forward_distance = 7;  // InvalidShortDistance, instantiate_fft_tests.hpp:350-355
    queue q;
    try {
      bad.commit(q);
      std::printf("expected invalid_configuration\n");
      ++fails;
    } catch (const invalid_configuration&) {
    }
    descriptor<float, domain::COMPLEX> good({8});
    auto c = good.commit(q);
    float* re = nullptr;
    try {
      c.compute_forward(re, re, re, re);  // split call on an interleaved descriptor
      std::printf("expected invalid_configuration for storage mismatch\n");
      ++fails;
    } catch (const invalid_configuration&) {
    }
    descriptor<float, domain::COMPLEX> d23({2, 3});
    d23.number_of_transforms = 2;
    d23.forward_strides = {8, 3};
    d23.forward_distance = 15;
    d23.forward_offset = 3;
    d23.backward_strides = {2, 4};
    d23.backward_distance = 1;
    d23.backward_offset = 5;
    if (d23.get_input_count(direction::FORWARD) != 33 || d23.get_input_count(direction::BACKWARD) != 17 ||
        d23.get_flattened_length() != 6) {
      std::printf("buffer count KAT failed\n");
      ++fails;
    }
    if (detail::get_layout(good, direction::FORWARD) != detail::layout::PACKED) ++fails;
  }
  std::printf(fails ? "FAILED\n" : "OK\n");
  return fails;
}
