// C++ smoke / parity test of the drop-in header: reads like a reference unit test
// (/root/reference/test/unit_test/fft_test_utils.hpp:276-347) with sycl::queue -> portfft::queue.
// Checks a forward/backward round trip and a naive DFT on the host, in-place interleaved and out-of-place split,
// plus the error convention (invalid_configuration from commit, storage mismatch from compute).
#include <portfft/portfft.hpp>
#include <portfft/distributed.hpp>

#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

using namespace portfft;

template <typename C>
static void pfft_compute_raw(C& c, const std::complex<float>* in, std::complex<float>* out) {
  c.compute_forward(in, out);  // (returns its completion event, as the reference does; dropped here)
}

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

// Copies of a committed descriptor own their workspaces (committed_descriptor_impl.hpp:774-803): the original and a
// copy run a GLOBAL-level plan (N = 65536: two passes through the plan's scratch) concurrently on two streams, many
// times over; every result must equal the one computed alone.
static int check_copies() {
  const std::size_t n = 65536, batch = 6;
  descriptor<float, domain::COMPLEX> desc({n});
  desc.number_of_transforms = batch;
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  int bad = 0;
  std::complex<float>*in1, *in2, *out1, *out2, *ref1, *ref2;
  const std::size_t bytes = sizeof(std::complex<float>) * n * batch;
  for (auto pp : {&in1, &in2, &out1, &out2, &ref1, &ref2}) cudaMalloc(pp, bytes);
  std::vector<std::complex<float>> h1(n * batch), h2(n * batch), r(n * batch), o(n * batch);
  for (std::size_t i = 0; i < h1.size(); ++i) {
    h1[i] = {float(std::sin(0.11 * i)), float(std::cos(0.7 * i))};
    h2[i] = {float(std::cos(0.13 * i + 1.0)), float(std::sin(0.3 * i))};
  }
  cudaMemcpy(in1, h1.data(), bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(in2, h2.data(), bytes, cudaMemcpyHostToDevice);
  {
    queue q1(s1), q2(s2);
    auto a = desc.commit(q1);
    auto b = a;  // copy: shared twiddles, own scratch
    b.set_queue(q2);
    if (a.get_workspace_bytes() == 0 || b.get_workspace_bytes() != a.get_workspace_bytes()) ++bad;
    a.compute_forward(in1, ref1).wait();
    a.compute_forward(in2, ref2).wait();  // results computed alone, one at a time
    event last_a, last_b;
    for (int it = 0; it < 20; ++it) {
      last_a = a.compute_forward(in1, out1);
      last_b = b.compute_forward(in2, out2);
    }
    event first = last_a;  // events stay valid however many computes follow
    for (int it = 0; it < 40; ++it) a.compute_forward(in1, out1);
    first.wait();
    last_b.wait();
    q1.wait();
    for (int k = 0; k < 2; ++k) {
      cudaMemcpy(r.data(), k ? ref2 : ref1, bytes, cudaMemcpyDeviceToHost);
      cudaMemcpy(o.data(), k ? out2 : out1, bytes, cudaMemcpyDeviceToHost);
      for (std::size_t i = 0; i < r.size(); ++i)
        if (r[i] != o[i]) {
          ++bad;
          break;
        }
    }
    auto c = std::move(b);  // moved-from objects release nothing
    committed_descriptor<float, domain::COMPLEX> d2 = c;
    d2 = a;
  }
  for (auto pp : {in1, in2, out1, out2, ref1, ref2}) cudaFree(pp);
  cudaStreamDestroy(s1);
  cudaStreamDestroy(s2);
  std::printf("copies on two streams: %s\n", bad ? "MISMATCH" : "identical");
  return bad;
}

// BASELINE config C1 (N = 64, batch 1024, out of place): host cost of one compute call through the C++ header and
// the rate at which back-to-back calls retire on the device.
static void c1_latency() {
  descriptor<float, domain::COMPLEX> desc({64});
  desc.number_of_transforms = 1024;
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  std::complex<float>*in, *out;
  cudaMalloc(&in, 8 * 65536);
  cudaMalloc(&out, 8 * 65536);
  cudaMemset(in, 0, 8 * 65536);
  {
    queue q(s);
    auto c = desc.commit(q);
    for (int i = 0; i < 200; ++i) pfft_compute_raw(c, in, out);
    q.wait();
    const int iters = 5000;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) pfft_compute_raw(c, in, out);
    auto t1 = std::chrono::steady_clock::now();
    q.wait();
    auto t2 = std::chrono::steady_clock::now();
    const double submit = std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
    const double total = std::chrono::duration<double, std::micro>(t2 - t0).count() / iters;
    // one call at a time: submit, wait, repeat (latency of a lone transform as a caller sees it)
    auto t3 = std::chrono::steady_clock::now();
    for (int i = 0; i < 500; ++i) c.compute_forward(in, out).wait();
    auto t4 = std::chrono::steady_clock::now();
    const double lone = std::chrono::duration<double, std::micro>(t4 - t3).count() / 500;
    std::printf("C1 n=64 batch=1024: host submit %.2f us/call, back-to-back %.2f us/call, submit+wait %.2f us/call\n",
                submit, total, lone);
  }
  cudaFree(in);
  cudaFree(out);
  cudaStreamDestroy(s);
}

// Multi-GPU header (include/portfft/distributed.hpp) with every rank on the current device: a slab-decomposed 3-D
// transform and a batch-sharded host-buffer transform, each against the ordinary single-GPU plan.
static int check_distributed() {
  using cf = std::complex<float>;
  int bad = 0;
  int ngpu = 0;
  cudaGetDeviceCount(&ngpu);
  {
    const std::size_t n0 = 16, n1 = 8, n2 = 32, total = n0 * n1 * n2;
    const int world = 4;
    std::vector<int> devices;
    for (int r = 0; r < world; ++r) devices.push_back(ngpu >= world ? r : 0);  // real GPUs when the box has them
    descriptor<float, domain::COMPLEX> desc({n0, n1, n2});
    std::vector<cf> host(total), ref(total), got(total);
    for (std::size_t i = 0; i < total; ++i) host[i] = {float(std::sin(0.11 * i)), float(std::cos(0.7 * i + 1))};
    cf *din, *dref;
    cudaMalloc(&din, total * sizeof(cf));
    cudaMalloc(&dref, total * sizeof(cf));
    cudaMemcpy(din, host.data(), total * sizeof(cf), cudaMemcpyHostToDevice);
    {
      queue q;
      auto single = desc.commit(q);
      single.compute_forward(din, dref).wait();
    }
    cudaMemcpy(ref.data(), dref, total * sizeof(cf), cudaMemcpyDeviceToHost);
    distributed::slab_descriptor<float> slab(desc, devices);
    const std::size_t xl = n0 / world, yb = n1 / world;
    std::vector<const cf*> xs;
    std::vector<cf*> owned;
    for (int r = 0; r < world; ++r) {
      cudaSetDevice(devices[r]);
      cf* p;
      cudaMalloc(&p, xl * n1 * n2 * sizeof(cf));
      cudaMemcpy(p, host.data() + r * xl * n1 * n2, xl * n1 * n2 * sizeof(cf), cudaMemcpyHostToDevice);
      xs.push_back(p);
      owned.push_back(p);
    }
    cudaSetDevice(0);
    std::vector<cf*> ys = slab.compute_forward(xs);
    slab.wait();
    double err = 0, nrm = 0;
    for (int r = 0; r < world; ++r) {
      std::vector<cf> part(n0 * yb * n2);
      cudaMemcpy(part.data(), ys[r], part.size() * sizeof(cf), cudaMemcpyDefault);
      for (std::size_t x = 0; x < n0; ++x)
        for (std::size_t y = 0; y < yb; ++y)
          for (std::size_t z = 0; z < n2; ++z) {
            const std::complex<double> w = ref[(x * n1 + r * yb + y) * n2 + z], g = part[(x * yb + y) * n2 + z];
            err += std::norm(g - w);
            nrm += std::norm(w);
          }
    }
    const double rel = std::sqrt(err / nrm);
    std::printf("distributed slab %zux%zux%zu over %d ranks (%s): relL2 vs single-GPU plan %.2e\n", n0, n1, n2, world,
                ngpu >= world ? "one GPU each" : "one GPU", rel);
    if (!(rel < 1e-6)) ++bad;
    for (cf* p : owned) cudaFree(p);
    cudaFree(din);
    cudaFree(dref);
  }
  {
    const std::size_t n = 1000, batch = 77;
    descriptor<float, domain::COMPLEX> desc({n});
    desc.number_of_transforms = batch;
    std::vector<cf> host(n * batch), ref(n * batch), got(n * batch);
    for (std::size_t i = 0; i < host.size(); ++i) host[i] = {float(std::sin(0.3 * i)), float(std::cos(0.9 * i))};
    cf* d;
    cudaMalloc(&d, host.size() * sizeof(cf));
    cudaMemcpy(d, host.data(), host.size() * sizeof(cf), cudaMemcpyHostToDevice);
    {
      queue q;
      auto single = desc.commit(q);
      single.compute_forward(d).wait();
    }
    cudaMemcpy(ref.data(), d, host.size() * sizeof(cf), cudaMemcpyDeviceToHost);
    cudaFree(d);
    std::vector<int> devices;
    for (int r = 0; r < 3; ++r) devices.push_back(ngpu >= 3 ? r : 0);
    distributed::sharded_descriptor<float, domain::COMPLEX> sharded(desc, devices);
    sharded.compute_forward(host.data(), got.data());
    std::size_t covered = 0;
    for (int r = 0; r < sharded.size(); ++r) covered += sharded.get_shard(r).count;
    const bool same = std::equal(ref.begin(), ref.end(), got.begin());
    std::printf("distributed batch sharding %zu x %zu over 3 ranks: %s\n", n, batch, same ? "identical" : "MISMATCH");
    if (!same || covered != batch) ++bad;
  }
  return bad;
}

int main() {
  int fails = 0;
  fails += check_copies();
  fails += check_distributed();
  c1_latency();
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
