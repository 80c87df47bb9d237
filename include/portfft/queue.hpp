// `portfft::queue` / `portfft::event`: the CUDA stand-ins for sycl::queue / sycl::event in the reference API
// (descriptor::commit(sycl::queue&), compute_*(..., const std::vector<sycl::event>&) -> sycl::event).
#ifndef PFFT_B200_PORTFFT_QUEUE_HPP
#define PFFT_B200_PORTFFT_QUEUE_HPP

#include <cuda_runtime.h>

namespace portfft {

/// An in-order queue: a device ordinal plus a CUDA stream. Implicitly constructible from a cudaStream_t.
class queue {
  int device_ = 0;
  cudaStream_t stream_ = nullptr;

 public:
  queue() { cudaGetDevice(&device_); }
  queue(cudaStream_t s) : stream_(s) { cudaGetDevice(&device_); }  // NOLINT: implicit by design
  queue(int device, cudaStream_t s) : device_(device), stream_(s) {}
  int device() const { return device_; }
  cudaStream_t stream() const { return stream_; }
  void wait() const { cudaStreamSynchronize(stream_); }
};

/// Completion marker of one compute call: an event recorded on the queue's stream.
class event {
  cudaEvent_t ev_ = nullptr;
  cudaStream_t stream_ = nullptr;

 public:
  event() = default;
  event(cudaEvent_t e, cudaStream_t s) : ev_(e), stream_(s) {}
  void wait() const {
    if (ev_)
      cudaEventSynchronize(ev_);
    else
      cudaStreamSynchronize(stream_);
  }
  cudaEvent_t native() const { return ev_; }
};

}  // namespace portfft
#endif
