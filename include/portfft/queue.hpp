// `portfft::queue` / `portfft::event`: the CUDA stand-ins for sycl::queue / sycl::event in the reference API
// (descriptor::commit(sycl::queue&), compute_*(..., const std::vector<sycl::event>&) -> sycl::event).
#ifndef PFFT_B200_PORTFFT_QUEUE_HPP
#define PFFT_B200_PORTFFT_QUEUE_HPP

#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

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

/// Completion marker of one compute call: an event recorded on the queue's stream.  The CUDA event is created per
/// call and owned (ref-counted) by the `event` object and its copies, like a sycl::event: it stays valid after further
/// compute calls and after the committed descriptor is gone.
namespace detail {
/// Free list of CUDA events per device, so that the event of every compute call costs a record, not a create/destroy.
class event_pool {
  std::mutex mu_;
  std::vector<std::pair<int, cudaEvent_t>> free_;

 public:
  static event_pool& instance() {
    static event_pool* p = new event_pool;  // never destroyed: events may be released during static destruction
    return *p;
  }
  cudaEvent_t get(int device) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      for (std::size_t i = free_.size(); i-- > 0;)
        if (free_[i].first == device) {
          cudaEvent_t e = free_[i].second;
          free_[i] = free_.back();
          free_.pop_back();
          return e;
        }
    }
    cudaEvent_t e = nullptr;
    int prev = device;
    cudaGetDevice(&prev);  // an event belongs to the device that is current when it is created
    if (prev != device) cudaSetDevice(device);
    const cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (prev != device) cudaSetDevice(prev);
    return rc == cudaSuccess ? e : nullptr;
  }
  void put(int device, cudaEvent_t e) {
    std::lock_guard<std::mutex> lock(mu_);
    if (free_.size() < 256)
      free_.emplace_back(device, e);
    else
      cudaEventDestroy(e);
  }
};
}  // namespace detail

class event {
  struct holder {
    cudaEvent_t ev = nullptr;
    int device = 0;
    ~holder() {
      if (ev) detail::event_pool::instance().put(device, ev);
    }
  };
  std::shared_ptr<holder> h_;
  cudaStream_t stream_ = nullptr;

 public:
  event() = default;
  static event record(cudaStream_t s, int device) {
    event e;
    e.stream_ = s;
    e.h_ = std::make_shared<holder>();
    e.h_->device = device;
    e.h_->ev = detail::event_pool::instance().get(device);
    if (e.h_->ev == nullptr || cudaEventRecord(e.h_->ev, s) != cudaSuccess)
      e.h_.reset();  // wait() falls back to synchronising the stream
    return e;
  }
  void wait() const {
    if (h_ && h_->ev)
      cudaEventSynchronize(h_->ev);
    else
      cudaStreamSynchronize(stream_);
  }
  cudaEvent_t native() const { return h_ ? h_->ev : nullptr; }
};

}  // namespace portfft
#endif
