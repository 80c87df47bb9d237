// WORKITEM level instantiations for double (see wi.cuh): one kernel per transform length.
#include "wi.cuh"

namespace pfft {

cudaError_t launch_wi_f64(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  switch (p.n) {
#define PFFT_WI(NN) \
  case NN:          \
    return launch_wi_n<NN, double>(p, il, swap, grid, stream);
    PFFT_WI(1)
    PFFT_WI(2)
    PFFT_WI(3)
    PFFT_WI(4)
    PFFT_WI(5)
    PFFT_WI(6)
    PFFT_WI(7)
    PFFT_WI(8)
    PFFT_WI(9)
    PFFT_WI(10)
    PFFT_WI(11)
    PFFT_WI(12)
    PFFT_WI(13)
    PFFT_WI(14)
    PFFT_WI(15)
    PFFT_WI(16)
#undef PFFT_WI
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pfft
