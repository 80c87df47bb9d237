// WORKITEM level instantiations for float (see wi.cuh): one kernel per transform length.
#include "wi.cuh"

namespace pfft {

cudaError_t launch_wi_f32(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  switch (p.n) {
#define PFFT_WI(NN) \
  case NN:          \
    return launch_wi_n<NN, float>(p, il, swap, grid, stream);
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
    PFFT_WI(17)
    PFFT_WI(18)
    PFFT_WI(19)
    PFFT_WI(20)
    PFFT_WI(21)
    PFFT_WI(22)
    PFFT_WI(23)
    PFFT_WI(24)
    PFFT_WI(25)
    PFFT_WI(26)
    PFFT_WI(27)
    PFFT_WI(28)
    PFFT_WI(29)
    PFFT_WI(30)
    PFFT_WI(31)
    PFFT_WI(32)
#undef PFFT_WI
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pfft
