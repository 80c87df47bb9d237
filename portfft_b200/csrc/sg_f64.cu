// SUBGROUP level instantiations for double (see sg.cuh): one kernel per points-per-lane count M.
#include "sg.cuh"

namespace pfft {

cudaError_t launch_sg_f64(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  switch (p.n / p.threads_per_fft) {
#define PFFT_SG(MM) \
  case MM:          \
    return launch_sg_m<MM, double>(p, il, swap, grid, stream);
    PFFT_SG(2)
    PFFT_SG(3)
    PFFT_SG(4)
    PFFT_SG(5)
    PFFT_SG(6)
    PFFT_SG(7)
    PFFT_SG(8)
    PFFT_SG(9)
    PFFT_SG(10)
    PFFT_SG(11)
    PFFT_SG(12)
    PFFT_SG(13)
    PFFT_SG(14)
    PFFT_SG(15)
    PFFT_SG(16)
#undef PFFT_SG
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pfft
