// SUBGROUP level instantiations for float (see sg.cuh): one kernel per points-per-lane count M.
#include "sg.cuh"

namespace pfft {

cudaError_t launch_sg_f32(const PassParams& p, bool il, bool swap, int grid, cudaStream_t stream) {
  switch (p.n / p.threads_per_fft) {
#define PFFT_SG(MM) \
  case MM:          \
    return launch_sg_m<MM, float>(p, il, swap, grid, stream);
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
    PFFT_SG(18)
    PFFT_SG(20)
    PFFT_SG(21)
    PFFT_SG(22)
    PFFT_SG(24)
    PFFT_SG(25)
    PFFT_SG(26)
    PFFT_SG(27)
    PFFT_SG(28)
    PFFT_SG(30)
    PFFT_SG(32)
#undef PFFT_SG
    default:
      return cudaErrorInvalidValue;
  }
}

bool sg_supports_m(int m, bool is_double) {
  if (m < 2) return false;
  if (m <= 16) return true;
  if (is_double) return false;
  switch (m) {
    case 18:
    case 20:
    case 21:
    case 22:
    case 24:
    case 25:
    case 26:
    case 27:
    case 28:
    case 30:
    case 32:
      return true;
    default:
      return false;
  }
}

}  // namespace pfft
