// mbarrier / TMA / cp.async.bulk primitives and the padded-exchange index helpers shared by the column-tile kernels
// (wg_col.cu, wg_fused.cu).
#pragma once
#include <cuda.h>

#include <cstdint>

#include "device_utils.cuh"
#include "pass.h"

namespace pfft {

// host: 5-D TMA view (column, row j, b1, b2, b3) of a pass input whose fastest batch dimension is contiguous
// (wg_col.cu); false when the geometry or the pointer alignment cannot be encoded
bool col_make_tensor_map(const PassParams& p, bool is_double, int C, int box_rows, CUtensorMap* map);
// the same view over one scalar plane of split storage (scalars = 1) or over interleaved pairs (scalars = 2), `base` =
// address of element ioff; promo = L2 promotion (0 none, 1 64 B, 2 128 B, 3 256 B)
bool col_make_tensor_map_plane(const PassParams& p, const void* base, int scalars, bool is_double, int C, int box_rows,
                               int promo, CUtensorMap* map);

namespace col {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// TMA tile load: box {C columns, rows, 1, 1, 1} at coordinates (c0, r0, b1, b2, b3)
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int r0, int b1, int b2, int b3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
      "%6}], [%7];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(b1), "r"(b2), "r"(b3), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename T>
__host__ __device__ constexpr int pad(int i) {
  return i + (i >> (sizeof(T) == 4 ? 4 : 3));
}
template <typename T>
__host__ __device__ constexpr int pitch(int n) {
  return (pad<T>(n - 1) + 1) | 1;
}
constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

}  // namespace col

// inter-factor twiddle w_M^m, m = g * k < M <= 2^40, from the two-level table: one 32 x 32 -> 64 bit multiply
template <typename T>
__device__ __forceinline__ cx<T> gtw_lookup(const PassParams& p, unsigned g, unsigned k, unsigned long long gmask) {
  const unsigned long long m = (unsigned long long)g * k;
  return cmul(ldg_cx<T>(p.gtw_hi, (long long)(m >> p.gtw_bits)), ldg_cx<T>(p.gtw_lo, (long long)(m & gmask)));
}


}  // namespace pfft
