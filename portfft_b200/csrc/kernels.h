// Host-side launch entry points of the CUDA kernels (one per level) used by the planner.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "pass.h"
#include "plan.h"

namespace pfft {

inline size_t wg_generic_smem_bytes(int ffts_per_block, int pitch, size_t scalar_bytes) {
  return (size_t)2 * ffts_per_block * pitch * 2 * scalar_bytes + (size_t)4 * ffts_per_block * sizeof(long long);
}

// WORKGROUP level, generic (wg_generic.cu)
cudaError_t launch_wg_generic(const PassParams& p, bool is_double, bool interleaved, bool swap, int grid,
                              cudaStream_t stream);

// WORKITEM level (wi.cuh, wi_f32.cu, wi_f64.cu): one thread per transform, n <= kWiMaxN
inline size_t wi_smem_bytes(int n, size_t scalar_bytes) {
  return (size_t)kWiBlock * (n | 1) * 2 * scalar_bytes + (size_t)kWiBlock * 2 * sizeof(long long);
}
cudaError_t launch_wi_f32(const PassParams& p, bool interleaved, bool swap, int grid, cudaStream_t stream);
cudaError_t launch_wi_f64(const PassParams& p, bool interleaved, bool swap, int grid, cudaStream_t stream);

// WORKITEM level, TMA tiles in and out (wi_tma.cu): packed interleaved rows of exactly 128 bytes (fp32 N = 16, fp64 N = 8)
cudaError_t launch_wi_tma(const PassParams& p, bool is_double, bool swap, cudaStream_t stream, bool* used);
// ... with the REAL-domain pre / post-processing in registers (real: 1 real-to-complex, 2 complex-to-real): real rows of
// exactly one line, dense half-spectrum rows
cudaError_t launch_wi_tma_real(const PassParams& p, bool is_double, int real, cudaStream_t stream);

// SUBGROUP level (sg.cuh, sg_f32.cu, sg_f64.cu): n = lanes * m, `lanes` (power of two <= 32) threads per transform,
// m <= kSgMaxM points per lane, cross-lane stages through __shfl_xor_sync
inline size_t sg_smem_bytes(int m, size_t scalar_bytes) { return (size_t)kSgBlock * (m | 1) * 2 * scalar_bytes; }
cudaError_t launch_sg_f32(const PassParams& p, bool interleaved, bool swap, int grid, cudaStream_t stream);
cudaError_t launch_sg_f64(const PassParams& p, bool interleaved, bool swap, int grid, cudaStream_t stream);

// WORKGROUP level, N = R^3 specialisation with TMA-fed persistent CTAs (wg_cube.cu). variant 0: TMA ring, 1: direct loads
// real: 0 complex; 1 real-to-complex, 2 complex-to-real fused into the kernel (p.tw2 = w_{2n}^k; TMA variant, no swap)
cudaError_t launch_wg_cube(const PassParams& p, bool is_double, bool swap, int variant, int grid, cudaStream_t stream,
                           int real = 0);

// WORKGROUP level, tiles of 16 (fp32) / 8 (fp64) transforms fed by TMA (wg_col.cu): n in {64,128,256,512}.
// variant bits 0-1: input mode (0 strided columns via TMA tensor tiles, 1 contiguous rows loaded directly, 2 contiguous
// rows via cp.async.bulk), bit 2: contiguous output rows (else output columns).
// *used == false with cudaSuccess: tensor map not encodable for these pointers, run the generic kernel.
// `cache` (optional, owned by the plan, one per pass): the encoded tensor map of the last call, reused while the input
// base address is the same (the geometry of a pass never changes after commit) -- no cuTensorMapEncodeTiled per call.
struct ColMapCache {
  const void* base = nullptr;  // input address the cached map was encoded for (nullptr: empty)
  alignas(64) unsigned char map[128];
};
cudaError_t launch_wg_col(const PassParams& p, bool is_double, bool swap, int variant, int grid, cudaStream_t stream,
                          bool* used, ColMapCache* cache = nullptr);

// WORKGROUP level, three compile-time radix passes for any layout / storage (wg_r3.cu): n in {1000, 1024, 1536, 2048,
// 3072, 4096}; geometry = p.ffts_per_block transforms per CTA, r3_supported's threads per transform
cudaError_t launch_wg_r3(const PassParams& p, bool is_double, bool interleaved, bool swap, int grid, cudaStream_t stream);

// WORKGROUP level, column tiles of any 31-smooth length, in place in shared memory (wg_colg.cu): p.ffts_per_block =
// columns per tile, p.threads_per_fft = butterfly threads per column; needs ibd[0] == obd[0] == 1
size_t colg_smem_bytes(int n, int columns, bool is_double);
// ... and its specialisation with three compile-time radices (wg_colr3.cu; pass variant 1): fixed tile geometry
// `cache`: four ColMapCache entries of the pass (input planes 0 / 1, output planes 0 / 1), or nullptr
struct ColMapCache;
cudaError_t launch_wg_colr3(const PassParams& p, bool is_double, bool interleaved, bool swap, int grid, cudaStream_t stream,
                            ColMapCache* cache);
cudaError_t launch_wg_colg(const PassParams& p, bool is_double, bool interleaved, bool swap, int grid, cudaStream_t stream);

// GLOBAL level, two consecutive 256-point passes fused into one persistent kernel with the intermediate result in an
// L2-resident ring (wg_fused.cu).  FusedGeom: chunking decided at commit time; FusedArgs: per-launch arguments.
struct FusedGeom {
  int mode = 0;              // how pass b enumerates its rows (0: rows in batch dim 0, 1: rows in batch dim 1)
  int group = 0;             // chunk = `group` consecutive values of the chunk index (pass a's batch dimension 1)
  int lead = 1, slots = 3;   // pass a runs `lead` chunks ahead of pass b; ring slots
  long long num_chunks = 0, unit = 0, tiles_a = 0, tiles_b = 0;
  size_t ring_bytes = 0;
};
struct FusedArgs {
  void* ring;
  unsigned long long* done_a;  // per chunk: tiles of pass a completed (monotone over launches)
  unsigned long long* done_b;
  unsigned long long epoch;    // 1-based launch count of this plan: the counters reach epoch * tiles
  int group, lead, slots;
  long long num_chunks, unit, tiles_a, tiles_b;
};
int fused2_grid(bool is_double);  // persistent grid on the current device (0: the kernel cannot run)
bool fused2_plan(const PassParams& a, int variant_a, const PassParams& b, int variant_b, bool is_double, int grid,
                 FusedGeom* g);
cudaError_t launch_wg_fused2(const PassParams& a, const PassParams& b, const FusedArgs& fa, int mode, bool is_double,
                             bool swap_a, bool swap_b, cudaStream_t stream, bool* used);

// element-wise pass with modifiers (ew.cu): n == 1, modifier index = index along batch dimension 0
cudaError_t launch_ew(const PassParams& p, bool is_double, bool interleaved_in, bool interleaved_out, bool swap, int grid,
                      cudaStream_t stream);

// REAL domain pre / post passes (real.cu).  variant 0: even length (half-length complex transform), 1: odd length
cudaError_t launch_real_pack(const PassParams& p, bool is_double, int variant, int grid, cudaStream_t stream);
cudaError_t launch_real_unpack(const PassParams& p, bool is_double, int variant, int grid, cudaStream_t stream);
cudaError_t launch_r2c_post(const PassParams& p, bool is_double, bool interleaved, int variant, int grid, cudaStream_t stream);
cudaError_t launch_c2r_pre(const PassParams& p, bool is_double, bool interleaved, int variant, int grid, cudaStream_t stream);

}  // namespace pfft
