// Host-side planner: descriptor validation, layout classification, level selection, factorisation into passes.
// Pure C++17 (no CUDA calls) so that it can be exercised on a machine without a GPU.
//
// Counterpart of the reference's commit-time logic:
//   /root/reference/src/portfft/descriptor_validation.hpp (whole file)
//   /root/reference/src/portfft/utils.hpp:94-132,190-246
//   /root/reference/src/portfft/committed_descriptor_impl.hpp:210-313 (prepare_implementation), :448-532
#pragma once
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pfft.h"
#include "pass.h"

namespace pfft {

// Error carrying a pfft_status; the C ABI maps it to status + thread-local message, the C++ header back to the
// reference's exception types (src/portfft/common/exceptions.hpp:32-77).
void set_last_error(const std::string& msg);  // thread-local message behind pfft_last_error (runtime.cu)

struct PlanError : std::runtime_error {
  pfft_status status;
  PlanError(pfft_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

struct DescHost {
  bool no_real_fuse = false;  // plan-internal: build the REAL-domain passes unfused (fallback for unaligned pointers)
  bool is_double = false;
  int domain = PFFT_DOMAIN_COMPLEX;
  std::vector<size_t> lengths;
  double forward_scale = 1.0, backward_scale = 1.0;
  size_t number_of_transforms = 1;
  int complex_storage = PFFT_INTERLEAVED_COMPLEX;
  int placement = PFFT_OUT_OF_PLACE;
  std::vector<size_t> forward_strides, backward_strides;
  size_t forward_distance = 1, backward_distance = 1;
  size_t forward_offset = 0, backward_offset = 0;
  // guru extension (pfft_commit_guru): extra batch dimensions; peer_last: the last one selects an output buffer
  struct ExtraDim {
    size_t count, forward_distance, backward_distance;
  };
  std::vector<ExtraDim> extra;
  bool peer_last = false;

  const std::vector<size_t>& strides(int dir) const { return dir == PFFT_FORWARD ? forward_strides : backward_strides; }
  size_t distance(int dir) const { return dir == PFFT_FORWARD ? forward_distance : backward_distance; }
  size_t offset(int dir) const { return dir == PFFT_FORWARD ? forward_offset : backward_offset; }
  double scale(int dir) const { return dir == PFFT_FORWARD ? forward_scale : backward_scale; }
  size_t flattened_length() const;
  size_t buffer_count(int dir) const;
  // lengths of the data in one domain: REAL descriptors hold lengths[last] / 2 + 1 complex elements along the last
  // dimension of the backward domain (the half spectrum, as numpy.fft.rfftn:
  // /root/reference/test/common/reference_data_wrangler.hpp:136-137,196), everything else equals `lengths`
  std::vector<size_t> domain_lengths(int dir) const;
  bool is_real() const { return domain == PFFT_DOMAIN_REAL; }
};

DescHost desc_from_c(const pfft_desc* d);
std::vector<size_t> default_strides(const std::vector<size_t>& lengths);
int get_layout(const DescHost& d, int dir);
void validate_descriptor(const DescHost& d);  // throws PlanError

enum BufSel : int { BUF_IN = 0, BUF_OUT = 1, BUF_SCRATCH = 2, BUF_SCRATCH2 = 3, BUF_SCRATCH3 = 4 };

enum KernelKind : int { KERNEL_WG_GENERIC = 0, KERNEL_WI = 1, KERNEL_SG = 2, KERNEL_WG_CUBE = 3, KERNEL_WG_COL = 4, KERNEL_WG_R3 = 5, KERNEL_EW = 6,
                        KERNEL_REAL_PACK = 7, KERNEL_R2C_POST = 8, KERNEL_C2R_PRE = 9, KERNEL_REAL_UNPACK = 10,
                        KERNEL_WG_COLG = 11 };

// One launch. `pp` holds everything except pointers / table addresses, which the runtime patches in.
struct PassHost {
  PassParams pp{};
  int kernel = KERNEL_WG_GENERIC;
  int level = LEVEL_WORKGROUP;
  int src = BUF_IN, dst = BUF_OUT;
  int grid = 1;
  int block = 1;
  size_t smem = 0;
  long long tw_n = 0;  // per-pass twiddle table w_n^k (0: none)
  int alt_grid = 0;    // launch geometry of the specialised kernel (the generic geometry stays valid as fallback)
  int variant = 0;
  // element-wise modifier tables (tables.h ModTable) for transform length mod_l / convolution length mod_m
  int lmod_kind = 0, smod_kind = 0;
  long long mod_l = 0, mod_m = 0;
  // REAL-domain plans: plan-internal data is always interleaved complex; `user_side` tells which side of the pass
  // (bit 0 input, bit 1 output) is user memory in the descriptor's complex storage.  `real_view`: the input (bit 0)
  // or output (bit 1) is the user's REAL buffer read / written as interleaved complex pairs (x[2j], x[2j+1]).
  int internal_storage = 0;  // bit 0 / bit 1: the input / output side is plan-internal, interleaved complex
  int real_view = 0;
  int force_swap = 0;        // REAL N-D backward: inverse complex pass on the (interleaved) workspace
  // REAL domain fused into the transform kernel (wg_cube.cu): 1 = the pass also does r2c_post (writes n + 1 outputs per
  // row), 2 = it also does c2r_pre (reads n + 1 inputs per row); tw2_n = length of the second twiddle table (2 n)
  int fuse_real = 0;
  long long tw2_n = 0;
};

// geometry limits of the thread- and warp-level kernels (wi.cuh, sg.cuh); sg_supports_m lives in sg_f32.cu
constexpr int kWiMaxNFloat = 32, kWiMaxNDouble = 16, kWiBlock = 128;
constexpr int kSgMaxMFloat = 32, kSgMaxMDouble = 16, kSgBlock = 256;
bool sg_supports_m(int m, bool is_double);
bool wi_tma_supported(int n, bool is_double);  // wi_tma.cu
bool wi_tma_real_supported(int n, bool is_double);  // wi_tma.cu: half lengths of the fused REAL forms
// column-tile kernel (wg_col.cu): supported lengths, shared memory and block size
bool col_supported(int n, bool is_double, int* n1, int* n2);
size_t col_smem_bytes(int n, bool is_double, bool ring);
int col_threads(int n, bool is_double);
int col_tile_columns(int n, bool is_double);
// N = R^3 kernel (wg_cube.cu): packed interleaved fp32 4096 (one transform per tile) and 512 (kCube512Tile per tile)
constexpr int kCube512Tile = 4;
bool cube_supported(int n, bool is_double, int* transforms_per_tile, int* ctas_per_sm);
// ... and the half lengths of REAL-domain transforms it takes (cube_supported's plus 256)
bool cube_real_supported(int n, bool is_double, int* transforms_per_tile, int* ctas_per_sm);
// generic in-place column-tile kernel (wg_colg.cu)
size_t colg_smem_bytes(int n, int columns, bool is_double);
// three-radix kernel (wg_r3.cu)
bool r3_supported(int n, bool is_double, int* threads_per_fft, int* pitch);
// wg_colr3.cu: column tiles with three compile-time radices (fixed tile geometry)
bool colr3_supported(int n, bool is_double, int* columns, int* threads_per_column, size_t* smem);

struct PlanHost {
  DescHost desc;
  std::vector<PassHost> passes[2];  // [direction]
  std::vector<int> dim_level;       // per dimension, PFFT_LEVEL_*
  size_t scratch_elems = 0;         // complex elements of plan-owned workspace
  size_t scratch2_elems = 0;        // second workspace (Bluestein with a multi-pass convolution length)
  size_t scratch3_elems = 0;        // REAL N-D backward: the half spectrum, packed
};

struct DeviceLimits {
  int num_sms = 148;
  size_t max_smem_per_block = 227 * 1024;
};

// radix sequence (largest first) for a block-level transform of length n; empty when n has a prime factor > 31
std::vector<int> choose_radices(size_t n);
size_t max_workgroup_length(bool is_double, const DeviceLimits& lim);
PlanHost build_plan(const DescHost& d, const DeviceLimits& lim);  // validates first; throws PlanError
std::string describe_plan(const PlanHost& plan, int direction);
std::string export_plan_json(const PlanHost& plan, int direction);

}  // namespace pfft
