// One kernel launch of a plan ("pass"): a batched, strided 1-D DFT with optional inter-factor twiddle on store.
//
//   for every batch multi-index b = (b0, b1, b2, b3), 0 <= b_d < nb[d]:
//     out[ooff + sum_d b_d*obd[d] + k*os] = scale * gtw(b, k) * sum_j in[ioff + sum_d b_d*ibd[d] + j*is] * w_n^{jk}
//
// Every algorithm of the library is a sequence of such passes: a plain batched 1-D transform is one pass; an N-D
// transform is one pass per dimension (the reference instead loops over `batch*outer` separate launches,
// /root/reference/src/portfft/committed_descriptor_impl.hpp:899-950); the GLOBAL level (four-step) is one pass per
// factor with gtw = w_{gtw_n}^{b[gtw_dim]*k} fused on store and the transposition fused into (os, obd)
// (the reference: one launch per factor per batch plus a chain of transpose kernels,
// /root/reference/src/portfft/dispatcher/global_dispatcher.hpp:312-412).
#pragma once
#include <cstdint>

namespace pfft {

constexpr int kMaxBatchDims = 4;
constexpr int kMaxRadices = 12;
constexpr int kMaxPeers = 16;

enum IoMode : int {
  IO_DIRECT = 0,       // butterfly-owning threads access global memory directly (coalesced when stride == 1)
  IO_STAGED_ELEM = 1,  // cooperative tile copy through shared memory, lanes run along the element index
  IO_STAGED_BATCH = 2  // cooperative tile copy through shared memory, lanes run along the batch index
};

enum Level : int { LEVEL_WORKITEM = 0, LEVEL_SUBGROUP = 1, LEVEL_WORKGROUP = 2, LEVEL_GLOBAL = 3 };

struct PassParams {
  // data pointers (scalar arrays). Interleaved storage: *_im == nullptr and *_re points at (re,im) pairs.
  const void* in_re;
  const void* in_im;
  void* out_re;
  void* out_im;
  // transform
  int n;
  int num_radices;
  int radix[kMaxRadices];
  int threads_per_fft;  // T
  int ffts_per_block;   // F
  int pitch;            // shared-memory pitch (complex elements) of one transform
  // batch geometry (in complex elements)
  long long batch_total;
  long long nb[kMaxBatchDims];
  long long ibd[kMaxBatchDims];
  long long obd[kMaxBatchDims];
  long long is, os;
  long long ioff, ooff;
  int in_mode, out_mode;
  // twiddles w_n^k, k in [0, n), complex<T>, device resident
  const void* tw;
  // REAL-domain pre / post-processing fused into a transform kernel: w_{2n}^k, k in [0, 2n) (nullptr otherwise)
  const void* tw2;
  // inter-factor twiddle on store (GLOBAL level): w_{gtw_n}^{b[gtw_dim] * k} = hi[m >> gtw_bits] * lo[m & mask]
  int gtw_dim;  // -1: none
  int gtw_bits;
  long long gtw_n;
  const void* gtw_lo;
  const void* gtw_hi;
  // peer output (multi-GPU slab exchange fused into the store): the index along batch dimension `peer_dim` selects
  // the output base pointers from out_tab_* (peer-mapped device memory of the destination GPU) instead of out_re/im
  int peer_dim;  // -1: none
  void* out_tab_re[kMaxPeers];
  void* out_tab_im[kMaxPeers];
  // scale applied on store (only when apply_scale != 0)
  double scale;
  int apply_scale;
  // element-wise modifiers (Bluestein chirp / convolution kernel; wg_generic.cu and ew.cu only, zero elsewhere):
  //   load : x_j  <- (j < valid_in  ? in[j] * lmod[j] : 0)
  //   store: y_k  -> swap_post(swap_pre(y_k) * smod[k]), written only for k < valid_out
  // lmod / smod == nullptr: no multiply; valid_* == 0: every element.  In the element-wise kernel (ew.cu, n == 1) the
  // index j = k is the index along batch dimension 0.
  int valid_in, valid_out;
  const void* lmod;
  const void* smod;
  int mod_flags;  // ModFlags
  // smod_mask != 0 (wg_col.cu only): the store modifier is indexed by the LINEAR position of the output element inside
  // its packed power-of-two row, (output offset) & smod_mask, instead of by the pass-local index k -- the last pass of a
  // multi-pass transform multiplies by a table over the whole transform (Bluestein: the transformed chirp)
  long long smod_mask;
  // smod_n1 != 0 (wg_col.cu only; last pass of a two-factor transform writing the user's layout): the store modifier is
  // indexed by the linear output index (index along batch dimension 0) + smod_n1 * k, and only the elements whose
  // linear index is below valid_out are written (Bluestein: chirp / M and truncation to the transform length)
  int smod_n1;
};

enum ModFlags : int {
  MOD_SWAP_PRE = 1,          // (re <-> im) before the smod multiply
  MOD_SWAP_POST = 2,         // (re <-> im) after the smod multiply
  MOD_NO_USER_SWAP_IN = 4,   // input is plan-internal data: the backward (re <-> im) swap does not apply on load
  MOD_NO_USER_SWAP_OUT = 8   // output is plan-internal data: no backward swap on store
};

}  // namespace pfft
