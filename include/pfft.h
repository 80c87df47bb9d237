/* pfft.h -- C ABI of the B200-native FFT library that stands in for portFFT's descriptor/commit/compute path.
 *
 * portFFT is header-only C++/SYCL, so it has no FFI of its own; these entry points are what a binding of its public
 * API binds to (all citations relative to the reference checkout):
 *
 *   pfft_desc              <- struct portfft::descriptor<Scalar, Domain>      src/portfft/descriptor.hpp:43-129
 *   pfft_validate          <- detail::validate::validate_descriptor           src/portfft/descriptor_validation.hpp:264-281
 *   pfft_commit            <- descriptor::commit(sycl::queue&)                src/portfft/descriptor.hpp:152-156
 *                             + committed_descriptor_impl ctor                src/portfft/committed_descriptor_impl.hpp:716-768
 *   pfft_compute           <- committed_descriptor::compute_forward/backward  src/portfft/committed_descriptor.hpp:171-310
 *                             (USM overloads; dispatch_direction              src/portfft/committed_descriptor_impl.hpp:852-878)
 *   pfft_destroy           <- ~committed_descriptor_impl                      src/portfft/committed_descriptor_impl.hpp:825-828
 *   pfft_get_buffer_count  <- descriptor::get_input_count / get_output_count  src/portfft/descriptor.hpp:172-183,262-270
 *   pfft_get_layout        <- detail::get_layout                              src/portfft/utils.hpp:237-246
 *   status codes           <- exception classes                               src/portfft/common/exceptions.hpp:32-77
 *
 * `sycl::queue` becomes a CUDA stream (passed as void* so that this header needs no CUDA include); USM pointers
 * become plain device pointers.  All strides / distances / offsets count complex elements for interleaved storage
 * and scalars-per-array for split storage, exactly as in the reference (committed_descriptor_impl.hpp:1105-1110).
 * The header-only C++ mirror of the reference API (include/portfft/portfft.hpp) is a thin layer over these calls.
 */
#ifndef PFFT_H_
#define PFFT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pfft_status {
  PFFT_OK = 0,
  PFFT_INVALID_CONFIGURATION = 1,     /* portfft::invalid_configuration     */
  PFFT_UNSUPPORTED_CONFIGURATION = 2, /* portfft::unsupported_configuration */
  PFFT_OUT_OF_LOCAL_MEMORY = 3,       /* portfft::out_of_local_memory_error */
  PFFT_INTERNAL_ERROR = 4,            /* portfft::internal_error            */
  PFFT_CUDA_ERROR = 5,
  PFFT_NCCL_ERROR = 6
} pfft_status;

/* enumerator order follows src/portfft/enums.hpp:26-39 */
enum { PFFT_DOMAIN_REAL = 0, PFFT_DOMAIN_COMPLEX = 1 };
enum { PFFT_INTERLEAVED_COMPLEX = 0, PFFT_SPLIT_COMPLEX = 1 };
enum { PFFT_IN_PLACE = 0, PFFT_OUT_OF_PLACE = 1 };
enum { PFFT_FORWARD = 0, PFFT_BACKWARD = 1 };
enum { PFFT_FLOAT = 0, PFFT_DOUBLE = 1 };
/* src/portfft/enums.hpp:44 and :46-69 */
enum { PFFT_LEVEL_WORKITEM = 0, PFFT_LEVEL_SUBGROUP = 1, PFFT_LEVEL_WORKGROUP = 2, PFFT_LEVEL_GLOBAL = 3 };
enum { PFFT_LAYOUT_PACKED = 0, PFFT_LAYOUT_UNPACKED = 1, PFFT_LAYOUT_BATCH_INTERLEAVED = 2 };

/* POD mirror of portfft::descriptor (src/portfft/descriptor.hpp:59-129). `lengths`, `forward_strides` and
 * `backward_strides` point at `rank` entries each (stride arrays may carry a different count through
 * `n_forward_strides` / `n_backward_strides` so that the "mismatching strides length" check can be exercised). */
typedef struct pfft_desc {
  int precision; /* PFFT_FLOAT / PFFT_DOUBLE  (template parameter Scalar) */
  int domain;    /* PFFT_DOMAIN_*             (template parameter Domain) */
  size_t rank;
  const size_t* lengths;
  double forward_scale;
  double backward_scale;
  size_t number_of_transforms;
  int complex_storage;
  int placement;
  size_t n_forward_strides;
  const size_t* forward_strides;
  size_t n_backward_strides;
  const size_t* backward_strides;
  size_t forward_distance;
  size_t backward_distance;
  size_t forward_offset;
  size_t backward_offset;
} pfft_desc;

typedef struct pfft_plan pfft_plan;

/* One additional independent batch dimension (FFTW-guru style "howmany" dimension) on top of number_of_transforms.
 * The reference has no counterpart (one batch dimension, one device: committed_descriptor_impl.hpp:109); these exist
 * for the multi-GPU slab decomposition, whose local passes read x-slabs and write blocks addressed by
 * (destination GPU, plane, row).  Distances count elements like forward_distance / backward_distance. */
typedef struct pfft_batch_dim {
  size_t count;
  size_t forward_distance;
  size_t backward_distance;
} pfft_batch_dim;

/* pfft_commit_guru flag: the LAST extra batch dimension does not address the output buffer by a distance but selects
 * one of several output buffers (pfft_compute_peer): entry i of the pointer table receives the transforms whose
 * index along that dimension is i.  With peer-mapped device memory the stores of the last pass ARE the exchange. */
enum { PFFT_GURU_PEER_LAST_DIM = 1 };

/* Host-only: the reference's commit-time validation. Never touches the GPU. */
pfft_status pfft_validate(const pfft_desc* desc);

/* Host-only helpers of the descriptor. `direction` selects the forward or backward domain. */
size_t pfft_get_flattened_length(const pfft_desc* desc);
size_t pfft_get_buffer_count(const pfft_desc* desc, int direction);
int pfft_get_layout(const pfft_desc* desc, int direction);

/* Host-only: run the planner without allocating device memory and describe the resulting passes (levels, radices,
 * launch geometry) as text. Returns the number of characters that the full description needs. */
pfft_status pfft_plan_describe(const pfft_desc* desc, int direction, char* buf, size_t buf_len, size_t* needed);

/* Host-only: the pass list of the plan as JSON (one object per kernel launch: transform length, element strides,
 * offsets, batch dimensions, buffers, inter-factor twiddle, element-wise modifiers, scale).  tests/plan_emulator.py
 * evaluates it with numpy to check the planner without a GPU.  Same buffer convention as pfft_plan_describe. */
pfft_status pfft_plan_export(const pfft_desc* desc, int direction, char* buf, size_t buf_len, size_t* needed);

/* Host-only: copy of a device-resident modifier table as the plan builds it (csrc/tables.h ModTable kinds: 1 chirp
 * exp(-i pi j^2 / L), 2 chirp / M, 3 FFT_M of the conjugate chirp) into `out` (interleaved complex of the given
 * precision; L entries for kinds 1 and 2, M entries for kind 3). */
pfft_status pfft_table_host(int precision, int kind, size_t transform_length, size_t convolution_length, void* out);

/* validate + plan + build device-resident twiddle tables and workspace on `device`. `stream` (cudaStream_t) is the
 * queue the plan is committed to; it is used for the one-off table uploads and as default stream of pfft_compute. */
pfft_status pfft_commit(const pfft_desc* desc, int device, void* stream, pfft_plan** plan_out);

/* Copy of a committed plan (copy constructor of committed_descriptor_impl,
 * src/portfft/committed_descriptor_impl.hpp:774-803): shares the immutable device tables (twiddles) with `plan` and
 * owns fresh workspaces, so that the two can compute concurrently on different streams. */
pfft_status pfft_clone(const pfft_plan* plan, pfft_plan** plan_out);

/* pfft_commit with `n_extra` additional batch dimensions (1-D descriptors only; at most 2). */
pfft_status pfft_commit_guru(const pfft_desc* desc, size_t n_extra, const pfft_batch_dim* extra, int flags, int device,
                             void* stream, pfft_plan** plan_out);

/* Asynchronous, stream-ordered transform on device pointers. Interleaved storage: `in` / `out` point at complex
 * arrays and the *_imag pointers must be NULL. Split storage: real and imaginary arrays. in == out is the in-place
 * call. A storage mismatch with the descriptor returns PFFT_INVALID_CONFIGURATION
 * (committed_descriptor_impl.hpp:862-871). `stream` may be NULL to use the commit stream. */
pfft_status pfft_compute(pfft_plan* plan, int direction, const void* in, const void* in_imag, void* out,
                         void* out_imag, void* stream);

/* pfft_compute for a plan committed with PFFT_GURU_PEER_LAST_DIM: `out[i]` (and `out_imag[i]` for split storage) is
 * the output buffer for index i of the last extra batch dimension; n_peers must equal that dimension's count.  The
 * buffers may live on other GPUs (peer access / symmetric memory). */
pfft_status pfft_compute_peer(pfft_plan* plan, int direction, const void* in, const void* in_imag, size_t n_peers,
                              void* const* out, void* const* out_imag, void* stream);

/* End-to-end convenience for callers holding HOST buffers: H2D of the input, pfft_compute, D2H of the output, one
 * stream synchronisation. Buffers are sized by pfft_get_buffer_count (elements) and staged through plan-owned
 * device buffers that persist between calls. */
pfft_status pfft_compute_host(pfft_plan* plan, int direction, const void* in, const void* in_imag, void* out,
                              void* out_imag);

pfft_status pfft_destroy(pfft_plan* plan);

/* ---------------------------------------------------------------------------------------------------------------
 * Multi-GPU (csrc/multi.cu).  The reference is single-device (one sycl::queue,
 * src/portfft/committed_descriptor_impl.hpp:109), so these have no reference counterpart beyond the descriptor
 * vocabulary.  Two shapes of the same thing:
 *   * batch sharding: the transforms of a descriptor are independent units (overlap is rejected at commit,
 *     src/portfft/descriptor_validation.hpp:162-204); GPU r of `world` transforms a contiguous batch range.  No
 *     data-path collective.
 *   * slab decomposition of ONE 3-D complex transform: three local passes and one exchange step that is fused into
 *     the stores of the middle pass (peer-mapped memory over NVLink) and closed by a device-side flag barrier.
 * Both work with one process per GPU (the caller moves 64-byte IPC handles between processes with whatever transport
 * it has) and with one process driving all GPUs.
 * --------------------------------------------------------------------------------------------------------------- */

/* Contiguous balanced split of n_items over `world` ranks: the first n_items % world ranks get one extra item. */
pfft_status pfft_partition(size_t n_items, int world, int rank, size_t* first, size_t* count);

typedef struct pfft_shard_info {
  size_t first;          /* first transform of the shard within the un-sharded batch */
  size_t count;          /* transforms in the shard (0: this rank has nothing to do; its plan transforms 1) */
  size_t forward_start;  /* where the shard begins inside the un-sharded forward-domain buffer, in elements */
  size_t backward_start; /* ... and inside the backward-domain buffer */
} pfft_shard_info;

/* pfft_commit of the rank-local shard of `desc`.  Batch-major layouts: the shard is the batch range starting at
 * first * distance (pass base pointers advanced by *_start, or rank-local buffers).  BATCH_INTERLEAVED layouts
 * (distance 1, stride == number_of_transforms): the shard is a column range, batch-interleaved with
 * stride == count in rank-local buffers. */
pfft_status pfft_commit_shard(const pfft_desc* desc, int world, int rank, int device, void* stream,
                              pfft_plan** plan_out, pfft_shard_info* info);

/* One process, n_dev GPUs, batch sharded: shard r lives on devices[r] (`streams` may be NULL: the library creates one
 * non-blocking stream per device; otherwise n_dev cudaStream_t). */
typedef struct pfft_multi pfft_multi;
pfft_status pfft_commit_multi(const pfft_desc* desc, int n_dev, const int* devices, void* const* streams,
                              pfft_multi** multi_out);
int pfft_multi_size(const pfft_multi* multi);
pfft_status pfft_multi_shard(const pfft_multi* multi, int r, pfft_shard_info* info, pfft_plan** plan);
/* Asynchronous: entry r of every pointer table is the rank-local buffer on devices[r] (the *_imag tables may be NULL
 * for interleaved storage). */
pfft_status pfft_multi_compute(pfft_multi* multi, int direction, const void* const* in, const void* const* in_imag,
                               void* const* out, void* const* out_imag);
/* Un-sharded HOST buffers (batch-major layouts): every GPU runs the pfft_compute_host pipeline on its batch range,
 * one host thread per GPU; returns when all have finished. */
pfft_status pfft_multi_compute_host(pfft_multi* multi, int direction, const void* in, const void* in_imag, void* out,
                                    void* out_imag);
pfft_status pfft_multi_sync(pfft_multi* multi);
pfft_status pfft_multi_destroy(pfft_multi* multi);

/* Slab-decomposed 3-D transform: `desc` = COMPLEX, INTERLEAVED_COMPLEX, rank 3, default strides, one transform;
 * lengths[0] and lengths[1] divisible by `world`.  Rank r holds the x-planes [r*XL, (r+1)*XL) of the input
 * ([XL][n1][n2], XL = n0 / world) and receives the y-rows [r*YB, (r+1)*YB) of the spectrum ([n0][YB][n2]).
 * Every rank owns an exchange window in device memory (receive buffer + arrival flags) that its peers map. */
typedef struct pfft_slab pfft_slab;
enum { PFFT_IPC_HANDLE_BYTES = 64 };
/* `stream`: the cudaStream_t every call of this rank is ordered on (NULL: the default stream, as in pfft_commit). */
pfft_status pfft_slab_commit(const pfft_desc* desc, int world, int rank, int device, void* stream,
                             pfft_slab** slab_out);
size_t pfft_slab_elems(const pfft_slab* slab); /* complex elements of every rank-local buffer: n0*n1*n2 / world */
/* Window plumbing.  Other process: pfft_slab_export here, move the 64 bytes, pfft_slab_import there.  Same process
 * (or memory that is already mapped, e.g. a symmetric-memory allocation): pfft_slab_window + pfft_slab_attach.
 * A rank must know the windows of ALL ranks (its own is attached at commit) before the first transform. */
pfft_status pfft_slab_window(pfft_slab* slab, void** base, size_t* bytes);
pfft_status pfft_slab_export(pfft_slab* slab, void* ipc_handle);
pfft_status pfft_slab_import(pfft_slab* slab, int peer_rank, const void* ipc_handle);
pfft_status pfft_slab_attach(pfft_slab* slab, int peer_rank, void* peer_window);
/* Replace the rank's own window by caller-owned device memory of at least the size pfft_slab_window reports (e.g. one
 * buffer of a symmetric-memory allocation that the peers already map); before the first transform only.  The caller
 * keeps ownership. */
pfft_status pfft_slab_use_window(pfft_slab* slab, void* base, size_t bytes);
/* One process, n_dev GPUs (entries of `devices` may repeat: several ranks on one GPU): commits every rank, enables
 * peer access and attaches all windows.  slabs_out receives n_dev objects.  `streams` NULL (or a NULL entry): the
 * library creates one non-blocking stream per rank -- ranks of one process must not share a stream, their barriers
 * wait for each other. */
pfft_status pfft_slab_commit_local(const pfft_desc* desc, int n_dev, const int* devices, void* const* streams,
                                   pfft_slab** slabs_out);
/* Exchange through a collective of the caller (e.g. an all-to-all over ncclSend / ncclRecv) instead of peer stores:
 * fn moves block d (block_bytes each) of `send` to rank d's `recv` at block index = the sender's rank, stream-ordered
 * on `stream`; returns 0 on success.  NULL restores the peer-store exchange. */
typedef int (*pfft_alltoall_fn)(void* user, const void* send, void* recv, size_t block_bytes, void* stream);
pfft_status pfft_slab_set_alltoall(pfft_slab* slab, pfft_alltoall_fn fn, void* user);
/* Asynchronous on the slab's stream; collective: every rank calls the same sequence.  Forward: x-slab in, *out_yslab =
 * the y-slab of the spectrum inside the rank's window (valid until the next call).  Backward: y-slab in (may be that
 * pointer), x-slab out, times backward_scale. */
pfft_status pfft_slab_forward(pfft_slab* slab, const void* in_xslab, void** out_yslab);
pfft_status pfft_slab_backward(pfft_slab* slab, const void* in_yslab, void* out_xslab);
/* Waits for the slab's stream; reports a barrier that timed out (a peer that never arrived) as PFFT_CUDA_ERROR. */
pfft_status pfft_slab_sync(pfft_slab* slab);
pfft_status pfft_slab_destroy(pfft_slab* slab);

/* Introspection. */
size_t pfft_workspace_bytes(const pfft_plan* plan);
/* Multi-pass plans whose per-transform workspace is small against the L2 run the batch in chunks of this many
 * transforms, so that the intermediate results stay L2 resident between the passes (0: the plan runs in one piece). */
size_t pfft_plan_chunk_transforms(const pfft_plan* plan);
int pfft_plan_level(const pfft_plan* plan, size_t dimension); /* PFFT_LEVEL_* of one dimension, -1 if out of range */
size_t pfft_plan_num_launches(const pfft_plan* plan, int direction); /* kernel launches per compute call */
unsigned long long pfft_total_launches(void);                        /* kernels launched by this process so far */

/* Thread-local message of the last non-OK status returned on this thread. */
const char* pfft_last_error(void);
const char* pfft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PFFT_H_ */
