"""ctypes loader of the C-ABI shared library (include/pfft.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFFT_LIB") or os.path.join(_HERE, "lib", "libpfft_b200.so")  # PFFT_LIB: A/B builds


class pfft_desc(ctypes.Structure):
    """POD mirror of `pfft_desc` (include/pfft.h) <- portfft::descriptor (src/portfft/descriptor.hpp:59-129)."""

    _fields_ = [
        ("precision", c_int),
        ("domain", c_int),
        ("rank", c_size_t),
        ("lengths", POINTER(c_size_t)),
        ("forward_scale", c_double),
        ("backward_scale", c_double),
        ("number_of_transforms", c_size_t),
        ("complex_storage", c_int),
        ("placement", c_int),
        ("n_forward_strides", c_size_t),
        ("forward_strides", POINTER(c_size_t)),
        ("n_backward_strides", c_size_t),
        ("backward_strides", POINTER(c_size_t)),
        ("forward_distance", c_size_t),
        ("backward_distance", c_size_t),
        ("forward_offset", c_size_t),
        ("backward_offset", c_size_t),
    ]


class pfft_batch_dim(ctypes.Structure):
    """`pfft_batch_dim` (include/pfft.h): one extra batch dimension of pfft_commit_guru."""

    _fields_ = [("count", c_size_t), ("forward_distance", c_size_t), ("backward_distance", c_size_t)]


class pfft_shard_info(ctypes.Structure):
    """`pfft_shard_info` (include/pfft.h): where a rank's batch shard lies inside the un-sharded buffers."""

    _fields_ = [("first", c_size_t), ("count", c_size_t), ("forward_start", c_size_t), ("backward_start", c_size_t)]


# pfft_alltoall_fn: int fn(void* user, const void* send, void* recv, size_t block_bytes, void* stream)
ALLTOALL_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p)

# every symbol include/pfft.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pfft_validate": (c_int, [POINTER(pfft_desc)]),
    "pfft_get_flattened_length": (c_size_t, [POINTER(pfft_desc)]),
    "pfft_get_buffer_count": (c_size_t, [POINTER(pfft_desc), c_int]),
    "pfft_get_layout": (c_int, [POINTER(pfft_desc), c_int]),
    "pfft_plan_describe": (c_int, [POINTER(pfft_desc), c_int, c_char_p, c_size_t, POINTER(c_size_t)]),
    "pfft_plan_export": (c_int, [POINTER(pfft_desc), c_int, c_char_p, c_size_t, POINTER(c_size_t)]),
    "pfft_table_host": (c_int, [c_int, c_int, c_size_t, c_size_t, c_void_p]),
    "pfft_commit": (c_int, [POINTER(pfft_desc), c_int, c_void_p, POINTER(c_void_p)]),
    "pfft_clone": (c_int, [c_void_p, POINTER(c_void_p)]),
    "pfft_commit_guru": (c_int, [POINTER(pfft_desc), c_size_t, POINTER(pfft_batch_dim), c_int, c_int, c_void_p,
                                 POINTER(c_void_p)]),
    "pfft_compute_peer": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, POINTER(c_void_p), POINTER(c_void_p),
                                  c_void_p]),
    "pfft_compute": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pfft_compute_host": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pfft_destroy": (c_int, [c_void_p]),
    # multi-GPU (csrc/multi.cu)
    "pfft_partition": (c_int, [c_size_t, c_int, c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "pfft_commit_shard": (c_int, [POINTER(pfft_desc), c_int, c_int, c_int, c_void_p, POINTER(c_void_p),
                                  POINTER(pfft_shard_info)]),
    "pfft_commit_multi": (c_int, [POINTER(pfft_desc), c_int, POINTER(c_int), POINTER(c_void_p), POINTER(c_void_p)]),
    "pfft_multi_size": (c_int, [c_void_p]),
    "pfft_multi_shard": (c_int, [c_void_p, c_int, POINTER(pfft_shard_info), POINTER(c_void_p)]),
    "pfft_multi_compute": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                   POINTER(c_void_p)]),
    "pfft_multi_compute_host": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pfft_multi_sync": (c_int, [c_void_p]),
    "pfft_multi_destroy": (c_int, [c_void_p]),
    "pfft_slab_commit": (c_int, [POINTER(pfft_desc), c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "pfft_slab_elems": (c_size_t, [c_void_p]),
    "pfft_slab_window": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    "pfft_slab_export": (c_int, [c_void_p, c_void_p]),
    "pfft_slab_import": (c_int, [c_void_p, c_int, c_void_p]),
    "pfft_slab_attach": (c_int, [c_void_p, c_int, c_void_p]),
    "pfft_slab_use_window": (c_int, [c_void_p, c_void_p, c_size_t]),
    "pfft_slab_commit_local": (c_int, [POINTER(pfft_desc), c_int, POINTER(c_int), POINTER(c_void_p), POINTER(c_void_p)]),
    "pfft_slab_set_alltoall": (c_int, [c_void_p, ALLTOALL_FN, c_void_p]),
    "pfft_slab_forward": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "pfft_slab_backward": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pfft_slab_sync": (c_int, [c_void_p]),
    "pfft_slab_destroy": (c_int, [c_void_p]),
    "pfft_workspace_bytes": (c_size_t, [c_void_p]),
    "pfft_plan_chunk_transforms": (c_size_t, [c_void_p]),
    "pfft_plan_level": (c_int, [c_void_p, c_size_t]),
    "pfft_plan_num_launches": (c_size_t, [c_void_p, c_int]),
    "pfft_total_launches": (c_ulonglong, []),
    "pfft_last_error": (c_char_p, []),
    "pfft_version": (c_char_p, []),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libpfft_b200.so (built in-tree by `make` / `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "portfft_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
