"""Python mirror of the reference's public API over the C ABI.

`descriptor`, `committed_descriptor`, the enums and the exception classes carry the names, fields, argument
meaning and error behaviour of /root/reference/src/portfft/descriptor.hpp:43-271,
committed_descriptor.hpp:35-311, enums.hpp:26-39 and common/exceptions.hpp:32-77, so that the parity tests read like
the reference's own (test/unit_test/fft_test_utils.hpp).  `sycl::queue` becomes a CUDA stream (an int handle /
`torch.cuda.Stream`), USM pointers become device pointers (ints or torch CUDA tensors).
"""
from __future__ import annotations

import ctypes
import enum
from typing import List, Optional, Sequence

from . import _lib


class domain(enum.IntEnum):
    REAL = 0
    COMPLEX = 1


class complex_storage(enum.IntEnum):
    INTERLEAVED_COMPLEX = 0
    SPLIT_COMPLEX = 1


class placement(enum.IntEnum):
    IN_PLACE = 0
    OUT_OF_PLACE = 1


class direction(enum.IntEnum):
    FORWARD = 0
    BACKWARD = 1


def inv(d: direction) -> direction:
    """src/portfft/enums.hpp:36-38"""
    return direction.BACKWARD if d == direction.FORWARD else direction.FORWARD


class level(enum.IntEnum):
    WORKITEM = 0
    SUBGROUP = 1
    WORKGROUP = 2
    GLOBAL = 3


class layout(enum.IntEnum):
    PACKED = 0
    UNPACKED = 1
    BATCH_INTERLEAVED = 2


class base_error(RuntimeError):
    pass


class internal_error(base_error):
    pass


class invalid_configuration(base_error):
    pass


class unsupported_configuration(base_error):
    pass


class out_of_local_memory_error(unsupported_configuration):
    pass


class cuda_error(base_error):
    pass


_STATUS_TO_EXC = {1: invalid_configuration, 2: unsupported_configuration, 3: out_of_local_memory_error,
                  4: internal_error, 5: cuda_error, 6: cuda_error}


def _check(status: int) -> None:
    if status != 0:
        msg = _lib.load().pfft_last_error()
        raise _STATUS_TO_EXC.get(status, internal_error)(msg.decode() if msg else f"pfft status {status}")


def get_default_strides(lengths: Sequence[int]) -> List[int]:
    """src/portfft/utils.hpp:190-201"""
    strides = [0] * len(lengths)
    total = 1
    for i in range(len(lengths) - 1, -1, -1):
        strides[i] = total
        total *= lengths[i]
    return strides


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):  # numpy (host) array, for compute_host
        return x.ctypes.data
    raise TypeError(f"cannot take a pointer from {type(x)}")


def _stream_handle(queue) -> Optional[int]:
    if queue is None:
        return None
    if isinstance(queue, int):
        return queue
    if hasattr(queue, "cuda_stream"):
        return queue.cuda_stream
    raise TypeError(f"queue must be a CUDA stream handle or torch.cuda.Stream, got {type(queue)}")


class descriptor:
    """portfft::descriptor<Scalar, Domain> (src/portfft/descriptor.hpp:43-271): public mutable fields."""

    def __init__(self, lengths: Sequence[int], scalar: str = "float", dom: domain = domain.COMPLEX):
        assert scalar in ("float", "double")
        self.scalar = scalar
        self.domain = dom
        self.lengths = [int(x) for x in lengths]
        self.forward_scale = 1.0
        self.backward_scale = 1.0
        self.number_of_transforms = 1
        self.complex_storage = complex_storage.INTERLEAVED_COMPLEX
        self.placement = placement.OUT_OF_PLACE
        self.forward_strides = get_default_strides(self.lengths)
        self.backward_strides = list(self.forward_strides)
        total = self.get_flattened_length()
        self.forward_distance = total
        self.backward_distance = total
        self.forward_offset = 0
        self.backward_offset = 0

    # -- getters (descriptor.hpp:161-251) -------------------------------------------------------------------------
    def get_flattened_length(self) -> int:
        t = 1
        for n in self.lengths:
            t *= n
        return t

    def get_strides(self, d):
        return self.forward_strides if d == direction.FORWARD else self.backward_strides

    def get_distance(self, d):
        return self.forward_distance if d == direction.FORWARD else self.backward_distance

    def get_offset(self, d):
        return self.forward_offset if d == direction.FORWARD else self.backward_offset

    def get_scale(self, d):
        return self.forward_scale if d == direction.FORWARD else self.backward_scale

    def get_input_count(self, d) -> int:
        c, keep = self._c_desc()
        return int(_lib.load().pfft_get_buffer_count(ctypes.byref(c), int(d)))

    def get_output_count(self, d) -> int:
        return self.get_input_count(inv(d))

    def get_layout(self, d) -> layout:
        c, keep = self._c_desc()
        return layout(_lib.load().pfft_get_layout(ctypes.byref(c), int(d)))

    # -- C ABI plumbing -------------------------------------------------------------------------------------------
    def _c_desc(self):
        def arr(v):
            return (ctypes.c_size_t * max(1, len(v)))(*[int(x) for x in v])

        lengths, fs, bs = arr(self.lengths), arr(self.forward_strides), arr(self.backward_strides)
        c = _lib.pfft_desc()
        c.precision = 1 if self.scalar == "double" else 0
        c.domain = int(self.domain)
        c.rank = len(self.lengths)
        c.lengths = lengths
        c.forward_scale = float(self.forward_scale)
        c.backward_scale = float(self.backward_scale)
        c.number_of_transforms = int(self.number_of_transforms)
        c.complex_storage = int(self.complex_storage)
        c.placement = int(self.placement)
        c.n_forward_strides = len(self.forward_strides)
        c.forward_strides = fs
        c.n_backward_strides = len(self.backward_strides)
        c.backward_strides = bs
        c.forward_distance = int(self.forward_distance)
        c.backward_distance = int(self.backward_distance)
        c.forward_offset = int(self.forward_offset)
        c.backward_offset = int(self.backward_offset)
        return c, (lengths, fs, bs)

    def validate(self) -> None:
        """detail::validate::validate_descriptor (descriptor_validation.hpp:264-281); host only."""
        c, keep = self._c_desc()
        _check(_lib.load().pfft_validate(ctypes.byref(c)))

    def describe_plan(self, d=direction.FORWARD) -> str:
        """Planner dry run (host only): the passes `commit` would build."""
        c, keep = self._c_desc()
        needed = ctypes.c_size_t(0)
        _check(_lib.load().pfft_plan_describe(ctypes.byref(c), int(d), None, 0, ctypes.byref(needed)))
        buf = ctypes.create_string_buffer(needed.value)
        _check(_lib.load().pfft_plan_describe(ctypes.byref(c), int(d), buf, needed.value, None))
        return buf.value.decode()

    def export_plan(self, d=direction.FORWARD) -> dict:
        """Planner dry run (host only): the pass list as a dict (pfft_plan_export JSON)."""
        import json

        c, keep = self._c_desc()
        needed = ctypes.c_size_t(0)
        _check(_lib.load().pfft_plan_export(ctypes.byref(c), int(d), None, 0, ctypes.byref(needed)))
        buf = ctypes.create_string_buffer(needed.value)
        _check(_lib.load().pfft_plan_export(ctypes.byref(c), int(d), buf, needed.value, None))
        return json.loads(buf.value.decode())

    def commit(self, queue=None, device: int = 0, extra=None, peer_last: bool = False) -> "committed_descriptor":
        """descriptor::commit(sycl::queue&) (descriptor.hpp:152-156): validate, then build the plan on `device`.

        `extra` (no reference counterpart; used by portfft_b200.distributed): additional batch dimensions as
        (count, forward_distance, backward_distance) tuples -> pfft_commit_guru; `peer_last`: the last one selects an
        output buffer of compute_forward_peer."""
        c, keep = self._c_desc()
        handle = ctypes.c_void_p()
        if extra:
            dims = (_lib.pfft_batch_dim * len(extra))(*[_lib.pfft_batch_dim(int(a), int(b), int(cc))
                                                        for a, b, cc in extra])
            _check(_lib.load().pfft_commit_guru(ctypes.byref(c), len(extra), dims, 1 if peer_last else 0, int(device),
                                                _stream_handle(queue), ctypes.byref(handle)))
        else:
            _check(_lib.load().pfft_commit(ctypes.byref(c), int(device), _stream_handle(queue), ctypes.byref(handle)))
        return committed_descriptor(self, handle, device)


class committed_descriptor:
    """portfft::committed_descriptor (src/portfft/committed_descriptor.hpp:35-311), USM overloads.

    compute_forward(inout) / (in, out) for interleaved storage; compute_forward(inout_re, inout_im) /
    (in_re, in_im, out_re, out_im) for split storage -- the argument count selects the overload exactly as in C++.
    """

    def __init__(self, desc: descriptor, handle, device: int):
        import copy

        self.params = copy.deepcopy(desc)
        self._handle = handle
        self.device = device

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def destroy(self):
        if getattr(self, "_handle", None):
            _lib.load().pfft_destroy(self._handle)
            self._handle = None

    def copy(self) -> "committed_descriptor":
        """Copy constructor of the reference (committed_descriptor_impl.hpp:774-803): shares the device tables, owns
        fresh workspaces (pfft_clone), so the copy may compute concurrently on another stream."""
        handle = ctypes.c_void_p()
        _check(_lib.load().pfft_clone(self._handle, ctypes.byref(handle)))
        return committed_descriptor(self.params, handle, self.device)

    __copy__ = copy

    def _dispatch(self, d: direction, args, queue):
        n = len(args)
        split = self.params.complex_storage == complex_storage.SPLIT_COMPLEX
        if self.params.domain == domain.REAL:
            # REAL domain (committed_descriptor.hpp:201-206,273-278): the forward-domain side is one scalar array.
            # forward: (inout) | (in, out) | split (in, out_re, out_im); backward: (inout) | (in, out) | split (in_re, in_im, out)
            fwd = d == direction.FORWARD
            if n == 1 and not split:
                a = (args[0], None, args[0], None)
            elif n == 2 and not split:
                a = (args[0], None, args[1], None)
            elif n == 3 and split:
                a = (args[0], None, args[1], args[2]) if fwd else (args[0], args[1], args[2], None)
            else:
                raise TypeError("REAL-domain compute_* takes (inout), (in, out) or, for split storage, 3 data arguments")
        elif n == 1:      # in-place interleaved  (committed_descriptor.hpp:171-176)
            a = (args[0], None, args[0], None)
        elif n == 2:
            if split:
                a = (args[0], args[1], args[0], args[1])  # in-place split (:186-192)
            else:
                a = (args[0], None, args[1], None)        # out-of-place interleaved (:242-246)
        elif n == 4:    # out-of-place split (:248-254)
            a = (args[0], args[1], args[2], args[3])
        else:
            raise TypeError("compute_* takes 1, 2 or 4 data arguments")
        _check(_lib.load().pfft_compute(self._handle, int(d), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]),
                                        _stream_handle(queue)))

    def compute_forward(self, *args, queue=None):
        self._dispatch(direction.FORWARD, args, queue)

    def compute_backward(self, *args, queue=None):
        self._dispatch(direction.BACKWARD, args, queue)

    def compute_forward_peer(self, in_, outs, in_imag=None, outs_imag=None, queue=None):
        """pfft_compute_peer: `outs[i]` receives the transforms whose index along the last extra batch dimension is i
        (buffers may be peer-mapped memory of other GPUs)."""
        n = len(outs)
        tab = (ctypes.c_void_p * n)(*[_ptr(o) for o in outs])
        tab_im = (ctypes.c_void_p * n)(*[_ptr(o) for o in outs_imag]) if outs_imag is not None else None
        _check(_lib.load().pfft_compute_peer(self._handle, int(direction.FORWARD), _ptr(in_), _ptr(in_imag), n, tab,
                                             tab_im, _stream_handle(queue)))

    def compute_host(self, d: direction, in_, in_imag, out, out_imag):
        """End-to-end call on HOST (numpy) buffers: H2D + compute + D2H through pfft_compute_host."""
        _check(_lib.load().pfft_compute_host(self._handle, int(d), _ptr(in_), _ptr(in_imag), _ptr(out),
                                             _ptr(out_imag)))

    # introspection
    def get_level(self, dimension: int = 0) -> level:
        return level(_lib.load().pfft_plan_level(self._handle, dimension))

    def workspace_bytes(self) -> int:
        return int(_lib.load().pfft_workspace_bytes(self._handle))

    def l2_chunk(self) -> int:
        """Transforms per L2-resident chunk (0: the plan runs over the whole batch in one piece)."""
        return int(_lib.load().pfft_plan_chunk_transforms(self._handle))

    def num_launches(self, d=direction.FORWARD) -> int:
        return int(_lib.load().pfft_plan_num_launches(self._handle, int(d)))


def mod_table(scalar: str, kind: int, transform_length: int, convolution_length: int):
    """Host copy of a modifier table as the plan builds it (pfft_table_host; kinds: csrc/tables.h ModTable)."""
    import numpy as np

    n = convolution_length if kind == 3 else transform_length
    out = np.empty(n, dtype=np.complex128 if scalar == "double" else np.complex64)
    _check(_lib.load().pfft_table_host(1 if scalar == "double" else 0, int(kind), int(transform_length),
                                       int(convolution_length), out.ctypes.data))
    return out


def total_launches() -> int:
    return int(_lib.load().pfft_total_launches())
