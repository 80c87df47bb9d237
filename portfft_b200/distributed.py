"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing.

The reference is single-device (one `sycl::queue`, /root/reference/src/portfft/committed_descriptor_impl.hpp:109; no
collective call sites anywhere), so nothing here has a reference counterpart except the descriptor vocabulary:

* **Batch sharding** (`partition`, `shard_descriptor`): batched transforms are independent units
  (overlap is rejected at commit, descriptor_validation.hpp:162-204), so rank r transforms the contiguous batch range
  `partition(number_of_transforms, world, r)` of the same descriptor on its own GPU.  No data-path collective.
* **Slab decomposition of one 3-D transform** (`slab_geometry`, `slab_fft3d`): rank r owns the x-planes
  [r*XL, (r+1)*XL).  Local pass 1 transforms along y, local pass 2 along z and writes its rows straight into the
  exchange layout (block for destination GPU d = rows with y in [d*YB, (d+1)*YB)), one exchange step moves block d to
  GPU d, local pass 3 transforms along x.  The result is left y-slab distributed: `out[kx, yl, kz]` on rank r is
  `X[kx, r*YB + yl, kz]`.
    - exchange "nccl": pass 2 writes a send buffer, `all_to_all_single` over NCCL/NVLink moves it;
    - exchange "peer": pass 2 stores directly into the destination GPUs' receive buffers (symmetric memory mapped
      through NVLink, `pfft_compute_peer`): the FFT kernel's stores ARE the all-to-all, tile by tile, and only a
      barrier follows.
  The three local passes are ordinary plans of the C ABI; pass 2 uses the guru batch dimensions of
  `pfft_commit_guru` (destination GPU, plane, row) so that no pack / unpack kernel exists.

The orchestration itself (passes, exchange, device-side flag barrier) is C++ behind the C ABI (`pfft_slab_*`,
`pfft_commit_shard`, `pfft_multi_*` in include/pfft.h, csrc/multi.cu); `slab_plan` / `slab_fft3d` / `commit_shard`
below are bindings.  `slab_geometry` and `shard_descriptor` restate the same geometry in Python so that the gloo
tests can check it on CPU (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple


def partition(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of `n_items` over `world_size` ranks -> (first item, item count) of `rank`.
    The first `n_items % world_size` ranks get one extra item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, rem = divmod(int(n_items), world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


@dataclass
class BatchShard:
    """Rank-local view of a batch-sharded descriptor: `desc` describes the local shard in rank-local buffers;
    `forward_start` / `backward_start` are where the shard begins inside the un-sharded buffers (elements), which is
    what a host-side scatter / gather needs."""
    desc: object
    first: int
    count: int
    forward_start: int
    backward_start: int


def shard_descriptor(desc, world_size: int, rank: int) -> BatchShard:
    """Batch-shard `desc` (any object with the reference's descriptor fields) for `rank`.

    PACKED / strided layouts (distance >= 1 per batch): the shard is the batch range at `first*distance`.
    BATCH_INTERLEAVED layouts (distance 1, stride == number_of_transforms): the shard is a column range; in
    rank-local buffers it becomes batch-interleaved with stride == local count."""
    import copy

    first, count = partition(desc.number_of_transforms, world_size, rank)
    local = copy.deepcopy(desc)
    local.number_of_transforms = max(count, 1)
    for dom in ("forward", "backward"):
        strides = getattr(desc, f"{dom}_strides")
        dist = getattr(desc, f"{dom}_distance")
        if len(desc.lengths) == 1 and dist == 1 and strides[-1] == desc.number_of_transforms and \
                desc.number_of_transforms > 1:
            setattr(local, f"{dom}_strides", [max(count, 1)])
    return BatchShard(local, first, count, first * desc.forward_distance, first * desc.backward_distance)


def commit_shard(desc, world_size: int, rank: int, device: int = 0, queue=None):
    """pfft_commit_shard: commit the rank-local shard of `desc` -> (committed_descriptor, BatchShard).  The shard
    geometry comes from the C library; `shard_descriptor` is its Python restatement (checked equal in the tests)."""
    import ctypes

    from . import _lib
    from .api import _check, _stream_handle, committed_descriptor

    c, keep = desc._c_desc()
    handle = ctypes.c_void_p()
    info = _lib.pfft_shard_info()
    _check(_lib.load().pfft_commit_shard(ctypes.byref(c), int(world_size), int(rank), int(device), _stream_handle(queue),
                                         ctypes.byref(handle), ctypes.byref(info)))
    sh = shard_descriptor(desc, world_size, rank)
    assert (sh.first, sh.count, sh.forward_start, sh.backward_start) == \
        (info.first, info.count, info.forward_start, info.backward_start)
    return committed_descriptor(sh.desc, handle, device), sh


# ---------------------------------------------------------------------------------------------------------------
# slab decomposition
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class PassGeom:
    """One local pass in the vocabulary of `pfft_desc` + `pfft_batch_dim` (include/pfft.h)."""
    name: str
    length: int
    number_of_transforms: int
    forward_stride: int
    backward_stride: int
    forward_distance: int
    backward_distance: int
    extra: List[Tuple[int, int, int]] = field(default_factory=list)   # (count, forward_distance, backward_distance)
    peer_last: bool = False
    in_place: bool = False
    src: str = ""
    dst: str = ""


@dataclass
class SlabGeom:
    lengths: Tuple[int, int, int]
    world_size: int
    rank: int
    xl: int          # local x-planes
    yb: int          # local y-rows after the exchange
    slab_elems: int  # elements of every rank-local buffer
    block_elems: int  # elements one rank sends to one rank
    passes: List[PassGeom] = field(default_factory=list)


def slab_geometry(lengths: Sequence[int], world_size: int, rank: int, peer: bool = False) -> SlabGeom:
    """Local passes of the slab-decomposed forward 3-D transform of `lengths` = (n0, n1, n2) on `rank`.

    buffers (rank-local, complex elements, all of size n0*n1*n2/world):
      in   [xl][y][z]          the rank's x-slab of the input (default strides)
      A    [xl][y][z]          after the y pass
      S    [d][xl][yl][z]      send layout after the z pass (exchange "nccl"); block d goes to rank d
      B    [x][yl][z]          receive layout: block from rank s lands at x = s*XL.. ; pass 3 runs in place on it
    With `peer` the z pass writes block d straight into rank d's B at element offset rank*block (no S)."""
    n0, n1, n2 = (int(v) for v in lengths)
    w = int(world_size)
    if n0 % w or n1 % w:
        raise ValueError(f"slab decomposition needs lengths[0] and lengths[1] divisible by the world size "
                         f"({n0}, {n1} vs {w})")
    xl, yb = n0 // w, n1 // w
    g = SlabGeom((n0, n1, n2), w, rank, xl, yb, xl * n1 * n2, xl * yb * n2)
    # pass 1: along y (element stride n2); batch = z (distance 1), extra = plane (distance n1*n2); in -> A
    g.passes.append(PassGeom("y", n1, n2, n2, n2, 1, 1, extra=[(xl, n1 * n2, n1 * n2)], src="in", dst="A"))
    # pass 2: along z (contiguous rows); batch = yl, extra = plane, destination GPU; A -> S (or peers' B)
    dest_bwd = 0 if peer else xl * yb * n2
    g.passes.append(PassGeom("z", n2, yb, 1, 1, n2, n2,
                             extra=[(xl, n1 * n2, yb * n2), (w, yb * n2, dest_bwd)], peer_last=peer,
                             src="A", dst="B@peer" if peer else "S"))
    # pass 3: along x (element stride yb*n2), batch-interleaved over (yl, z); in place on B
    g.passes.append(PassGeom("x", n0, yb * n2, yb * n2, yb * n2, 1, 1, in_place=True, src="B", dst="B"))
    return g


def _make_descriptor(pf, geom: PassGeom, scalar: str):
    d = pf.descriptor([geom.length], scalar)
    d.number_of_transforms = geom.number_of_transforms
    d.forward_strides, d.backward_strides = [geom.forward_stride], [geom.backward_stride]
    d.forward_distance, d.backward_distance = geom.forward_distance, geom.backward_distance
    d.placement = pf.placement.IN_PLACE if geom.in_place else pf.placement.OUT_OF_PLACE
    return d


def _tensor_view(ptr: int, shape, dtype, device):
    """Zero-copy torch view of device memory owned by the C library (CUDA array interface)."""
    import torch

    typestr = {torch.complex64: "<c8", torch.complex128: "<c16", torch.uint8: "|u1"}[dtype]

    class _Mem:
        __cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr,
                                    "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_Mem(), device=device)


class slab_plan:
    """Binding of `pfft_slab` (include/pfft.h, csrc/multi.cu): one rank's part of a slab-decomposed 3-D transform.
    All the orchestration -- the three local passes, the exchange fused into the z pass's stores, the device-side flag
    barrier -- lives behind the C ABI; this class only moves pointers."""

    def __init__(self, handle, lengths, world: int, rank: int, scalar: str, device):
        import torch

        self._handle = handle
        self.lengths = tuple(int(v) for v in lengths)
        self.world, self.rank, self.scalar = world, rank, scalar
        self.device = torch.device(device)
        self.cdtype = torch.complex128 if scalar == "double" else torch.complex64
        self.geom = slab_geometry(self.lengths, world, rank, peer=True)
        self._keep = []  # ctypes callbacks / symmetric-memory allocations that must outlive the plan

    @classmethod
    def commit(cls, desc, world: int, rank: int, device, queue=None) -> "slab_plan":
        import ctypes

        import torch

        from . import _lib
        from .api import _check, _stream_handle

        device = torch.device(device)
        c, keep = desc._c_desc()
        handle = ctypes.c_void_p()
        _check(_lib.load().pfft_slab_commit(ctypes.byref(c), int(world), int(rank), int(device.index or 0),
                                            _stream_handle(queue), ctypes.byref(handle)))
        return cls(handle, desc.lengths, world, rank, desc.scalar, device)

    @classmethod
    def commit_local(cls, desc, devices: Sequence[int], queues=None) -> List["slab_plan"]:
        """One process, len(devices) ranks (entries may repeat): pfft_slab_commit_local."""
        import ctypes

        from . import _lib
        from .api import _check, _stream_handle

        n = len(devices)
        c, keep = desc._c_desc()
        devs = (ctypes.c_int * n)(*[int(d) for d in devices])
        streams = (ctypes.c_void_p * n)(*[_stream_handle(q) for q in queues]) if queues is not None else None
        out = (ctypes.c_void_p * n)()
        _check(_lib.load().pfft_slab_commit_local(ctypes.byref(c), n, devs, streams, out))
        return [cls(ctypes.c_void_p(out[r]), desc.lengths, n, r, desc.scalar, f"cuda:{int(devices[r])}") for r in range(n)]

    # -- window plumbing ------------------------------------------------------------------------------------------
    def window(self) -> Tuple[int, int]:
        import ctypes

        from . import _lib
        from .api import _check

        base, size = ctypes.c_void_p(), ctypes.c_size_t()
        _check(_lib.load().pfft_slab_window(self._handle, ctypes.byref(base), ctypes.byref(size)))
        return int(base.value), int(size.value)

    def export_handle(self) -> bytes:
        import ctypes

        from . import _lib
        from .api import _check

        buf = ctypes.create_string_buffer(64)
        _check(_lib.load().pfft_slab_export(self._handle, buf))
        return buf.raw

    def import_handle(self, peer_rank: int, handle: bytes) -> None:
        import ctypes

        from . import _lib
        from .api import _check

        _check(_lib.load().pfft_slab_import(self._handle, int(peer_rank), ctypes.create_string_buffer(handle, 64)))

    def attach(self, peer_rank: int, window_ptr: int) -> None:
        from . import _lib
        from .api import _check

        _check(_lib.load().pfft_slab_attach(self._handle, int(peer_rank), int(window_ptr)))

    def use_window(self, ptr: int, nbytes: int) -> None:
        from . import _lib
        from .api import _check

        _check(_lib.load().pfft_slab_use_window(self._handle, int(ptr), int(nbytes)))

    def set_alltoall(self, fn) -> None:
        """`fn(send_ptr, recv_ptr, block_bytes, stream_handle)` replaces the peer-store exchange (None restores it)."""
        from . import _lib
        from .api import _check

        if fn is None:
            cb = _lib.ALLTOALL_FN()
        else:
            def _cb(user, send, recv, block_bytes, stream):
                try:
                    fn(int(send), int(recv), int(block_bytes), int(stream or 0))
                    return 0
                except Exception:  # reported through the C status (PFFT_NCCL_ERROR)
                    import traceback

                    traceback.print_exc()
                    return 1

            cb = _lib.ALLTOALL_FN(_cb)
        self._keep.append(cb)
        _check(_lib.load().pfft_slab_set_alltoall(self._handle, cb, None))

    # -- transforms -----------------------------------------------------------------------------------------------
    def forward(self, x_slab):
        import ctypes

        from . import _lib
        from .api import _check, _ptr

        g = self.geom
        assert x_slab.is_contiguous() and x_slab.numel() == g.slab_elems and x_slab.dtype == self.cdtype
        out = ctypes.c_void_p()
        _check(_lib.load().pfft_slab_forward(self._handle, _ptr(x_slab), ctypes.byref(out)))
        return _tensor_view(out.value, (g.lengths[0], g.yb, g.lengths[2]), self.cdtype, self.device)

    def backward(self, y_slab, out=None):
        import torch

        from . import _lib
        from .api import _check, _ptr

        g = self.geom
        assert y_slab.is_contiguous() and y_slab.numel() == g.slab_elems and y_slab.dtype == self.cdtype
        if out is None:
            out = torch.empty(g.xl, g.lengths[1], g.lengths[2], dtype=self.cdtype, device=self.device)
        _check(_lib.load().pfft_slab_backward(self._handle, _ptr(y_slab), _ptr(out)))
        return out

    def sync(self) -> None:
        from . import _lib
        from .api import _check

        _check(_lib.load().pfft_slab_sync(self._handle))

    def destroy(self) -> None:
        from . import _lib

        if getattr(self, "_handle", None):
            _lib.load().pfft_slab_destroy(self._handle)
            self._handle = None
        self._keep = []

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class slab_fft3d:
    """3-D C2C transform of `lengths`, slab-decomposed over the ranks of `group` (one process and one GPU per rank).

        plan = slab_fft3d([512, 512, 512], "float", exchange="peer")
        spectrum_yslab = plan.forward(x_slab)        # x_slab: complex tensor [XL, n1, n2] on this rank's GPU
        x_again = plan.backward(spectrum_yslab)      # times backward_scale

    A binding: the transform itself is `pfft_slab_*` (csrc/multi.cu); `torch.distributed` only carries the 64-byte IPC
    handles of the exchange windows at construction ("peer") or serves as the caller's collective ("nccl":
    `all_to_all_single` through pfft_slab_set_alltoall).  "peer" maps the windows with CUDA IPC; if that is not
    possible on this machine the windows are taken from a symmetric-memory allocation instead
    (`torch.distributed._symmetric_memory`, pfft_slab_use_window) -- both are peer memory over NVLink, and if neither
    can be set up the constructor raises: there is no silent fallback to another exchange."""

    def __init__(self, lengths: Sequence[int], scalar: str = "float", group=None, device=None, exchange: str = "nccl",
                 stream=None, backward_scale: float = 1.0):
        import torch
        import torch.distributed as dist

        import portfft_b200 as pf

        assert exchange in ("nccl", "peer")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.exchange = exchange
        self.scalar = scalar
        self.cdtype = torch.complex128 if scalar == "double" else torch.complex64
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        d = pf.descriptor([int(v) for v in lengths], scalar)
        d.backward_scale = backward_scale
        self.plan = slab_plan.commit(d, self.world, self.rank, self.device, self.stream)
        self.geom = self.plan.geom
        self.window_mapping = "local"
        if exchange == "peer" and self.world > 1:
            self._map_windows()
        elif exchange == "nccl":
            self.plan.set_alltoall(self._all_to_all)

    def _map_windows(self):
        import torch
        import torch.distributed as dist

        ok = 1
        try:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.plan.export_handle(), group=self.group)
            for r, h in enumerate(handles):
                if r != self.rank:
                    self.plan.import_handle(r, h)
        except Exception as exc:
            self._ipc_error = repr(exc)
            ok = 0
        flag = torch.tensor([ok], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.window_mapping = "cuda_ipc"
            return
        import torch.distributed._symmetric_memory as symm_mem

        _, nbytes = self.plan.window()
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        hdl = symm_mem.rendezvous(buf, self.group)
        self.plan._keep += [buf, hdl]
        self.plan.use_window(buf.data_ptr(), nbytes)
        for r in range(self.world):
            if r != self.rank:
                self.plan.attach(r, int(hdl.buffer_ptrs[r]))
        self.window_mapping = "symmetric_memory"
        dist.barrier(group=self.group)

    def _all_to_all(self, send: int, recv: int, block_bytes: int, stream: int):
        import torch
        import torch.distributed as dist

        n = self.world * block_bytes
        s = _tensor_view(send, (self.world, block_bytes), torch.uint8, self.device)
        r = _tensor_view(recv, (self.world, block_bytes), torch.uint8, self.device)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=self.device) if stream else self.stream):
            dist.all_to_all_single(r, s, group=self.group)
        del n

    def forward(self, x_slab):
        return self.plan.forward(x_slab)

    def backward(self, y_slab, out=None):
        return self.plan.backward(y_slab if y_slab.is_contiguous() else y_slab.contiguous(), out)

    def sync(self):
        self.plan.sync()

    def destroy(self):
        if self.plan is not None:
            self.plan.destroy()
            self.plan = None
