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


class slab_fft3d:
    """Forward 3-D C2C transform of `lengths`, slab-decomposed over the ranks of `group` (one GPU per rank).

        plan = slab_fft3d([512, 512, 512], "float", exchange="peer")
        spectrum_yslab = plan.forward(x_slab)        # x_slab: complex tensor [XL, n1, n2] on this rank's GPU

    `exchange`: "nccl" (all_to_all_single) or "peer" (FFT stores into peer memory + barrier).  "peer" needs
    symmetric memory (`torch.distributed._symmetric_memory`, NVLink P2P); if it cannot be set up the constructor
    raises -- there is no silent fallback."""

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
        self.geom = slab_geometry(lengths, self.world, self.rank, peer=(exchange == "peer"))
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        g = self.geom
        self.A = torch.empty(g.slab_elems, dtype=self.cdtype, device=self.device)
        self._symm = None
        if exchange == "peer":
            import torch.distributed._symmetric_memory as symm_mem

            self.B = symm_mem.empty(g.slab_elems, dtype=self.cdtype, device=self.device)
            self._symm = symm_mem.rendezvous(self.B, self.group)
            esize = 16 if scalar == "double" else 8
            # block from this rank lands at x = rank*XL in every destination's B
            self.peer_ptrs = [int(p) + self.rank * g.block_elems * esize for p in self._symm.buffer_ptrs]
            self.S = None
        else:
            self.B = torch.empty(g.slab_elems, dtype=self.cdtype, device=self.device)
            self.S = torch.empty(g.slab_elems, dtype=self.cdtype, device=self.device)
        self.plans = []
        for pg in g.passes:
            d = _make_descriptor(pf, pg, scalar)
            if pg.name == "y":
                d.backward_scale = backward_scale  # the y pass runs last in the backward transform
            self.plans.append(d.commit(self.stream, self.device.index, extra=pg.extra, peer_last=pg.peer_last))
        self._pz_back = None  # peer mode: the backward z pass reads an ordinary exchange buffer (built on first use)

    def forward(self, x_slab):
        import torch
        import torch.distributed as dist

        g = self.geom
        assert x_slab.is_contiguous() and x_slab.numel() == g.slab_elems and x_slab.dtype == self.cdtype
        py, pz, px = self.plans
        with torch.cuda.stream(self.stream):
            py.compute_forward(x_slab, self.A, queue=self.stream)
            if self.exchange == "peer":
                # nobody may still be reading its B (previous call's x pass) when remote stores begin
                self._symm.barrier(channel=0)
                pz.compute_forward_peer(self.A, self.peer_ptrs, queue=self.stream)
                self._symm.barrier(channel=1)  # every block has landed everywhere
            else:
                pz.compute_forward(self.A, self.S, queue=self.stream)
                dist.all_to_all_single(torch.view_as_real(self.B).view(self.world, -1),
                                       torch.view_as_real(self.S).view(self.world, -1), group=self.group)
            px.compute_forward(self.B, queue=self.stream)
        return self.B.view(g.lengths[0], g.yb, g.lengths[2])

    def backward(self, y_slab, out=None):
        """Inverse of `forward`: y-slab of the spectrum [n0, YB, n2] in, x-slab [XL, n1, n2] out (times
        `backward_scale`).  Mirror image of the forward pipeline: x pass, one exchange step, z pass reading the
        exchange layout through the same guru batch dimensions, y pass."""
        import torch
        import torch.distributed as dist

        import portfft_b200 as pf

        g = self.geom
        assert y_slab.numel() == g.slab_elems and y_slab.dtype == self.cdtype
        py, pz, px = self.plans
        if out is None:
            out = torch.empty(g.xl, g.lengths[1], g.lengths[2], dtype=self.cdtype, device=self.device)
        if self.S is None:
            self.S = torch.empty(g.slab_elems, dtype=self.cdtype, device=self.device)
        if self.exchange == "peer" and self._pz_back is None:
            pg = slab_geometry(g.lengths, self.world, self.rank, peer=False).passes[1]
            self._pz_back = _make_descriptor(pf, pg, self.scalar).commit(self.stream, self.device.index, extra=pg.extra)
        with torch.cuda.stream(self.stream):
            if y_slab.data_ptr() != self.B.data_ptr():
                self.B.copy_(y_slab.reshape(-1))
            px.compute_backward(self.B, queue=self.stream)
            if self.exchange == "peer":
                self._symm.barrier(channel=0)  # every rank's x pass is complete
                for s in range(self.world):   # pull block `rank` of every peer's B over NVLink
                    src = self._symm.get_buffer(s, (g.slab_elems,), self.cdtype)
                    self.S[s * g.block_elems:(s + 1) * g.block_elems].copy_(
                        src[self.rank * g.block_elems:(self.rank + 1) * g.block_elems], non_blocking=True)
                self._symm.barrier(channel=1)  # nobody overwrites its B while peers still read it
                self._pz_back.compute_backward(self.S, self.A, queue=self.stream)
            else:
                dist.all_to_all_single(torch.view_as_real(self.S).view(self.world, -1),
                                       torch.view_as_real(self.B).view(self.world, -1), group=self.group)
                pz.compute_backward(self.S, self.A, queue=self.stream)
            py.compute_backward(self.A, out, queue=self.stream)
        return out

    def destroy(self):
        for p in self.plans:
            p.destroy()
        if self._pz_back is not None:
            self._pz_back.destroy()
            self._pz_back = None
        self.plans = []
