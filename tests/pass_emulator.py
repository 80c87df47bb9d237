"""numpy emulation of ONE local pass of portfft_b200.distributed (a `PassGeom`): the address formula of pass.h /
pfft_commit_guru evaluated on the host.  Test infrastructure only (CPU gloo tests of the multi-GPU host logic)."""
from __future__ import annotations

import numpy as np


def emulate_pass_backward(pg, src: np.ndarray, dst: np.ndarray, scale: float = 1.0) -> None:
    """The same plan run in the BACKWARD direction: reads the backward-domain layout, writes the forward-domain layout,
    unnormalised inverse transform times `scale` (non-peer geometries only)."""
    assert not pg.peer_last
    dims = [(pg.number_of_transforms, pg.forward_distance, pg.backward_distance)] + [tuple(e) for e in pg.extra]
    grids = np.meshgrid(*[np.arange(d[0]) for d in dims], indexing="ij")
    in_base = sum(g * d[2] for g, d in zip(grids, dims))
    out_base = sum(g * d[1] for g, d in zip(grids, dims))
    j = np.arange(pg.length)
    data = src[in_base[..., None] + j * pg.backward_stride]
    data = (np.fft.ifft(data.astype(np.complex128), axis=-1) * pg.length * scale).astype(src.dtype)
    dst[out_base[..., None] + j * pg.forward_stride] = data


def emulate_pass(pg, src: np.ndarray, dsts, block_offset: int = 0) -> None:
    """out[sum_d b_d*bwd_d + k*bwd_stride] = FFT_j(in[sum_d b_d*fwd_d + j*fwd_stride]) for every batch multi-index.
    `dsts`: one flat array, or (peer_last) one flat array per index of the last extra dimension; `block_offset` is
    added to every output address of the peer buffers (where this rank's block lands)."""
    dims = [(pg.number_of_transforms, pg.forward_distance, pg.backward_distance)] + [tuple(e) for e in pg.extra]
    counts = [d[0] for d in dims]
    grids = np.meshgrid(*[np.arange(c) for c in counts], indexing="ij")
    in_base = sum(g * d[1] for g, d in zip(grids, dims))
    out_dims = dims[:-1] if pg.peer_last else dims
    out_base = sum(g * d[2] for g, d in zip(grids[:len(out_dims)], out_dims))
    j = np.arange(pg.length)
    data = src[in_base[..., None] + j * pg.forward_stride]
    data = np.fft.fft(data.astype(np.complex128), axis=-1).astype(src.dtype)
    out_idx = out_base[..., None] + j * pg.backward_stride
    if pg.peer_last:
        for i in range(counts[-1]):
            dsts[i][out_idx[..., i, :] + block_offset] = data[..., i, :]
    else:
        dst = dsts if isinstance(dsts, np.ndarray) else dsts[0]
        dst[out_idx] = data
