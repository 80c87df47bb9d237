"""CPU tests of the multi-GPU host logic (portfft_b200/distributed.py): batch partitioning, descriptor sharding and
the slab-decomposed 3-D transform, with numpy standing in for the local CUDA passes (tests/pass_emulator.py) and
`gloo` (world_size 2, 127.0.0.1) for the exchange.  The same geometry objects drive the CUDA plans on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import portfft_oracle as oracle  # noqa: E402
from pass_emulator import emulate_pass, emulate_pass_backward  # noqa: E402
from portfft_b200.distributed import partition, shard_descriptor, slab_geometry  # noqa: E402


def test_partition_is_contiguous_balanced_and_complete():
    for n in [0, 1, 7, 8, 65536, 100000]:
        for w in [1, 2, 3, 4, 8]:
            nxt, sizes = 0, []
            for r in range(w):
                first, count = partition(n, w, r)
                assert first == nxt
                nxt += count
                sizes.append(count)
            assert nxt == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition(4, 2, 2)


def _full_and_sharded(od, world):
    """transform the whole batch with the oracle, and shard by shard; both in the descriptor's own layouts"""
    host_in, host_ref = oracle.expected_io(od, oracle.FORWARD)
    out = np.full_like(host_ref, complex(oracle.PADDING_VALUE, oracle.PADDING_VALUE))
    for r in range(world):
        sh = shard_descriptor(od, world, r)
        if sh.count == 0:
            continue
        loc = sh.desc
        bi_f = oracle.get_layout(od, oracle.FORWARD) == oracle.BATCH_INTERLEAVED
        bi_b = oracle.get_layout(od, oracle.BACKWARD) == oracle.BATCH_INTERLEAVED
        n = od.lengths[0]
        # gather the shard's elements from the global input (what a host-side scatter does)
        b = np.arange(sh.first, sh.first + sh.count)
        j = np.arange(n)
        gin = od.forward_offset + b[:, None] * od.forward_distance + j[None, :] * od.forward_strides[0]
        x = host_in[gin]
        y = np.fft.fft(x.astype(np.complex128), axis=1).astype(host_ref.dtype) * od.forward_scale
        gout = od.backward_offset + b[:, None] * od.backward_distance + j[None, :] * od.backward_strides[0]
        out[gout] = y
        if sh.count > 1:  # the rank-local descriptor keeps the global layout class
            assert (oracle.get_layout(loc, oracle.FORWARD) == oracle.BATCH_INTERLEAVED) == bi_f
            assert (oracle.get_layout(loc, oracle.BACKWARD) == oracle.BATCH_INTERLEAVED) == bi_b
        assert loc.number_of_transforms == sh.count
        assert sh.forward_start == sh.first * od.forward_distance and sh.backward_start == sh.first * od.backward_distance
    return host_ref, out


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_batch_sharding_reassembles_the_unsharded_result(world):
    for kwargs in [dict(lengths=[64], number_of_transforms=37),
                   dict(lengths=[16], number_of_transforms=24, forward_strides=[24], forward_distance=1,
                        backward_strides=[24], backward_distance=1),
                   dict(lengths=[20], number_of_transforms=9, forward_strides=[2], forward_distance=50,
                        backward_strides=[1], backward_distance=20, forward_offset=3, backward_offset=1)]:
        od = oracle.OracleDescriptor(**kwargs)
        ref, got = _full_and_sharded(od, world)
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-4)


def _slab_single_process(lengths, world, peer):
    """all ranks in one process: exercises geometry + emulator without a process group"""
    rng = np.random.default_rng(5)
    n0, n1, n2 = lengths
    x = (rng.uniform(-1, 1, lengths) + 1j * rng.uniform(-1, 1, lengths)).astype(np.complex64)
    geoms = [slab_geometry(lengths, world, r, peer=peer) for r in range(world)]
    g0 = geoms[0]
    A = [np.zeros(g0.slab_elems, np.complex64) for _ in range(world)]
    S = [np.zeros(g0.slab_elems, np.complex64) for _ in range(world)]
    B = [np.zeros(g0.slab_elems, np.complex64) for _ in range(world)]
    for r, g in enumerate(geoms):
        slab = np.ascontiguousarray(x[r * g.xl:(r + 1) * g.xl]).reshape(-1)
        emulate_pass(g.passes[0], slab, A[r])
        if peer:
            emulate_pass(g.passes[1], A[r], B, block_offset=r * g.block_elems)
        else:
            emulate_pass(g.passes[1], A[r], S[r])
    if not peer:  # all-to-all: block d of rank s -> rank d, position s
        for s in range(world):
            for d in range(world):
                B[d][s * g0.block_elems:(s + 1) * g0.block_elems] = S[s][d * g0.block_elems:(d + 1) * g0.block_elems]
    ref = np.fft.fftn(x.astype(np.complex128))
    for r, g in enumerate(geoms):
        emulate_pass(g.passes[2], B[r], B[r])
        got = B[r].reshape(n0, g.yb, n2)
        np.testing.assert_allclose(got, ref[:, r * g.yb:(r + 1) * g.yb, :], rtol=0, atol=1e-3)
    # backward pipeline (slab_fft3d.backward): x pass, exchange back, z pass, y pass; round trip = identity
    gb = [slab_geometry(lengths, world, r, peer=False) for r in range(world)]
    for r in range(world):
        emulate_pass_backward(gb[r].passes[2], B[r], B[r])
    S2 = [np.zeros(g0.slab_elems, np.complex64) for _ in range(world)]
    for s in range(world):
        for d in range(world):
            S2[d][s * g0.block_elems:(s + 1) * g0.block_elems] = B[s][d * g0.block_elems:(d + 1) * g0.block_elems]
    for r in range(world):
        back_a = np.zeros(g0.slab_elems, np.complex64)
        out = np.zeros(g0.slab_elems, np.complex64)
        emulate_pass_backward(gb[r].passes[1], S2[r], back_a)
        emulate_pass_backward(gb[r].passes[0], back_a, out, scale=1.0 / (n0 * n1 * n2))
        np.testing.assert_allclose(out.reshape(gb[r].xl, n1, n2), x[r * gb[r].xl:(r + 1) * gb[r].xl], rtol=0, atol=1e-4)


@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("lengths,world", [((4, 4, 8), 1), ((4, 4, 8), 2), ((8, 12, 5), 4), ((16, 8, 6), 8)])
def test_slab_geometry_single_process(lengths, world, peer):
    _slab_single_process(lengths, world, peer)


def test_slab_geometry_rejects_indivisible_lengths():
    with pytest.raises(ValueError):
        slab_geometry((6, 8, 4), 4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, lengths, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n0, n1, n2 = lengths
        rng = np.random.default_rng(11)  # same seed everywhere: every rank can slice its own slab
        x = (rng.uniform(-1, 1, lengths) + 1j * rng.uniform(-1, 1, lengths)).astype(np.complex64)
        g = slab_geometry(lengths, world, rank, peer=False)
        slab = np.ascontiguousarray(x[rank * g.xl:(rank + 1) * g.xl]).reshape(-1)
        A = np.zeros(g.slab_elems, np.complex64)
        S = np.zeros(g.slab_elems, np.complex64)
        emulate_pass(g.passes[0], slab, A)
        emulate_pass(g.passes[1], A, S)
        send = torch.view_as_real(torch.from_numpy(S)).reshape(world, -1).contiguous()
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)  # the exchange step of slab_fft3d.forward (exchange="nccl")
        B = torch.view_as_complex(recv.reshape(-1, 2)).numpy().copy()
        emulate_pass(g.passes[2], B, B)
        ref = np.fft.fftn(x.astype(np.complex128))[:, rank * g.yb:(rank + 1) * g.yb, :]
        err = float(np.abs(B.reshape(n0, g.yb, n2) - ref).max())
        # batch sharding: each rank transforms its batch range, results are gathered (no data-path collective
        # inside the transform itself)
        first, count = partition(10, world, rank)
        xb = (rng.uniform(-1, 1, (10, 32)) + 0j).astype(np.complex64)
        mine = torch.from_numpy(np.fft.fft(xb[first:first + count].astype(np.complex128), axis=1).astype(np.complex64))
        parts = [torch.empty(partition(10, world, r)[1], 32, dtype=torch.complex64) for r in range(world)]
        dist.all_gather(parts, mine) if all(p.shape == parts[0].shape for p in parts) else None
        ok_batch = True
        if all(p.shape == parts[0].shape for p in parts):
            full = torch.cat(parts).numpy()
            ok_batch = bool(np.abs(full - np.fft.fft(xb.astype(np.complex128), axis=1)).max() < 1e-3)
        q.put((rank, err, ok_batch))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lengths", [(8, 6, 10), (16, 16, 16)])
def test_slab_exchange_over_gloo_world_size_2(lengths):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ok_batch in results:
        assert err < 1e-3, (rank, err)
        assert ok_batch
