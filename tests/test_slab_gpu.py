"""GPU tests of the slab-decomposed 3-D transform (portfft_b200/distributed.py) on ONE GPU: the W ranks are played in
turn by the same device ("virtual ranks"), each with its own buffers.  The local passes are the real CUDA plans
(pfft_commit_guru: extra batch dimensions; pfft_compute_peer: the z pass stores straight into the destination ranks'
receive buffers), the exchange is either a block copy (what all_to_all_single does) or nothing at all (peer stores).
Checked against numpy.fft.fftn, the oracle the reference's own tests use."""
import numpy as np
import pytest

import portfft_b200 as pf
from portfft_b200.distributed import _make_descriptor, slab_geometry

pytestmark = pytest.mark.gpu


def _run(lengths, world, peer, scalar):
    import torch

    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    cdt = torch.complex128 if scalar == "double" else torch.complex64
    n0, n1, n2 = lengths
    rng = np.random.Generator(np.random.SFC64(0))
    x = (rng.uniform(-1, 1, lengths) + 1j * rng.uniform(-1, 1, lengths)).astype(
        np.complex128 if scalar == "double" else np.complex64)
    geoms = [slab_geometry(lengths, world, r, peer=peer) for r in range(world)]
    g0 = geoms[0]
    A = [torch.zeros(g0.slab_elems, dtype=cdt, device=dev) for _ in range(world)]
    S = [torch.zeros(g0.slab_elems, dtype=cdt, device=dev) for _ in range(world)]
    B = [torch.zeros(g0.slab_elems, dtype=cdt, device=dev) for _ in range(world)]
    esize = 16 if scalar == "double" else 8
    plans = []
    for r, g in enumerate(geoms):
        ps = [_make_descriptor(pf, pg, scalar).commit(stream, 0, extra=pg.extra, peer_last=pg.peer_last)
              for pg in g.passes]
        plans.append(ps)
        slab = torch.from_numpy(np.ascontiguousarray(x[r * g.xl:(r + 1) * g.xl]).reshape(-1)).to(dev)
        ps[0].compute_forward(slab, A[r])
        if peer:
            ptrs = [B[d].data_ptr() + r * g.block_elems * esize for d in range(world)]
            ps[1].compute_forward_peer(A[r], ptrs)
        else:
            ps[1].compute_forward(A[r], S[r])
    if not peer:
        for s in range(world):
            for d in range(world):
                B[d][s * g0.block_elems:(s + 1) * g0.block_elems] = S[s][d * g0.block_elems:(d + 1) * g0.block_elems]
    ref = np.fft.fftn(x.astype(np.complex128))
    nflat = n0 * n1 * n2
    bound = (1e-13 if scalar == "double" else 1e-5) * np.log2(nflat)
    for r, g in enumerate(geoms):
        plans[r][2].compute_forward(B[r])
        torch.cuda.synchronize(dev)
        got = B[r].cpu().numpy().reshape(n0, g.yb, n2)
        want = ref[:, r * g.yb:(r + 1) * g.yb, :]
        rel = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert rel <= bound, (r, rel, bound)
    # backward pipeline of slab_fft3d.backward with the same plans run in the BACKWARD direction (non-peer z plan)
    if not peer:
        for r in range(world):
            plans[r][2].compute_backward(B[r])
        for s in range(world):
            for d in range(world):
                S[d][s * g0.block_elems:(s + 1) * g0.block_elems] = B[s][d * g0.block_elems:(d + 1) * g0.block_elems]
        for r, g in enumerate(geoms):
            out = torch.zeros(g0.slab_elems, dtype=cdt, device=dev)
            plans[r][1].compute_backward(S[r], A[r])
            plans[r][0].compute_backward(A[r], out)
            torch.cuda.synchronize(dev)
            got = out.cpu().numpy().reshape(g.xl, n1, n2) / nflat
            want = x[r * g.xl:(r + 1) * g.xl]
            rel = np.linalg.norm(got - want) / np.linalg.norm(want)
            assert rel <= 2 * bound, (r, rel, bound)
    for ps in plans:
        for p in ps:
            p.destroy()


@pytest.mark.parametrize("scalar", ["float", "double"])
@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("lengths,world", [((4, 4, 8), 1), ((8, 12, 5), 4), ((16, 8, 64), 8), ((64, 64, 512), 8),
                                            ((32, 16, 1000), 2), ((8, 8, 4096), 2)])
def test_slab_virtual_ranks(lengths, world, peer, scalar):
    _run(lengths, world, peer, scalar)


def test_peer_plan_rejects_plain_compute():
    import torch

    g = slab_geometry((8, 8, 64), 2, 0, peer=True)
    pg = g.passes[1]
    plan = _make_descriptor(pf, pg, "float").commit(torch.cuda.current_stream(), 0, extra=pg.extra, peer_last=True)
    a = torch.zeros(g.slab_elems, dtype=torch.complex64, device="cuda")
    with pytest.raises(pf.invalid_configuration):
        plan.compute_forward(a, a.clone())
    plan.destroy()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("kind", ["packed", "batch_interleaved", "strided"])
def test_batch_sharding_virtual_ranks(world, kind):
    """Batch sharding (portfft_b200.distributed.shard_descriptor) with the real CUDA plans: every virtual rank commits
    its shard descriptor and transforms its slice; the reassembled result equals the un-sharded transform."""
    import torch

    import portfft_oracle as oracle
    from portfft_b200.distributed import shard_descriptor

    n, batch = 256, 37
    d = pf.descriptor([n])
    d.number_of_transforms = batch
    if kind == "batch_interleaved":
        d.forward_strides = d.backward_strides = [batch]
        d.forward_distance = d.backward_distance = 1
    elif kind == "strided":
        d.forward_strides, d.forward_distance, d.forward_offset = [2], 2 * n + 3, 5
        d.backward_strides, d.backward_distance, d.backward_offset = [1], n + 1, 2
    od = oracle.OracleDescriptor(lengths=[n], number_of_transforms=batch, forward_strides=list(d.forward_strides),
                                 backward_strides=list(d.backward_strides), forward_distance=d.forward_distance,
                                 backward_distance=d.backward_distance, forward_offset=d.forward_offset,
                                 backward_offset=d.backward_offset)
    host_in, host_ref = oracle.expected_io(od, oracle.FORWARD)
    dev = torch.device("cuda", 0)
    out = np.full_like(host_ref, complex(oracle.PADDING_VALUE, oracle.PADDING_VALUE))
    j = np.arange(n)
    for r in range(world):
        sh = shard_descriptor(d, world, r)
        if sh.count == 0:
            continue
        b = np.arange(sh.first, sh.first + sh.count)
        # host-side scatter: the shard's elements into a rank-local buffer laid out by the shard descriptor
        loc = sh.desc
        gin = d.forward_offset + b[:, None] * d.forward_distance + j[None, :] * d.forward_strides[0]
        lb = np.arange(sh.count)
        lin = loc.forward_offset + lb[:, None] * loc.forward_distance + j[None, :] * loc.forward_strides[0]
        local_in = np.zeros(loc.get_input_count(pf.direction.FORWARD), dtype=host_in.dtype)
        local_in[lin] = host_in[gin]
        t_in = torch.from_numpy(local_in).to(dev)
        t_out = torch.zeros(loc.get_output_count(pf.direction.FORWARD), dtype=t_in.dtype, device=dev)
        plan = loc.commit(torch.cuda.current_stream(dev), 0)
        plan.compute_forward(t_in, t_out)
        torch.cuda.synchronize(dev)
        plan.destroy()
        lout = loc.backward_offset + lb[:, None] * loc.backward_distance + j[None, :] * loc.backward_strides[0]
        gout = d.backward_offset + b[:, None] * d.backward_distance + j[None, :] * d.backward_strides[0]
        out[gout] = t_out.cpu().numpy()[lout]
    addressed = np.zeros(host_ref.shape, bool)
    allb = np.arange(batch)
    addressed[d.backward_offset + allb[:, None] * d.backward_distance + j[None, :] * d.backward_strides[0]] = True
    rel = np.linalg.norm(out[addressed] - host_ref[addressed]) / np.linalg.norm(host_ref[addressed])
    assert rel < 1e-5 * np.log2(n), rel
