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
