"""GPU parity tests: the reference's FFT test grid (/root/reference/test/unit_test/instantiate_fft_tests.hpp:95-319,
SURVEY.md App. B) re-expressed in pytest, run through the C ABI against the numpy oracle, for float and double."""
import pytest

from fft_check import BI, P, U, CaseParams, run_case
from grid_cases import (ALL_LAYOUTS, BOTH_DIR, GLOBAL_LAYOUTS, MD_LAYOUTS, SCALARS, STORAGES, SUITES, basic, fuzz_cases,
                        offsets, real, scaled)

pytestmark = pytest.mark.gpu

CASES = [pytest.param(tp, id=f"{suite}-{tp.ident()}") for suite, tps in SUITES.items() for tp in tps]


@pytest.mark.parametrize("tp", CASES)
def test_reference_grid(tp):
    run_case(tp)


# ---- every level's kernel stays under test: the planner's default choice prefers the TMA tile kernels, so the
# ---- warp-level (SUBGROUP) kernel and the generic block-level kernel are forced for a slice of the grid
FORCED = {
    "subgroup": ({"PFFT_FORCE_LEVEL": "1"},
                 basic(ALL_LAYOUTS[:3], BOTH_DIR, STORAGES, [1, 131], [64, 96, 128, 256, 512, 1024])),
    "workgroup_generic": ({"PFFT_FORCE_LEVEL": "2", "PFFT_NO_COL": "1", "PFFT_NO_R3": "1", "PFFT_CUBE_VARIANT": "-1",
                           "PFFT_NO_COLG": "1"},
                          basic(ALL_LAYOUTS, BOTH_DIR, STORAGES, [3], [16, 64, 100, 512, 1000, 4096]) +
                          basic(GLOBAL_LAYOUTS, ["fwd"], ["interleaved"], [3], [32768, 65536]) +
                          basic(MD_LAYOUTS, ["fwd"], ["interleaved"], [3], [[16, 512], [64, 64, 64]])),
    # L2-resident execution (batch chunks of multi-pass plans, plane chunks of N-D transforms): forced on small
    # problems by shrinking the L2 budget
    "l2_chunked": ({"PFFT_L2_CHUNK_BYTES": "300000"},
                   basic(ALL_LAYOUTS, BOTH_DIR, STORAGES, [5], [16384]) +
                   basic(GLOBAL_LAYOUTS, BOTH_DIR, ["interleaved"], [7], [4099, 9800]) +
                   basic(MD_LAYOUTS, BOTH_DIR, STORAGES, [1, 3], [[64, 64, 64], [40, 16, 512], [300, 256], [37, 8, 64]]) +
                   offsets(MD_LAYOUTS, BOTH_DIR, [2], [[64, 64, 64]], [(3, 3), (16, 16)]) +
                   offsets([("OOP", P, P)], BOTH_DIR, [2], [[64, 64, 64]], [(0, 5), (7, 2)]) +
                   scaled("fwd", [[64, 64, 64], 16384], -1.0, 2.0) + scaled("bwd", [[64, 64, 64], 16384], -1.0, 2.0) +
                   real(BOTH_DIR, STORAGES, [5], [65536])),
    # two GLOBAL-level passes fused into one persistent kernel with an L2-resident ring (wg_fused.cu): small chunks so
    # that modest batches already run many chunks through the ring slots and both arrival counters
    "fused_small_chunks": ({"PFFT_FUSE": "1", "PFFT_FUSE_CHUNK_KB": "512"},
                           basic(GLOBAL_LAYOUTS, BOTH_DIR, ["interleaved"], [8, 37], [65536]) +
                           offsets([("OOP", P, P)], BOTH_DIR, [9], [65536], [(0, 6), (4, 0)]) +
                           scaled("fwd", [65536], -1.0, 2.0) + scaled("bwd", [65536], -1.0, 2.0) +
                           basic([("OOP", P, P)], BOTH_DIR, ["interleaved"], [2], [1 << 24])),
    "fused_lead2": ({"PFFT_FUSE": "1", "PFFT_FUSE_CHUNK_KB": "1024", "PFFT_FUSE_LEAD": "2"},
                    basic([("OOP", P, P)], BOTH_DIR, ["interleaved"], [40], [65536])),
    "fused_default_chunks": ({"PFFT_FUSE": "1"}, [c for c in basic([("OOP", P, P)], BOTH_DIR, ["interleaved"], [128], [65536])
                                                  if c.scalar == "float" or c.dir == "fwd"]),
    "cube_direct_loads": ({"PFFT_CUBE_VARIANT": "1"}, basic([("IP", P, P), ("OOP", P, P)], BOTH_DIR, ["interleaved"],
                                                            [5], [4096])),
}
def _applies(env, tp):
    # the fp64 warp-level kernel covers N <= 512: larger sizes have no forced sub-group path (not a case, not a skip)
    return not (env.get("PFFT_FORCE_LEVEL") == "1" and tp.scalar == "double" and max(tp.lengths) > 512)


FORCED_CASES = [pytest.param(env, tp, id=f"{name}-{tp.ident()}") for name, (env, tps) in FORCED.items() for tp in tps
                if _applies(env, tp)]


@pytest.mark.parametrize("env,tp", FORCED_CASES)
def test_forced_kernel_paths(env, tp, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    run_case(tp)


@pytest.mark.parametrize("scalar", SCALARS)
@pytest.mark.parametrize("n,batch", [(8, 5), (30, 5), (81, 5), (1000, 5), (4096, 5), (16384, 5),
                                     # fused REAL forms on rows that are only 8-byte aligned, steady state of the ring
                                     (512, 5001), (1024, 2501), (8192, 1301), (4096, 1300)])
def test_real_in_place(n, batch, scalar):
    """REAL domain, IN_PLACE (committed_descriptor.hpp:201-206: `compute_forward(Scalar* inout)`): rows padded to
    2 * (n // 2 + 1) reals hold the real input and then the half spectrum; every pass that reads the input finishes
    before the first pass writes the output, so no relation between the two layouts is needed beyond this padding.
    Forward against numpy rfft, then backward in place against the input (backward_scale = 1 / n)."""
    import numpy as np
    import torch

    import portfft_b200 as pf
    import portfft_oracle as oracle

    h = n // 2 + 1
    d = pf.descriptor([n], scalar, pf.domain.REAL)
    d.number_of_transforms = batch
    d.placement = pf.placement.IN_PLACE
    d.forward_distance, d.backward_distance = 2 * h, h
    d.backward_scale = 1.0 / n
    rdt = np.float64 if scalar == "double" else np.float32
    rng = np.random.Generator(np.random.SFC64(0))
    x = rng.uniform(-1, 1, (batch, n)).astype(rdt)
    host = np.full((batch, 2 * h), oracle.PADDING_VALUE, dtype=rdt)
    host[:, :n] = x
    buf = torch.from_numpy(host.reshape(-1)).cuda()
    c = d.commit(torch.cuda.current_stream(), 0)
    c.compute_forward(buf)
    torch.cuda.synchronize()
    spec = buf.cpu().numpy().view(np.complex128 if scalar == "double" else np.complex64).reshape(batch, h)
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    bound = oracle.rel_l2_bound(n, scalar == "double")
    err = np.max(np.linalg.norm(spec - ref, axis=1) / np.linalg.norm(ref, axis=1))
    assert err <= bound, (err, bound)
    c.compute_backward(buf)
    torch.cuda.synchronize()
    back = buf.cpu().numpy().reshape(batch, 2 * h)[:, :n]
    err = np.max(np.linalg.norm(back - x, axis=1) / np.linalg.norm(x, axis=1))
    assert err <= bound, (err, bound)
    c.destroy()


@pytest.mark.parametrize("tp", fuzz_cases(400, 17), ids=lambda tp: tp.ident())
def test_random_layouts(tp):
    """The planner fuzz of tests/test_plan_emulator.py (random ranks, lengths incl. primes, nested padded layouts in both
    domains, offsets, scales, storage, precision, complex and REAL) through the CUDA kernels: the GPU coverage of N-D
    transforms with non-default strides."""
    run_case(tp)


@pytest.mark.parametrize("n,scalar", [(65536, "float"), (1 << 20, "double"), (4099, "float"), (8192, "float")])
def test_copies_compute_concurrently(n, scalar):
    """pfft_clone (the reference's copy constructor, committed_descriptor_impl.hpp:774-803): a copy shares the twiddles
    and owns its workspaces.  The original and the copy run multi-pass plans (GLOBAL level, Bluestein, REAL) on two
    streams at the same time, repeatedly; every result equals the one computed alone."""
    import torch

    import portfft_b200 as pf

    real = n == 8192
    d = pf.descriptor([n], scalar, pf.domain.REAL if real else pf.domain.COMPLEX)
    d.number_of_transforms = 4
    if real:
        d.backward_distance = n // 2 + 1
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a = d.commit(s1, 0)
    b = a.copy()
    if not real:
        assert a.workspace_bytes() > 0 and b.workspace_bytes() == a.workspace_bytes()
    cdt = torch.complex128 if scalar == "double" else torch.complex64
    rdt = torch.float64 if scalar == "double" else torch.float32
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    n_out = 4 * (n // 2 + 1 if real else n)

    def rand():
        if real:
            return torch.rand(4 * n, dtype=rdt, device="cuda", generator=g) * 2 - 1
        return torch.view_as_complex(torch.rand(4 * n, 2, dtype=rdt, device="cuda", generator=g) * 2 - 1)

    x1, x2 = rand(), rand()
    ref1, ref2 = (torch.empty(n_out, dtype=cdt, device="cuda") for _ in range(2))
    o1, o2 = (torch.empty(n_out, dtype=cdt, device="cuda") for _ in range(2))
    torch.cuda.synchronize()
    a.compute_forward(x1, ref1, queue=s1)
    s1.synchronize()
    a.compute_forward(x2, ref2, queue=s1)
    s1.synchronize()
    for _ in range(25):
        a.compute_forward(x1, o1, queue=s1)
        b.compute_forward(x2, o2, queue=s2)
    torch.cuda.synchronize()
    assert torch.equal(o1, ref1) and torch.equal(o2, ref2)
    b.destroy()
    a.compute_forward(x1, o1, queue=s1)  # the tables outlive the copy
    torch.cuda.synchronize()
    assert torch.equal(o1, ref1)
    a.destroy()
