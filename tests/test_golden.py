"""Committed golden fixtures (tests/golden, made by tools/make_golden.py in the authoring container):
numpy_* = the reference test generator's input/expected pairs, refcode_* = outputs of the reference's own wi_dft /
sg_dft code.  CPU tests keep the oracle honest on the GPU box; the GPU test runs the CUDA path against them."""
import glob
import os

import numpy as np
import pytest

import portfft_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _numpy_fixtures():
    return sorted(glob.glob(os.path.join(GOLD, "numpy_*.npz")))


def _real_fixtures():
    return sorted(glob.glob(os.path.join(GOLD, "numpyre_*.npz")))


def test_fixtures_present():
    assert len(_numpy_fixtures()) == 14 and len(_real_fixtures()) == 8
    for name in ("refcode_f32.npz", "refcode_f64.npz", "refcode_wg_f32.npz", "refcode_wg_f64.npz"):
        assert os.path.exists(os.path.join(GOLD, name)), name


@pytest.mark.parametrize("path", _numpy_fixtures(), ids=os.path.basename)
def test_oracle_reproduces_numpy_fixture(path):
    """Same bytes for the input stream, same transform to rounding, whatever numpy this machine has."""
    f = np.load(path)
    x, y = f["input"], f["output"]
    dbl = x.dtype == np.complex128
    x2, y2 = o.gen_data(x.shape[0], list(x.shape[1:]), dbl)
    assert np.array_equal(x, x2)
    assert np.linalg.norm(y - y2) <= 4 * np.finfo(x.real.dtype).eps * np.linalg.norm(y)


@pytest.mark.parametrize("path", _real_fixtures(), ids=os.path.basename)
def test_oracle_reproduces_real_fixture(path):
    f = np.load(path)
    x, y = f["input"], f["output"]
    dbl = x.dtype == np.float64
    x2, y2 = o.gen_data(x.shape[0], list(x.shape[1:]), dbl, is_real=True)
    assert np.array_equal(x, x2) and y.shape == y2.shape and y.shape[-1] == x.shape[-1] // 2 + 1
    assert np.linalg.norm(y - y2) <= 4 * np.finfo(x.dtype).eps * np.linalg.norm(y)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_refcode_fixture_agrees_with_oracle(tag):
    f = np.load(os.path.join(GOLD, f"refcode_{tag}.npz"))
    dbl = tag == "f64"
    pos = 0
    for n in f["sizes"]:
        n = int(n)
        x = f["inputs"][pos:pos + n]
        out = f["outputs"][pos:pos + n]
        pos += n
        x2, y2 = o.gen_data(1, [n], dbl)
        assert np.array_equal(x, x2.reshape(-1))
        assert np.linalg.norm(out - y2.reshape(-1)) <= o.rel_l2_bound(n, dbl) * np.linalg.norm(y2)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_refcode_wg_fixture_agrees_with_oracle(tag):
    """The reference's WORKGROUP-level code (wg_dft, workgroup.hpp:319-346) on N = 1000 ... 4096 against numpy."""
    f = np.load(os.path.join(GOLD, f"refcode_wg_{tag}.npz"))
    dbl = tag == "f64"
    pos = 0
    for n in f["sizes"]:
        n = int(n)
        x, out = f["inputs"][pos:pos + n], f["outputs"][pos:pos + n]
        pos += n
        x2, y2 = o.gen_data(1, [n], dbl)
        assert np.array_equal(x, x2.reshape(-1))
        assert np.linalg.norm(out - y2.reshape(-1)) <= o.rel_l2_bound(n, dbl) * np.linalg.norm(y2)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_cuda_matches_reference_wg_code_outputs(tag):
    """CUDA path vs what the reference's own WORKGROUP-level code produced on the same inputs -- the lengths of
    BASELINE configs C2 (4096) and C3 (1000) among them: north_star's 'match the reference's own implementation on
    identical inputs' with the relative-L2 bound 1e-5 log2 N (fp32) / 1e-13 log2 N (fp64)."""
    import torch

    import portfft_b200 as pf

    f = np.load(os.path.join(GOLD, f"refcode_wg_{tag}.npz"))
    dbl = tag == "f64"
    pos = 0
    for n in f["sizes"]:
        n = int(n)
        x, ref = f["inputs"][pos:pos + n], f["outputs"][pos:pos + n]
        pos += n
        d = pf.descriptor([n], "double" if dbl else "float")
        c = d.commit(torch.cuda.current_stream(), 0)
        tin = torch.from_numpy(x.copy()).cuda()
        tout = torch.empty_like(tin)
        c.compute_forward(tin, tout)
        torch.cuda.synchronize()
        got = tout.cpu().numpy()
        assert np.linalg.norm(got - ref) <= o.rel_l2_bound(n, dbl) * np.linalg.norm(ref), n
        c.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_cuda_matches_reference_code_outputs(tag):
    """CUDA path vs what the reference's own wi_dft / sg_dft code produced on the same inputs."""
    import torch

    import portfft_b200 as pf

    f = np.load(os.path.join(GOLD, f"refcode_{tag}.npz"))
    dbl = tag == "f64"
    pos = 0
    for n in f["sizes"]:
        n = int(n)
        x = f["inputs"][pos:pos + n]
        ref = f["outputs"][pos:pos + n]
        pos += n
        d = pf.descriptor([n], "double" if dbl else "float")
        c = d.commit(torch.cuda.current_stream(), 0)
        tin = torch.from_numpy(x.copy()).cuda()
        tout = torch.empty_like(tin)
        c.compute_forward(tin, tout)
        torch.cuda.synchronize()
        got = tout.cpu().numpy()
        assert np.linalg.norm(got - ref) <= o.rel_l2_bound(n, dbl) * np.linalg.norm(ref), n
        c.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("path", _numpy_fixtures(), ids=os.path.basename)
def test_cuda_matches_numpy_fixture(path):
    import torch

    import portfft_b200 as pf

    f = np.load(path)
    x, y = f["input"], f["output"]
    dbl = x.dtype == np.complex128
    d = pf.descriptor(list(x.shape[1:]), "double" if dbl else "float")
    d.number_of_transforms = x.shape[0]
    c = d.commit(torch.cuda.current_stream(), 0)
    tin = torch.from_numpy(x.reshape(-1).copy()).cuda()
    tout = torch.empty_like(tin)
    c.compute_forward(tin, tout)
    torch.cuda.synchronize()
    n = int(np.prod(x.shape[1:]))
    got = tout.cpu().numpy().reshape(x.shape[0], -1)
    yy = y.reshape(x.shape[0], -1)
    err = np.max(np.linalg.norm(got - yy, axis=1) / np.linalg.norm(yy, axis=1))
    assert err <= o.rel_l2_bound(n, dbl)
    c.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("path", _real_fixtures(), ids=os.path.basename)
def test_cuda_matches_real_fixture(path):
    """REAL domain: real input -> half spectrum (dense rows), then back with backward_scale = 1 / N."""
    import torch

    import portfft_b200 as pf

    f = np.load(path)
    x, y = f["input"], f["output"]
    dbl = x.dtype == np.float64
    dims = list(x.shape[1:])
    n = int(np.prod(dims))
    cdims = dims[:-1] + [dims[-1] // 2 + 1]
    d = pf.descriptor(dims, "double" if dbl else "float", pf.domain.REAL)
    d.number_of_transforms = x.shape[0]
    d.backward_strides = pf.get_default_strides(cdims)
    d.backward_distance = int(np.prod(cdims))
    d.backward_scale = 1.0 / n
    c = d.commit(torch.cuda.current_stream(), 0)
    tin = torch.from_numpy(x.reshape(-1).copy()).cuda()
    tout = torch.empty(y.size, dtype=torch.complex128 if dbl else torch.complex64, device="cuda")
    c.compute_forward(tin, tout)
    torch.cuda.synchronize()
    got = tout.cpu().numpy().reshape(x.shape[0], -1)
    yy = y.reshape(x.shape[0], -1)
    assert np.max(np.linalg.norm(got - yy, axis=1) / np.linalg.norm(yy, axis=1)) <= o.rel_l2_bound(n, dbl)
    back = torch.zeros_like(tin)
    c.compute_backward(tout, back)
    torch.cuda.synchronize()
    xx = x.reshape(x.shape[0], -1)
    bb = back.cpu().numpy().reshape(x.shape[0], -1)
    assert np.max(np.linalg.norm(bb - xx, axis=1) / np.linalg.norm(xx, axis=1)) <= o.rel_l2_bound(n, dbl)
    c.destroy()
