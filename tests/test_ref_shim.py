"""Pins the oracle against the REFERENCE's own code: oracle/_ref/libportfft_ref.so is compiled from
/root/reference/src/portfft/common/{workitem,subgroup}.hpp (in place, through oracle/ref_shim) by `make -C oracle ref`.
The .so is prebuilt in the authoring container and travels to the GPU box; nothing here reads /root/reference."""
import ctypes
import os

import numpy as np
import pytest

import portfft_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libportfft_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")


def _ref():
    lib = ctypes.CDLL(REF)
    for f in ("refshim_factorize", "refshim_wi_temps", "refshim_factorize_sg"):
        getattr(lib, f).restype = ctypes.c_longlong
    lib.refshim_factorize.argtypes = [ctypes.c_longlong]
    lib.refshim_wi_temps.argtypes = [ctypes.c_longlong]
    lib.refshim_factorize_sg.argtypes = [ctypes.c_longlong, ctypes.c_int]
    lib.refshim_fits_in_wi.argtypes = [ctypes.c_longlong, ctypes.c_int]
    lib.refshim_fits_in_sg.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    for f in ("ref_wi_dft_f32", "ref_wi_dft_f64"):
        getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    for f in ("ref_sg_dft_f32", "ref_sg_dft_f64"):
        getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return lib


def _port():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    a = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_double,
         ctypes.c_int]
    lib.pfft_oracle_fft_f32.argtypes = a
    lib.pfft_oracle_fft_f64.argtypes = a
    return lib


def test_planner_predicates_match_reference_code():
    ref = _ref()
    for n in list(range(1, 600)) + [992, 1000, 1024, 4096, 9800, 68640, 1 << 24]:
        assert ref.refshim_factorize(n) == o.ref_factorize(n)
        assert ref.refshim_factorize_sg(n, 32) == o.ref_factorize_sg(n, 32)
        if n < 2000:
            assert ref.refshim_wi_temps(n) == o.ref_wi_temps(n)
            for dbl in (0, 1):
                assert bool(ref.refshim_fits_in_wi(n, dbl)) == o.ref_fits_in_wi(n, bool(dbl))
                assert bool(ref.refshim_fits_in_sg(n, 32, dbl)) == o.ref_fits_in_sg(n, 32, bool(dbl))


@pytest.mark.parametrize("dbl", [False, True], ids=["float", "double"])
def test_reference_wi_dft_matches_numpy_and_port(dbl):
    """workitem.hpp:200-219 run as-is: agrees with numpy to rounding and with the C port of the same algorithm to the
    last bits (same operation order; twiddles rounded from double in both)."""
    ref, port = _ref(), _port()
    sizes = [n for n in range(1, 57) if o.ref_fits_in_wi(n, dbl)]
    assert sizes[-1] == (13 if dbl else 31)
    for n in sizes:
        x, y = o.gen_data(1, [n], dbl)
        out_ref = np.empty_like(x)
        (ref.ref_wi_dft_f64 if dbl else ref.ref_wi_dft_f32)(x.ctypes.data, out_ref.ctypes.data, n)
        bound = o.rel_l2_bound(n, dbl)
        assert np.linalg.norm(out_ref - y) <= bound * max(np.linalg.norm(y), 1e-30), n
        out_port = np.empty_like(x)
        (port.pfft_oracle_fft_f64 if dbl else port.pfft_oracle_fft_f32)(x.ctypes.data, out_port.ctypes.data, n, 1, 0,
                                                                         1.0, 1)
        eps = np.finfo(np.float64 if dbl else np.float32).eps
        assert np.max(np.abs(out_port - out_ref)) <= 8 * eps * max(1.0, np.max(np.abs(y))), n


@pytest.mark.parametrize("dbl", [False, True], ids=["float", "double"])
@pytest.mark.parametrize("n", [32, 64, 96, 100, 128, 256, 512, 75, 85, 104, 70])
def test_reference_sg_dft_matches_numpy_and_port(n, dbl):
    """subgroup.hpp:271-291 on 32 lock-step emulated lanes: agrees with numpy and with the C port's sg_dft."""
    if not o.ref_fits_in_sg(n, 32, dbl) or o.ref_fits_in_wi(n, dbl):
        pytest.skip("size is not SUBGROUP level in this precision")
    ref, port = _ref(), _port()
    f_sg = o.ref_factorize_sg(n, 32)
    f_wi = n // f_sg
    x, y = o.gen_data(1, [n], dbl)
    out_ref = np.empty_like(x)
    (ref.ref_sg_dft_f64 if dbl else ref.ref_sg_dft_f32)(x.ctypes.data, out_ref.ctypes.data, f_wi, f_sg)
    bound = o.rel_l2_bound(n, dbl)
    assert np.linalg.norm(out_ref - y) <= bound * np.linalg.norm(y)
    out_port = np.empty_like(x)
    (port.pfft_oracle_fft_f64 if dbl else port.pfft_oracle_fft_f32)(x.ctypes.data, out_port.ctypes.data, n, 1, 0, 1.0, 1)
    assert np.linalg.norm(out_port - out_ref) <= bound * np.linalg.norm(y)
