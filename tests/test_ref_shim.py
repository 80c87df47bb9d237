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


@pytest.mark.parametrize("dbl", [False, True], ids=["float", "double"])
@pytest.mark.parametrize("n", [1000, 1024, 2048, 3072, 4096, 8192])
def test_reference_wg_dft_matches_numpy_and_port(n, dbl):
    """workgroup.hpp:319-346 (wg_dft / dimension_dft) run as it is on an emulated work-group of 2 x 32 lock-step host
    threads, with the twiddle layout of workgroup_dispatcher.hpp:382-443: agrees with numpy and with the C port."""
    ref, port = _ref(), _port()
    f = ref.ref_wg_dft_f64 if dbl else ref.ref_wg_dft_f32
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    x, y = o.gen_data(1, [n], dbl)
    out_ref = np.empty_like(x)
    f(x.ctypes.data, out_ref.ctypes.data, n)
    bound = o.rel_l2_bound(n, dbl)
    assert np.linalg.norm(out_ref - y) <= bound * np.linalg.norm(y)
    out_port = np.empty_like(x)
    (port.pfft_oracle_fft_f64 if dbl else port.pfft_oracle_fft_f32)(x.ctypes.data, out_port.ctypes.data, n, 1, 0, 1.0, 1)
    assert np.linalg.norm(out_port - out_ref) <= bound * np.linalg.norm(y)


def _arr(v):
    return (ctypes.c_size_t * max(1, len(v)))(*[int(x) for x in v])


def _ref_validate(lib, lengths, fs, bs, fd, bd, batch, placement, dbl=False):
    return lib.refshim_validate(int(dbl), placement, len(lengths), _arr(lengths), len(fs), _arr(fs), len(bs), _arr(bs),
                                fd, bd, batch)


def test_validation_matches_reference_code():
    """descriptor_validation.hpp + utils.hpp::get_layout run AS THEY ARE (compiled from /root/reference through the
    shim) against the numpy oracle and against the library's pfft_validate / pfft_get_layout:
      * invalid_configuration: the same verdict from all three, on the reference's own invalid list
        (instantiate_fft_tests.hpp:322-373) and on 4000 random descriptors;
      * unsupported_configuration (validate_layout, :57-81): the oracle's `reference_layout_limits` switch reproduces
        it; the library accepts those (it lifts the restriction) but never accepts an invalid one;
      * layout classification identical."""
    import random

    import portfft_b200 as pf
    from test_oracle import INVALID

    lib = _ref()
    szp = ctypes.POINTER(ctypes.c_size_t)
    lib.refshim_validate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, szp, ctypes.c_size_t, szp,
                                     ctypes.c_size_t, szp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
    lib.refshim_get_layout.argtypes = [ctypes.c_int, ctypes.c_size_t, szp, szp, szp, ctypes.c_size_t, ctypes.c_size_t,
                                       ctypes.c_size_t]
    for lengths, fs, bs, fd, bd, batch, pl in [c[:7] for c in INVALID]:
        assert _ref_validate(lib, lengths, fs, bs, fd, bd, batch, pl) == 1, (lengths, fs, bs, fd, bd, batch, pl)

    rng = random.Random(11)
    lay = {o.PACKED: 0, o.UNPACKED: 1, o.BATCH_INTERLEAVED: 2}
    counts = {0: 0, 1: 0, 2: 0}
    for _ in range(4000):
        rank = rng.choice([1, 1, 1, 2, 3])
        lengths = [rng.choice([1, 2, 3, 4, 5, 8, 16, 75, 100]) for _ in range(rank)]
        batch = rng.choice([1, 2, 3, 7, 33])
        pl = rng.choice([0, 1])

        def dom():
            kind = rng.random()
            if kind < 0.35:
                st = o.get_default_strides(lengths)
                return st, int(np.prod(lengths))
            if kind < 0.5 and rank == 1:
                return [batch], 1
            return [rng.choice([1, 2, 3, 4, 6, 8, 16, 33, 64]) for _ in range(rank)], rng.choice(
                [0, 1, 2, 3, 5, 8, 16, 40, 64, 300, 2000])

        fs, fd = dom()
        bs, bd = (fs, fd) if (pl == 0 and rng.random() < 0.7) else dom()
        verdict = _ref_validate(lib, lengths, fs, bs, fd, bd, batch, pl)
        counts[verdict] += 1
        od = o.OracleDescriptor(lengths, number_of_transforms=batch, placement=pl, forward_strides=fs,
                                backward_strides=bs, forward_distance=fd, backward_distance=bd)
        try:
            o.validate_descriptor(od, reference_layout_limits=True)
            oracle_verdict = 0
        except o.InvalidConfiguration:
            oracle_verdict = 1
        except o.UnsupportedConfiguration:
            oracle_verdict = 2
        assert oracle_verdict == verdict, (lengths, fs, bs, fd, bd, batch, pl, verdict, oracle_verdict)
        d = pf.descriptor(lengths)
        d.number_of_transforms, d.placement = batch, pf.placement(pl)
        d.forward_strides, d.backward_strides, d.forward_distance, d.backward_distance = list(fs), list(bs), fd, bd
        if verdict == 1:
            with pytest.raises(pf.invalid_configuration):
                d.validate()
        else:
            d.validate()  # accepted, including what the reference calls unsupported
            for dr in (0, 1):
                ref_layout = lib.refshim_get_layout(dr, rank, _arr(lengths), _arr(fs), _arr(bs), fd, bd, batch)
                assert ref_layout == lay[o.get_layout(od, dr)] == int(d.get_layout(pf.direction(dr)))
    assert counts[0] > 500 and counts[1] > 500 and counts[2] > 50, counts
