"""CPU tests that pin the oracle (numpy restatement + C port) against every literal known answer the reference's own
tests hold for this path and against each other."""
import ctypes
import os

import numpy as np
import pytest

import portfft_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_first_element_and_dtype():
    # SFC64(0), uniform(-1,1), real draw then imaginary draw (reference_data_wrangler.hpp:130-133)
    x, y = o.gen_data(3, [8], False)
    assert x.dtype == np.complex64 and y.dtype == np.complex64
    assert np.isclose(x[0, 0], 0.1373785 + 0.7557978j, atol=1e-7)
    xd, yd = o.gen_data(3, [8], True)
    assert xd.dtype == np.complex128
    # same stream in both precisions
    assert np.allclose(x, xd.astype(np.complex64))
    assert np.allclose(np.fft.fft(xd, axis=1), yd)


def test_descriptor_known_answers():
    # test/unit_test/descriptor.cpp:32,84-108
    d = o.OracleDescriptor([2, 3], number_of_transforms=2, forward_strides=[8, 3], forward_distance=15,
                           forward_offset=3, backward_strides=[2, 4], backward_distance=1, backward_offset=5)
    assert d.get_flattened_length() == 6
    assert d.get_input_count(o.FORWARD) == 33 and d.get_output_count(o.BACKWARD) == 33
    assert d.get_input_count(o.BACKWARD) == 17 and d.get_output_count(o.FORWARD) == 17
    assert o.get_default_strides([512, 512, 512]) == [262144, 512, 1]


def test_layout_classification():
    d = o.OracleDescriptor([64], number_of_transforms=7)
    assert o.get_layout(d, o.FORWARD) == o.PACKED
    d = o.OracleDescriptor([64], number_of_transforms=7, forward_strides=[7], forward_distance=1)
    assert o.get_layout(d, o.FORWARD) == o.BATCH_INTERLEAVED and o.get_layout(d, o.BACKWARD) == o.PACKED
    d = o.OracleDescriptor([64], number_of_transforms=7, forward_strides=[2], forward_distance=128)
    assert o.get_layout(d, o.FORWARD) == o.UNPACKED


# instantiate_fft_tests.hpp:322-373: (lengths, fwd_strides, bwd_strides, fwd_dist, bwd_dist, batch, placement)
INVALID = [
    ([0], [1], [1], 1, 1, 1, o.OUT_OF_PLACE),
    ([1], [1], [1], 1, 1, 0, o.OUT_OF_PLACE),
    ([5], [5], [1], 0, 5, 2, o.OUT_OF_PLACE),
    ([5], [1], [5], 5, 0, 2, o.OUT_OF_PLACE),
    ([5], [0], [1], 0, 5, 1, o.OUT_OF_PLACE),
    ([5], [1], [0], 5, 0, 1, o.OUT_OF_PLACE),
    ([5, 12], [12, 1], [12, 0], 60, 0, 1, o.OUT_OF_PLACE),
    ([8], [1], [1], 7, 8, 2, o.OUT_OF_PLACE),
    ([8, 4], [8, 2], [4, 1], 24, 24, 2, o.OUT_OF_PLACE),
    ([8], [2], [1], 16, 8, 2, o.IN_PLACE),
    ([8, 4], [8, 2], [8, 2], 48, 50, 2, o.IN_PLACE),
    ([4], [1], [1], 1, 4, 3, o.OUT_OF_PLACE),
    ([4], [1], [2], 4, 3, 3, o.OUT_OF_PLACE),
    ([8], [3333333], [3333333], 1, 1, 3333334, o.OUT_OF_PLACE),
    ([8], [2], [2], 2, 2, 2, o.OUT_OF_PLACE),
    ([8], [1], [1], 1, 1, 2, o.OUT_OF_PLACE),
]


@pytest.mark.parametrize("case", INVALID, ids=[str(i) for i in range(len(INVALID))])
def test_invalid_configurations_rejected(case):
    lengths, fs, bs, fd, bd, batch, pl = case
    d = o.OracleDescriptor(lengths, number_of_transforms=batch, placement=pl, forward_strides=fs,
                           backward_strides=bs, forward_distance=fd, backward_distance=bd)
    with pytest.raises(o.InvalidConfiguration):
        o.validate_descriptor(d)


def test_valid_reference_layouts_accepted():
    # a sample of the layouts the reference's positive tests use (instantiate_fft_tests.hpp:237-319)
    for n, fs, bs, fd, bd, batch in [(3, 4, 7, 12, 21, 33000), (9, 3, 4, 30, 40, 3), (8, 33, 99, 1, 3, 33),
                                     (8, 2, 66, 16, 2, 33), (85, 13, 13, 12, 12, 13), (4, 4, 4, 3, 3, 4),
                                     (96, 40, 40, 1, 1, 33), (8, 2, 2, 2, 2, 1), (75, 66, 66, 2, 2, 33)]:
        d = o.OracleDescriptor([n], number_of_transforms=batch, forward_strides=[fs], backward_strides=[bs],
                               forward_distance=fd, backward_distance=bd)
        o.validate_descriptor(d)


def test_reference_planner_predicates():
    # SURVEY.md App. A (computed there with the reference's own functions)
    assert max(n for n in range(1, 57) if o.ref_fits_in_wi(n, False)) == 31
    assert not o.ref_fits_in_wi(27, False)
    assert max(n for n in range(1, 57) if o.ref_fits_in_wi(n, True)) == 13
    assert [o.ref_factorize(n) for n in (64, 4096, 1000, 1 << 24, 512)] == [8, 64, 25, 4096, 16]
    assert [o.ref_factorize_sg(n, 32) for n in (64, 512, 1000, 40)] == [32, 32, 25, 20]
    assert o.ref_fits_in_sg(64, 32, False) and not o.ref_fits_in_sg(1000, 32, False)
    assert max(n for n in range(1, 2000) if o.ref_fits_in_sg(n, 32, False)) == 992
    assert max(n for n in range(1, 2000) if o.ref_fits_in_sg(n, 32, True)) == 416


def _c_oracle():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    args = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_double,
            ctypes.c_int]
    lib.pfft_oracle_fft_f32.argtypes = args
    lib.pfft_oracle_fft_f64.argtypes = args
    lib.ref_select_level.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_longlong,
                                     ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)]
    for f in ("ref_factorize", "ref_wi_temps", "ref_factorize_sg"):
        getattr(lib, f).restype = ctypes.c_longlong
    lib.ref_factorize.argtypes = [ctypes.c_longlong]
    lib.ref_wi_temps.argtypes = [ctypes.c_longlong]
    lib.ref_factorize_sg.argtypes = [ctypes.c_longlong, ctypes.c_int]
    lib.ref_fits_in_wi.argtypes = [ctypes.c_longlong, ctypes.c_int]
    lib.ref_fits_in_sg.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    return lib


def test_c_port_predicates_match_python_restatement():
    lib = _c_oracle()
    for n in list(range(1, 300)) + [512, 992, 1000, 1024, 4096, 9800, 68640]:
        assert lib.ref_factorize(n) == o.ref_factorize(n)
        assert lib.ref_wi_temps(n) == o.ref_wi_temps(n)
        assert lib.ref_factorize_sg(n, 32) == o.ref_factorize_sg(n, 32)
        for dbl in (0, 1):
            assert bool(lib.ref_fits_in_wi(n, dbl)) == o.ref_fits_in_wi(n, bool(dbl))
            assert bool(lib.ref_fits_in_sg(n, 32, dbl)) == o.ref_fits_in_sg(n, 32, bool(dbl))


def test_c_port_level_selection_matches_survey_appendix_a():
    lib = _c_oracle()
    fac = (ctypes.c_longlong * 64)()
    nf = ctypes.c_int()
    expect = {(64, 0): 1, (4096, 0): 2, (1000, 0): 2, (512, 0): 1, (1 << 24, 1): 3, (8, 0): 0, (31, 0): 0, (16, 1): 1}
    for (n, dbl), lvl in expect.items():
        assert lib.ref_select_level(n, dbl, 49152, fac, ctypes.byref(nf)) == lvl, (n, dbl)
    # 32 KiB local memory: N=4096 fp32 falls to the GLOBAL level
    assert lib.ref_select_level(4096, 0, 32768, fac, ctypes.byref(nf)) == 3
    prod = 1
    for i in range(nf.value):
        prod *= fac[i]
    assert prod == 4096


SIZES = [1, 2, 3, 4, 8, 9, 16, 31, 32, 64, 80, 96, 100, 128, 256, 512, 1000, 1024, 1536, 2048, 3072, 4096, 8192, 9800,
         15360, 16384, 32768, 65536, 68640]


@pytest.mark.parametrize("dbl", [False, True], ids=["float", "double"])
@pytest.mark.parametrize("n", SIZES)
def test_c_port_matches_numpy_oracle(n, dbl):
    """The C restatement of the reference's algorithm agrees with the numpy oracle on the reference's test sizes,
    forward and backward (unnormalised inverse * backward_scale)."""
    lib = _c_oracle()
    batch = 3
    x, y = o.gen_data(batch, [n], dbl)
    fn = lib.pfft_oracle_fft_f64 if dbl else lib.pfft_oracle_fft_f32
    out = np.empty_like(x)
    assert fn(x.ctypes.data, out.ctypes.data, n, batch, 0, 2.0, 2) == 0
    bound = o.rel_l2_bound(n, dbl)
    assert np.linalg.norm(out - 2.0 * y) / max(np.linalg.norm(2.0 * y), 1e-30) <= bound
    back = np.empty_like(x)
    assert fn(y.ctypes.data, back.ctypes.data, n, batch, 1, 1.0 / n, 2) == 0
    assert np.linalg.norm(back - x) / np.linalg.norm(x) <= bound


def test_verify_dft_catches_errors():
    d = o.OracleDescriptor([16], number_of_transforms=3, backward_offset=4, backward_strides=[2], backward_distance=40)
    _, ref = o.expected_io(d, o.FORWARD)
    o.verify_dft(d, o.FORWARD, ref, ref.copy())
    bad = ref.copy()
    bad[1] = 0  # prefix before the offset must be bit identical
    with pytest.raises(AssertionError):
        o.verify_dft(d, o.FORWARD, ref, bad)
    bad = ref.copy()
    bad[5] = 1  # padding between strided elements must stay -5
    with pytest.raises(AssertionError):
        o.verify_dft(d, o.FORWARD, ref, bad)
    bad = ref.copy()
    bad[4] += 1e-2
    with pytest.raises(AssertionError):
        o.verify_dft(d, o.FORWARD, ref, bad)


def test_backward_expected_is_scaled_input():
    # reference_data_wrangler.hpp:202-210
    d = o.OracleDescriptor([8], number_of_transforms=2, backward_scale=0.5)
    x, y = o.gen_data(2, [8], False)
    inp, ref = o.expected_io(d, o.BACKWARD)
    assert np.array_equal(inp, y.reshape(-1))
    assert np.allclose(ref, x.reshape(-1) * 0.5 * 8)
