"""CPU tests of the planner: the exported pass list (pfft_plan_export) replayed in numpy (tests/plan_emulator.py) must
reproduce the oracle's expected output on the same seeded inputs -- pass geometry, offsets, inter-factor twiddles,
digit reversal of the GLOBAL level, N-D passes, and the Bluestein passes with their chirp tables.  No GPU needed: the
per-pass DFT is numpy's, so this checks WHAT the kernels are asked to do; tests/test_fft_gpu.py checks the kernels."""
import numpy as np
import pytest

import portfft_oracle as oracle
from fft_check import BI, P, U, CaseParams, make_descriptors
from grid_cases import fuzz_cases
from plan_emulator import run_plan

CASES = [
    # single pass, every layout family
    CaseParams([64], 5, "OOP", P, P, "fwd"),
    CaseParams([1000], 3, "IP", BI, BI, "bwd", backward_scale=1e-3),
    CaseParams([96], 4, "OOP", U, U, "fwd", forward_strides=[3], backward_strides=[2], forward_distance=300,
               backward_distance=200, forward_offset=7, backward_offset=3),
    # GLOBAL level: two and three factors, both directions, offsets, batch-interleaved and strided layouts
    CaseParams([16384], 2, "OOP", P, P, "fwd"),
    CaseParams([32768], 2, "IP", P, P, "bwd", backward_scale=0.5),
    CaseParams([9800], 3, "OOP", P, P, "fwd", forward_offset=5, backward_offset=9),
    CaseParams([8192 * 4], 3, "OOP", BI, P, "fwd"),
    CaseParams([8192 * 2], 3, "OOP", P, BI, "bwd"),
    CaseParams([16384], 2, "OOP", U, U, "fwd", forward_strides=[2], backward_strides=[3], forward_distance=40000,
               backward_distance=50000),
    CaseParams([1 << 18], 1, "OOP", P, P, "fwd", scalar="double"),
    # N-D
    CaseParams([4, 6, 10], 2, "OOP", P, P, "fwd"),
    CaseParams([16, 512], 2, "IP", P, P, "bwd"),
    CaseParams([8, 16384], 1, "OOP", P, P, "fwd"),
    # Bluestein: prime lengths, one CTA per convolution (M <= 8192 fp32 / 4096 fp64) and the multi-pass form
    CaseParams([37], 3, "OOP", P, P, "fwd"),
    CaseParams([1031], 2, "OOP", P, P, "bwd", backward_scale=1.0 / 1031),
    CaseParams([2 * 1031], 2, "IP", P, P, "fwd", scalar="double"),
    CaseParams([67], 5, "OOP", BI, BI, "fwd", storage="split"),
    CaseParams([4099], 2, "OOP", P, P, "fwd"),
    CaseParams([4099], 2, "OOP", P, P, "bwd", scalar="double"),
    CaseParams([65537], 1, "OOP", P, P, "fwd"),
    CaseParams([6, 37], 2, "OOP", P, P, "fwd"),
    CaseParams([37, 6], 2, "OOP", P, P, "bwd"),
    # REAL domain: even lengths on the pair view and through pack / unpack (strided, odd offsets), odd lengths,
    # multi-pass half-length transforms, split complex storage
    CaseParams([8], 3, "OOP", P, P, "fwd", domain="real"),
    CaseParams([8], 3, "OOP", P, P, "bwd", domain="real"),
    CaseParams([2], 3, "OOP", P, P, "fwd", domain="real"),
    CaseParams([1], 3, "OOP", P, P, "bwd", domain="real"),
    CaseParams([9], 3, "OOP", P, P, "fwd", domain="real"),
    CaseParams([15], 2, "OOP", P, P, "bwd", domain="real", backward_scale=1.0 / 15),
    CaseParams([4096], 2, "OOP", P, P, "fwd", domain="real", storage="split"),
    CaseParams([1000], 3, "OOP", P, P, "bwd", domain="real", scalar="double", backward_scale=1e-3),
    CaseParams([96], 4, "OOP", U, U, "fwd", domain="real", forward_strides=[3], backward_strides=[2],
               forward_distance=300, backward_distance=100, forward_offset=7, backward_offset=3),
    CaseParams([96], 4, "OOP", U, U, "bwd", domain="real", forward_strides=[1], backward_strides=[2],
               forward_distance=97, backward_distance=100, forward_offset=1, backward_offset=3),
    CaseParams([65536], 2, "OOP", P, P, "fwd", domain="real"),
    CaseParams([65536], 2, "OOP", P, P, "bwd", domain="real", storage="split"),
    CaseParams([32768], 2, "OOP", U, U, "fwd", domain="real", forward_strides=[2], backward_strides=[1],
               forward_distance=70000, backward_distance=16385),
    CaseParams([3 * 16384], 1, "OOP", U, U, "bwd", domain="real", forward_strides=[2], backward_strides=[1],
               forward_distance=100000, backward_distance=30000),
    # REAL N-D: real-to-complex rows, then complex passes on the half spectrum (forward); workspace + inverse passes,
    # then complex-to-real rows (backward)
    CaseParams([4, 8], 3, "OOP", P, P, "fwd", domain="real"),
    CaseParams([4, 8], 3, "OOP", P, P, "bwd", domain="real", storage="split"),
    CaseParams([6, 9], 2, "OOP", P, P, "bwd", domain="real", backward_scale=1.0 / 54),
    CaseParams([2, 3, 4], 2, "OOP", P, P, "fwd", domain="real", storage="split"),
    CaseParams([5, 4, 6], 1, "OOP", P, P, "bwd", domain="real", scalar="double"),
    CaseParams([37, 8], 2, "OOP", P, P, "bwd", domain="real"),
    CaseParams([8, 16384], 1, "OOP", P, P, "fwd", domain="real"),
    CaseParams([8, 16384], 1, "OOP", P, P, "bwd", domain="real"),
    # REAL pre / post-processing fused into the transform pass (PassHost::fuse_real): tile-kernel rows (half lengths
    # 64 .. 8192), dense and default half-spectrum distances, N-D, the
    # thread-level form needs a large batch (one 128-byte line per row)
    CaseParams([4096], 5, "OOP", U, U, "fwd", domain="real", forward_strides=[1], backward_strides=[1],
               forward_distance=4096, backward_distance=2049),
    CaseParams([4096], 5, "OOP", U, U, "bwd", domain="real", forward_strides=[1], backward_strides=[1],
               forward_distance=4096, backward_distance=2049, backward_scale=1.0 / 4096),
    # (rows that are 8-byte aligned only, as in the in-place layout: n + 2 reals = n / 2 + 1 pairs apart; the in-place
    # call itself aliases a real and a complex view of one buffer and is covered on the GPU, test_real_in_place)
    CaseParams([1024], 7, "OOP", U, U, "fwd", domain="real", forward_strides=[1], backward_strides=[1],
               forward_distance=1026, backward_distance=513),
    CaseParams([1024], 7, "OOP", U, U, "bwd", domain="real", forward_strides=[1], backward_strides=[1],
               forward_distance=1026, backward_distance=513),
    CaseParams([128], 9, "OOP", P, P, "fwd", domain="real"),
    CaseParams([256], 9, "OOP", P, P, "bwd", domain="real", scalar="double"),
    CaseParams([16, 512], 2, "OOP", P, P, "fwd", domain="real"),
    CaseParams([16, 512], 2, "OOP", P, P, "bwd", domain="real"),
    CaseParams([32], 4100, "OOP", U, U, "fwd", domain="real", forward_strides=[1], backward_strides=[1],
               forward_distance=32, backward_distance=17),
    CaseParams([32], 4100, "OOP", P, P, "bwd", domain="real"),
]


@pytest.mark.parametrize("tp", CASES, ids=[tp.ident() for tp in CASES])
def test_exported_plan_matches_oracle(tp):
    d, od = make_descriptors(tp)
    dr = oracle.FORWARD if tp.dir == "fwd" else oracle.BACKWARD
    host_in, host_ref = oracle.expected_io(od, dr)
    if tp.placement == "IP":
        out = host_in.copy()
        run_plan(d, dr, out, out)
    else:
        pad = oracle.PADDING_VALUE
        out = np.full(host_ref.shape, complex(pad, pad) if np.iscomplexobj(host_ref) else pad, dtype=host_ref.dtype)
        run_plan(d, dr, host_in.copy(), out)
    oracle.verify_dft(od, dr, host_ref, out)


@pytest.mark.parametrize("scalar,L", [("float", 37), ("double", 1031), ("float", 4099)])
def test_bluestein_tables_match_numpy(scalar, L):
    """pfft_table_host: the chirp w_j = exp(-i pi j^2 / L) and the transformed convolution kernel FFT_M(conj w) the
    plan uploads (csrc/tables.cpp, long double on the host) against a float64 numpy evaluation."""
    from portfft_b200 import api

    M = 1
    while M < 2 * L - 1:
        M *= 2
    j = np.arange(L, dtype=np.int64)
    w = np.exp(-1j * np.pi * ((j * j) % (2 * L)) / L)
    eps = 1e-6 if scalar == "float" else 1e-14
    np.testing.assert_allclose(api.mod_table(scalar, 1, L, M), w, atol=eps)
    np.testing.assert_allclose(api.mod_table(scalar, 2, L, M), w / M, atol=eps / M)
    b = np.zeros(M, dtype=np.complex128)
    b[:L] = np.conj(w)
    b[M - L + 1:] = np.conj(w[1:])[::-1]
    ref = np.fft.fft(b)
    got = api.mod_table(scalar, 3, L, M)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < (1e-6 if scalar == "float" else 1e-13)


def _grid_sample():
    """Every 3rd case of the GPU parity grid (tests/test_fft_gpu.py) whose buffers stay small: the planner's pass lists
    for the reference's own test matrix -- layouts, offsets, scales, storages, N-D, GLOBAL, Bluestein, REAL -- are
    replayed on the CPU, so a planner regression shows up without a GPU."""
    import grid_cases as grid

    out = []
    i = 0
    for suite, tps in grid.SUITES.items():
        for tp in tps:
            i += 1
            n = 1
            for l in tp.lengths:
                n *= l
            if i % 3 == 0 and n * tp.batch <= (1 << 18) and n <= (1 << 17):
                out.append(pytest.param(tp, id=f"{suite}-{tp.ident()}"))
    return out


@pytest.mark.parametrize("tp", _grid_sample())
def test_gpu_grid_sample_on_the_emulator(tp):
    test_exported_plan_matches_oracle(tp)


@pytest.mark.parametrize("tp", fuzz_cases(), ids=lambda tp: tp.ident())
def test_random_layouts_on_the_emulator(tp):
    """Seeded fuzz of the planner: random ranks, lengths (powers of two, mixed radix, primes), nested-and-padded
    layouts in both domains, offsets, scales, storage, precision, complex and REAL domain."""
    test_exported_plan_matches_oracle(tp)
