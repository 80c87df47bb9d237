"""GPU parity tests of the end-to-end entry point pfft_compute_host (the call bench.py's `e2e` figure times):
host buffers in, host buffers out, chunk-pipelined H2D / kernels / D2H.  PFFT_HOST_CHUNK_BYTES is lowered so that
small cases exercise many chunks, ragged last chunks and the monolithic fallback (batch-interleaved layouts)."""
import os

import pytest

from fft_check import BI, P, U, CaseParams, run_case_host

pytestmark = pytest.mark.gpu

CASES = [
    CaseParams([4096], 37, "IP", P, P, "fwd", "interleaved", "float"),
    CaseParams([4096], 37, "IP", P, P, "bwd", "interleaved", "float", backward_scale=1.0 / 4096),
    CaseParams([4096], 64, "OOP", P, P, "fwd", "interleaved", "double"),
    CaseParams([1000], 101, "OOP", P, P, "fwd", "split", "float"),
    CaseParams([1000], 101, "OOP", U, U, "bwd", "split", "float", forward_strides=[2], backward_strides=[1],
               forward_distance=2048, backward_distance=1024, forward_offset=7, backward_offset=3,
               backward_scale=1e-3),
    CaseParams([1000], 101, "OOP", U, U, "fwd", "split", "float", forward_strides=[2], backward_strides=[1],
               forward_distance=2048, backward_distance=1024, forward_offset=7, backward_offset=3),
    CaseParams([256], 555, "OOP", P, BI, "fwd", "interleaved", "float"),      # not batch-major: monolithic path
    CaseParams([256], 555, "IP", BI, BI, "fwd", "split", "double"),
    CaseParams([16, 512], 9, "OOP", P, P, "fwd", "interleaved", "float"),
    CaseParams([65536], 5, "OOP", P, P, "fwd", "interleaved", "float"),      # GLOBAL level sub-plans own scratch
    CaseParams([2048], 33, "OOP", P, P, "bwd", "interleaved", "float", forward_offset=2047, backward_offset=2049),
    CaseParams([8], 1, "OOP", P, P, "fwd", "interleaved", "float"),
    # lengths with a large prime factor and the REAL domain through the host entry point
    CaseParams([1031], 9, "OOP", P, P, "fwd", "interleaved", "float"),
    CaseParams([4096], 21, "OOP", P, P, "fwd", "interleaved", "float", domain="real"),
    CaseParams([4096], 21, "OOP", P, P, "bwd", "split", "float", domain="real", backward_scale=1.0 / 4096),
    CaseParams([30], 77, "OOP", P, P, "fwd", "split", "double", domain="real"),
    CaseParams([6, 16], 5, "OOP", P, P, "bwd", "interleaved", "double", domain="real"),
    # REAL through the chunk pipeline: dense half spectrum (rows of n / 2 + 1: the fused kernels, 8-byte aligned chunk
    # starts), split half spectrum (one real plane in, two planes out)
    CaseParams([8192], 301, "OOP", U, U, "fwd", "interleaved", "float", forward_strides=[1], backward_strides=[1],
               forward_distance=8192, backward_distance=4097, domain="real"),
    CaseParams([8192], 301, "OOP", U, U, "bwd", "interleaved", "float", forward_strides=[1], backward_strides=[1],
               forward_distance=8192, backward_distance=4097, domain="real", backward_scale=1.0 / 8192),
    CaseParams([512], 1001, "OOP", U, U, "fwd", "split", "float", forward_strides=[1], backward_strides=[1],
               forward_distance=512, backward_distance=257, domain="real"),
    CaseParams([512], 1001, "OOP", U, U, "bwd", "split", "double", forward_strides=[1], backward_strides=[1],
               forward_distance=512, backward_distance=257, domain="real", backward_scale=1.0 / 512),
]


@pytest.mark.parametrize("chunk_bytes", [1 << 16, 1 << 20])
@pytest.mark.parametrize("tp", CASES, ids=[c.ident() for c in CASES])
def test_compute_host_matches_oracle(tp, chunk_bytes, monkeypatch):
    monkeypatch.setenv("PFFT_HOST_CHUNK_BYTES", str(chunk_bytes))
    run_case_host(tp)
