#!/usr/bin/env python
"""Parity cases for compute-sanitizer (racecheck / memcheck / synccheck): every persistent TMA-ring kernel with more
tiles than its grid holds CTAs x ring stages, small enough to finish under the tool's 10-100x slowdown.  Each case is
also compared with the numpy oracle (tests/fft_check.run_case).

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [name ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from fft_check import BI, P, U, CaseParams, run_case  # noqa: E402


def _real(n, batch, direction, scalar="float"):
    """REAL domain, packed rows (the layout of the reference's real benchmark set)."""
    return CaseParams([n], batch, "OOP", U, U, direction, "interleaved", scalar, forward_strides=[1], backward_strides=[1],
                      forward_distance=n, backward_distance=n // 2 + 1, domain="real",
                      backward_scale=(1.0 / n if direction == "bwd" else None))


CASES = {
    # wg_cube<float,16,1> (C2's kernel): 2 CTAs x 148 SMs x 2 stages = 592 tiles in flight
    "cube4096_ip": CaseParams([4096], 1300, "IP", P, P, "fwd", "interleaved", "float"),
    "cube4096_bwd": CaseParams([4096], 700, "OOP", P, P, "bwd", "interleaved", "float"),
    # wg_cube<float,8,4> (N = 512 rows, C5's z pass): 4 CTAs x 148 x 2 stages x 4 transforms
    "cube512": CaseParams([512], 10000, "OOP", P, P, "fwd", "interleaved", "float"),
    # wg_rows3 (1024 / 2048 / 8192)
    "rows1024": CaseParams([1024], 5000, "IP", P, P, "fwd", "interleaved", "float"),
    "rows2048": CaseParams([2048], 2500, "OOP", P, P, "bwd", "interleaved", "float"),
    "rows8192": CaseParams([8192], 700, "OOP", P, P, "fwd", "interleaved", "float"),
    # fp64 variants of the same kernels
    "cube4096_f64": CaseParams([4096], 700, "OOP", P, P, "fwd", "interleaved", "double"),
    "cube512_f64": CaseParams([512], 5000, "IP", P, P, "fwd", "interleaved", "double"),
    "rows2048_f64": CaseParams([2048], 1300, "OOP", P, P, "fwd", "interleaved", "double"),
    # wg_col in place (N = 256 columns, three CTAs per SM x 2 stages): 256 x 16384 = 1024 tiles ... and rows -> columns
    "col256_inplace": CaseParams([256, 16384], 2, "IP", P, P, "fwd", "interleaved", "float"),
    "global65536": CaseParams([65536], 64, "OOP", P, P, "fwd", "interleaved", "float"),
    "global65536_f64": CaseParams([65536], 32, "OOP", P, P, "bwd", "interleaved", "double"),
    # wg_col512 (two groups, three-stage ring, full / freed barriers): 512 x 16384 = 1024 tiles per transform
    "col512": CaseParams([512, 16384], 2, "IP", P, P, "fwd", "interleaved", "float"),
    "col512_bi": CaseParams([512], 30000, "OOP", BI, BI, "bwd", "interleaved", "float"),
    # wi_tma (thread-level, TMA tiles in and out, three stages): 128 lines per tile, 4 x 148 CTAs
    "wi_tma16": CaseParams([16], 400000, "OOP", P, P, "fwd", "interleaved", "float"),
    "wi_tma8_f64": CaseParams([8], 300000, "IP", P, P, "bwd", "interleaved", "double"),
    # three-radix column tiles (wg_colr3.cu: tensor loads / stores, ring of three tile buffers; C3B's kernel):
    # 3001 columns = 376 tiles of 8 on <= 148 CTAs, ragged last tile; the register form on an unaligned layout
    "colr3_1000_split": CaseParams([1000], 3001, "OOP", BI, BI, "fwd", "split", "float"),
    "colr3_1000_inter": CaseParams([1000], 3001, "IP", BI, BI, "bwd", "interleaved", "float"),
    "colr3_1024_f64": CaseParams([1024], 1501, "OOP", BI, BI, "fwd", "split", "double"),
    "colr3_regs": CaseParams([1000], 1203, "OOP", BI, BI, "fwd", "split", "float"),
    # REAL domain fused into the tile kernels (wg_cube.cu REAL = 1 / 2): odd and even row starts, last row even
    "r2c_8192": _real(8192, 1301, "fwd"),
    "c2r_8192": _real(8192, 1301, "bwd"),
    "r2c_512": _real(512, 10001, "fwd"),
    "c2r_512": _real(512, 10001, "bwd"),
    "c2r_2048_f64": _real(2048, 1301, "bwd", "double"),
    # Bluestein, multi-pass form: the column kernel's store with a table over the whole transform (position index,
    # linear index with truncation at L and clamped look-ups), user layout written by the last pass
    "bluestein_4099": CaseParams([4099], 67, "OOP", P, P, "fwd", "interleaved", "float"),
    "bluestein_4099_bwd_f64": CaseParams([4099], 33, "OOP", P, P, "bwd", "interleaved", "double"),
    "bluestein_65537": CaseParams([65537], 3, "OOP", P, P, "fwd", "interleaved", "float"),
    # fused two-pass kernel (opt-in; PFFT_FUSE=1 is set below for this case only)
    "fused65536": CaseParams([65536], 24, "OOP", P, P, "fwd", "interleaved", "float"),
}


def main():
    names = sys.argv[1:] or list(CASES)
    for n in names:
        if n.startswith("fused"):
            os.environ["PFFT_FUSE"] = "1"
            os.environ["PFFT_FUSE_CHUNK_KB"] = "512"
        err = run_case(CASES[n])
        os.environ.pop("PFFT_FUSE", None)
        print(f"{n:18s} ok rel_l2={err:.2e}", flush=True)


if __name__ == "__main__":
    main()
