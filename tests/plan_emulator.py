"""numpy evaluation of a PLAN (the pass list `pfft_plan_export` returns): every kernel launch of csrc/pass.h replayed on
the host with the same address arithmetic, inter-factor twiddles, element-wise modifiers, (re <-> im) swaps and scale,
the per-pass DFT itself taken from numpy.  Test infrastructure only: it checks the planner (pass geometry of the
GLOBAL level, N-D passes, Bluestein, real-domain pre/post passes) on a machine without a GPU; the kernels are checked
against the oracle on the GPU by tests/test_fft_gpu.py.

    out[ooff + sum_d b_d*obd[d] + k*os] = scale * M_store(gtw(b,k) * sum_j M_load(in[ioff + sum_d b_d*ibd[d] + j*is]) w_n^{jk})
"""
from __future__ import annotations

import numpy as np

import portfft_b200 as pf
from portfft_b200 import api

BUF_IN, BUF_OUT, BUF_SCRATCH, BUF_SCRATCH2 = 0, 1, 2, 3
KERNEL_EW = 6
MOD_SWAP_PRE, MOD_SWAP_POST, MOD_NO_USER_SWAP_IN, MOD_NO_USER_SWAP_OUT = 1, 2, 4, 8


def _swap(a):
    return a.imag + 1j * a.real


def run_plan(desc: "pf.descriptor", direction, in_buf: np.ndarray, out_buf: np.ndarray) -> np.ndarray:
    """Run the exported pass list of `desc` in `direction` on flat complex host arrays (interleaved view of the data;
    split storage only differs in how the kernels address memory).  Returns out_buf (modified in place).
    in_buf may be out_buf (in-place)."""
    plan = desc.export_plan(direction)
    dt = np.complex128
    cdt = np.complex128 if plan["is_double"] else np.complex64
    scalar = "double" if plan["is_double"] else "float"
    bwd = int(direction) == 1
    bufs = {BUF_IN: in_buf, BUF_OUT: out_buf,
            BUF_SCRATCH: np.full(max(1, plan["scratch_elems"]), np.nan + 0j, dtype=cdt),
            BUF_SCRATCH2: np.full(max(1, plan["scratch2_elems"]), np.nan + 0j, dtype=cdt)}
    for ps in plan["passes"]:
        src, dst = bufs[ps["src"]], bufs[ps["dst"]]
        nb, ibd, obd = ps["nb"], ps["ibd"], ps["obd"]
        grids = np.meshgrid(*[np.arange(c, dtype=np.int64) for c in nb], indexing="ij")
        ib = ps["ioff"] + sum(g * d for g, d in zip(grids, ibd))
        ob = ps["ooff"] + sum(g * d for g, d in zip(grids, obd))
        internal = MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT
        flags = ps["mod_flags"]
        swap_in = bwd and not (flags & MOD_NO_USER_SWAP_IN)
        swap_out = bwd and not (flags & MOD_NO_USER_SWAP_OUT)
        lmod = api.mod_table(scalar, ps["lmod"], ps["mod_l"], ps["mod_m"]) if ps["lmod"] else None
        smod = api.mod_table(scalar, ps["smod"], ps["mod_l"], ps["mod_m"]) if ps["smod"] else None
        n = ps["n"]
        assert ps["peer_dim"] < 0
        if ps["kernel"] == KERNEL_EW:
            assert n == 1
            j = grids[0]
            vi = ps["valid_in"] or nb[0]
            vo = ps["valid_out"] or nb[0]
            live_in = j < vi
            v = np.zeros(j.shape, dtype=dt)
            v[live_in] = src[ib[live_in]]
            if swap_in:
                v = _swap(v)
            if lmod is not None:
                v[live_in] = v[live_in] * lmod[j[live_in]]
            if flags & MOD_SWAP_PRE:
                v = _swap(v)
            if smod is not None:
                jj = np.minimum(j, len(smod) - 1)
                v = v * smod[jj]
            if flags & MOD_SWAP_POST:
                v = _swap(v)
            if ps["apply_scale"]:
                v = v * ps["scale"]
            if swap_out:
                v = _swap(v)
            live_out = j < vo
            dst[ob[live_out]] = v[live_out].astype(cdt)
            continue
        jv = np.arange(n, dtype=np.int64)
        vi = ps["valid_in"] or n
        vo = ps["valid_out"] or n
        x = np.zeros(ib.shape + (n,), dtype=dt)
        x[..., :vi] = src[ib[..., None] + jv[:vi] * ps["is"]]
        assert not np.isnan(x).any(), "pass reads workspace that no earlier pass wrote"
        if swap_in:
            x = _swap(x)
        if lmod is not None:
            x[..., :vi] = x[..., :vi] * lmod[:vi]
        y = np.fft.fft(x, axis=-1)
        if ps["gtw_dim"] >= 0:
            c = grids[ps["gtw_dim"]]
            m = (c[..., None] * jv) % ps["gtw_n"]
            y = y * np.exp(-2j * np.pi * m / ps["gtw_n"])
        if flags & MOD_SWAP_PRE:
            y = _swap(y)
        if smod is not None:
            y[..., :vo] = y[..., :vo] * smod[:vo]
        if flags & MOD_SWAP_POST:
            y = _swap(y)
        if ps["apply_scale"]:
            y = y * ps["scale"]
        if swap_out:
            y = _swap(y)
        dst[ob[..., None] + jv[:vo] * ps["os"]] = y[..., :vo].astype(cdt)
    return out_buf
