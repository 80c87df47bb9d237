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

BUF_IN, BUF_OUT, BUF_SCRATCH, BUF_SCRATCH2, BUF_SCRATCH3 = 0, 1, 2, 3, 4
KERNEL_EW, KERNEL_REAL_PACK, KERNEL_R2C_POST, KERNEL_C2R_PRE, KERNEL_REAL_UNPACK = 6, 7, 8, 9, 10
MOD_SWAP_PRE, MOD_SWAP_POST, MOD_NO_USER_SWAP_IN, MOD_NO_USER_SWAP_OUT = 1, 2, 4, 8


def _swap(a):
    return a.imag + 1j * a.real


class _PairView:
    """A REAL user buffer addressed as interleaved complex pairs (x[2j], x[2j+1]) -- what the complex passes of a
    REAL-domain plan see when the real rows have unit stride (csrc/real.cu)."""

    def __init__(self, real: np.ndarray):
        self.real = real

    def __getitem__(self, idx):
        return self.real[2 * idx] + 1j * self.real[2 * idx + 1]

    def __setitem__(self, idx, v):
        self.real[2 * idx] = v.real
        self.real[2 * idx + 1] = v.imag


def _real_rows(ps, src, dst, grids, ib, ob, cdt):
    """The four REAL-domain row kernels of csrc/real.cu, restated with the same index arithmetic."""
    n, variant, k = ps["n"], ps["variant"], ps["kernel"]
    h = n // 2
    w = np.exp(-2j * np.pi * np.arange(h + 1) / n)
    ibx, obx = ib[..., None], ob[..., None]
    if k == KERNEL_REAL_PACK:
        if variant == 0:
            m = np.arange(h)
            dst[obx + m] = (src[ibx + 2 * m * ps["is"]] + 1j * src[ibx + (2 * m + 1) * ps["is"]]).astype(cdt)
        else:
            m = np.arange(n)
            dst[obx + m] = src[ibx + m * ps["is"]].astype(cdt)
    elif k == KERNEL_REAL_UNPACK:
        s = ps["scale"] if ps["apply_scale"] else 1.0
        if variant == 0:
            m = np.arange(h)
            z = src[ibx + m]
            dst[obx + 2 * m * ps["os"]] = z.real * s
            dst[obx + (2 * m + 1) * ps["os"]] = z.imag * s
        else:
            m = np.arange(n)
            dst[obx + m * ps["os"]] = src[ibx + m].real * s
    elif k == KERNEL_R2C_POST:
        s = ps["scale"] if ps["apply_scale"] else 1.0
        kk = np.arange(h + 1)
        if variant == 0:
            a = src[ibx + np.where(kk == h, 0, kk)].astype(np.complex128)
            b = np.conj(src[ibx + np.where(kk == 0, 0, h - kk)].astype(np.complex128))
            x = (a + b) / 2 + w * (a - b) / 2j
        else:
            x = src[ibx + kk]
        dst[obx + kk * ps["os"]] = (x * s).astype(dst.dtype)
    elif k == KERNEL_C2R_PRE:
        if variant == 0:
            kk = np.arange(h)
            a = src[ibx + kk * ps["is"]].astype(np.complex128)
            b = src[ibx + (h - kk) * ps["is"]].astype(np.complex128)
            a[..., 0] = a[..., 0].real
            b[..., 0] = b[..., 0].real
            b = np.conj(b)
            z = (a + b) + 1j * np.conj(w[:h]) * (a - b)
            dst[obx + np.where(kk == 0, 0, h - kk)] = z.astype(cdt)
        else:
            kk = np.arange((n + 1) // 2)
            a = src[ibx + kk * ps["is"]].astype(np.complex128)
            a[..., 0] = a[..., 0].real
            dst[obx + kk] = np.conj(a).astype(cdt)
            dst[obx + (n - kk[1:])] = a[..., 1:].astype(cdt)


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
            BUF_SCRATCH2: np.full(max(1, plan["scratch2_elems"]), np.nan + 0j, dtype=cdt),
            BUF_SCRATCH3: np.full(max(1, plan["scratch3_elems"]), np.nan + 0j, dtype=cdt)}
    if plan["is_real"]:
        bwd = False  # REAL plans never use the (re <-> im) swap
    for ps in plan["passes"]:
        src, dst = bufs[ps["src"]], bufs[ps["dst"]]
        if ps["real_view"] & 1:
            src = _PairView(src)
        if ps["real_view"] & 2:
            dst = _PairView(dst)
        nb, ibd, obd = ps["nb"], ps["ibd"], ps["obd"]
        grids = np.meshgrid(*[np.arange(c, dtype=np.int64) for c in nb], indexing="ij")
        ib = ps["ioff"] + sum(g * d for g, d in zip(grids, ibd))
        ob = ps["ooff"] + sum(g * d for g, d in zip(grids, obd))
        internal = MOD_NO_USER_SWAP_IN | MOD_NO_USER_SWAP_OUT
        flags = ps["mod_flags"]
        swapped = bwd or bool(ps["force_swap"])
        swap_in = swapped and not (flags & MOD_NO_USER_SWAP_IN)
        swap_out = swapped and not (flags & MOD_NO_USER_SWAP_OUT)
        lmod = api.mod_table(scalar, ps["lmod"], ps["mod_l"], ps["mod_m"]) if ps["lmod"] else None
        smod = api.mod_table(scalar, ps["smod"], ps["mod_l"], ps["mod_m"]) if ps["smod"] else None
        n = ps["n"]
        assert ps["peer_dim"] < 0
        if KERNEL_REAL_PACK <= ps["kernel"] <= KERNEL_REAL_UNPACK:
            _real_rows(ps, src, dst, grids, ib, ob, cdt)
            continue
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
        if ps.get("fuse_real"):
            # REAL-domain pre / post-processing fused into the transform kernel (csrc/wg_cube.cu; formulas: real.cu);
            # n = half the real length
            h = n
            w = np.exp(-2j * np.pi * np.arange(h + 1) / (2 * h))
            s = ps["scale"] if ps["apply_scale"] else 1.0
            if ps["fuse_real"] == 1:
                z = np.fft.fft(np.asarray(src[ib[..., None] + jv * ps["is"]], dtype=dt), axis=-1)
                kk = np.arange(h + 1)
                a = z[..., np.where(kk == h, 0, kk)]
                b = np.conj(z[..., np.where(kk == 0, 0, h - kk)])
                dst[ob[..., None] + kk * ps["os"]] = (((a + b) / 2 + w * (a - b) / 2j) * s).astype(cdt)
            else:
                kk = np.arange(h)
                a = np.asarray(src[ib[..., None] + kk * ps["is"]], dtype=dt)
                b = np.asarray(src[ib[..., None] + (h - kk) * ps["is"]], dtype=dt)
                a[..., 0] = a[..., 0].real
                b[..., 0] = b[..., 0].real
                b = np.conj(b)
                zin = np.zeros(ib.shape + (h,), dtype=dt)
                zin[..., np.where(kk == 0, 0, h - kk)] = (a + b) + 1j * np.conj(w[:h]) * (a - b)
                dst[ob[..., None] + jv * ps["os"]] = np.fft.fft(zin, axis=-1) * s
            continue
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
        lin_keep = None
        if smod is not None and ps.get("smod_mask"):
            # table over the whole multi-pass transform: indexed by the position inside the packed power-of-two row
            y = y * smod[(ob[..., None] + jv * ps["os"]) & ps["smod_mask"]]
        elif smod is not None and ps.get("smod_n1"):
            # ... or by the linear output index (index along batch dimension 0) + n1 * k, truncated at valid_out
            lin = grids[0][..., None] + ps["smod_n1"] * jv
            lin_keep = lin < ps["valid_out"]
            y = y * smod[np.minimum(lin, len(smod) - 1)]
        elif smod is not None:
            y[..., :vo] = y[..., :vo] * smod[:vo]
        if flags & MOD_SWAP_POST:
            y = _swap(y)
        if ps["apply_scale"]:
            y = y * ps["scale"]
        if swap_out:
            y = _swap(y)
        if lin_keep is not None:
            addr = ob[..., None] + jv * ps["os"]
            dst[addr[lin_keep]] = y[lin_keep].astype(cdt)
        else:
            dst[ob[..., None] + jv[:vo] * ps["os"]] = y[..., :vo].astype(cdt)
    return out_buf
