"""Parity harness: the pytest counterpart of the reference's `check_fft` / `run_test`
(/root/reference/test/unit_test/fft_test_utils.hpp:276-347,436-479).

Builds a descriptor from test parameters exactly as `get_descriptor` does (:211-260), gets input and expected
output from the numpy oracle (oracle/portfft_oracle.py, the restatement of reference_data_wrangler.hpp), runs the
CUDA library through the C ABI on device buffers pre-filled with the padding value -5 (:452) and verifies with
`verify_dft` (exact prefix / untouched padding, element tolerance 2*eps*N*log2N, and the north_star relative-L2
bound 1e-5*log2(N) fp32 / 1e-13*log2(N) fp64).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

import portfft_oracle as oracle
import portfft_b200 as pf

P, BI, U = "PACKED", "BATCH_INTERLEAVED", "UNPACKED"


@dataclass
class CaseParams:
    lengths: Sequence[int]
    batch: int = 1
    placement: str = "OOP"          # "IP" / "OOP"
    input_layout: str = P
    output_layout: str = P
    dir: str = "fwd"                # "fwd" / "bwd"
    storage: str = "interleaved"    # "interleaved" / "split"
    scalar: str = "float"
    forward_scale: Optional[float] = None
    backward_scale: Optional[float] = None
    forward_offset: Optional[int] = None
    backward_offset: Optional[int] = None
    # explicit layout (layout_params, fft_test_utils.hpp:52-78)
    forward_strides: Optional[List[int]] = None
    backward_strides: Optional[List[int]] = None
    forward_distance: Optional[int] = None
    backward_distance: Optional[int] = None
    domain: str = "complex"         # "complex" / "real" (REAL: forward domain real, backward domain half spectrum)

    def ident(self) -> str:
        s = f"{'re-' if self.domain == 'real' else ''}{self.scalar}-{self.placement}-{self.input_layout[:2]}to{self.output_layout[:2]}-{self.dir}-{self.storage[:5]}"
        s += f"-b{self.batch}-n" + "x".join(str(x) for x in self.lengths)
        if self.forward_strides is not None:
            s += "-fs" + "_".join(map(str, self.forward_strides)) + "-bs" + "_".join(map(str, self.backward_strides))
            s += f"-fd{self.forward_distance}-bd{self.backward_distance}"
        if self.forward_scale is not None:
            s += f"-fsc{self.forward_scale}-bsc{self.backward_scale}"
        if self.forward_offset is not None:
            s += f"-fo{self.forward_offset}-bo{self.backward_offset}"
        return s


def make_descriptors(tp: CaseParams):
    """-> (portfft_b200.descriptor, oracle.OracleDescriptor), fields set as `get_descriptor` does."""
    d = pf.descriptor(list(tp.lengths), tp.scalar, pf.domain.REAL if tp.domain == "real" else pf.domain.COMPLEX)
    d.number_of_transforms = tp.batch
    d.placement = pf.placement.IN_PLACE if tp.placement == "IP" else pf.placement.OUT_OF_PLACE
    d.complex_storage = (pf.complex_storage.INTERLEAVED_COMPLEX if tp.storage == "interleaved"
                         else pf.complex_storage.SPLIT_COMPLEX)
    fdir = pf.direction.FORWARD if tp.dir == "fwd" else pf.direction.BACKWARD

    def apply_layout(lay, dr):
        if lay == BI:
            if dr == pf.direction.FORWARD:
                d.forward_strides, d.forward_distance = [tp.batch], 1
            else:
                d.backward_strides, d.backward_distance = [tp.batch], 1

    apply_layout(tp.input_layout, fdir)
    apply_layout(tp.output_layout, pf.inv(fdir))
    if tp.forward_scale is not None:
        d.forward_scale = tp.forward_scale
    if tp.backward_scale is not None:
        d.backward_scale = tp.backward_scale
    if tp.forward_offset is not None:
        d.forward_offset = tp.forward_offset
    if tp.backward_offset is not None:
        d.backward_offset = tp.backward_offset
    if tp.forward_strides is not None:
        d.forward_strides = list(tp.forward_strides)
        d.backward_strides = list(tp.backward_strides)
        fd, bd = tp.forward_distance, tp.backward_distance
        if fd is None:  # layout_params 3-argument ctor (fft_test_utils.hpp:57-68)
            fd = bd = 1
            for n, fs, bs in zip(tp.lengths, tp.forward_strides, tp.backward_strides):
                fd *= n * fs
                bd *= n * bs
        d.forward_distance, d.backward_distance = fd, bd
    od = oracle.OracleDescriptor(
        lengths=list(d.lengths), forward_scale=d.forward_scale, backward_scale=d.backward_scale,
        number_of_transforms=d.number_of_transforms, complex_storage=int(d.complex_storage),
        placement=int(d.placement), forward_strides=list(d.forward_strides),
        backward_strides=list(d.backward_strides), forward_distance=d.forward_distance,
        backward_distance=d.backward_distance, forward_offset=d.forward_offset, backward_offset=d.backward_offset,
        is_double=(tp.scalar == "double"), is_real=(tp.domain == "real"))
    return d, od


def run_case(tp: CaseParams, device: int = 0, rel_l2_tol: Optional[float] = None) -> float:
    """Run one parity case on the GPU; returns the max relative L2 error.  Raises AssertionError on mismatch."""
    import torch

    d, od = make_descriptors(tp)
    dr = oracle.FORWARD if tp.dir == "fwd" else oracle.BACKWARD
    host_in, host_ref = oracle.expected_io(od, dr)
    cdt = torch.complex128 if tp.scalar == "double" else torch.complex64
    dev = torch.device("cuda", device)
    in_place = tp.placement == "IP"
    split = tp.storage == "split"
    committed = d.commit(torch.cuda.current_stream(dev), device)
    fn = committed.compute_forward if tp.dir == "fwd" else committed.compute_backward
    pad = oracle.PADDING_VALUE
    if tp.domain == "real":
        # forward: real array -> half spectrum; backward: half spectrum -> real array (out of place)
        assert not in_place
        fwd = tp.dir == "fwd"
        if fwd or not split:
            args_in = [torch.from_numpy(host_in).to(dev)]
        else:
            args_in = [torch.from_numpy(np.ascontiguousarray(host_in.real)).to(dev),
                       torch.from_numpy(np.ascontiguousarray(host_in.imag)).to(dev)]
        n_out = host_ref.shape[0]
        rdt = torch.float64 if tp.scalar == "double" else torch.float32
        if not fwd:
            outs = [torch.full((n_out,), pad, dtype=rdt, device=dev)]
        elif split:
            outs = [torch.full((n_out,), pad, dtype=rdt, device=dev), torch.full((n_out,), pad, dtype=rdt, device=dev)]
        else:
            outs = [torch.full((n_out,), complex(pad, pad), dtype=cdt, device=dev)]
        fn(*args_in, *outs)
        torch.cuda.synchronize(dev)
        if fwd and split:
            actual = (outs[0].cpu().numpy() + 1j * outs[1].cpu().numpy()).astype(host_ref.dtype)
        else:
            actual = outs[0].cpu().numpy()
        committed.destroy()
        return oracle.verify_dft(od, dr, host_ref, actual, rel_l2_tol)
    if in_place:
        assert host_in.shape == host_ref.shape
    if not split:
        t_in = torch.from_numpy(host_in).to(dev)
        if in_place:
            fn(t_in)
            t_out = t_in
        else:
            t_out = torch.full((host_ref.shape[0],), complex(pad, pad), dtype=cdt, device=dev)
            fn(t_in, t_out)
        torch.cuda.synchronize(dev)
        actual = t_out.cpu().numpy()
    else:
        in_re = torch.from_numpy(np.ascontiguousarray(host_in.real)).to(dev)
        in_im = torch.from_numpy(np.ascontiguousarray(host_in.imag)).to(dev)
        if in_place:
            fn(in_re, in_im)
            out_re, out_im = in_re, in_im
        else:
            out_re = torch.full((host_ref.shape[0],), pad, dtype=in_re.dtype, device=dev)
            out_im = torch.full((host_ref.shape[0],), pad, dtype=in_re.dtype, device=dev)
            fn(in_re, in_im, out_re, out_im)
        torch.cuda.synchronize(dev)
        actual = (out_re.cpu().numpy() + 1j * out_im.cpu().numpy()).astype(host_ref.dtype)
    committed.destroy()
    return oracle.verify_dft(od, dr, host_ref, actual, rel_l2_tol)


def run_case_host(tp: CaseParams, device: int = 0) -> float:
    """Same parity case through the end-to-end entry point `pfft_compute_host` (HOST buffers, H2D/D2H inside the
    call; chunk-pipelined when the layout is batch-major).  The output buffer is pre-filled with the padding value,
    which must survive wherever the descriptor addresses nothing."""
    import torch

    d, od = make_descriptors(tp)
    dr = oracle.FORWARD if tp.dir == "fwd" else oracle.BACKWARD
    host_in, host_ref = oracle.expected_io(od, dr)
    in_place = tp.placement == "IP"
    split = tp.storage == "split"
    committed = d.commit(torch.cuda.current_stream(torch.device("cuda", device)), device)
    pdir = pf.direction.FORWARD if tp.dir == "fwd" else pf.direction.BACKWARD
    pad = oracle.PADDING_VALUE
    if tp.domain == "real":
        # the real side of the call is one scalar array; the complex side follows the descriptor's storage
        assert not in_place
        fwd = tp.dir == "fwd"
        if fwd:
            h_in = np.ascontiguousarray(host_in)
            if split:
                out_re = np.full(host_ref.shape, pad, dtype=h_in.dtype)
                out_im = np.full(host_ref.shape, pad, dtype=h_in.dtype)
                committed.compute_host(pdir, h_in, None, out_re, out_im)
                actual = (out_re + 1j * out_im).astype(host_ref.dtype)
            else:
                actual = np.full(host_ref.shape, complex(pad, pad), dtype=host_ref.dtype)
                committed.compute_host(pdir, h_in, None, actual, None)
        else:
            actual = np.full(host_ref.shape, pad, dtype=host_ref.dtype)
            if split:
                in_re, in_im = np.ascontiguousarray(host_in.real), np.ascontiguousarray(host_in.imag)
                committed.compute_host(pdir, in_re, in_im, actual, None)
            else:
                committed.compute_host(pdir, np.ascontiguousarray(host_in), None, actual, None)
        committed.destroy()
        return oracle.verify_dft(od, dr, host_ref, actual)
    if not split:
        h_in = np.ascontiguousarray(host_in)
        h_out = h_in if in_place else np.full(host_ref.shape, complex(pad, pad), dtype=host_ref.dtype)
        committed.compute_host(pdir, h_in, None, h_out, None)
        actual = h_out
    else:
        in_re, in_im = np.ascontiguousarray(host_in.real), np.ascontiguousarray(host_in.imag)
        if in_place:
            out_re, out_im = in_re, in_im
        else:
            out_re = np.full(host_ref.shape, pad, dtype=in_re.dtype)
            out_im = np.full(host_ref.shape, pad, dtype=in_re.dtype)
        committed.compute_host(pdir, in_re, in_im, out_re, out_im)
        actual = (out_re + 1j * out_im).astype(host_ref.dtype)
    committed.destroy()
    return oracle.verify_dft(od, dr, host_ref, actual)
