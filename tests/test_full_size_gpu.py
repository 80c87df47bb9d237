"""Full-size parity of the BASELINE.json configurations that bench.py only checks through round trips:

* C5 (3-D fp32 512^3, default strides, out of place) against the oracle's generator on the whole array
  (`numpy.fft.fftn` in complex128 as /root/reference/test/common/reference_data_wrangler.hpp:117-145 does), both
  directions;
* C2 (1-D fp32 N=4096, batch 65536, in place) at the full batch: every 16th transform (4096 of them, spread over every
  CTA of the persistent grid and every stage of its ring) against `numpy.fft.fft` in complex128, forward and backward;
* C4 (1-D fp64 N=2^24, batch 8) at the full batch: transforms 0, 3 and 7 against numpy.

These are the sizes at which the persistent TMA-ring kernels are deep in their steady state."""
import numpy as np
import pytest

import portfft_oracle as oracle

pytestmark = pytest.mark.gpu


def _rel_l2_blocked(actual: np.ndarray, ref: np.ndarray, block: int = 1 << 24):
    """(relative L2, max |diff|) of two large 1-D complex arrays without materialising complex128 copies"""
    num = den = 0.0
    worst = 0.0
    a, r = actual.reshape(-1), ref.reshape(-1)
    for i in range(0, a.size, block):
        d = a[i:i + block].astype(np.complex128) - r[i:i + block]
        num += float(np.vdot(d, d).real)
        den += float(np.vdot(r[i:i + block], r[i:i + block]).real)
        worst = max(worst, float(np.max(np.abs(d))))
    return (num / den) ** 0.5, worst


def test_c5_full_size_against_numpy():
    import torch

    import portfft_b200 as pf

    n = 512
    x, ref = oracle.gen_data(1, [n, n, n], False)  # complex64 input, complex64 cast of the complex128 fftn
    d = pf.descriptor([n, n, n], "float")
    d.backward_scale = 1.0 / n ** 3
    c = d.commit(torch.cuda.current_stream(), 0)
    t_in = torch.from_numpy(x.reshape(-1)).cuda()
    t_out = torch.full((n ** 3,), complex(oracle.PADDING_VALUE, oracle.PADDING_VALUE), dtype=torch.complex64, device="cuda")
    c.compute_forward(t_in, t_out)
    torch.cuda.synchronize()
    err, worst = _rel_l2_blocked(t_out.cpu().numpy(), ref.reshape(-1))
    bound = oracle.rel_l2_bound(n ** 3, False)
    assert err <= bound, (err, bound)
    assert worst <= oracle.reference_elem_tolerance(n ** 3, False), worst
    # backward of the expected spectrum gives the input back (backward_scale = 1 / N)
    t_spec = torch.from_numpy(ref.reshape(-1)).cuda()
    c.compute_backward(t_spec, t_out)
    torch.cuda.synchronize()
    err, _ = _rel_l2_blocked(t_out.cpu().numpy(), x.reshape(-1))
    assert err <= bound, (err, bound)
    c.destroy()


def test_c2_full_batch_sampled_against_numpy():
    import torch

    import portfft_b200 as pf

    n, batch, every = 4096, 65536, 16
    d = pf.descriptor([n], "float")
    d.number_of_transforms = batch
    d.placement = pf.placement.IN_PLACE
    d.backward_scale = 1.0 / n
    c = d.commit(torch.cuda.current_stream(), 0)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    buf = torch.view_as_complex(torch.rand(batch, n, 2, device="cuda", generator=g) * 2 - 1)
    x = buf[::every].cpu().numpy()
    ref = np.fft.fft(x.astype(np.complex128), axis=1)
    c.compute_forward(buf.view(-1))
    torch.cuda.synchronize()
    got = buf[::every].cpu().numpy().astype(np.complex128)
    bound = oracle.rel_l2_bound(n, False)
    err = np.max(np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1))
    assert err <= bound, (err, bound)
    assert np.max(np.abs(got - ref)) <= oracle.reference_elem_tolerance(n, False)
    c.compute_backward(buf.view(-1))
    torch.cuda.synchronize()
    back = buf[::every].cpu().numpy()
    err = np.max(np.linalg.norm(back - x, axis=1) / np.linalg.norm(x, axis=1))
    assert err <= bound, (err, bound)
    c.destroy()


def test_c4_full_batch_sampled_against_numpy():
    import torch

    import portfft_b200 as pf

    n, batch = 1 << 24, 8
    d = pf.descriptor([n], "double")
    d.number_of_transforms = batch
    d.backward_scale = 1.0 / n
    c = d.commit(torch.cuda.current_stream(), 0)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    buf = torch.view_as_complex(torch.rand(batch, n, 2, device="cuda", generator=g, dtype=torch.float64) * 2 - 1)
    out = torch.empty_like(buf)
    c.compute_forward(buf.view(-1), out.view(-1))
    torch.cuda.synchronize()
    bound = oracle.rel_l2_bound(n, True)
    for b in (0, 3, 7):
        ref = np.fft.fft(buf[b].cpu().numpy())
        got = out[b].cpu().numpy()
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert err <= bound, (b, err, bound)
    back = torch.empty_like(buf)
    c.compute_backward(out.view(-1), back.view(-1))
    torch.cuda.synchronize()
    err = float(torch.linalg.vector_norm(back - buf) / torch.linalg.vector_norm(buf))
    assert err <= bound, (err, bound)
    c.destroy()
