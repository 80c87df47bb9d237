"""REAL-domain transforms whose pre / post-processing is fused into the TMA tile kernel (csrc/wg_cube.cu): buffers the
fused kernels cannot take (not 16-byte aligned) must run the unfused plan with identical results; expected values =
numpy.fft.rfft / irfft as in the reference's generator (test/common/reference_data_wrangler.hpp:136-137)."""
import numpy as np
import pytest

import portfft_b200 as pf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1024, 8192])
@pytest.mark.parametrize("shift_in,shift_out", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_fused_and_unfused_agree(n, shift_in, shift_out):
    import torch

    batch, h = 37, n // 2
    rng = np.random.Generator(np.random.SFC64(0))
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    d = pf.descriptor([n], "float", pf.domain.REAL)
    d.number_of_transforms = batch
    d.backward_distance = h + 1
    d.backward_scale = 1.0 / n
    plan = d.commit(torch.cuda.current_stream(), 0)
    assert plan.num_launches(pf.direction.FORWARD) == 1 and plan.num_launches(pf.direction.BACKWARD) == 1
    # buffers shifted by 8 bytes (two floats / one complex): still valid for the descriptor, not for cp.async.bulk
    xin = torch.zeros(batch * n + 2, dtype=torch.float32, device="cuda")
    xin[2 * shift_in:2 * shift_in + batch * n] = torch.from_numpy(x.reshape(-1)).cuda()
    spec = torch.zeros(batch * (h + 1) + 1, dtype=torch.complex64, device="cuda")
    xi = xin[2 * shift_in:2 * shift_in + batch * n]
    sp = spec[shift_out:shift_out + batch * (h + 1)]
    plan.compute_forward(xi, sp)
    torch.cuda.synchronize()
    want = np.fft.rfft(x.astype(np.float64), axis=-1)
    got = sp.cpu().numpy().reshape(batch, h + 1)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5 * np.log2(n)
    back = torch.zeros(batch * n + 2, dtype=torch.float32, device="cuda")
    bo = back[2 * shift_in:2 * shift_in + batch * n]
    plan.compute_backward(sp, bo)
    torch.cuda.synchronize()
    assert np.linalg.norm(bo.cpu().numpy().reshape(batch, n) - x) / np.linalg.norm(x) < 2e-5 * np.log2(n)
    # nothing outside the addressed elements was written
    assert float(back[:2 * shift_in].abs().sum()) == 0 and float(back[2 * shift_in + batch * n:].abs().sum()) == 0
    plan.destroy()
