#!/usr/bin/env python
"""Run one REAL-domain transform a few times (profiling aid): tools/run_real.py N BATCH [fwd|bwd] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import portfft_b200 as pf  # noqa: E402

n, batch = int(sys.argv[1]), int(sys.argv[2])
direction = sys.argv[3] if len(sys.argv) > 3 else "fwd"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
d = pf.descriptor([n], "float", pf.domain.REAL)
d.number_of_transforms = batch
plan = d.commit(torch.cuda.current_stream(), 0)
x = torch.rand(batch * n, device="cuda")
y = torch.zeros(batch * n, dtype=torch.complex64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps + 2):
    if i == 2:
        e0.record()
    if direction == "fwd":
        plan.compute_forward(x, y)
    else:
        plan.compute_backward(y, x)
e1.record()
torch.cuda.synchronize()
print(f"real n={n} batch={batch} {direction}: {e0.elapsed_time(e1) / reps:.4f} ms per call, launches {plan.num_launches()}")
