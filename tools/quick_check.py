"""Ad-hoc GPU sanity run (development aid): a handful of parity cases, printed errors."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from fft_check import CaseParams, run_case, P, BI

cases = [
    CaseParams([8], 3), CaseParams([64], 1024), CaseParams([4096], 3, "IP"), CaseParams([4096], 64, "IP", dir="bwd"),
    CaseParams([1000], 7, storage="split"), CaseParams([2], 5), CaseParams([1], 5), CaseParams([3], 33000),
    CaseParams([96], 555, "OOP", P, BI), CaseParams([256], 131, "OOP", BI, P, storage="split"),
    CaseParams([2048], 3, "IP", BI, BI), CaseParams([8192], 3), CaseParams([16384], 3), CaseParams([65536], 3),
    CaseParams([68640], 3), CaseParams([9800], 3, dir="bwd"), CaseParams([16, 512], 3), CaseParams([2, 3, 2, 3], 3),
    CaseParams([64, 64, 64], 1), CaseParams([4096], 8, scalar="double"), CaseParams([1 << 18], 2, scalar="double"),
    CaseParams([31], 100), CaseParams([1024], 4, forward_scale=2.0, backward_scale=-1.0),
    CaseParams([2048], 33, forward_offset=2047, backward_offset=2049),
    CaseParams([85], 13, "IP", forward_strides=[13], backward_strides=[13], forward_distance=12, backward_distance=12),
]
fails = 0
for tp in cases:
    t0 = time.time()
    try:
        e = run_case(tp)
        print(f"OK   {tp.ident():70s} relL2={e:.2e} ({time.time()-t0:.2f}s)", flush=True)
    except Exception as ex:
        fails += 1
        print(f"FAIL {tp.ident():70s} {type(ex).__name__}: {str(ex)[:200]}", flush=True)
print("failures:", fails)
sys.exit(1 if fails else 0)
