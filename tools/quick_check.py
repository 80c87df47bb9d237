"""Ad-hoc GPU sanity run (development aid): a handful of parity cases, printed errors."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from fft_check import TestParams, run_case, P, BI

cases = [
    TestParams([8], 3), TestParams([64], 1024), TestParams([4096], 3, "IP"), TestParams([4096], 64, "IP", dir="bwd"),
    TestParams([1000], 7, storage="split"), TestParams([2], 5), TestParams([1], 5), TestParams([3], 33000),
    TestParams([96], 555, "OOP", P, BI), TestParams([256], 131, "OOP", BI, P, storage="split"),
    TestParams([2048], 3, "IP", BI, BI), TestParams([8192], 3), TestParams([16384], 3), TestParams([65536], 3),
    TestParams([68640], 3), TestParams([9800], 3, dir="bwd"), TestParams([16, 512], 3), TestParams([2, 3, 2, 3], 3),
    TestParams([64, 64, 64], 1), TestParams([4096], 8, scalar="double"), TestParams([1 << 18], 2, scalar="double"),
    TestParams([31], 100), TestParams([1024], 4, forward_scale=2.0, backward_scale=-1.0),
    TestParams([2048], 33, forward_offset=2047, backward_offset=2049),
    TestParams([85], 13, "IP", forward_strides=[13], backward_strides=[13], forward_distance=12, backward_distance=12),
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
