"""Generates tests/golden/*.npz (committed): small fixtures that travel to the GPU box, where /root/reference and
oracle/_ref's sources do not exist.

  numpy_*.npz  -- the reference test generator's input/expected pairs (reference_data_wrangler.hpp:117-145): protects
                  the oracle against numpy version drift between this container and the GPU box.
  numpyre_*.npz -- the same for the generator's REAL-domain branch (real input, rfftn output).
  refcode_*.npz -- outputs of the REFERENCE's own wi_dft / sg_dft / wg_dft (compiled from /root/reference through
                  oracle/ref_shim into oracle/_ref) on the same SFC64(0) inputs.

Run in the authoring container:  make -C oracle ref && python tools/make_golden.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import portfft_oracle as o  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

for dbl in (False, True):
    tag = "f64" if dbl else "f32"
    for batch, dims in [(3, [8]), (2, [64]), (1, [1000]), (2, [4096]), (1, [2, 3, 6]), (1, [16, 32]), (1, [1031])]:
        x, y = o.gen_data(batch, dims, dbl)
        name = f"numpy_{tag}_b{batch}_n" + "x".join(map(str, dims)) + ".npz"
        np.savez_compressed(os.path.join(OUT, name), input=x, output=y)
    # REAL domain: the generator's `is_complex = False` branch (real input, rfftn output; :130-137)
    for batch, dims in [(2, [64]), (1, [30]), (1, [81]), (1, [4, 6])]:
        x, y = o.gen_data(batch, dims, dbl, is_real=True)
        name = f"numpyre_{tag}_b{batch}_n" + "x".join(map(str, dims)) + ".npz"
        np.savez_compressed(os.path.join(OUT, name), input=x, output=y)

ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libportfft_ref.so"))
for dbl in (False, True):
    tag = "f64" if dbl else "f32"
    wi = ref.ref_wi_dft_f64 if dbl else ref.ref_wi_dft_f32
    wi.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    sg = ref.ref_sg_dft_f64 if dbl else ref.ref_sg_dft_f32
    sg.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    ins, outs, sizes, levels = [], [], [], []
    for n in [2, 3, 4, 5, 7, 8, 9, 11, 13, 16, 25, 31, 32, 64, 75, 96, 100, 128, 256, 512]:
        x, _ = o.gen_data(1, [n], dbl)
        out = np.empty_like(x)
        if o.ref_fits_in_wi(n, dbl):
            wi(x.ctypes.data, out.ctypes.data, n)
            lvl = 0
        elif o.ref_fits_in_sg(n, 32, dbl):
            f_sg = o.ref_factorize_sg(n, 32)
            sg(x.ctypes.data, out.ctypes.data, n // f_sg, f_sg)
            lvl = 1
        else:
            continue
        ins.append(x.reshape(-1))
        outs.append(out.reshape(-1))
        sizes.append(n)
        levels.append(lvl)
    np.savez_compressed(os.path.join(OUT, f"refcode_{tag}.npz"), sizes=np.array(sizes), levels=np.array(levels),
                        inputs=np.concatenate(ins), outputs=np.concatenate(outs))
# the reference's WORKGROUP level (wg_dft on an emulated work-group): the lengths of BASELINE configs C2 and C3 among them
for dbl in (False, True):
    tag = "f64" if dbl else "f32"
    wg = ref.ref_wg_dft_f64 if dbl else ref.ref_wg_dft_f32
    wg.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    ins, outs, sizes = [], [], []
    for n in [1000, 1024, 2048, 3072, 4096]:
        x, _ = o.gen_data(1, [n], dbl)
        out = np.empty_like(x)
        wg(x.ctypes.data, out.ctypes.data, n)
        ins.append(x.reshape(-1))
        outs.append(out.reshape(-1))
        sizes.append(n)
    np.savez_compressed(os.path.join(OUT, f"refcode_wg_{tag}.npz"), sizes=np.array(sizes), inputs=np.concatenate(ins),
                        outputs=np.concatenate(outs))
print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
