#!/usr/bin/env python
"""bench_manual_float / bench_manual_double / bench_float of the reference, over the B200 library:

    python tools/bench_manual.py [--double] "d=cpx,n=64,b=1024" "d=cpx,n=512x512,p=ip" ...
    python tools/bench_manual.py --canned            # the four bench_float configurations
    python tools/bench_manual.py --reference-set     # reference_dft_set.hpp: complex (incl. 65537) and real float sets

Prints one JSON object per benchmark with the reference's names and counters (portfft_b200/bench_cli.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from portfft_b200 import bench_cli  # noqa: E402


def main():
    args = [a for a in sys.argv[1:]]
    scalar = "float"
    if "--double" in args:
        scalar = "double"
        args.remove("--double")
    if "--help" in args or not args:
        print(__doc__)
        print("keys:", ", ".join(f"'{a}', '{b}'" for a, b in bench_cli.ARG_KEYS))
        return 0
    import portfft_b200 as pf

    configs = []
    if "--canned" in args:
        args.remove("--canned")
        for name, lengths, batch in bench_cli.CANNED_FLOAT:
            d = pf.descriptor(lengths, "float")
            d.number_of_transforms = batch
            configs.append((d, name))
    if "--reference-set" in args:
        args.remove("--reference-set")
        for name, lengths, batch in bench_cli.CANNED_COMPLEX_SET:
            d = pf.descriptor(lengths, "float")
            d.number_of_transforms = batch
            configs.append((d, name))
        for name, lengths, batch in bench_cli.CANNED_REAL_SET:
            d = pf.descriptor(lengths, "float", pf.domain.REAL)
            d.number_of_transforms = batch
            configs.append((d, name))
    for a in args:
        d = bench_cli.parse_manual_args(a, scalar)
        tname = "f" if scalar == "float" else "d"  # typeid(FType).name() of the reference's suffix
        configs.append((d, f"{tname}:{a}"))
    for d, suffix in configs:
        for r in bench_cli.run_host_device_benchmark(d, suffix):
            print(json.dumps(r), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
