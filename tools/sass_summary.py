#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library: instruction counts and the mnemonics that prove the data path
(UBLKCP / UTMALDG = TMA, SYNCS = mbarrier, LDGSTS = cp.async, SHFL = warp shuffles, BAR = block barriers).
Usage: python tools/sass_summary.py [regex]  ->  text on stdout (committed under profiles/)."""
import collections
import re
import subprocess
import sys

LIB = "portfft_b200/lib/libpfft_b200.so"
KEYS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "SHFL", "BAR", "LDG", "STG", "LDS", "STS", "FFMA", "FMUL",
        "FADD", "DFMA", "DMUL", "DADD", "LDL", "STL"]


def main():
    pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    name, counts, total = None, None, 0
    rows = []

    def flush():
        if name is not None and (pat is None or pat.search(name)):
            rows.append((name, total, dict(counts)))

    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            counts, total = collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name is not None:
            total += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + "_") or (k in ("LDG", "STG", "LDS", "STS", "LDL", "STL") and op == k):
                    counts[k] += 1
                    break
    flush()
    print(f"{'kernel':90s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
    for nm, tot, c in sorted(rows):
        print(f"{nm[:90]:90s} {tot:7d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEYS))


if __name__ == "__main__":
    main()
