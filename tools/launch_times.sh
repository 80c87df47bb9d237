#!/bin/bash
# per-kernel launch times of one bench configuration (ncu, cold cache, serialised): tools/launch_times.sh <tag> [ENV=VAL ...] -- <bench args>
O=gpurun_out/lt; mkdir -p $O
tag=$1; shift
envs=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
env "${envs[@]}" X=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/$tag.csv python bench.py "$@" --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - $O/$tag.csv $tag <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
acc = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        acc.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
for k, v in acc.items():
    v = v[len(v)//2:]  # steady half
    print(f"{sys.argv[2]:14s} {sum(v)/len(v)/1000:9.1f} us x{len(v):3d}  {k}")
PY
