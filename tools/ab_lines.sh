#!/bin/bash
# A/B timing helper (run under gpurun): tools/ab_lines.sh OUTDIR "name config ENV=.. ENV=.." ...
# Each argument is one bench.py run (kernel-only: no CPU baseline, no e2e); prints name, ms per step, roofline fraction.
set -u
O=gpurun_out/$1; shift
mkdir -p $O
for line in "$@"; do
  set -- $line
  name=$1 cfg=$2; shift 2
  env X=1 "$@" python bench.py --config $cfg --steps 30 --warmup 4 --no-cpu-baseline --no-e2e > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} ms={d['ms_per_step']:.4f} launches/step={d['roofline']['kernel_launches_per_step']} frac={d['roofline']['frac']:.3f} rt={d.get('roundtrip_rel_l2')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
