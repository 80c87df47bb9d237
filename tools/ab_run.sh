#!/bin/bash
# A/B timing of kernel variants on one B200 (run under gpurun): each line = config, environment, ms per step.
set -u
O=gpurun_out/ab
mkdir -p $O
run() {  # name config env...
  local name=$1 cfg=$2; shift 2
  env "$@" python bench.py --config $cfg --steps 30 --warmup 4 --no-cpu-baseline --no-e2e > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} ms={d['ms_per_step']:.4f} frac={d['roofline']['frac']:.3f} rt={d.get('roundtrip_rel_l2')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run C3B_colg C3B X=1
run C3B_generic C3B PFFT_NO_COLG=1
run C3 C3 X=1
run C5 C5 X=1
