#!/bin/bash
set -u
O=gpurun_out/abd; mkdir -p $O
run() {  # name config env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --steps 30 --warmup 4 --no-cpu-baseline --no-e2e > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} ms={d['ms_per_step']:.4f} launches/step={d['roofline']['kernel_launches_per_step']} frac={d['roofline']['frac']:.3f} rt={d.get('roundtrip_rel_l2')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for c in D512 D1024 D2048 D4096; do run ${c}_new $c X=1; run ${c}_old $c PFFT_NO_CUBE_F64=1; done
