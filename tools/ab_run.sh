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
run C5_promo128 C5 X=1
run C5_promo256 C5 PFFT_COL_L2PROMO=3
run C5_promo0 C5 PFFT_COL_L2PROMO=0
run C4_promo128 C4 X=1
run C4_promo256 C4 PFFT_COL_L2PROMO=3
run L1D_promo128 L1D X=1
run L1D_promo256 L1D PFFT_COL_L2PROMO=3
