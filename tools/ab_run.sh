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
# the variants behind the numbers of DESIGN.md section 5 (each pair: default / previous kernel)
run C2 C2 X=1
run M512_cube M512 X=1
run M512_tile M512 PFFT_NO_CUBE512=1
run M1024_rows3 M1024 X=1
run M1024_r3 M1024 PFFT_NO_ROWS3=1
run M2048_rows3 M2048 X=1
run M2048_r3 M2048 PFFT_NO_ROWS3=1
run M8192_rows3 M8192 X=1
run M8192_generic M8192 PFFT_NO_ROWS3=1
run S16_tma S16 X=1
run S16_cpasync S16 PFFT_NO_WI_TMA=1
run C5_groups C5 X=1
run C5_nogroups C5 PFFT_COL512_GROUPS=0
run L1D_inplace L1D X=1
run L1D_exchange_buffer L1D PFFT_COL_INPLACE=0
run C3B_colg C3B X=1
run C3B_generic C3B PFFT_NO_COLG=1
