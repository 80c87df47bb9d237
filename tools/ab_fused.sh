#!/bin/bash
# A/B timing of the fused GLOBAL-level kernel (wg_fused.cu) on one B200 (run under gpurun).
set -u
O=gpurun_out/abf
mkdir -p $O
run() {  # name config env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --steps 30 --warmup 4 --no-cpu-baseline --no-e2e > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"{sys.argv[2]:28s} ms={d['ms_per_step']:.4f} launches/step={d['roofline']['kernel_launches_per_step']} rt={d.get('roundtrip_rel_l2')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run L1D_unfused L1D PFFT_NO_FUSE=1
for kb in 2048 8192; do for t in 10 20 30; do
run L1D_fused_${kb}K_t${t} L1D PFFT_FUSE_CHUNK_KB=$kb PFFT_FUSE_LEAD_TENTHS=$t
done; done
run L1D_fused_2cta L1D PFFT_FUSE_CTAS_PER_SM=2
run C4_unfused C4 PFFT_NO_FUSE=1
for kb in 8192; do for t in 10 20; do
run C4_fused_${kb}K_t${t} C4 PFFT_FUSE_CHUNK_KB=$kb PFFT_FUSE_LEAD_TENTHS=$t
done; done
