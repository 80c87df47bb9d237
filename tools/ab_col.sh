#!/bin/bash
# timing of the configurations that run through the column-tile kernels (run under gpurun)
set -u
O=gpurun_out/abc
mkdir -p $O
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
OLD=$PWD/portfft_b200/lib/libpfft_old.so
run L1D_unfused L1D PFFT_NO_FUSE=1
run L1D_old L1D PFFT_NO_FUSE=1 PFFT_LIB=$OLD
run L1D_fused L1D PFFT_FUSE_CHUNK_KB=2048 PFFT_FUSE_LEAD_TENTHS=30
run C4_unfused C4 PFFT_NO_FUSE=1
run C4_old C4 PFFT_NO_FUSE=1 PFFT_LIB=$OLD
run C5 C5 X=1
run C5_old C5 PFFT_LIB=$OLD
run M256 M256 X=1
run M256_old M256 PFFT_LIB=$OLD
run C1 C1 X=1
run C1_old C1 PFFT_LIB=$OLD
