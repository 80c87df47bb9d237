#!/bin/bash
# Refresh of the evidence for the configs whose kernels changed late in the round (run under gpurun); same outputs as
# tools/gpu_evidence.sh, into gpurun_out/evidence/.
set -u
O=gpurun_out/evidence
mkdir -p $O
python bench.py > $O/bench_default_C2.json 2> $O/bench_default_C2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_C2.json 2>> $O/bench_default_C2.err
for c in "$@"; do
  python bench.py --config $c --steps 20 --warmup 4 > $O/bench_$c.json 2> $O/bench_$c.err
done
python tools/bench_manual.py --reference-set > $O/bench_manual_reference_set.jsonl 2> $O/bench_manual_reference_set.err
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/full_$name \
      python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python tools/ncu_summary.py $O/full_$name.ncu-rep > $O/full_${name}_ncu_summary.txt 2>/dev/null
  rm -f $O/full_$name.ncu-rep
}
cap m8192_wg_rows3 wg_rows3 3 2 --config M8192
cap d4096_wg_cube wg_cube 3 2 --config D4096
cap c4_wg_col wg_col 9 3 --config C4
cap r8192_wg_cube wg_cube 3 2 --config R8192
ls $O | wc -l
