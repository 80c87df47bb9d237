#!/bin/bash
# One ncu --set full capture (run under gpurun): tools/ncu_cap.sh <name> <kernel-regex> <skip> <count> [ENV=VAL ...] -- <bench args>
# The text summary (tools/ncu_summary.py) is written next to the report under gpurun_out/ncu/; the report itself is
# kept only with KEEP_NCU_REP=1.
set -u
O=gpurun_out/ncu
mkdir -p $O
name=$1; rx=$2; skip=$3; cnt=$4; shift 4
envs=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
shift
env "${envs[@]}" X=1 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/$name \
    python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/$name.log 2>&1
python tools/ncu_summary.py $O/$name.ncu-rep > $O/${name}_summary.txt 2>/dev/null
if [ "${KEEP_NCU_REP:-0}" != "1" ]; then rm -f $O/$name.ncu-rep; fi
