#!/bin/bash
# Round evidence on one B200 (run under gpurun): bench lines for every config, the reference arm, the reference's
# own benchmark sets, the ncu launch list of the default bench command and ncu --set full captures of the hot kernels.
# Output: gpurun_out/evidence/.
set -u
O=gpurun_out/evidence
mkdir -p $O
python bench.py > $O/bench_default_C2.json 2> $O/bench_default_C2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_C2.json 2>> $O/bench_default_C2.err
for c in C1 C3 C3B C4 C5 L1D S16 M256 M512 M1024 M2048 M8192 D512 D1024 D2048 D4096 R32 R512 R8192 R131072; do
  python bench.py --config $c --steps 20 --warmup 4 > $O/bench_$c.json 2> $O/bench_$c.err
done
python bench.py --config C1 --graph --steps 200 --warmup 4 --no-cpu-baseline --no-e2e > $O/bench_C1_graph.json 2>> $O/bench_C1.err
python tools/bench_manual.py --reference-set > $O/bench_manual_reference_set.jsonl 2> $O/bench_manual_reference_set.err
# launch list of the default command (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_C2.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# full captures: C2 hot kernel, N=512 tile variant (M512), column kernels of C5 and C4, r3 (C3), TMA thread-level (S16)
cap() {  # name kernel-regex skip count bench-args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $O/full_$name \
      python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  # only the text summary travels back (gpurun merges at most 64 MiB; the reports are 5-15 MB each)
  python tools/ncu_summary.py $O/full_$name.ncu-rep > $O/full_${name}_ncu_summary.txt 2>/dev/null
  if [ "${KEEP_NCU_REP:-0}" != "1" ]; then rm -f $O/full_$name.ncu-rep; fi
}
cap c2_wg_cube wg_cube 3 2
cap m512_wg_cube wg_cube 3 2 --config M512
cap m256_wg_col wg_col 3 2 --config M256
cap c5_all "wg_c" 9 3 --config C5
cap c4_wg_col wg_col 9 3 --config C4
cap c3_wg_r3 wg_r3 3 2 --config C3
cap c3b_wg_colr3 colr3 3 2 --config C3B
cap m8192_wg_rows3 wg_rows3 3 2 --config M8192
cap d4096_wg_cube wg_cube 3 2 --config D4096
cap s16_wi_tma wi_tma 3 2 --config S16
ls -la $O
