#!/bin/bash
# Round evidence on one B200 (run under gpurun): bench lines for every config, the reference arm, the ncu launch
# list of the default bench command and ncu --set full captures of the hot kernels.  Output: gpurun_out/evidence/.
set -u
O=gpurun_out/evidence
mkdir -p $O
python bench.py > $O/bench_default_C2.json 2> $O/bench_default_C2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_C2.json 2>> $O/bench_default_C2.err
for c in C1 C3 C4 C5 L1D S16 M256; do
  python bench.py --config $c --steps 20 --warmup 4 > $O/bench_$c.json 2> $O/bench_$c.err
done
python bench.py --config C1 --graph --steps 200 --warmup 4 --no-cpu-baseline --no-e2e > $O/bench_C1_graph.json 2>> $O/bench_C1.err
# launch list of the default command (per-launch times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default_C2.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# full captures: C2 hot kernel, tile kernel (M256 rows->rows), r3 (C3), wi (S16)
ncu --set full --clock-control none --import-source on -k regex:wg_cube -s 3 -c 2 -o $O/full_c2_wg_cube \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wg_col -s 3 -c 2 -o $O/full_m256_wg_col \
    python bench.py --config M256 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wg_r3 -s 3 -c 2 -o $O/full_c3_wg_r3 \
    python bench.py --config C3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wi_kernel -s 3 -c 2 -o $O/full_s16_wi \
    python bench.py --config S16 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wg_col -s 9 -c 3 -o $O/full_c4_wg_col \
    python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la $O
