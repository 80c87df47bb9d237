#!/bin/bash
# Multi-GPU validation (run under gpurun --gpus N): slab transform through the C layer (CUDA IPC windows + flag barrier,
# and the NCCL all-to-all hook) against numpy, then the default multi-GPU bench line.
set -u
N=${1:-2}
O=gpurun_out/mg$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for ex in peer nccl; do
  timeout 300 $TR tools/slab_run.py --lengths 128 128 128 --exchange $ex > $O/slab128_$ex.json 2> $O/slab128_$ex.err
  timeout 300 $TR tools/slab_run.py --lengths 512 512 512 --exchange $ex > $O/slab512_$ex.json 2> $O/slab512_$ex.err
  tail -n1 $O/slab128_$ex.json; tail -n1 $O/slab512_$ex.json; grep -iE "error|Traceback" $O/slab*_$ex.err | head -5
done
timeout 300 $TR tools/slab_run.py --lengths 64 64 64 --exchange peer --scalar double > $O/slab64_f64.json 2> $O/slab64_f64.err; tail -n1 $O/slab64_f64.json
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 6 > $O/bench.json 2> $O/bench.err; tail -n1 $O/bench.json | cut -c1-3000
timeout 120 ./build/api_smoke | head -4
