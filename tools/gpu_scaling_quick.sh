#!/bin/bash
# Short version of gpu_scaling.sh for a tight GPU budget: C4 at N=1 (kernel only) and N=NG, same box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/q_bench_c4_n1.json 2> gpurun_out/q_bench_c4_n1.err
timeout 300 $TR --nproc-per-node $NG --master-port 29561 bench.py --gpus $NG --steps 100 --warmup 5 > gpurun_out/q_bench_c4_n$NG.json 2> gpurun_out/q_bench_c4_n$NG.err
for f in gpurun_out/q_bench_c4_n*.json; do cut -c1-170 $f; done
tail -n 3 gpurun_out/q_bench_c4_n$NG.err
