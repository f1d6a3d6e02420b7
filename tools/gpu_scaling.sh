#!/bin/bash
# Reproduces the multi-GPU evidence (run with: gpurun --gpus 8 -- 'bash tools/gpu_scaling.sh'):
# bit-identity of the x-slab decomposition on real GPUs, then the strong (C4) and weak (C5) scaling lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 $TR --nproc-per-node $NG --master-port 29521 tools/check_multigpu.py > gpurun_out/check_multigpu_n$NG.txt 2>&1; echo "rc=$?" >> gpurun_out/check_multigpu_n$NG.txt
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/bench_c4_n$n.json 2> gpurun_out/bench_c4_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
timeout 600 $TR --nproc-per-node $NG --master-port 29541 bench.py --gpus $NG --workload c5 --steps 100 --warmup 5 --no-e2e > gpurun_out/bench_c5_n$NG.json 2> gpurun_out/bench_c5_n$NG.err
tail -n 2 gpurun_out/check_multigpu_n$NG.txt
for f in gpurun_out/bench_c4_n*.json gpurun_out/bench_c5_n*.json; do cut -c1-160 $f; done
