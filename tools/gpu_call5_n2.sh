#!/bin/bash
# 2-GPU call: real multi-process halo exchange -- parity, then bench lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py > gpurun_out/check_multigpu_n2.txt 2>&1; echo "rc=$?" >> gpurun_out/check_multigpu_n2.txt
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "multi_gpu or slab" > gpurun_out/pytest_gpu_n2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; echo "rc=$?" >> gpurun_out/bench_c4_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --steps 100 --warmup 5 --no-e2e > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err
timeout 600 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4_n1_strict.json 2> gpurun_out/bench_c4_n1_strict.err
tail -n 12 gpurun_out/check_multigpu_n2.txt; tail -n 3 gpurun_out/pytest_gpu_n2.txt
cat gpurun_out/bench_c4_n2.json gpurun_out/bench_c5_n2.json gpurun_out/bench_c4_n1_strict.json | cut -c1-400
tail -n 5 gpurun_out/bench_c4_n2.err
