#!/bin/bash
# Reproduces the multi-GPU evidence (run with: gpurun --gpus 8 -- 'bash tools/gpu_scaling.sh'):
# bit-identity of the x-slab decomposition on real GPUs (one-update and two-update kernels), then the strong (C4)
# scaling lines at N = 8, 4, 2, 1 on the same box and the weak (C5) line at N = 8.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 $TR --nproc-per-node $NG --master-port 29521 tools/check_multigpu.py > gpurun_out/r2_check_multigpu_n$NG.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_check_multigpu_n$NG.txt
tail -n 3 gpurun_out/r2_check_multigpu_n$NG.txt
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r2_scaling_c4_n$n.json 2> gpurun_out/r2_scaling_c4_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scaling_c4_n1.json 2> gpurun_out/r2_scaling_c4_n1.err
timeout 600 $TR --nproc-per-node $NG --master-port 29541 bench.py --gpus $NG --workload c5 --steps 100 --warmup 5 --no-e2e > gpurun_out/r2_scaling_c5_n$NG.json 2> gpurun_out/r2_scaling_c5_n$NG.err
timeout 600 $TR --nproc-per-node $NG --master-port 29542 bench.py --gpus $NG --steps 100 --warmup 5 --tb2 off --no-e2e > gpurun_out/r2_scaling_c4_n${NG}_one_update_kernel.json 2> /dev/null
for f in gpurun_out/r2_scaling_c4_n*.json gpurun_out/r2_scaling_c5_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"], "checksum", d["checks"]["checksum"], d["config"]["kernel"][:46])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
