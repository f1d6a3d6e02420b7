#!/bin/bash
# Quick 8-GPU confirmation (gpurun --gpus 8 -- 'bash tools/gpu_scaling_quick.sh'): bit-identity of the slabs
# (check_multigpu.py, with CHECK=1), C4 at N = 8 and N = 1 on the same box and C5 at N = 8 and N = 1.  N = 4, 2 and the one-update kernel: tools/gpu_scaling.sh.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
[ "${CHECK:-0}" = 1 ] && timeout 600 $TR --nproc-per-node $NG --master-port 29521 tools/check_multigpu.py > gpurun_out/r2_final4_check_multigpu_n$NG.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_final4_check_multigpu_n$NG.txt
[ "${CHECK:-0}" = 1 ] && { grep -c "bit-identical" gpurun_out/r2_final4_check_multigpu_n$NG.txt; grep -i "mismatch\|rc=" gpurun_out/r2_final4_check_multigpu_n$NG.txt | head -5; }
timeout 400 $TR --nproc-per-node $NG --master-port 29538 bench.py --gpus $NG --steps 60 --warmup 5 --no-e2e > gpurun_out/r2_final4_scaling_c4_n$NG.json 2> gpurun_out/r2_final4_scaling_c4_n$NG.err
timeout 400 python bench.py --gpus 1 --steps 60 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_final4_scaling_c4_n1.json 2> gpurun_out/r2_final4_scaling_c4_n1.err
timeout 400 $TR --nproc-per-node $NG --master-port 29541 bench.py --gpus $NG --workload c5 --steps 60 --warmup 5 --no-e2e > gpurun_out/r2_final4_scaling_c5_n$NG.json 2> gpurun_out/r2_final4_scaling_c5_n$NG.err
timeout 400 python bench.py --gpus 1 --workload c5 --steps 60 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_final4_scaling_c5_n1.json 2> gpurun_out/r2_final4_scaling_c5_n1.err
for f in gpurun_out/r2_final4_scaling_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "launches", d["gpu_launches"], "checksum", d["checks"]["checksum"], d["config"]["kernel"][:52], d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
