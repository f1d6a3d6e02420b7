#!/bin/bash
# 2 GPUs: bit-identity of slabs on real devices (both kernels), then C4 at N=2 and N=1 on the same box
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $NG --master-port 29521 tools/check_multigpu.py > gpurun_out/r2_check_multigpu_n$NG.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_check_multigpu_n$NG.txt
tail -n 30 gpurun_out/r2_check_multigpu_n$NG.txt
timeout 600 $TR --nproc-per-node $NG --master-port 29532 bench.py --gpus $NG --steps 100 --warmup 5 > gpurun_out/r2_bench_c4_n$NG.json 2> gpurun_out/r2_bench_c4_n$NG.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1_samebox.json 2> gpurun_out/r2_bench_c4_n1_samebox.err
for f in gpurun_out/r2_bench_c4_n$NG.json gpurun_out/r2_bench_c4_n1_samebox.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "ms", d["ms_per_step"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"], "checksum", d["checks"]["checksum"], d["config"]["kernel"][:40])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -5 gpurun_out/r2_bench_c4_n$NG.err
