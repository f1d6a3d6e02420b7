#!/bin/bash
# two GPUs, last tree: C4 at N=2 and N=1 on the same box (128-row segments on both shapes)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/r2_final5_bench_c4_n2.json 2> $O/r2_final5_bench_c4_n2.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_final5_bench_c4_n1.json 2> $O/r2_final5_bench_c4_n1.err
for f in $O/r2_final5_bench_c4_n2.json $O/r2_final5_bench_c4_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "launches", d["gpu_launches"], "checksum", d["checks"]["checksum"], d["config"]["kernel"][:60])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
