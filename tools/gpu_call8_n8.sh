#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 tools/check_multigpu.py > gpurun_out/check_multigpu_n8.txt 2>&1; echo "rc=$?" >> gpurun_out/check_multigpu_n8.txt
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 100 --warmup 5 --no-e2e > gpurun_out/bench2_c4_n$n.json 2> gpurun_out/bench2_c4_n$n.err; echo "rc=$?" >> gpurun_out/bench2_c4_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench2_c4_n1.json 2> gpurun_out/bench2_c4_n1.err
timeout 600 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 100 --warmup 5 --math fast --no-e2e > gpurun_out/bench2_c4_n8_fast.json 2> gpurun_out/bench2_c4_n8_fast.err
timeout 600 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench2_c4_n8_e2e.json 2> gpurun_out/bench2_c4_n8_e2e.err
tail -n 3 gpurun_out/check_multigpu_n8.txt
for f in bench2_c4_n1 bench2_c4_n2 bench2_c4_n4 bench2_c4_n8 bench2_c4_n8_fast bench2_c4_n8_e2e; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), "MLUPS", d["ms_per_step"], "ms/step  frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"] and round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("$f", "ERR", e)
PY
done
