#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "streamed" > gpurun_out/r2_seventh_tests.txt 2>&1
tail -12 gpurun_out/r2_seventh_tests.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_n1_steps20.json 2> gpurun_out/r2_bench_c4_n1_steps20.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_c4_n1_steps20.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", d["ms_per_step"], "e2e", d["e2e"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved","frac","traffic","frac_on_measured_traffic","launch_ms")})
print("cpu", d["cpu_baseline"])
PY
tail -3 gpurun_out/r2_bench_c4_n1_steps20.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -4 gpurun_out/r2_smoke.txt
