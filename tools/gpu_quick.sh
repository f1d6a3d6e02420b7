#!/bin/bash
# Quick single-GPU verification: smoke, full GPU test-suite, default bench line (the driver's command).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
tail -n 3 gpurun_out/smoke.txt; tail -n 4 gpurun_out/pytest_gpu.txt | cut -c1-250
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["config"]["kernel"][:60])
print({k: d["roofline"][k] for k in ("achieved","frac","traffic","frac_on_measured_traffic","launch_ms","updates_per_launch")})
print(d["checks"]["checksum"], d["clocks"], d["cpu_baseline"]["value"])
PY
cat gpurun_out/bench_default.time
if [ "${1:-}" = "more" ]; then   # the other workloads' lines
  for w in c2 c5 pub c3; do
    timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
    python - $w <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["config"]["kernel"][:75])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  done
fi
