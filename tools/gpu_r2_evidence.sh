#!/bin/bash
# Round-2 single-GPU evidence (gpurun -- 'bash tools/gpu_r2_evidence.sh'): smoke, the full -m gpu suite, ncu --set full
# of the shipped two-update kernel on the C4 lattice (fp32) and on C5's per-GPU lattice (fp64), the launch list of
# the bench command, every workload's bench line, the CPU reference arm, compute-sanitizer on the two-update tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.txt 2>&1; echo "smoke rc=$?" >> $O/r2_final_smoke.txt
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/r2_final_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/r2_final_pytest_gpu.txt
tail -n 2 $O/r2_final_smoke.txt; tail -n 4 $O/r2_final_pytest_gpu.txt | cut -c1-200
# ncu: the shipped shape on the benchmark lattices (one launch = two updates)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fused_march -s 2 -c 1 -o $O/r2_final_ncu_march_f32_strict_c4 \
   python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 4 --reps 1 --shapes auto > $O/r2_final_ncu_c4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:fused_march -s 2 -c 1 -o $O/r2_final_ncu_march_f64_strict_c5 \
   python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --no-mask --steps 4 --reps 1 --shapes auto > $O/r2_final_ncu_c5.log 2>&1
for r in c4 c5; do f=$O/r2_final_ncu_march_$([ $r = c4 ] && echo f32 || echo f64)_strict_$r; ncu -i $f.ncu-rep --page details > ${f}_details.txt 2>/dev/null; ncu -i $f.ncu-rep --page raw --csv > ${f}_raw.csv 2>/dev/null; done
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final_launches_bench_c4.csv \
   python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_final_launches_bench.log 2>&1
# bench lines
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_final_bench_c4_n1.json 2> $O/r2_final_bench_c4_n1.err
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > $O/r2_final_bench_c4_n1_steps100.json 2> /dev/null
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --tb2 off --no-cpu-baseline --no-e2e > $O/r2_final_bench_c4_n1_one_update.json 2> /dev/null
for w in c1 c2 c3 c5 pub; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps $([ $w = c1 ] && echo 2000 || echo 200) --warmup 20 > $O/r2_final_bench_$w.json 2> $O/r2_final_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 100 --warmup 2 > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err
for f in $O/r2_final_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(sys.argv[1].split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "launches", d.get("gpu_launches"), "frac", r.get("frac"), "frac_traffic", r.get("frac_on_measured_traffic"), "clk", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
# sanitizer: the two-update tests (small lattices) under memcheck
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "two_update or self_ring or half_as_many or halo_timeout" > $O/r2_final_memcheck.txt 2>&1; echo "rc=$?" >> $O/r2_final_memcheck.txt
tail -n 4 $O/r2_final_memcheck.txt
