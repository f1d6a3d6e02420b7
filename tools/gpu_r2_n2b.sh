#!/bin/bash
# 2 GPUs, final tree: bit-identity of slabs on real devices (one-, two-, three-update launches, fp32 and fp64, slab widths
# whose published columns spread over two strips), the 2-GPU pytest, C4 and C5 at N=2 and N=1 on the same box
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29521 tools/check_multigpu.py > $O/r2_final2_check_multigpu_n2.txt 2>&1; echo "rc=$?" >> $O/r2_final2_check_multigpu_n2.txt
grep -c "bit-identical" $O/r2_final2_check_multigpu_n2.txt; grep -i "mismatch\|error\|rc=" $O/r2_final2_check_multigpu_n2.txt | head
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "real_multi_gpu" > $O/r2_final2_pytest_2gpu.txt 2>&1; tail -n 2 $O/r2_final2_pytest_2gpu.txt
timeout 600 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_final2_bench_c4_n2.json 2> $O/r2_final2_bench_c4_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_final2_bench_c4_n1.json 2> $O/r2_final2_bench_c4_n1.err
timeout 600 $TR --nproc-per-node 2 --master-port 29533 bench.py --gpus 2 --workload c5 --steps 30 --warmup 5 --no-e2e > $O/r2_final2_bench_c5_n2.json 2> $O/r2_final2_bench_c5_n2.err
timeout 600 python bench.py --gpus 1 --workload c5 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_final2_bench_c5_n1.json 2> $O/r2_final2_bench_c5_n1.err
for f in $O/r2_final2_bench_c4_n2.json $O/r2_final2_bench_c4_n1.json $O/r2_final2_bench_c5_n2.json $O/r2_final2_bench_c5_n1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"], "checksum", d["checks"]["checksum"], d["config"]["kernel"][:46])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
