#!/bin/bash
# full GPU validation after the training-step work: every GPU test, smoke, the default bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --maxfail=15 > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-140
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"; head -c 260 gpurun_out/r2_bench_n1_final.json; echo
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1_final.json"))
print(json.dumps(d.get("train"), indent=None)[:900])
PY
