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
BENCH="python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline"
$BENCH > gpurun_out/r2_ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2400 --csv --log-file gpurun_out/r2_ncu_launches.csv $BENCH > gpurun_out/r2_ncu_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/gpu_hbm_kernels_bench.py 2>&1 | tail -16
echo "--- training step, fp16 forward (default) and bf16 forward (FZ_TRAIN_ACT=bf16)"
STEPS=3 timeout 300 python tools/gpu_train_step_bench.py 2>&1 | tail -1 | cut -c1-260
FZ_TRAIN_ACT=bf16 STEPS=3 timeout 300 python tools/gpu_train_step_bench.py 2>&1 | tail -1 | cut -c1-260
GRAPH=0 timeout 600 python tools/gpu_train_step_profile.py > gpurun_out/r2_train_step_profile_b.txt 2>&1; echo "profile rc=$?"
