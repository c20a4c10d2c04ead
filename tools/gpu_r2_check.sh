#!/bin/bash
# Round-2 GPU check: full GPU test suite (prints kept), smoke, bench with fp16 operands (product) and with the bf16 A/B
# build, and the precision budget at full tile size.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s --maxfail=25 -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_f16.json 2> gpurun_out/r2_bench_f16.err; echo "bench f16 rc=$?"; tail -c 600 gpurun_out/r2_bench_f16.json
FZ_OPERANDS=bf16 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_bf16.json 2> gpurun_out/r2_bench_bf16.err; echo "bench bf16 rc=$?"; tail -c 300 gpurun_out/r2_bench_bf16.json
FZ_OPERANDS=bf16 timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke_bf16.log 2>&1; tail -1 gpurun_out/r2_smoke_bf16.log
timeout 600 python tests/error_budget.py --tile 512 --seeds 3 --device cuda > gpurun_out/r2_error_budget.txt 2>&1; echo "budget rc=$?"; head -12 gpurun_out/r2_error_budget.txt
