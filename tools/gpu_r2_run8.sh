#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=15 > gpurun_out/r2_pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu8.log
timeout 600 python tools/gpu_hbm_kernels_bench.py > gpurun_out/r2_hbm_kernels.log 2>&1; echo "hbm rc=$?"; tail -14 gpurun_out/r2_hbm_kernels.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_n1_c.json; echo; grep -o '"kernel_time_shares_eager.*' gpurun_out/r2_bench_n1_c.json | head -c 700
