#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_convnext.py tests/test_gpu_pipeline.py tests/test_gpu_gemm.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
timeout 600 python tools/gpu_hbm_kernels_bench.py > gpurun_out/r2_hbm_kernels.log 2>&1; echo "hbm rc=$?"; tail -14 gpurun_out/r2_hbm_kernels.log
timeout 600 python tools/gpu_train_step_profile.py > gpurun_out/r2_train_profile.txt 2>&1; echo "profile rc=$?"; head -45 gpurun_out/r2_train_profile.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-train > gpurun_out/r2_bench_n1_grn.json 2> gpurun_out/r2_bench_n1_grn.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_n1_grn.json; echo; grep -o '"kernel_time_shares_eager.*' gpurun_out/r2_bench_n1_grn.json | head -c 600
