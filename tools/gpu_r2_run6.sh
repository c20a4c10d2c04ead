#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_convnext.py tests/test_training_ops.py tests/test_gpu_block_backward.py tests/test_gpu_train_step.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider --maxfail=10 2>&1 | tail -12
timeout 600 python tools/gpu_hbm_kernels_bench.py > gpurun_out/r2_hbm_kernels.log 2>&1; echo "hbm rc=$?"; tail -13 gpurun_out/r2_hbm_kernels.log
timeout 600 python tools/gpu_train_step_profile.py > gpurun_out/r2_train_profile2.txt 2>&1; echo "profile rc=$?"; sed -n 3,22p gpurun_out/r2_train_profile2.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_bench_n1_b.json; echo; grep -o '"kernel_time_shares_eager.*' gpurun_out/r2_bench_n1_b.json | head -c 1500
