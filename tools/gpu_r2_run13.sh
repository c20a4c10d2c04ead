#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_training_ops.py tests/test_gpu_block_backward.py tests/test_gpu_train_step.py -m gpu -q -p no:cacheprovider --maxfail=10 2>&1 | tail -6
timeout 600 python tools/gpu_train_step_profile.py > gpurun_out/r2_train_profile3.txt 2>&1; echo "profile rc=$?"; sed -n 3,16p gpurun_out/r2_train_profile3.txt | cut -c1-150
B=16 STEPS=5 timeout 600 python tools/gpu_train_step_bench.py 2>&1 | tail -2
