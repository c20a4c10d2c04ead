#!/bin/bash
# training-step check on a B200: parity tests of the backward, kernel breakdown, step time (graph and eager)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_block_backward.py tests/test_gpu_train_step.py tests/test_training_ops.py -m gpu -q -p no:cacheprovider --maxfail=8 -s 2>&1 | grep -v "^$" | tail -40
GRAPH=0 timeout 600 python tools/gpu_train_step_profile.py > gpurun_out/r2_train_step_profile_b.txt 2>&1; echo "profile rc=$?"; head -40 gpurun_out/r2_train_step_profile_b.txt
echo "--- graph"; timeout 600 python tools/gpu_train_step_bench.py 2>&1 | tail -6
echo "--- eager"; GRAPH=0 timeout 600 python tools/gpu_train_step_bench.py 2>&1 | tail -3
