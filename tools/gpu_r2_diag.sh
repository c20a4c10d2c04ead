#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tests/diag/gpu_stage_errors.py > gpurun_out/r2_stage_errors.log 2>&1; echo "stage rc=$?"; tail -25 gpurun_out/r2_stage_errors.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_convnext.py tests/test_model_golden.py tests/test_gpu_gemm.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; grep -n "agree\|passed\|failed" gpurun_out/r2_pytest_gpu2.log | tail -20
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python tests/error_budget.py --tile 512 --seeds 2 --device cuda > gpurun_out/r2_error_budget.txt 2>&1; echo "budget rc=$?"; head -14 gpurun_out/r2_error_budget.txt
