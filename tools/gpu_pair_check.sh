#!/bin/bash
# bring-up of the CTA-pair GEMM: parity tests under a timeout, then the shape benchmark with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k pair > gpurun_out/pair_test.log 2>&1; echo pair_test=$?
tail -15 gpurun_out/pair_test.log
FZ_GEMM_PAIR=0 timeout 120 python tools/gpu_gemm_shapes.py > gpurun_out/shapes_pair0.log 2>&1; cat gpurun_out/shapes_pair0.log
FZ_GEMM_PAIR=1 timeout 120 python tools/gpu_gemm_shapes.py > gpurun_out/shapes_pair1.log 2>&1; cat gpurun_out/shapes_pair1.log
