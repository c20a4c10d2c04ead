#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_convnext.py tests/test_gpu_swin.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider --maxfail=8 2>&1 | tail -5
timeout 300 python tools/gpu_gemm_shapes.py 2>&1 | tail -14
echo "--- without TMA store"; FZ_GEMM_TMA_STORE=0 timeout 300 python tools/gpu_gemm_shapes.py 2>&1 | tail -14
timeout 900 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2_bench_tma.json 2> gpurun_out/r2_bench_tma.err; echo "bench rc=$?"; head -c 200 gpurun_out/r2_bench_tma.json; echo
FZ_GEMM_TMA_STORE=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2_bench_notma.json 2> gpurun_out/r2_bench_notma.err; echo "bench(no tma store) rc=$?"; head -c 200 gpurun_out/r2_bench_notma.json; echo
