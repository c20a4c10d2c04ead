#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench n1 rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --steps 2 --warmup 1 --zone 60000 --no-cpu-baseline --no-train > gpurun_out/r2_bench_60k_1gpu.json 2> gpurun_out/r2_bench_60k_1gpu.err; echo "bench 60k rc=$?"; head -c 700 gpurun_out/r2_bench_60k_1gpu.json; tail -5 gpurun_out/r2_bench_60k_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; head -c 900 gpurun_out/r2_bench_ref.json
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_model_golden.py tests/test_gpu_resnet.py tests/test_gpu_swin.py tests/test_multimodal_zone.py tests/test_rescale.py tests/test_gpu_fusion.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
