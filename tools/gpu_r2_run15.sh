#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_convnext.py tests/test_model_golden.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
echo "--- pair kernel"; timeout 300 python tools/gpu_dwconv_bench.py 2>&1 | tail -5
echo "--- round-1 kernel"; FZ_DWCONV_PAIR=0 timeout 300 python tools/gpu_dwconv_bench.py 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2_bench_pair.json 2> gpurun_out/r2_bench_pair.err; echo "bench rc=$?"; head -c 200 gpurun_out/r2_bench_pair.json; echo
FZ_DWCONV_PAIR=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2_bench_nopair.json 2> gpurun_out/r2_bench_nopair.err; echo "bench(no pair) rc=$?"; head -c 200 gpurun_out/r2_bench_nopair.json; echo
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-120
