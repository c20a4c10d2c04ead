#!/bin/bash
set -u
FZ_DWCONV_PAIR_SH=4 timeout 600 python -m pytest tests/test_gpu_convnext.py -m gpu -q -p no:cacheprovider -k "dwconv or engine" 2>&1 | tail -3
echo "--- pair 4x8"; FZ_DWCONV_PAIR_SH=4 timeout 300 python tools/gpu_dwconv_bench.py 2>&1 | tail -4
echo "--- pair 2x8"; timeout 300 python tools/gpu_dwconv_bench.py 2>&1 | tail -4
echo "--- round-1"; FZ_DWCONV_PAIR=0 timeout 300 python tools/gpu_dwconv_bench.py 2>&1 | tail -4
