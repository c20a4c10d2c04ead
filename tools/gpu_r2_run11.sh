#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_convnext.py tests/test_gpu_pipeline.py tests/test_gpu_train_step.py tests/test_gpu_postprocess.py tests/test_multimodal_zone.py tests/test_gpu_swin.py -m gpu -q -p no:cacheprovider --maxfail=10 2>&1 | tail -6
timeout 600 python tools/gpu_hbm_kernels_bench.py > gpurun_out/r2_hbm_kernels.log 2>&1; echo "hbm rc=$?"; tail -12 gpurun_out/r2_hbm_kernels.log | cut -c1-140
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; echo "bench rc=$?"; head -c 260 gpurun_out/r2_bench_n1_d.json; echo; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_d.json').read().strip().splitlines()[-1])
print("e2e", d['e2e']['value'], "roof", d['roofline']['frac'], d['roofline']['avg_launch_us'], "other", d.get('other_configs'), "train ms", d['train'].get('ms_per_step'))
print(d['kernel_time_shares_eager'])
PY
tail -3 gpurun_out/r2_bench_n1_d.err
