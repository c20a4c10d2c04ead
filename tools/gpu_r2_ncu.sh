#!/bin/bash
# Round-2 ncu evidence (one gpurun call): launch list of the bench command, and --set full captures of the stage-2 depthwise
# 7x7 + LayerNorm kernel and of the stage-2 fc1 GEMM (fp16 operands).  Reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline"
$BENCH > gpurun_out/r2_ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 2400 --csv --log-file gpurun_out/r2_ncu_launches.csv $BENCH > gpurun_out/r2_ncu_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
python tools/gpu_ncu_targets.py dwconv > gpurun_out/r2_ncu_plain_dwconv.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dwconv_ln -s 1 -c 2 -o gpurun_out/r2_prof_dwconv -f python tools/gpu_ncu_targets.py dwconv > gpurun_out/r2_ncu_dwconv.log 2>&1
echo "dwconv rc=$?"
python tools/gpu_ncu_targets.py gemm > gpurun_out/r2_ncu_plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_pair -s 1 -c 2 -o gpurun_out/r2_prof_gemm -f python tools/gpu_ncu_targets.py gemm > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_ncu_launches.csv 2>&1 | head
