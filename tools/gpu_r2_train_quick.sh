#!/bin/bash
# quick look at the training step: kernel breakdown (eager, under the profiler) and the graph-replayed step time
set -u
mkdir -p gpurun_out
GRAPH=0 timeout 600 python tools/gpu_train_step_profile.py > gpurun_out/r2_train_step_profile_b.txt 2>&1; echo "profile rc=$?"; grep -v Warn gpurun_out/r2_train_step_profile_b.txt | head -${1:-30}
timeout 600 python tools/gpu_train_step_bench.py 2>&1 | tail -2
