#!/bin/bash
# (the reports of one call must stay under 64 MiB to travel back: a dozen kernels)
# ncu --set full captures of the training step's memory-/issue-bound kernels at their stage-2 shapes (one gpurun call).
set -u
mkdir -p gpurun_out
CMD="python tools/gpu_train_step_bench.py"
export GRAPH=0 STEPS=1
$CMD > gpurun_out/r2_ncu_train_plain.log 2>&1; rc=$?; echo "plain rc=$rc"; tail -1 gpurun_out/r2_ncu_train_plain.log | cut -c1-200
[ $rc -eq 0 ] || exit 1
ncu --set full --clock-control none --import-source on -k regex:'dwconv7_f32_tile48_kernel|colreduce_vec_kernel|grn_apply_rows_kernel' \
    --launch-skip 24 --launch-count 3 -o gpurun_out/r2_prof_train_fwd -f $CMD > gpurun_out/r2_ncu_train_fwd.log 2>&1; echo "fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'grn_gelu_bwd_rows_kernel|dwconv7_wgrad_tile48_kernel|ln_bwd_vec_kernel' \
    --launch-skip 10 --launch-count 3 -o gpurun_out/r2_prof_train_bwd -f $CMD > gpurun_out/r2_ncu_train_bwd.log 2>&1; echo "bwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_small_fwd_kernel|conv3x3_small_wgrad_kernel' \
    --launch-skip 3 --launch-count 4 -o gpurun_out/r2_prof_train_conv -f $CMD > gpurun_out/r2_ncu_train_conv.log 2>&1; echo "conv rc=$?"
ls -la gpurun_out/r2_prof_train_*.ncu-rep
