#!/bin/bash
# usage: gpu_r2_train_multi.sh N   -- the training step under torchrun on N GPUs: graph replay vs eager, all-reduce overlap
set -u
N=${1:-2}
mkdir -p gpurun_out
for G in 1 0; do
  echo "--- GRAPH=$G"
  GRAPH=$G STEPS=4 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$G \
    tools/gpu_train_step_bench.py 2>&1 | grep -v "^$" | tail -4
done
