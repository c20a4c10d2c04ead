#!/bin/bash
# usage: gpu_r2_multi.sh N [steps] [warmup]   -- bench.py under torchrun on N GPUs of this box (configs[3]: ONE 60k zone, strips)
set -u
N=${1:-2}; STEPS=${2:-2}; WARM=${3:-1}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps $STEPS --warmup $WARM > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench n$N rc=$?"; tail -c 2500 gpurun_out/r2_bench_n$N.json; grep -v "^$" gpurun_out/r2_bench_n$N.err | tail -12
