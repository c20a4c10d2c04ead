#!/bin/bash
# compute-sanitizer over the hot path (race detection / sanitizers): memcheck and racecheck on smoke()
# (one 1000 x 700 zone through feeder, encoder / decoder kernels, the head epilogue) and on the GEMM unit tests; initcheck on
# the post-processing kernels.  Needs a GPU: run under gpurun, e.g.
#   gpurun --timeout 900 -- 'bash tools/gpu_sanitizer.sh > gpurun_out/sanitizer.log 2>&1'
# Each tool slows kernels 10-100x; the selection below takes a few minutes on a B200.  NOT run in round 2 (the GPU budget
# was spent on kernels, measurements and tests); kept so that the next round starts with it.
set -u
cd "$(dirname "$0")/.."
export FZ_CUDA_GRAPH=0
run() { echo "=== $*"; "$@" 2>&1 | tail -25; }
run compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()"
run compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_gemm.py -x -q -k "not sustained"
run compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_postprocess.py -x -q
run compute-sanitizer --tool memcheck python -m pytest tests/test_polygonize.py -x -q -m gpu
