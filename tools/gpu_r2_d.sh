#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 tools/experimental/mma_rate_probe > gpurun_out/r2_mma_probe2.log 2>&1; cat gpurun_out/r2_mma_probe2.log
bash tools/gpu_r2_c.sh
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3x3" > gpurun_out/r2d_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2d_tests.log
