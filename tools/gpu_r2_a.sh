#!/bin/bash
# round 2, call A: GPU test suite, the new bench line (all configs), per-op times of the weak models
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r2a_tests.log
timeout 900 python bench.py > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; tail -c 600 gpurun_out/r2a_bench.err
for a in "ResNet34 fp16 64 200" "CAMPPlus bf16 64 200" "ECAPA_TDNN_c512 bf16 256 200" "ECAPA_TDNN_c512 tf32x3 256 200" "ECAPA_TDNN_c1024 bf16 256 200"; do
  timeout 300 python tools/op_times.py $a >> gpurun_out/r2a_ops.log 2>&1
done
tail -12 gpurun_out/r2a_ops.log
