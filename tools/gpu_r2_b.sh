#!/bin/bash
# round 2, call B: halo-resident 3x3 kernel tests + per-op times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3x3" -s > gpurun_out/r2b_tests.log 2>&1; echo "tests exit $?"; grep -v "^conv3x3 .*max|err|" gpurun_out/r2b_tests.log | tail -25
rm -f gpurun_out/r2b_ops.log
for a in "ResNet34 fp16 64 200" "CAMPPlus bf16 64 200"; do
  timeout -k 10 300 python tools/op_times.py $a >> gpurun_out/r2b_ops.log 2>&1
done
tail -6 gpurun_out/r2b_ops.log | cut -c1-1500
