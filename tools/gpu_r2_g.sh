#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for a in "ResNet34 fp16 64 200" "CAMPPlus bf16 64 200" "ECAPA_TDNN_c512 bf16 256 200" "ECAPA_TDNN_c1024 bf16 256 200"; do
  timeout -k 10 300 python tools/op_times.py $a 2>&1 | tail -2
done
} > gpurun_out/r2g.log 2>&1
cat gpurun_out/r2g.log | cut -c1-1800
