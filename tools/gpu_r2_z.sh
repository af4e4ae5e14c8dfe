#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 200 python tools/op_times.py ResNet34 fp16 64 200 2>&1 | grep -E "F=20|sum|rror" | head -16
echo "== no multicast"
WS_C3_NO_MC=1 timeout -k 10 200 python tools/op_times.py ResNet34 fp16 64 200 2>&1 | grep -E "F=20|sum|rror" | head -4
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3x3 or resnet or ResNet or masked or f4" 2>&1 | tail -4
} > gpurun_out/r2z.log 2>&1
cut -c1-220 gpurun_out/r2z.log
