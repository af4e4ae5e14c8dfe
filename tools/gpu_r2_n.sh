#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3x3 or golden or precisions or bench_size or invariance or config4 or campplus" 2>&1 | tail -4
for a in "ResNet34 fp16 64 200" "CAMPPlus bf16 64 200" "ECAPA_TDNN_c512 bf16 256 200" "ECAPA_TDNN_c1024 bf16 256 200"; do
  timeout -k 10 300 python tools/op_times.py $a 2>&1 | tail -2
done
} > gpurun_out/r2n.log 2>&1
cut -c1-900 gpurun_out/r2n.log
