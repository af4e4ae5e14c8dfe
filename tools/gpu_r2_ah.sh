#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 200 python tools/op_times.py CAMPPlus bf16 64 200 2>&1 | grep -E "conv_tc2|sum"
timeout -k 10 200 python tools/op_times.py ResNet34 fp16 64 200 2>&1 | grep -E "conv_tc2|conv_tc3|sum"
timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_GLOB_c512 bf16 256 200 2>&1 | grep -E "tstats|sum"
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv or lean or campplus or CAMPPlus or resnet or ResNet or golden or masked or f4" 2>&1 | tail -4
} > gpurun_out/r2ah.log 2>&1
cut -c1-200 gpurun_out/r2ah.log
