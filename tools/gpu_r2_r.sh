#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 300 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200
timeout -k 10 300 python tools/op_times.py ECAPA_TDNN_c512 bf16 256 200
timeout -k 10 300 python tools/op_times.py ResNet34 fp16 64 200
timeout -k 10 300 python tools/op_times.py CAMPPlus bf16 64 200
timeout -k 10 300 python tools/op_times.py CAMPPlus bf16 64 600
} > gpurun_out/r02_op_times.md 2>&1
cut -c1-200 gpurun_out/r02_op_times.md
