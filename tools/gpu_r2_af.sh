#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_GLOB_c512 bf16 256 200 2>&1 | tail -32
timeout -k 10 200 python tools/op_times.py CAMPPlus bf16 148 200 2>&1 | grep -E "cam_dense|sum|stem|conv_tc2"
} > gpurun_out/r2af.log 2>&1
cut -c1-200 gpurun_out/r2af.log
