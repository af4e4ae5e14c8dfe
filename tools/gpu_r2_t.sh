#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== CL=4 op times"
timeout -k 10 180 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "conv_tc3|ops, sum|GEMM-like|rror"
echo "== CL=2 op times"
WS_TC3_CL=2 timeout -k 10 180 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "conv_tc3|ops, sum|GEMM-like|rror"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
} > gpurun_out/r2t.log 2>&1
cut -c1-250 gpurun_out/r2t.log
