#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "res2 or fused_astp or masked or ragged" -s 2>&1 | grep -E "res2 fused|passed|failed|Error|error" | tail -20
timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_c1024 bf16 64 800 2>&1 | grep -E "res2|sum"
WS_NO_RES2_FUSED=1 timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_c1024 bf16 64 800 2>&1 | grep -E "sum"
} > gpurun_out/r2ae.log 2>&1
cut -c1-250 gpurun_out/r2ae.log
