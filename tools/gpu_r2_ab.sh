#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_c512 bf16 256 200 2>&1 | grep -E "res2|sum|rror"
echo "== EW8"
WS_RES2_EW8=1 timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_c512 bf16 256 200 2>&1 | grep -E "res2|sum|rror"
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "ecapa or ECAPA or res2 or masked" 2>&1 | tail -4
} > gpurun_out/r2ab.log 2>&1
cut -c1-220 gpurun_out/r2ab.log
