#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "fused_astp" -s 2>&1 | grep -v "^$" | tail -12
timeout -k 10 300 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "astp|sum"
WS_ASTP_PROF=1 timeout -k 10 120 python tools/prof_astp.py 2>&1 | grep -E "astp prof|rror" | tail -1
} > gpurun_out/r2u.log 2>&1
cut -c1-400 gpurun_out/r2u.log
