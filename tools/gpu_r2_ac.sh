#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "K=1536 N=128|sum|rror"
timeout -k 10 200 python tools/op_times.py ECAPA_TDNN_GLOB_c512 bf16 256 200 2>&1 | grep -E "K=1536 N=128|sum|rror"
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_speaker_engine.py -m gpu -q -x -p no:cacheprovider -k "ecapa or ECAPA or golden or lean or masked or config or engine" 2>&1 | tail -4
} > gpurun_out/r2ac.log 2>&1
cut -c1-220 gpurun_out/r2ac.log
