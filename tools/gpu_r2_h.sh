#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "campplus_fused or CAMPPlus or config4" -s 2>&1 | grep -v "^$" | tail -25
timeout -k 10 300 python tools/op_times.py CAMPPlus bf16 64 200 2>&1 | tail -2
} > gpurun_out/r2h.log 2>&1
cat gpurun_out/r2h.log | cut -c1-1500
