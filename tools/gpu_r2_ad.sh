#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "resample or extract_dropin" 2>&1 | tail -12
} > gpurun_out/r2ad.log 2>&1
cut -c1-250 gpurun_out/r2ad.log
