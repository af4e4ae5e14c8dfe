#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "length_masked" -s 2>&1 | grep -v "^$" | tail -40
} > gpurun_out/r2o.log 2>&1
cut -c1-250 gpurun_out/r2o.log
