#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3x3" 2>&1 | tail -2
timeout -k 10 300 python tools/prof_c3.py --dbg ${DBGS:-0,11} --tc 4 --prof 2>&1 | awk '/c3 prof/{c++; if (c%4==1) print; next} {print}'
} > gpurun_out/r2f_prof.log 2>&1
cat gpurun_out/r2f_prof.log
