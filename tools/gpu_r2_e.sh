#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3x3" 2>&1 | tail -2
for shp in "64 80 200 32" "64 40 100 64" "64 20 50 128"; do
  for dbg in 0 1 2 3 8 11; do
    echo "shape $shp dbg $dbg"
    WS_C3_DBG=$dbg WS_C3_PROF=1 timeout -k 10 120 python tools/prof_c3.py $shp 4 0 2>&1 | tail -2 | head -1
    WS_C3_DBG=$dbg timeout -k 10 120 python tools/prof_c3.py $shp 4 0 2>&1 | tail -1
  done
done
} > gpurun_out/r2e_prof.log 2>&1
cat gpurun_out/r2e_prof.log
