#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
for shp in "64 80 200 32" "64 40 100 64" "64 20 50 128"; do
  WS_C3_PROF=1 timeout -k 10 120 python tools/prof_c3.py $shp 4 1 2>&1 | tail -3
  timeout -k 10 120 python tools/prof_c3.py $shp 4 1 2>&1 | tail -1
  timeout -k 10 120 python tools/prof_c3.py $shp 4 0 2>&1 | tail -1
  timeout -k 10 120 python tools/prof_c3.py $shp 3 1 2>&1 | tail -1
done
} > gpurun_out/r2c_prof.log 2>&1
cat gpurun_out/r2c_prof.log
