#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== multicast"
timeout -k 10 200 python tools/prof_c3.py "64 20 50 128" "256 20 50 128" --prof 2>&1 | grep -E "c3 prof|conv3x3" | cut -c1-420
echo "== no multicast"
WS_C3_NO_MC=1 timeout -k 10 200 python tools/prof_c3.py "64 20 50 128" "256 20 50 128" --prof 2>&1 | grep -E "c3 prof|conv3x3" | cut -c1-420
} > gpurun_out/r2aa.log 2>&1
cat gpurun_out/r2aa.log
