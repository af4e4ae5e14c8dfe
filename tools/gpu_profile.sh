#!/bin/bash
# ncu passes on the B200 box: (1) per-launch device time of one bench run, (2) --set full on the dominant kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ws_conv_gemm_tc -s 3 -c 2 -f -o gpurun_out/prof_conv \
    python tools/prof_conv.py > gpurun_out/prof_conv.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out
