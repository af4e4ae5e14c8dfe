#!/bin/bash
# bench (full, default args) + ncu --set full of the dominant kernel + launch list -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_reference.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ws_conv_gemm_tc3 -s 3 -c 1 -f -o gpurun_out/prof_dom_tc3 python tools/prof_conv.py > gpurun_out/prof_dom_tc3.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-plda > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
