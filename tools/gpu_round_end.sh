#!/bin/bash
# round-end evidence: full GPU test suite, default bench, reference arm, ncu --set full of the dominant kernel, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 400 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_reference.log | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ws_conv_gemm_tc3 -s 3 -c 1 -f -o gpurun_out/prof_dom_tc3 python tools/prof_conv.py > gpurun_out/prof_dom_tc3.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-plda > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 400 ncu --metrics $M --clock-control none -k regex:"astp_stats|scale_residual|se_gate|fbank_kernel|cmn_kernel|linear_rows|res2_fused|convert" -s 30 -c 16 --csv --log-file gpurun_out/hbm_kernels.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plda > /dev/null 2>&1; echo "hbm exit $?"
