#!/bin/bash
# ncu captures of the bandwidth-bound kernels (pooling / SE / fbank) and the PLDA scorer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum"
timeout 600 ncu --metrics $M --clock-control none -k regex:"astp_stats|scale_residual|tmean8|fbank_kernel|cmn_kernel|linear_rows|res2_fused" -s 40 -c 14 --csv --log-file gpurun_out/hbm_kernels.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plda > /dev/null 2>&1; echo "hbm exit $?"
timeout 600 ncu --metrics $M --clock-control none -k regex:"dgemm_nt|prep_|center_norm" -s 4 -c 4 --csv --log-file gpurun_out/plda_kernels.csv \
    python -c "
import sys; sys.path.insert(0,'.')
import torch, bench
print(bench.plda_bench(torch.device('cuda:0'), iters=1))" > gpurun_out/plda_prof.log 2>&1; echo "plda exit $?"
