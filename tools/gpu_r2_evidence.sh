#!/bin/bash
# round-2 evidence: ncu --set full of the four dominant kernels, launch lists of the three model workloads, HBM-bound kernel metrics
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N="ncu --set full --clock-control none --import-source on -f"
timeout 500 $N -k regex:ws_conv_gemm_tc3 -s 3 -c 1 -o gpurun_out/r02_full_tc3 python tools/prof_conv.py > gpurun_out/r02_full_tc3.log 2>&1; echo "tc3 $?"
timeout 500 $N -k regex:ws_conv3x3 -s 1 -c 2 -o gpurun_out/r02_full_c3 python tools/op_times.py ResNet34 fp16 64 200 > gpurun_out/r02_full_c3.log 2>&1; echo "c3 $?"
timeout 500 $N -k regex:ws_conv3x3 -s 20 -c 1 -o gpurun_out/r02_full_c3_l3 python tools/op_times.py ResNet34 fp16 64 200 > gpurun_out/r02_full_c3_l3.log 2>&1; echo "c3 l3 $?"
timeout 500 $N -k regex:ws_cam_dense -s 1 -c 1 -o gpurun_out/r02_full_cam python tools/op_times.py CAMPPlus bf16 64 200 > gpurun_out/r02_full_cam.log 2>&1; echo "cam $?"
timeout 500 $N -k regex:dgemm_nt_dmma -s 1 -c 1 -o gpurun_out/r02_full_dgemm python tools/plda_full.py 65536 > gpurun_out/r02_full_dgemm.log 2>&1; echo "dgemm $?"
timeout 500 $N -k regex:res2_fused -s 1 -c 1 -o gpurun_out/r02_full_res2 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 > gpurun_out/r02_full_res2.log 2>&1; echo "res2 $?"
L="ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv"
timeout 500 $L --log-file gpurun_out/r02_launches_ecapa.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-plda --no-configs --sustained-s 0 > /dev/null 2>&1; echo "ll ecapa $?"
timeout 500 $L --log-file gpurun_out/r02_launches_resnet.csv python bench.py --workload resnet34_fp16_b64 --steps 2 --warmup 1 --no-cpu-baseline --no-plda --no-configs --sustained-s 0 > /dev/null 2>&1; echo "ll resnet $?"
timeout 500 $L --log-file gpurun_out/r02_launches_campp.csv python bench.py --workload campplus_bf16_b64 --steps 2 --warmup 1 --no-cpu-baseline --no-plda --no-configs --sustained-s 0 > /dev/null 2>&1; echo "ll campp $?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 400 ncu --metrics $M --clock-control none -k regex:"astp_stats|scale_residual|se_gate|fbank_kernel|cmn_kernel|linear_rows|convert|stem_kernel|tstats|zero_tail" -s 30 -c 20 --csv --log-file gpurun_out/r02_hbm_kernels.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plda --no-configs --sustained-s 0 > /dev/null 2>&1; echo "hbm $?"
ls -la gpurun_out/r02_*
