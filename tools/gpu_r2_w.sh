#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 500 ncu --set full --clock-control none --import-source on -f -k regex:ws_astp_fused -s 1 -c 1 -o gpurun_out/r02_full_astp python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 > gpurun_out/r02_full_astp.log 2>&1; echo "astp $?"
