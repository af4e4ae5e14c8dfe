#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
for d in 0 1 2 4 7; do
  for g in 3; do
    echo "== dbg $d g $g"; WS_ASTP_DBG=$d WS_ASTP_G=$g timeout -k 10 120 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "astp_fused|rror"
  done
done
echo "== g sweep"; for g in 1 2 6 12; do WS_ASTP_G=$g timeout -k 10 120 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "astp_fused|rror"; done
} > gpurun_out/r2v.log 2>&1
cut -c1-200 gpurun_out/r2v.log
