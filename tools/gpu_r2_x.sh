#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
for d in 0 1 7; do
echo "== dbg $d"; WS_ASTP_PROF=1 WS_ASTP_DBG=$d timeout -k 10 120 python tools/prof_astp.py 2>&1 | grep -E "astp prof|rror" | tail -1
done
} > gpurun_out/r2x.log 2>&1
cut -c1-400 gpurun_out/r2x.log
