#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "speaker_engine or extract_dropin or eval_sv or scalar_api" -s 2>&1 | grep -v "^$" | tail -15
timeout -k 10 300 python __graft_entry__.py smoke 2>&1 | tail -3
} > gpurun_out/r2l.log 2>&1
cut -c1-300 gpurun_out/r2l.log
