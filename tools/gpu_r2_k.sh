#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "ResNet50 or ResNet1 or ResNet2 or XVEC or embedding_processing" -s 2>&1 | grep -v "^$" | tail -40
} > gpurun_out/r2k.log 2>&1
cut -c1-220 gpurun_out/r2k.log
