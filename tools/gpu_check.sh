#!/bin/bash
# Run on the B200 box (via gpurun): staged GPU checks, each stage under its own timeout so a hung kernel
# cannot take the rest of the run with it.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
stage() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -${TAILN:-6} gpurun_out/$name.log; }
TMO=900 stage t1_simt_fbank_plda python -m pytest tests -m gpu -q -s -k "simt or fbank or plda or fp32_matches or config1" -p no:cacheprovider
TMO=600 stage t2_tc_conv python -m pytest tests -m gpu -q -s -k "conv_operator and tc" -p no:cacheprovider
TMO=900 stage t3_tc_models python -m pytest tests -m gpu -q -s -k "tensor_core or batch_invariance or tc_v1" -p no:cacheprovider
TMO=600 stage smoke python __graft_entry__.py smoke
TMO=900 stage bench python bench.py --steps 10 --warmup 3
