#!/bin/bash
# Tensor-core stages only (conv operator v1/v2, TC model tests, smoke, bench) + CPU thread probe.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
stage() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -${TAILN:-6} gpurun_out/$name.log; }
TMO=600 stage t2_tc_conv python -m pytest tests -m gpu -q -s -k "conv_operator and tc" -p no:cacheprovider
TMO=900 stage t3_tc_models python -m pytest tests -m gpu -q -s -k "tensor_core or batch_invariance or tc_v1" -p no:cacheprovider
TMO=600 stage smoke python __graft_entry__.py smoke
TMO=900 TAILN=3 stage bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=300 TAILN=8 stage cpu_probe python tools/cpu_threads_probe.py
