#!/bin/bash
# quick loop: conv-operator + lean-epilogue + fused-Res2 + model tests, per-op times, short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "${KEXPR:-conv_operator or lean or res2 or tensor_core or tf32x3 or batch_invariance or tc_v1}" > gpurun_out/t_quick.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t_quick.log
timeout 300 python tools/op_times.py > gpurun_out/ops_ecapa.log 2>&1; tail -2 gpurun_out/ops_ecapa.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-plda > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -c 900 gpurun_out/bench.log
