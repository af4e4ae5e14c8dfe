#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "${KEXPR:-column_sums or colsum or batch_invariance or tensor_core}" > gpurun_out/t_quick.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t_quick.log
for v in 0 1 2; do echo "astp variant $v"; WS_ASTP_VARIANT=$v timeout 300 python tools/op_times.py 2>&1 | tail -1; done
