#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== CL=4 op times"
timeout -k 10 180 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "conv_tc3|ops, sum|GEMM-like|rror"
echo "== CL=2 op times"
WS_TC3_CL=2 timeout -k 10 180 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 2>&1 | grep -E "conv_tc3|ops, sum|GEMM-like|rror"
echo "== tests (CL=4 default)"
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "ecapa or ECAPA or conv or tc3 or masked" 2>&1 | tail -5
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --no-plda 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('step_frac_of_sustained'), d.get('parity'))
"
} > gpurun_out/r2s.log 2>&1
cut -c1-250 gpurun_out/r2s.log
