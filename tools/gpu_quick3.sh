#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "fbank or config1 or extract or config4" > gpurun_out/t_quick.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/t_quick.log
for v in "" 1; do
  if [ -n "$v" ]; then export WS_FBANK_V1=1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-plda > gpurun_out/bench_fb$v.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/bench_fb$v.log') if x.startswith('{')][-1]
d=json.loads(l); print('fbank_v1=$v', d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
