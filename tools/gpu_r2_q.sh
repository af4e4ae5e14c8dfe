#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
timeout -k 10 900 python bench.py > gpurun_out/bench_r02_full.json 2> gpurun_out/bench_r02_full.err
tail -3 gpurun_out/bench_r02_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_full.json").read().strip().splitlines()[-1])
def show(k, v, ind=0):
    if isinstance(v, dict):
        print(" " * ind + k + ":")
        for kk, vv in v.items(): show(kk, vv, ind + 2)
    elif isinstance(v, list) and v and isinstance(v[0], dict):
        for i, x in enumerate(v): show(f"{k}[{i}]", x, ind)
    else:
        print(" " * ind + f"{k}: {v}")
for k, v in d.items(): show(k, v)
PY
} > gpurun_out/r2q.log 2>&1
cut -c1-260 gpurun_out/r2q.log | tail -150
