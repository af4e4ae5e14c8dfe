#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout -k 10 600 python bench.py --no-cpu-baseline --no-plda > gpurun_out/bench_r02_astp.json 2> gpurun_out/bench_r02_astp.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_astp.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"].get("step_frac_of_sustained"))
print("parity", d["parity"]["parity_rel_l2"], "sustained", d.get("sustained", {}).get("value"), d.get("sustained", {}).get("step_frac_of_sustained"))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"], 1), v.get("ms_per_step", v.get("ms_per_pass")), v.get("step_frac_of_sustained"), v.get("parity_rel_l2"))
PY
} > gpurun_out/r2y.log 2>&1
cut -c1-250 gpurun_out/r2y.log
