#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 10 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-configs --no-plda > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err; echo "n1 exit $?"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err; echo "n2 exit $?"
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no json", e); continue
    print(n, {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d.get("clocks", {}).get("sm_mhz"))
    for k, v in d.get("configs", {}).items():
        print("   ", k, round(v["value"], 1), v.get("parity_rel_l2", v.get("max_abs_err_vs_fp64_oracle")))
PY
tail -3 gpurun_out/r02_scale_n2.err
