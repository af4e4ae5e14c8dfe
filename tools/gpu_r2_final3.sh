#!/bin/bash
# round-2 (second session) final GPU evidence: Res2Net / ERes2Net tests, the whole GPU suite, smoke, the default bench (with the
# "extras" block), per-op tables of the two new base models
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
timeout -k 5 120 python -m pytest tests/test_res2net.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02c_res2net.log 2>&1; echo "res2net tests exit $? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/r02c_res2net.log | cut -c1-300
timeout -k 5 200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_res2net.py > gpurun_out/r02c_tests.log 2>&1; echo "full suite exit $? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/r02c_tests.log | cut -c1-300
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02c_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02c_smoke.log | cut -c1-200
timeout -k 5 150 python bench.py > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err; echo "bench exit $? ($(( $(date +%s) - t0 )) s)"
{
timeout -k 5 60 python tools/op_times.py Res2Net34_Base fp16 64 200
timeout -k 5 60 python tools/op_times.py ERes2Net34_Base fp16 64 200
} > gpurun_out/r02c_op_times_res2net.md 2>&1; echo "op times done ($(( $(date +%s) - t0 )) s)"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c_bench_default.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"])
    for k, v in d.get("configs", {}).items():
        print(k, round(v["value"], 1), v.get("ms_per_step", v.get("ms_per_pass")), v.get("parity_rel_l2", v.get("max_abs_err_vs_fp64_oracle")))
    for k, v in (d.get("extras") or {}).items():
        print("extra", k, v.get("value"), v.get("ms_per_step"), v.get("parity_rel_l2"), v.get("error"))
except Exception as ex:
    print("bench line unreadable:", ex)
PY
