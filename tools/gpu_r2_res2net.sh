#!/bin/bash
# round-2 (second session) GPU check: the Res2Net / ERes2Net tests first, then the whole GPU suite, then the default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t0=$(date +%s)
timeout -k 5 170 python -m pytest tests/test_res2net.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02b_res2net.log 2>&1; echo "res2net tests exit $? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/r02b_res2net.log | cut -c1-300
timeout -k 5 ${FULL_S:-330} python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_res2net.py --durations=15 > gpurun_out/r02b_tests.log 2>&1; echo "full suite exit $? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r02b_tests.log | cut -c1-300
timeout -k 5 ${BENCH_S:-170} python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err; echo "bench exit $? ($(( $(date +%s) - t0 )) s)"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02b_bench_default.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"])
    for k, v in d.get("configs", {}).items():
        print(k, round(v["value"], 1), v.get("ms_per_step", v.get("ms_per_pass")), v.get("parity_rel_l2", v.get("max_abs_err_vs_fp64_oracle")))
except Exception as ex:
    print("bench line unreadable:", ex)
PY
