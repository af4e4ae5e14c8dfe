#!/bin/bash
# round-2 end-of-round evidence (final): tests, smoke, bench, reference arm, op tables, ncu of the fused ASTP kernel, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r02_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r02_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench exit $?"
timeout -k 10 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref exit $?"
{
timeout -k 10 300 python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200
timeout -k 10 300 python tools/op_times.py ECAPA_TDNN_c512 bf16 256 200
timeout -k 10 300 python tools/op_times.py ECAPA_TDNN_GLOB_c512 bf16 256 200
timeout -k 10 300 python tools/op_times.py ResNet34 fp16 64 200
timeout -k 10 300 python tools/op_times.py CAMPPlus bf16 64 200
} > gpurun_out/r02_op_times_final.md 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:ws_astp_fused -s 1 -c 1 -o gpurun_out/r02_full_astp python tools/op_times.py ECAPA_TDNN_c1024 bf16 256 200 > gpurun_out/r02_full_astp.log 2>&1; echo "ncu astp $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_ecapa.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-plda --no-configs --sustained-s 0 > /dev/null 2>&1; echo "ll ecapa $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], d["roofline"].get("step_frac_of_sustained"))
print("parity", d["parity"]["parity_rel_l2"], "sustained", d.get("sustained", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
for k, v in d.get("configs", {}).items():
    print(k, round(v["value"], 1), v.get("ms_per_step", v.get("ms_per_pass")), v.get("step_frac_of_sustained"), v.get("parity_rel_l2", v.get("max_abs_err_vs_fp64_oracle")))
PY
