#!/bin/bash
# full GPU test suite + bench + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
stage() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -${TAILN:-4} gpurun_out/$name.log; }
TMO=1500 stage t_all python -m pytest tests -m gpu -q -p no:cacheprovider
TMO=900 TAILN=2 stage bench python bench.py --steps 10 --warmup 3 ${BENCH_ARGS}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"
