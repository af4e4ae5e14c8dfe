#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
stage() { name=$1; shift; echo "=== $name"; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -${TAILN:-4} gpurun_out/$name.log | cut -c1-${CUT:-900}; }
TMO=900 TAILN=12 stage t_new python -m pytest tests -m gpu -q -s -k "config4 or dropin" -p no:cacheprovider
for wl in resnet34_fp16_b64 campplus_bf16_b64 ecapa512_bf16_b256; do
  TMO=600 TAILN=1 CUT=700 stage bench_$wl python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-plda --workload $wl
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_resnet.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plda --workload resnet34_fp16_b64 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_campp.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-plda --workload campplus_bf16_b64 > /dev/null 2>&1
echo done
