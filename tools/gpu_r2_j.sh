#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "plda or score or embedding_processing" 2>&1 | tail -2
timeout -k 10 300 python - <<'PY'
import sys, os
sys.path.insert(0, '.')
import torch, numpy as np
import bench
from wespeaker_b200 import parallel
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
peaks = bench.measured_peaks()
for simt in (0, 1):
    if simt: os.environ["WS_PLDA_SIMT"] = "1"
    r = bench.plda_leg(dev, 0, 1, parallel, peaks, n_enroll=262144, cpu_loop=False)
    print("simt" if simt else "dmma", {k: r[k] for k in ("value", "ms_total", "tflops_f64", "max_abs_err_vs_fp64_oracle", "parity_ok")}, r["roofline"]["frac"], r["roofline"]["peak"])
PY
} > gpurun_out/r2j.log 2>&1
cat gpurun_out/r2j.log
