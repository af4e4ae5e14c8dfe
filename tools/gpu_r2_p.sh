#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "ragged or length_masked_wav" -s 2>&1 | grep -v "^$" | tail -30
timeout -k 10 600 python - <<'PY'
import json, torch, numpy as np, bench
from wespeaker_b200 import parallel
dev = torch.device("cuda", 0)
peaks = bench.measured_peaks()
r = bench.varlen_leg(20, 3, dev, 0, 1, parallel, peaks, None)
print(json.dumps(r, indent=1))
PY
} > gpurun_out/r2p.log 2>&1
cut -c1-300 gpurun_out/r2p.log | tail -80
