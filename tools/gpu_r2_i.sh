#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "campplus_fused or config4 or batch_invariance" 2>&1 | tail -2
WS_CAM_PROF=1 timeout -k 10 300 python -c "
import torch, sys
sys.path.insert(0,'.')
from wespeaker_b200.models import from_synthetic
from wespeaker_b200 import synthetic as syn
m = from_synthetic('CAMPPlus', 0, precision='bf16').to('cuda:0')
m.set_option('cuda_graph', 0)
x = torch.from_numpy(syn.make_feats(64, 200, 80, seed=1)).cuda()
m.embed(x); torch.cuda.synchronize()
print('---- second pass', file=sys.stderr)
m.embed(x); torch.cuda.synchronize()
" 2>&1 | awk '/second pass/{f=1} f' | sed -n '2,3p;36,37p;52,53p'
timeout -k 10 300 python - <<'PY'
import os, sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from wespeaker_b200 import lib, synthetic as syn
from wespeaker_b200.models import from_synthetic
for B in (64, 148):
    for blk in (0, 1):
        m = from_synthetic("CAMPPlus", 0, precision="bf16").to("cuda:0")
        m.set_option("cam_block", blk)
        x = torch.from_numpy(syn.make_feats(B, 200, 80, seed=1)).cuda()
        for _ in range(3): m.embed(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): m.embed(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        buf = (C.c_float * 1024)()
        n = lib.load().ws_engine_profile_ops(m._engine, B, 200, 3, buf, 1024)
        t = np.array(buf[:n]) * 1e3
        print(f"CAMPPlus bf16 B{B} cam_block={blk}: graph step {ms*1e3:.0f} us ({B/ms*1e3:.0f} utt/s), {n} ops, op sum {t.sum():.0f} us: " + " ".join(f"{v:.0f}" for v in t))
PY
} > gpurun_out/r2i.log 2>&1
cut -c1-700 gpurun_out/r2i.log
