"""Per-op device times of one plan in sequence context (warm L2), via ws_engine_profile_ops; writes a markdown table."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wespeaker_b200 import lib, synthetic as syn
from wespeaker_b200.models import from_synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "ECAPA_TDNN_c1024"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
T = int(sys.argv[4]) if len(sys.argv) > 4 else 200
m = from_synthetic(name, 0, precision=prec).to("cuda:0")
m.embed(torch.from_numpy(syn.make_feats(B, T, 80, seed=1)).cuda())
L = lib.load()
buf = (C.c_float * 2048)()
n = L.ws_engine_profile_ops(m._engine, B, T, 10, buf, 2048)
assert n > 0, L.ws_last_error()
t = np.array(buf[:n]) * 1e3
print(f"### {name} {prec} B={B} T={T}: {n} ops, sum {t.sum():.1f} us (events between ops, warm L2, no graph)\n")
print("| # | op | us | share | TFLOP/s |\n|---|---|---|---|---|")
fl = C.c_double(0.0)
tot_fl = 0.0
for i in range(n):
    nm = L.ws_engine_plan_op_name(m._engine, B, T, i, C.byref(fl))
    nm = nm.decode() if nm else "?"
    tf = f"{fl.value / (t[i] * 1e-6) / 1e12:.0f}" if fl.value > 0 else ""
    tot_fl += fl.value
    print(f"| {i} | {nm} | {t[i]:.1f} | {100 * t[i] / t.sum():.1f}% | {tf} |")
print(f"\nGEMM-like FLOPs {tot_fl / 1e9:.1f} G -> {tot_fl / (t.sum() * 1e-6) / 1e12:.0f} TFLOP/s over the op sum\n")
