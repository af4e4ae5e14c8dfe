"""Per-op device times of one plan in sequence context (warm L2), via ws_engine_profile_ops."""
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
buf = (C.c_float * 1024)()
n = lib.load().ws_engine_profile_ops(m._engine, B, T, 5, buf, 1024)
assert n > 0, lib.load().ws_last_error()
t = np.array(buf[:n]) * 1e3
print(f"{name} {prec} B{B} T{T}: {n} ops, total {t.sum():.1f} us")
print(" ".join(f"{x:.0f}" for x in t))
