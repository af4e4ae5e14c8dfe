#!/usr/bin/env python
"""Where the 3xTF32 path's distance from the fp32 reference comes from (CPU only, no GPU needed).

Re-evaluates the tf32x3 launch plan of a model on the host (ws_engine_plan_trace + tests/plan_interp.py) under three
arithmetic models of the tensor core and prints the rel-L2 distance of the embedding from the fp64 oracle:

  split only       operands truncated to tf32, x_lo*W + x*W_lo + x*W, exact accumulation   -> the cost of the operand split
  + RN accumulate  the same products added to an fp32 accumulator per K = 8 step, round-to-nearest
  + RZ accumulate  ... with every accumulate TRUNCATED toward zero, as tensor cores do

Measured on the B200 (tests / bench): ECAPA-TDNN-512 2.3e-5, ERes2Net34_Base 3.2e-5, ERes2Net34_aug 1.4e-4.  The split alone
predicts 2.4e-6 and RN accumulation adds nothing; truncating accumulation predicts 3.1e-5 for ECAPA-TDNN-512: the path's error
is the tensor core's fp32 accumulation (K/8 truncating adds per output and pass, a coherent shrink of every dot product), not
the 3xTF32 operand split.  The remedy is a kernel change (round-robin the k-blocks over 2-4 TMEM accumulators and add them
with round-to-nearest in the epilogue: the bias falls with the number of accumulators), not a different split.

    python tools/tf32x3_error_budget.py [MODEL [FRAMES]]
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plan_interp  # noqa: E402
from oracle import models_torch  # noqa: E402
from wespeaker_b200 import synthetic as syn  # noqa: E402
from wespeaker_b200.models import from_synthetic  # noqa: E402

KSTEP = 8   # K of one tcgen05.mma kind::tf32 instruction


def rz32(a):
    f = a.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(a)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float64)


def rn32(a):
    return a.astype(np.float32).astype(np.float64)


def make_run_conv(accumulate):
    def split(a):
        a32 = np.ascontiguousarray(a, dtype=np.float32)
        hi = plan_interp._tf32_trunc(a32)
        return hi.astype(np.float64), plan_interp._tf32_trunc(a32 - hi).astype(np.float64)

    def run_conv(mem, tr):
        es = tr["es"]
        B, F, T, Cout, K = tr["B"], tr["F"], tr["T"], tr["Cout"], tr["Ktot"]
        W = mem.vec(tr["W"], Cout * K).reshape(Cout, K)
        srcs = [mem.strided(s["p"], es, (s["B"], s["F"], s["T"], s["C"]), (s["sB"], s["sF"], s["sT"], 1)) for s in tr["src"]]
        acc = np.zeros((B, F, T, Cout))
        for si, c0, dt, df, wk, nch in tr["taps"]:
            xh, xl = split(plan_interp.shifted(srcs[si][..., c0:c0 + nch], df, dt, F, T))
            wh, wl = split(W[:, wk:wk + nch].T)
            for k0 in range(0, nch, KSTEP):
                s = slice(k0, k0 + KSTEP)
                for a, b in ((xl, wh), (xh, wl), (xh, wh)):      # the kernel's pass order within a k-block
                    acc = accumulate(acc + a[..., s] @ b[s])
        if tr["bias"]:
            acc = acc + mem.vec(tr["bias"], Cout)
        if tr["rowbias"]:
            acc = acc + mem.strided(tr["rowbias"], 4, (B, Cout), (tr["rowbias_ld"], 1))[:, None, None, :]
        acc = plan_interp.act(acc, tr["act1"])
        if tr["scale"]:
            acc = acc * mem.vec(tr["scale"], Cout) + mem.vec(tr["shift"], Cout)
        if tr["res"]:
            acc = acc + mem.view(dict(p=tr["res"], B=B, F=F, T=T, C=Cout, ld=tr["res_ld"]), es)
        acc = rn32(plan_interp.act(acc, tr["act2"]))
        mem.view(dict(p=tr["out"], B=B, F=F, T=T, C=Cout, ld=tr["out_ld"]), es, write=True)[...] = acc
        if tr["out2"]:
            add2 = mem.view(dict(p=tr["add2"], B=B, F=F, T=T, C=Cout, ld=tr["add2_ld"]), es)
            mem.view(dict(p=tr["out2"], B=B, F=F, T=T, C=Cout, ld=tr["out2_ld"]), es, write=True)[...] = rn32(acc + add2)
    return run_conv


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ECAPA_TDNN_c512"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    torch.set_num_threads(os.cpu_count() or 1)
    m = from_synthetic(name, precision="tf32x3")
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "plan.bin")
        m.plan_trace(path, 1, T)
        feats = syn.make_feats(1, T, 80, seed=17 * T)
        ref = models_torch.forward(name, syn.make_state_dict(name, 0), torch.from_numpy(feats).double()).numpy()
        orig = plan_interp.run_conv
        rows = []
        try:
            plan_interp.ROUND = "tf32x3_trunc_lo"
            emb, _ = plan_interp.run_plan(path, feats)
            plan_interp.ROUND = None
            rows.append(("split only (exact accumulation)", emb))
            for label, fn in (("+ fp32 accumulate per K=8 step, round to nearest", rn32), ("+ fp32 accumulate per K=8 step, TRUNCATED (tensor core)", rz32)):
                plan_interp.run_conv = make_run_conv(fn)
                t0 = time.time()
                emb, _ = plan_interp.run_plan(path, feats)
                rows.append((label + f"  [{time.time() - t0:.0f} s]", emb))
        finally:
            plan_interp.run_conv = orig
            plan_interp.ROUND = None
    print(f"{name}, 1 x {T} frames, tf32x3 plan: rel-L2 of the embedding to the fp64 oracle")
    for label, emb in rows:
        print(f"  {label:70s} {float(np.linalg.norm(emb - ref) / np.linalg.norm(ref)):.3e}")


if __name__ == "__main__":
    main()
