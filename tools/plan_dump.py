#!/usr/bin/env python
"""Print the launch plan the engine builds for a model and input shape - on any host, no GPU needed (plan-check engine:
nothing is computed).  Also the quickest way to see whether a checkpoint fits the engine.

    python tools/plan_dump.py MODEL [PRECISION [BATCH [FRAMES [CHECKPOINT.pt]]]]
    python tools/plan_dump.py ERes2Net34_Base fp16 64 200
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wespeaker_b200.models import from_synthetic, get_speaker_model, load_checkpoint  # noqa: E402
from wespeaker_b200.synthetic import DEFAULT_MODEL_ARGS  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ECAPA_TDNN_c1024"
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 200
    if len(sys.argv) > 5:
        m = get_speaker_model(name)(precision=prec, **DEFAULT_MODEL_ARGS[name])
        load_checkpoint(m, sys.argv[5])
    else:
        m = from_synthetic(name, precision=prec)       # random weights of the reference architecture
    ops = m.plan_check(B, T)
    tot = sum(f for _, f in ops)
    print(f"{name} {prec}, batch {B} x {T} frames: {len(ops)} ops, {tot / B / 1e9:.3f} GFLOP per utterance in GEMM-like ops")
    for i, (label, fl) in enumerate(ops):
        print(f"{i:4d}  {label}" + (f"   [{fl / 1e9:.2f} GFLOP]" if fl else ""))
    kinds = collections.Counter(label.split()[0] for label, _ in ops)
    print("by kind:", ", ".join(f"{k} x{v}" for k, v in kinds.most_common()))


if __name__ == "__main__":
    main()
