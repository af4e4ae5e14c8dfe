#!/usr/bin/env python
"""Goldens for the SURVEY.md section 8(f) rank-4 model families (Bottleneck ResNets, XVEC) from the REAL reference modules
(`wespeaker/models/resnet.py:72-107,223-260`, `wespeaker/models/tdnn.py:23-117`).  Build container only.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_f4.py

Same recipe as make_golden.py: namespace-bypass import, synthetic checkpoints loaded with strict=True (which pins the
key/shape specs of wespeaker_b200.synthetic), outputs stored in tests/golden/models_f4.npz; inputs are regenerated from
seeds at test time."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/wespeaker"


def main():
    from wespeaker_b200 import synthetic as syn
    for pkg, path in [("wespeaker", REF), ("wespeaker.models", REF + "/models")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    import wespeaker.models.resnet as resnet
    import wespeaker.models.tdnn as tdnn
    torch.set_num_threads(8)
    ctors = {"ResNet50": resnet.ResNet50, "ResNet101": resnet.ResNet101, "ResNet152": resnet.ResNet152,
             "ResNet221": resnet.ResNet221, "ResNet293": resnet.ResNet293, "XVEC": tdnn.XVEC}
    cases = [("ResNet50", 0, 2, 200), ("ResNet50", 0, 2, 99), ("ResNet101", 0, 1, 120), ("ResNet152", 0, 1, 64),
             ("ResNet221", 0, 1, 64), ("ResNet293", 0, 1, 48), ("XVEC", 0, 3, 200), ("XVEC", 0, 2, 61)]
    out = {}
    for name, seed, B, T in cases:
        model = ctors[name](**syn.DEFAULT_MODEL_ARGS[name])
        sd_np = syn.make_state_dict(name, seed)
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
        ref_keys = list(model.state_dict().keys())
        assert ref_keys == list(sd.keys()), (name, set(ref_keys) ^ set(sd.keys()))
        model.load_state_dict(sd, strict=True)
        model.eval()
        feats = torch.from_numpy(syn.make_feats(B, T, 80, seed=seed + 17 * T))
        with torch.no_grad():
            o = model(feats)
            o = o[-1] if isinstance(o, tuple) else o
        key = f"{name}__s{seed}_B{B}_T{T}"
        out[key] = o.numpy().astype(np.float32)
        print(f"{key}: emb {tuple(o.shape)} |e|={o.norm(dim=1).mean():.4f} params={sum(p.numel() for p in model.parameters())/1e6:.2f}M")
    np.savez_compressed(os.path.join(HERE, "models_f4.npz"), **out)


if __name__ == "__main__":
    main()
