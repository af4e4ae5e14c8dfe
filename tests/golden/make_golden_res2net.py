#!/usr/bin/env python
"""Goldens for the Res2Net / ERes2Net families from the REAL reference modules
(`wespeaker/models/res2net.py`, `wespeaker/models/eres2net.py`).  Build container only.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_res2net.py

Same recipe as make_golden_f4.py: namespace-bypass import, synthetic checkpoints loaded with strict=True (which pins the
key/shape specs of wespeaker_b200.synthetic.res2net_spec), outputs stored in tests/golden/models_res2net.npz; inputs are
regenerated from seeds at test time.  The "hot" cases multiply the input features by 6 so that the families' Hardtanh(0, 20)
"ReLU" (`eres2net.py:43-52`) actually clips (the fraction of clipped stem-block outputs is printed)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/wespeaker"

# (model, seed, B, T, input gain)
CASES = [("Res2Net34_Base", 0, 2, 200, 1.0), ("Res2Net34_Base", 1, 2, 99, 6.0), ("Res2Net34_Large", 0, 1, 120, 1.0),
         ("ERes2Net34_Base", 0, 2, 200, 1.0), ("ERes2Net34_Base", 1, 2, 99, 6.0), ("ERes2Net34_Large", 0, 1, 120, 1.0),
         ("ERes2Net34_aug", 0, 1, 104, 1.0)]


def case_key(name, seed, B, T, gain):
    return f"{name}__s{seed}_B{B}_T{T}_g{int(gain)}"


def main():
    from wespeaker_b200 import synthetic as syn
    for pkg, path in [("wespeaker", REF), ("wespeaker.models", REF + "/models")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    import wespeaker.models.eres2net as eres2net
    import wespeaker.models.res2net as res2net
    torch.set_num_threads(8)
    out = {}
    for name, seed, B, T, gain in CASES:
        mod = eres2net if name.startswith("ERes2Net") else res2net
        model = getattr(mod, name)(**syn.DEFAULT_MODEL_ARGS[name])
        sd_np = syn.make_state_dict(name, seed)
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
        ref_keys = list(model.state_dict().keys())
        assert ref_keys == list(sd.keys()), (name, set(ref_keys) ^ set(sd.keys()))
        model.load_state_dict(sd, strict=True)
        model.eval()
        feats = torch.from_numpy(syn.make_feats(B, T, 80, seed=seed + 17 * T)) * gain
        clipped = []
        hooks = [blk.register_forward_hook(lambda _m, _i, o: clipped.append(float((o >= 20.0).float().mean())))
                 for layer in (model.layer1, model.layer2, model.layer3, model.layer4) for blk in layer]
        with torch.no_grad():
            o = model(feats)
            o = o[-1] if isinstance(o, tuple) else o
        for h in hooks:
            h.remove()
        key = case_key(name, seed, B, T, gain)
        out[key] = o.numpy().astype(np.float32)
        print(f"{key}: emb {tuple(o.shape)} |e|={o.norm(dim=1).mean():.4f} params={sum(p.numel() for p in model.parameters())/1e6:.2f}M "
              f"clipped at 20: max over blocks {max(clipped):.4f}, mean {np.mean(clipped):.5f}")
    np.savez_compressed(os.path.join(HERE, "models_res2net.npz"), **out)


if __name__ == "__main__":
    main()
