"""Res2Net / ERes2Net: the oracle pinned on goldens from the real reference modules
(tests/golden/make_golden_res2net.py), the plan builder checked without a device, and - on the GPU - the engine against
the same goldens per precision."""
import os

import numpy as np
import pytest
import torch

from oracle import models_torch
from wespeaker_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
with np.load(os.path.join(HERE, "golden", "models_res2net.npz")) as _z:
    G = {k: _z[k] for k in _z.files}


def parse_case(key):
    name, rest = key.split("__")
    s, b, t, g = rest.split("_")
    return name, int(s[1:]), int(b[1:]), int(t[1:]), float(g[1:])


def case_inputs(key):
    name, seed, B, T, gain = parse_case(key)
    return name, syn.make_state_dict(name, seed), syn.make_feats(B, T, 80, seed=seed + 17 * T) * np.float32(gain)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)


@pytest.mark.parametrize("key", sorted(G))
def test_oracle_matches_reference(key):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    name, sd, feats = case_inputs(key)
    emb = models_torch.forward(name, sd, feats).numpy()
    assert rel_l2(emb, G[key]).max() < 5e-6, key   # fp32 restatement vs the fp32 reference modules


def test_param_counts():
    # parameter counts printed by the reference modules' own __main__ blocks (res2net.py:224-226, eres2net.py:441-443)
    def nparams(name):
        return sum(int(np.prod(s)) for k, s in syn.state_dict_spec(name, **syn.DEFAULT_MODEL_ARGS[name]).items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert abs(nparams("Res2Net34_Base") / 1e6 - 4.69) < 0.01
    assert abs(nparams("ERes2Net34_Base") / 1e6 - 9.89) < 0.01
