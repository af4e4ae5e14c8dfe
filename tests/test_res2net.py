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


# ------------------------------------------------------------------------------------------ plan builder, no device
from plan_interp import run_plan  # noqa: E402  (tests/ is on sys.path under pytest's rootdir conftest)
from wespeaker_b200.models import from_synthetic  # noqa: E402

PLAN_CASES = [("Res2Net34_Base", "fp32", 1, 48, 1.0), ("Res2Net34_Base", "bf16", 2, 57, 6.0), ("Res2Net34_Large", "fp16", 1, 40, 1.0),
              ("ERes2Net34_Base", "tf32x3", 1, 48, 1.0), ("ERes2Net34_Base", "bf16", 1, 56, 6.0), ("ERes2Net34_Large", "bf16", 1, 40, 3.0),
              ("ERes2Net34_aug", "bf16", 1, 40, 1.0)]


@pytest.mark.parametrize("name,prec,B,T,gain", PLAN_CASES)
def test_plan_arithmetic_matches_oracle_on_the_host(name, prec, B, T, gain, tmp_path):
    """The launch plan the engine builds for these families (channel padding to 32/64/128, split / cat as channel slices,
    summed chain inputs as repeated-weight K ranges, AFF, merged shortcuts, folded BN, Hardtanh) re-evaluated on the host
    from ws_engine_plan_trace equals the oracle - for the fp32 (FFMA), 3xTF32 and 16-bit (halo-resident 3x3 kernel) plan
    variants, which differ in the ops they contain.  No device, nothing computed by the library."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    m = from_synthetic(name, precision=prec)
    path = str(tmp_path / "plan.bin")
    m.plan_trace(path, B, T)
    feats = syn.make_feats(B, T, 80, seed=5) * np.float32(gain)
    emb, meta = run_plan(path, feats)
    assert meta["model"] == name and meta["precision"] == prec
    kinds = {op["trace"]["kind"] for op in meta["ops"]}
    assert ("conv3x3" in kinds) == (prec in ("bf16", "fp16")) and ("aff_combine" in kinds) == name.startswith("ERes2Net")
    ref = models_torch.forward(name, syn.make_state_dict(name, 0), feats).numpy()
    assert rel_l2(emb, ref).max() < 2e-5, rel_l2(emb, ref)   # float64 re-evaluation vs the fp32 oracle
