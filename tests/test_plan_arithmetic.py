"""CPU: the launch plans of the hot-path families re-evaluated on the host (ws_engine_plan_trace + tests/plan_interp.py).

A plan-check engine builds the same launch plan a real engine builds (same builder code, placeholder addresses, no device,
nothing computed by the library) and writes it out as data.  Re-evaluating that data in float64 checks everything the
plan BUILDER decides - BN folding, weight packing for the fused Res2 / ASTP / CAM++ kernels, torch.cat as channel slices,
global-context attention as a per-utterance bias row, merged shortcuts, parity planes, TSTP index order - against the oracle;
the kernels that execute the same plans are checked on the GPU (tests/test_gpu_parity.py)."""
import os

import numpy as np
import pytest
import torch

import plan_interp
from oracle import models_torch
from wespeaker_b200 import synthetic as syn
from wespeaker_b200.models import from_synthetic


def _run(name, prec, B, T, opts=None, seed=5, round_as=None, model_args=None):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model_args = model_args or {}
    m = from_synthetic(name, precision=prec, **model_args)
    for k, v in (opts or {}).items():
        m.set_option(k, v)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "plan.bin")
        m.plan_trace(path, B, T)
        feats = syn.make_feats(B, T, 80, seed=seed)
        plan_interp.ROUND = round_as
        try:
            emb, meta = plan_interp.run_plan(path, feats)
        finally:
            plan_interp.ROUND = None
    ref = models_torch.forward(name, syn.make_state_dict(name, 0, **model_args), feats, **model_args).numpy()
    rel = np.linalg.norm(emb - ref, axis=1) / np.linalg.norm(ref, axis=1)
    return rel.max(), {op["trace"]["kind"] for op in meta["ops"]}


CASES = [
    ("ECAPA_TDNN_c1024", "bf16", 1, 300, None, {"res2_fused", "astp_fused", "se_gate", "scale_residual"}),   # the benchmarked plan (T > 256: tiled Res2)
    ("ECAPA_TDNN_c512", "bf16", 2, 150, None, {"res2_fused", "astp_fused"}),
    ("ECAPA_TDNN_c512", "bf16", 2, 150, {"res2_fused": 0, "astp_fused": 0, "se_fused": 0}, {"astp_stats", "tstats"}),   # the unfused cross-check plan
    ("ECAPA_TDNN_GLOB_c512", "fp32", 2, 99, None, {"astp_stats", "tstats"}),          # global context as a per-utterance bias row
    ("ECAPA_TDNN_GLOB_c1024", "tf32x3", 1, 120, None, {"se_gate"}),
    ("CAMPPlus", "bf16", 2, 260, None, {"cam_dense", "conv3x3", "bnrelu"}),            # whole dense blocks per launch, 3 context segments
    ("CAMPPlus", "fp32", 1, 455, None, {"cam_gate", "bnrelu"}),                       # unfused dense layers
    ("ResNet18", "bf16", 1, 48, None, {"conv3x3", "stem"}),
]


@pytest.mark.parametrize("name,prec,B,T,opts,must_have", CASES)
def test_hot_path_plan_arithmetic_matches_oracle(name, prec, B, T, opts, must_have):
    rel, kinds = _run(name, prec, B, T, opts)
    assert must_have <= kinds, (kinds, must_have)
    assert rel < 5e-6, rel          # float64 re-evaluation vs the fp32 oracle (measured 2e-7 .. 5e-7)


@pytest.mark.parametrize("name,prec,B,T,bar,gpu_measured", [
    ("ECAPA_TDNN_c1024", "bf16", 2, 200, 1e-2, 2.6e-3), ("ResNet34", "fp16", 2, 200, 1e-2, 3.6e-4), ("CAMPPlus", "bf16", 1, 455, 1e-2, 2.2e-3)])
def test_16bit_storage_roundings_explain_the_16bit_distance(name, prec, B, T, bar, gpu_measured):
    """Rounding activations and conv weights to the plan's 16-bit storage type at the points where the kernels round them
    (plan_interp.ROUND) predicts the distance of the 16-bit paths from the oracle: the prediction stays inside the GPU
    tests' bar and within 2x of what the B200 measured at bench size (profiles/r02_bench_full_ecapa1024_bf16.json) - the
    16-bit deviation is the arithmetic of 16-bit storage, not a property of the kernels."""
    rel, _ = _run(name, prec, B, T, None, seed=23, round_as=prec)
    assert rel < bar and 0.5 * gpu_measured < rel < 2.0 * gpu_measured, (rel, gpu_measured)


@pytest.mark.parametrize("name,prec,B,T,model_args", [
    ("ResNet34", "fp16", 1, 64, dict(two_emb_layer=True)),          # seg_bn_1 (affine=False) folded into seg_2
    ("ECAPA_TDNN_c512", "bf16", 1, 99, dict(emb_bn=True)),           # bn2 folded into the embedding linear
    ("ERes2Net34_Base", "fp16", 1, 48, dict(two_emb_layer=True)),
    ("ResNet221", "bf16", 1, 40, {}),                                # 73 Bottleneck blocks
])
def test_model_arg_variants_and_deep_plans(name, prec, B, T, model_args):
    rel, _ = _run(name, prec, B, T, model_args=model_args)
    assert rel < 5e-6, rel


@pytest.mark.parametrize("name,prec,lens", [
    ("ResNet34", "fp16", [90, 37, 64, 21]), ("ResNet18", "fp32", [90, 37, 64, 21]), ("ResNet50", "bf16", [90, 37, 64, 21]),
    ("Res2Net34_Base", "fp16", [90, 37, 64, 21]), ("ERes2Net34_Base", "fp16", [90, 37, 64, 21]), ("ERes2Net34_Base", "tf32x3", [90, 37, 64, 21]),
    ("ECAPA_TDNN_c512", "bf16", [200, 137, 64, 256]),        # fused Res2 chain / ASTP tail / SE gate over each utterance's own frames
    ("ECAPA_TDNN_c1024", "bf16", [301, 137]),                # tiled Res2 chain (T > 256)
    ("ECAPA_TDNN_GLOB_c512", "fp32", [90, 37, 64, 21]),      # global-context statistics per utterance
    ("CAMPPlus", "bf16", [455, 137, 200, 301]),              # context segments (ceil mode) of each utterance's own length
])
def test_length_masked_plans_equal_each_utterance_alone(name, prec, lens, tmp_path):
    """The length-masked plans (ws_engine_forward_masked): where the builder zeroes the rows behind an utterance's end, which
    stride level's frame counts every op gets, statistics over the utterance's own frames.  Utterances of different lengths
    padded with large garbage into one batch, re-evaluated on the host, must equal the ORACLE run on each utterance alone,
    unpadded."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    lens = np.array(lens)
    B, T = len(lens), int(lens.max())
    g = np.random.default_rng(1)
    feats = syn.make_feats(B, T, 80, seed=9)
    for b in range(B):
        feats[b, lens[b]:] = 50.0 * g.standard_normal((T - lens[b], 80))
    m = from_synthetic(name, precision=prec)
    path = str(tmp_path / "plan.bin")
    m.plan_trace(path, B, T, masked=True)
    emb, meta = plan_interp.run_plan(path, feats, n_frames=lens)
    assert meta["ops"][0]["trace"]["kind"] == "lens_derive"
    if name != "CAMPPlus":      # (CAM++'s 16-bit kernels all take the frame counts themselves)
        assert any(op["trace"]["kind"] == "zero_tail" for op in meta["ops"])
    sd = syn.make_state_dict(name, 0)
    for b in range(B):
        ref = models_torch.forward(name, sd, feats[b:b + 1, :lens[b]]).numpy()[0]
        assert np.linalg.norm(emb[b] - ref) / np.linalg.norm(ref) < 5e-6, (name, b)


def test_tf32x3_error_is_the_truncating_accumulation_not_the_split(tmp_path):
    """tools/tf32x3_error_budget.py in small: on the 3xTF32 plan of ECAPA-TDNN-512 the operand split alone (x_lo*W + x*W_lo + x*W
    with tf32-truncated operands, exact accumulation) stays at fp32 level, while truncating every fp32 accumulate of a K = 8 MMA
    step - what tensor cores do - lands where the B200 measures the path (2.3e-5 at 200 frames)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tf32x3_error_budget", os.path.join(root, "tools", "tf32x3_error_budget.py"))
    eb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(eb)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    name, T = "ECAPA_TDNN_c512", 64
    m = from_synthetic(name, precision="tf32x3")
    path = str(tmp_path / "plan.bin")
    m.plan_trace(path, 1, T)
    feats = syn.make_feats(1, T, 80, seed=17 * T)
    ref = models_torch.forward(name, syn.make_state_dict(name, 0), torch.from_numpy(feats).double()).numpy()

    def rel(emb):
        return float(np.linalg.norm(emb - ref) / np.linalg.norm(ref))
    orig = plan_interp.run_conv
    try:
        plan_interp.ROUND = "tf32x3_trunc_lo"
        split = rel(plan_interp.run_plan(path, feats)[0])
        plan_interp.ROUND = None
        plan_interp.run_conv = eb.make_run_conv(eb.rn32)
        rn = rel(plan_interp.run_plan(path, feats)[0])
        plan_interp.run_conv = eb.make_run_conv(eb.rz32)
        rz = rel(plan_interp.run_plan(path, feats)[0])
    finally:
        plan_interp.run_conv = orig
        plan_interp.ROUND = None
    assert split < 6e-6 and rn < 6e-6, (split, rn)
    assert 5 * split < rz < 1e-4, (split, rz)
