"""Cosine scoring + S-norm / AS-norm.

CPU part: the numpy oracle against goldens produced by the reference's own score.py / score_norm.py
(tests/golden/make_golden_score.py).  GPU part: wespeaker_b200.score (ws_score.cu through the C ABI) against the oracle
and against the reference's output files, line by line."""
import os

import numpy as np
import pytest
import torch

from oracle import score_np
from wespeaker_b200 import kaldi_io, lib, score

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "score.npz"))
gpu = pytest.mark.gpu


def _rows(text):
    return [ln.split() for ln in str(text).strip().split("\n")]


def _same_lines(got, want, num_cols, tol):
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert len(a) == len(b)
        for c, (x, y) in enumerate(zip(a, b)):
            if c in num_cols:
                assert abs(float(x) - float(y)) <= tol[c], (a, b)
            else:
                assert x == y, (a, b)


NORM_COLS = {2: 5e-5, 4: 2e-4, 5: 2e-4, 6: 2e-4, 7: 2e-4}   # printed with 5 / 4 decimals by fp32 reference code


def _tables():
    utts, coh = [str(u) for u in G["utts"]], [str(c) for c in G["coh"]]
    return dict(zip(utts, G["evals"])), dict(zip(coh, G["cohort"])), {u: i for i, u in enumerate(utts)}


# ------------------------------------------------------------------------------------------------- CPU: oracle pinned
def test_oracle_get_mean_std_matches_reference():
    m, s = score_np.get_mean_std(G["evals"] - G["mean_vec"], G["cohort"] - G["mean_vec"], 30)
    assert np.abs(m - G["topn_mean"]).max() < 1e-6 and np.abs(s - G["topn_std"]).max() < 1e-6


@pytest.mark.parametrize("tag", ["mean", "nomean"])
def test_oracle_cosine_trials_match_reference_file(tag):
    _, _, idx = _tables()
    want = _rows(G[f"cos_{tag}"])
    sc = score_np.cosine_trials(G["evals"], [idx[r[0]] for r in want], [idx[r[1]] for r in want],
                                G["mean_vec"] if tag == "mean" else None)
    got = [[r[0], r[1], f"{v:.5f}", r[3]] for r, v in zip(want, sc)]
    _same_lines(got, want, {2}, {2: 2e-5})


@pytest.mark.parametrize("tag", ["mean", "nomean"])
@pytest.mark.parametrize("method", ["asnorm", "snorm"])
def test_oracle_score_norm_matches_reference_file(method, tag):
    ev, co, _ = _tables()
    lines = score_np.score_norm_lines(method, 30, _rows(G[f"cos_{tag}"]), ev, co, G["mean_vec"] if tag == "mean" else None)
    _same_lines([ln.split() for ln in lines], _rows(G[f"{method}_{tag}"]), set(NORM_COLS), NORM_COLS)


def test_score_host_helpers(tmp_path):
    """File-side helpers of score.py (no arithmetic on the device): mean vector of an scp, table reader."""
    scp = str(tmp_path / "x.scp")
    with kaldi_io.VectorWriter(str(tmp_path / "x.ark"), scp) as w:
        for u, v in zip(G["utts"][:7], G["evals"][:7]):
            w(str(u), v)
    m = score.calculate_mean_from_kaldi_vec(scp)                      # score.py:25-36
    assert m.dtype == np.float32 and np.allclose(m, G["evals"][:7].mean(axis=0), atol=1e-6)
    t = tmp_path / "t.txt"
    t.write_text("a b 0.5 target\n  c   d -1 nontarget \n")
    assert score.read_table(str(t)) == [["a", "b", "0.5", "target"], ["c", "d", "-1", "nontarget"]]
    got = dict(kaldi_io.load_scp_sequential(scp))
    assert list(got) == [str(u) for u in G["utts"][:7]] and np.array_equal(got[str(G["utts"][3])], G["evals"][3])


def test_score_needs_gpu_and_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(lib.B200Error):
        score.get_mean_std(G["evals"], G["cohort"], 30)


# ------------------------------------------------------------------------------------------------- GPU parity
@gpu
@pytest.mark.parametrize("top_n", [1, 30, 199, 200, 1000])
def test_gpu_get_mean_std_matches_oracle_and_reference(top_n):
    e, c = G["evals"] - G["mean_vec"], G["cohort"] - G["mean_vec"]   # cohort rows 3, 17, 101 are exact ties
    m, s = score.get_mean_std(e, c, top_n)
    mo, so = score_np.get_mean_std(e, c, top_n)
    assert np.abs(m - mo).max() < 1e-6 and np.abs(s - so).max() < 1e-6
    if top_n == 30:
        assert np.abs(m - G["topn_mean"]).max() < 1e-6 and np.abs(s - G["topn_std"]).max() < 1e-6


@gpu
def test_gpu_cohort_stats_large_and_tiled():
    rng = np.random.default_rng(0)
    e = rng.standard_normal((700, 192)).astype(np.float32)
    c = rng.standard_normal((3000, 192)).astype(np.float32)
    c[5] = c[4]
    mo, so = score_np.get_mean_std(e, c, 300)
    m, s = score.get_mean_std(e, c, 300)
    assert np.abs(m - mo).max() < 1e-6 and np.abs(s - so).max() < 1e-6
    dev = torch.device("cuda:0")
    eu, _ = score._unit_rows(e, None, dev)
    cu, _ = score._unit_rows(c, None, dev)
    m2, s2 = score._cohort_stats(eu, cu, 300, dev, tile_bytes=4 * 3000 * 97)   # 8 row tiles, ragged last tile
    assert np.array_equal(m2.cpu().numpy(), m) and np.array_equal(s2.cpu().numpy(), s)


@gpu
@pytest.mark.parametrize("tag", ["mean", "nomean"])
def test_gpu_score_files_match_reference(tmp_path, tag):
    utts, coh = [str(u) for u in G["utts"]], [str(c) for c in G["coh"]]
    eval_scp, coh_scp = str(tmp_path / "eval.scp"), str(tmp_path / "cohort.scp")
    with kaldi_io.VectorWriter(str(tmp_path / "eval.ark"), eval_scp) as w:
        for u, v in zip(utts, G["evals"]):
            w(u, v)
    with kaldi_io.VectorWriter(str(tmp_path / "cohort.ark"), coh_scp) as w:
        for u, v in zip(coh, G["cohort"]):
            w(u, v)
    mv = None
    if tag == "mean":
        mv = str(tmp_path / "mean_vec.npy")
        np.save(mv, G["mean_vec"])
    trials, trials2 = str(tmp_path / "trials.kaldi"), str(tmp_path / "trials_nolabel")
    with open(trials, "w") as f:
        for a, b, l in G["pairs"]:
            f.write(f"{a} {b} {l}\n")
    with open(trials2, "w") as f:
        for a, b, _ in G["pairs"][:20]:
            f.write(f"{a} {b}\n")
    sd = tmp_path / "scores"
    sd.mkdir()
    score.trials_cosine_score(eval_scp, str(sd), mv, (trials, trials2))
    _same_lines(_rows(open(sd / "trials.kaldi.score").read()), _rows(G[f"cos_{tag}"]), {2}, {2: 2e-5})
    _same_lines(_rows(open(sd / "trials_nolabel.score").read()), _rows(G[f"cos_nolabel_{tag}"]), {2}, {2: 2e-5})
    # normalise the REFERENCE's cosine score file so both sides start from identical inputs (score_norm.py reads text)
    ref_scores = str(tmp_path / "ref.score")
    open(ref_scores, "w").write(str(G[f"cos_{tag}"]))
    for method in ("asnorm", "snorm"):
        dst = str(tmp_path / f"norm_{method}")
        score.score_norm_main(method, 30, ref_scores, dst, coh_scp, eval_scp, mv)
        _same_lines(_rows(open(dst).read()), _rows(G[f"{method}_{tag}"]), set(NORM_COLS), NORM_COLS)
