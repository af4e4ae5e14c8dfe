"""`TwoCovPLDA.eval_sv` + Kaldi `<Plda>` model IO.

Goldens: the reference's own `eval_sv` score FILES for both `multisession_avg` values, with / without `indomain_scp`,
with / without length normalisation, and the reference `read_plda` parse of two binary `<Plda>` files
(tests/golden/make_golden_plda_eval.py).  CPU: the numpy oracle and the `<Plda>` reader against them.  GPU: the device
`eval_sv` (ws_plda.cu through the C ABI) against the reference's files line by line."""
import os

import numpy as np
import pytest

from oracle import plda_np
from wespeaker_b200 import kaldi_io, synthetic as syn
from wespeaker_b200.plda import TwoCovPLDA

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "plda_eval.npz"))
gpu = pytest.mark.gpu
CASES = [(nl, avg, ind) for nl in (1, 0) for avg in (1, 0) for ind in (1, 0)]
# scores are printed with 5 decimals; north-star bar for PLDA scores: 1e-5 * max(1, |s|)
TOL = lambda s: 1e-5 * max(1.0, abs(s)) + 0.6e-5   # noqa: E731  (+ the print rounding of both sides)


def _rows(text):
    return [ln.split() for ln in str(text).strip().split("\n")]


def _check(got_rows, want_rows):
    assert len(got_rows) == len(want_rows)
    worst = 0.0
    for a, b in zip(got_rows, want_rows):
        assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3], (a, b)
        d = abs(float(a[2]) - float(b[2]))
        assert d <= TOL(float(b[2])), (a, b)
        worst = max(worst, d)
    return worst


def _write_inputs(td):
    def write(name, keys, vecs):
        with kaldi_io.VectorWriter(os.path.join(td, name + ".ark"), os.path.join(td, name + ".scp")) as w:
            for k, v in zip(keys, vecs):
                w(str(k), v)
        return os.path.join(td, name + ".scp")
    e = write("enroll", G["enroll_keys"], G["enroll_vecs"])
    t = write("test", G["test_keys"], G["test_vecs"])
    i = write("indomain", [f"in{j}" for j in range(len(G["indomain_vecs"]))], G["indomain_vecs"])
    u2s = os.path.join(td, "utt2spk")
    with open(u2s, "w") as f:
        for k, s in zip(G["enroll_keys"], G["enroll_spk"]):
            if str(s):
                f.write(f"{k} {s}\n")
    tr = os.path.join(td, "trials")
    with open(tr, "w") as f:
        for a, b, l in G["trials"]:
            f.write(f"{a} {b} {l}\n")
    return e, u2s, t, tr, i


# ------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("nl,avg,ind", CASES)
def test_oracle_eval_sv_matches_reference_file(nl, avg, ind):
    pm = syn.make_plda(256, seed=3, normalize_length=bool(nl))
    mean_vec = G["indomain_vecs"].mean(0) if ind else None
    sess, spks = {}, []
    for k, s, v in zip(G["enroll_keys"], G["enroll_spk"], G["enroll_vecs"]):
        if str(s):
            if str(s) not in sess:
                spks.append(str(s))
            sess.setdefault(str(s), []).append(v)
    e_t, counts = plda_np.prepare_enroll(pm, [sess[s] for s in spks], mean_vec, multisession_avg=bool(avg))
    t_t = plda_np.prepare_test(pm, G["test_vecs"], mean_vec)
    si, ti = {s: i for i, s in enumerate(spks)}, {str(k): i for i, k in enumerate(G["test_keys"])}
    want = _rows(G[f"scores_nl{nl}_avg{avg}_ind{ind}"])
    got = [[r[0], r[1], "{:.5f}".format(plda_np.log_likelihood_ratio(pm, e_t[si[r[0]]], t_t[ti[r[1]]], counts[si[r[0]]])), r[3]]
           for r in want]
    _check(got, want)


def test_read_kaldi_plda_binary_matches_reference_parse(tmp_path):
    for tag, sfx in (("kaldi_plda_bin", "kaldi_plda"), ("kaldi_plda_f32", "kaldi_plda_f32")):
        p = tmp_path / tag
        p.write_bytes(G[tag].tobytes())
        mu, tr, psi = kaldi_io.read_plda(str(p))
        assert mu.dtype == G[sfx + "_mu"].dtype and tr.dtype == G[sfx + "_transform"].dtype
        assert np.array_equal(mu, G[sfx + "_mu"]) and np.array_equal(tr, G[sfx + "_transform"])
        assert np.array_equal(psi, G[sfx + "_psi"])
    # the writer reproduces the fixture byte for byte
    p2 = tmp_path / "rewritten"
    kaldi_io.write_plda(str(p2), G["kaldi_plda_mu"], G["kaldi_plda_transform"], G["kaldi_plda_psi"], binary=True)
    assert p2.read_bytes() == G["kaldi_plda_bin"].tobytes()


def test_read_kaldi_plda_text_roundtrip_and_load_model(tmp_path):
    pm = syn.make_plda(16, seed=5)
    p = tmp_path / "plda.txt"
    kaldi_io.write_plda(str(p), pm["mu"], pm["transform"], pm["psi"], binary=False)
    assert p.read_bytes().startswith(b"<Plda>  [ ")
    mu, tr, psi = kaldi_io.read_plda(str(p))
    assert np.allclose(mu, pm["mu"], atol=0) and np.allclose(psi, pm["psi"], atol=0)
    assert np.abs(tr - pm["transform"]).max() < 1e-6          # text matrices parse into float32, like kaldi_io's reader
    m = TwoCovPLDA.load_model(str(p), from_kaldi=True)        # two_cov_plda.py:344-347: offset = -transform @ mu
    assert m.dim == 16 and np.allclose(m.offset, -1.0 * m.transform @ m.mu)
    with pytest.raises(ValueError):
        (tmp_path / "bad").write_bytes(b"\0B<Nope> ")
        kaldi_io.read_plda(str(tmp_path / "bad"))


def test_save_model_npz_exact_path_and_h5_error(tmp_path):
    pm = syn.make_plda(16, seed=5, normalize_length=True)
    m = TwoCovPLDA.from_arrays(**pm)
    p = tmp_path / "model.npz"
    m.save_model(str(p))
    assert p.exists() and not (tmp_path / "model.npz.npz").exists()
    m2 = TwoCovPLDA.load_model(str(p))
    assert np.array_equal(m2.transform, m.transform) and m2.normalize_length is True
    try:
        import h5py  # noqa: F401
        m.save_model(str(tmp_path / "model.h5"))
        m3 = TwoCovPLDA.load_model(str(tmp_path / "model.h5"))
        assert np.array_equal(m3.psi, m.psi)
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            m.save_model(str(tmp_path / "model.h5"))
        with pytest.raises(ImportError, match="h5py"):
            TwoCovPLDA.load_model(str(tmp_path / "model.h5"))


# ------------------------------------------------------------------------------------------------- GPU
@gpu
@pytest.mark.parametrize("nl,avg,ind", CASES)
def test_gpu_eval_sv_matches_reference_file(nl, avg, ind, tmp_path):
    e, u2s, t, tr, i = _write_inputs(str(tmp_path))
    m = TwoCovPLDA.from_arrays(**syn.make_plda(256, seed=3, normalize_length=bool(nl)))
    sf = str(tmp_path / "scores")
    m.eval_sv(e, u2s, t, tr, sf, multisession_avg=bool(avg), indomain_scp=i if ind else None)
    worst = _check(_rows(open(sf).read()), _rows(G[f"scores_nl{nl}_avg{avg}_ind{ind}"]))
    print(f"eval_sv nl{nl} avg{avg} ind{ind}: worst |diff| {worst:.1e}")


@gpu
def test_gpu_scalar_api_is_fp64():
    pm = syn.make_plda(256, seed=3, normalize_length=True)
    m = TwoCovPLDA.from_arrays(**pm)
    x = np.random.default_rng(0).standard_normal(256)
    y = m.transform_embedding(x)
    assert np.abs(y - plda_np.transform_embedding(pm, x)).max() < 1e-12
