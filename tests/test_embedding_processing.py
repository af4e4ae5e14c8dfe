"""Embedding processing chain against goldens produced by the reference's own classes
(tests/golden/make_golden_embproc.py)."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from wespeaker_b200 import embedding_processing as ep
from wespeaker_b200 import kaldi_io

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "embproc.npz"))


def test_chain_string_parsing_matches_reference():
    got = ep.chain_string_to_dict("mean-subtract --scp a.scp | length-norm | lda  --scp b.scp --utt2spk=u2s --dim 100 | length-norm")
    assert repr(got) == str(G["parsed"])
    assert ep.chain_string_to_dict(None) == []


def _files(tmp_path):
    scp, u2s = str(tmp_path / "x.scp"), str(tmp_path / "utt2spk")
    keys = [f"k{i:04d}" for i in range(len(G["X"]))]
    with kaldi_io.VectorWriter(str(tmp_path / "x.ark"), scp) as w:
        for k, v in zip(keys, G["X"]):
            w(k, v)
    with open(u2s, "w") as f:
        for k, s in zip(keys, G["spk"]):
            f.write(f"{k} {s}\n")
    return scp, u2s


def _cols_aligned(y, ref):
    return y * np.sign(np.sum(y * ref, axis=0, keepdims=True))


def _run(tmp_path, device):
    scp, u2s = _files(tmp_path)
    with contextlib.redirect_stdout(io.StringIO()):
        c = ep.EmbeddingProcessingChain(f"mean-subtract --scp {scp} | length-norm | lda --scp {scp} --utt2spk={u2s} --dim 8 | length-norm",
                                        device=device)
        lda = c.chain_of_classes[2]
        for eq, tag in ((True, "eq"), (False, "w")):
            m, bc, wc = lda.compute_mean_and_lda_scatter_matrices(scp, u2s, equal_speaker_weight=eq, current_chain=lambda e: e)
            for got, key in ((m, "mean"), (bc, "bc"), (wc, "wc")):
                ref = G[f"{tag}_{key}"]
                assert np.abs(got.cpu().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (tag, key)
    assert np.abs(c.chain_of_classes[0].mean - G["mean0"]).max() < 1e-6
    assert np.abs(lda.m - G["lda_m"]).max() < 1e-6
    # LDA directions are eigenvectors: defined up to sign per column
    assert np.abs(_cols_aligned(lda.lda, G["lda_mat"]) - G["lda_mat"]).max() < 1e-4 * np.abs(G["lda_mat"]).max()
    for x, key in ((G["TX"], "y"), (G["X"], "y_train")):
        y = c(x.copy())
        assert y.shape == G[key].shape
        assert np.abs(_cols_aligned(y, G[key]) - G[key]).max() < 1e-4
        assert np.abs(y @ y.T - G[key] @ G[key].T).max() < 1e-4          # what downstream cosine / PLDA scoring sees
    # pickle round trip and link replacement keep the call contract
    p = str(tmp_path / "chain.pkl")
    with contextlib.redirect_stdout(io.StringIO()):
        c.save(p)
        c2 = ep.EmbeddingProcessingChain(None, device=device)
        c2.load(p)
        assert np.array_equal(c2(G["TX"].copy()), c(G["TX"].copy()))
        c2.update_link(3, "length-norm")
        assert np.array_equal(c2(G["TX"].copy()), c(G["TX"].copy()))


def test_chain_matches_reference_cpu_device(tmp_path):
    _run(tmp_path, "cpu")


@pytest.mark.gpu
def test_chain_matches_reference_on_gpu(tmp_path):
    _run(tmp_path, "cuda:0")


def test_default_device_is_the_gpu():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        ep.EmbeddingProcessingChain("length-norm")


def test_loads_chains_pickled_by_the_reference_classes(tmp_path):
    """A chain saved by the reference is a pickle of `wespeaker.utils.embedding_processing.<Class>` objects holding numpy
    arrays; it must load here (by attribute) and a chain saved here carries no device handle."""
    import pickle
    import sys
    import types
    fake = types.ModuleType("wespeaker.utils.embedding_processing")
    for name in ("Lda", "Length_norm", "MeanSubtraction"):
        cls = type(name, (), {})
        cls.__module__ = fake.__name__
        setattr(fake, name, cls)
    saved = {k: sys.modules.get(k) for k in ("wespeaker", "wespeaker.utils", fake.__name__)}
    sys.modules.setdefault("wespeaker", types.ModuleType("wespeaker"))
    sys.modules.setdefault("wespeaker.utils", types.ModuleType("wespeaker.utils"))
    sys.modules[fake.__name__] = fake
    try:
        rng = np.random.default_rng(0)
        ms, ln, lda = fake.MeanSubtraction(), fake.Length_norm(), fake.Lda()
        ms.mean = rng.standard_normal(6)
        lda.m, lda.lda = rng.standard_normal(6), rng.standard_normal((6, 3))
        p = tmp_path / "ref_chain.pkl"
        with open(p, "wb") as f:
            pickle.dump([ms, ln, lda], f)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    with contextlib.redirect_stdout(io.StringIO()):
        c = ep.EmbeddingProcessingChain(None, device="cpu")
        c.load(str(p))
    x = rng.standard_normal((5, 6))
    y = x - ms.mean
    y = y / np.sqrt((y ** 2).sum(1, keepdims=True))
    y = (y - lda.m) @ lda.lda
    assert [type(l).__name__ for l in c.chain_of_classes] == ["MeanSubtraction", "Length_norm", "Lda"]
    assert np.abs(c(x) - y).max() < 1e-12
    yt = c(torch.from_numpy(x))                     # tensor in -> tensor out (stays on its device)
    assert torch.is_tensor(yt) and np.abs(yt.numpy() - y).max() < 1e-12
    with contextlib.redirect_stdout(io.StringIO()):
        c.save(str(tmp_path / "ours.pkl"))
    blob = (tmp_path / "ours.pkl").read_bytes()
    assert b"cuda" not in blob and b"torch" not in blob
