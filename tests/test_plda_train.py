"""PLDA training (EM) and adaptation against goldens produced by the reference's own
TwoCovPLDA.train / adapt (tests/golden/make_golden_plda_train.py).  The reference accumulates its class statistics in
fp32 (Kaldi float vectors, `np.matmul(tmp.T, tmp)` on float32), this implementation in fp64: tolerances are set by that."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import plda_np
from wespeaker_b200 import plda_train

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "plda_train.npz"))


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _sign_align(t, ref):
    s = np.sign(np.sum(t * ref, axis=1, keepdims=True))
    return t * s


def _llr_matrix(mu, transform, psi, offset, normalize_length, e, t):
    p = dict(mu=mu, transform=transform, psi=psi, offset=offset, dim=len(mu), normalize_length=normalize_length)
    return plda_np.llr_matrix(p, plda_np.transform_embedding(p, e), plda_np.transform_embedding(p, t), 1)


def _run(tag, device):
    sub, nl = (False, False) if tag == "plain" else (True, True)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = plda_train.TwoCovPLDATrainer(embed_dim=24, subtract_train_set_mean=sub, normalize_length=nl, device=device,
                                          embeddings=G["X"], labels=[str(s) for s in G["spk"]])
        assert _rel(tr.stats.offset_scatter.cpu().numpy(), G[f"{tag}_scatter"]) < 1e-5
        assert _rel(tr.stats.sum_.cpu().numpy(), G[f"{tag}_sum"]) < 1e-5
        for it in range(4):
            tr.em_one_iter()
            assert _rel(tr.B.cpu().numpy(), G[f"{tag}_B"][it]) < 1e-5, it
            assert _rel(tr.W.cpu().numpy(), G[f"{tag}_W"][it]) < 1e-5, it
        tr.get_output()
    assert _rel(tr.mu, G[f"{tag}_mu"]) < 1e-5
    assert _rel(tr.psi, G[f"{tag}_psi"]) < 1e-5
    # eigenvectors are defined up to sign: compare rows after sign alignment, and the scores they produce
    assert _rel(_sign_align(tr.transform, G[f"{tag}_transform"]), G[f"{tag}_transform"]) < 1e-4
    e, t = G["AX"][:20].astype(np.float64), G["AX"][20:50].astype(np.float64)
    ref_s = _llr_matrix(G[f"{tag}_mu"], G[f"{tag}_transform"], G[f"{tag}_psi"], G[f"{tag}_offset"], nl, e, t)
    got_s = _llr_matrix(tr.mu, tr.transform, tr.psi, tr.offset, nl, e, t)
    assert np.abs(got_s - ref_s).max() < 1e-4 * max(1.0, np.abs(ref_s).max())
    # adaptation, from the REFERENCE's trained model so both sides start from identical inputs
    mu, tf, psi, off = plda_train.adapt(G[f"{tag}_mu"], G[f"{tag}_transform"], G[f"{tag}_psi"], G["AX"], normalize_length=nl,
                                        device=device)
    assert np.abs(mu - G[f"{tag}_adapt_mu"]).max() < 1e-5      # ~0 after mean subtraction (fp32 noise in the reference)
    assert _rel(psi, G[f"{tag}_adapt_psi"]) < 1e-5
    ref_a = _llr_matrix(G[f"{tag}_adapt_mu"], G[f"{tag}_adapt_transform"], G[f"{tag}_adapt_psi"], G[f"{tag}_adapt_offset"], nl, e, t)
    got_a = _llr_matrix(mu, tf, psi, off, nl, e, t)
    assert np.abs(got_a - ref_a).max() < 1e-4 * max(1.0, np.abs(ref_a).max())


@pytest.mark.parametrize("tag", ["plain", "norm"])
def test_train_and_adapt_match_reference_cpu_device(tag):
    _run(tag, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "norm"])
def test_train_and_adapt_match_reference_on_gpu(tag):
    _run(tag, "cuda:0")


def test_default_device_is_the_gpu():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        plda_train.TwoCovPLDATrainer(embed_dim=24, embeddings=G["X"], labels=[str(s) for s in G["spk"]])
