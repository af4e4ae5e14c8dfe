"""CPU: pin the oracle (oracle/) against fixtures produced by the REAL reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import fbank_np, models_torch, plda_np
from wespeaker_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
class _Goldens(dict):
    """models.npz (hot-path families) + models_f4.npz (section 8(f) rank-4 families) behind the NpzFile interface."""
    @property
    def files(self):
        return list(self.keys())


G_MODELS = _Goldens()
for _f in ("models.npz", "models_f4.npz"):
    with np.load(os.path.join(HERE, "golden", _f)) as _z:
        G_MODELS.update({k: _z[k] for k in _z.files})
G_FBANK = np.load(os.path.join(HERE, "golden", "fbank.npz"))
G_PLDA = np.load(os.path.join(HERE, "golden", "plda.npz"))


def parse_case(key):
    name, rest = key.split("__")
    s, b, t = rest.split("_")
    return name, int(s[1:]), int(b[1:]), int(t[1:])


@pytest.mark.parametrize("key", list(G_MODELS.files))
def test_model_oracle_matches_reference(key):
    name, seed, B, T = parse_case(key)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = syn.make_state_dict(name, seed)
    feats = syn.make_feats(B, T, 80, seed=seed + 17 * T)
    emb = models_torch.forward(name, sd, feats).numpy()
    ref = G_MODELS[key]
    rel = np.linalg.norm(emb - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 5e-6, (key, rel)  # fp32 restatement vs fp32 reference (2.7e-7 floor)


def test_spec_param_counts():
    # README param counts reproduced by the survey: ECAPA_GLOB_c512 6.19M, ResNet34 6.63M, CAM++ 7.18M
    def nparams(name):
        return sum(int(np.prod(s)) for k, s in syn.state_dict_spec(name, **syn.DEFAULT_MODEL_ARGS[name]).items()
                   if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert abs(nparams("ECAPA_TDNN_GLOB_c512") / 1e6 - 6.19) < 0.01
    assert abs(nparams("ResNet34") / 1e6 - 6.63) < 0.01
    assert abs(nparams("CAMPPlus") / 1e6 - 7.18) < 0.01


@pytest.mark.parametrize("wt", ["hamming", "povey"])
def test_fbank_oracle_matches_torchaudio(wt):
    wavs = syn.make_wavs(3, 32000, seed=0)
    ref = G_FBANK[f"fbank_{wt}"]
    for b in range(3):
        out = fbank_np.fbank(wavs[b], window_type=wt)
        assert out.shape == ref[b].shape == (198, 80)
        # log-mel domain; fp32 FFT implementations differ in rounding only
        assert np.abs(out - ref[b]).max() < 2e-3
        assert np.abs(out - ref[b]).mean() < 2e-5


def test_fbank_short_and_cmn():
    short = syn.make_wavs(1, 16000 + 77, seed=5)[0]
    ref = G_FBANK["fbank_hamming_short"]
    out = fbank_np.fbank(short)
    assert out.shape == ref.shape == (98, 80)
    assert np.abs(out - ref).max() < 2e-3
    assert fbank_np.fbank(np.zeros(399, np.float32)).shape == (0, 80)  # shorter than one frame -> empty
    cm = fbank_np.cmn(G_FBANK["fbank_hamming"])
    assert np.abs(cm - G_FBANK["cmvn_hamming"]).max() < 5e-5  # fp32 mean over 198 frames of values ~15


@pytest.mark.parametrize("tag", ["norm", "raw"])
def test_plda_oracle_matches_reference(tag):
    nl = tag == "norm"
    pm = syn.make_plda(256, seed=3, normalize_length=nl)
    enroll = syn.make_embeddings(48, 256, seed=3).astype(np.float64)
    test = syn.make_embeddings(40, 256, seed=4).astype(np.float64)
    e_t = plda_np.prepare_test(pm, enroll)  # one session per speaker == prepare_test arithmetic
    t_t = plda_np.prepare_test(pm, test)
    assert np.abs(e_t - G_PLDA[f"enroll_t_{tag}"]).max() < 1e-11
    assert np.abs(t_t - G_PLDA[f"test_t_{tag}"]).max() < 1e-11
    s1 = plda_np.llr_matrix(pm, e_t, t_t, 1)
    counts = (np.arange(48) % 5) + 1
    sn = plda_np.llr_matrix(pm, e_t, t_t, counts)
    assert np.abs(s1 - G_PLDA[f"scores_n1_{tag}"]).max() < 1e-9
    assert np.abs(sn - G_PLDA[f"scores_nvar_{tag}"]).max() < 1e-9
    # literal per-trial restatement on a few trials
    for i, j in [(0, 0), (5, 7), (47, 39)]:
        v = plda_np.log_likelihood_ratio(pm, e_t[i], t_t[j], int(counts[i]))
        assert abs(v - G_PLDA[f"scores_nvar_{tag}"][i, j]) < 1e-9


def test_plda_prepare_enroll_multisession():
    pm = syn.make_plda(256, seed=3, normalize_length=True)
    emb = syn.make_embeddings(12, 256, seed=9).astype(np.float64)
    sessions = [emb[0:3], emb[3:4], emb[4:9], emb[9:12]]
    e_avg, c_avg = plda_np.prepare_enroll(pm, sessions, multisession_avg=True)
    e_cnt, c_cnt = plda_np.prepare_enroll(pm, sessions, multisession_avg=False)
    assert list(c_avg) == [1, 1, 1, 1] and list(c_cnt) == [3, 1, 5, 3]
    assert np.allclose(e_avg, e_cnt)
    assert np.allclose(np.linalg.norm(e_avg, axis=1), 16.0)  # sqrt(256) length-norm
