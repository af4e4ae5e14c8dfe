#!/usr/bin/env python
"""Generate golden fixtures from the REAL reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the reference leaf modules from /root/reference with the namespace-bypass
recipe of SURVEY.md §8(c) (``import wespeaker`` itself fails on silero_vad / s3prl),
loads the synthetic checkpoints of ``wespeaker_b200.synthetic`` with
``load_state_dict(strict=True)`` (which also pins the key/shape specs), runs the
reference forwards / torchaudio fbank / TwoCovPLDA on seeded inputs and writes the
outputs to tests/golden/*.npz.  Inputs are regenerated from seeds at test time, only
outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/wespeaker"


def _import_reference():
    for pkg, path in [("wespeaker", REF), ("wespeaker.models", REF + "/models"),
                      ("wespeaker.dataset", REF + "/dataset"), ("wespeaker.utils", REF + "/utils"),
                      ("wespeaker.utils.plda", REF + "/utils/plda")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    for name in ("h5py", "kaldiio", "kaldi_io", "kaldi_io.kaldi_io", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    kk = sys.modules["kaldi_io.kaldi_io"]
    for mod in (kk, sys.modules["kaldi_io"]):
        for a in ("open_or_fd", "BadSampleSize", "UnknownMatrixHeader", "_read_compressed_mat",
                  "_read_mat_ascii"):
            if not hasattr(mod, a):
                setattr(mod, a, None)
    sys.modules["kaldi_io"].kaldi_io = kk
    import wespeaker.models.ecapa_tdnn as ecapa
    import wespeaker.models.resnet as resnet
    import wespeaker.models.campplus as campplus
    import wespeaker.dataset.dataset_utils as du
    from wespeaker.utils.plda.two_cov_plda import TwoCovPLDA
    from wespeaker.utils.plda.plda_utils import norm_embeddings
    return ecapa, resnet, campplus, du, TwoCovPLDA, norm_embeddings


def main():
    from wespeaker_b200 import synthetic as syn
    ecapa, resnet, campplus, du, TwoCovPLDA, norm_embeddings = _import_reference()
    torch.set_num_threads(8)
    ctors = {
        "ECAPA_TDNN_c512": ecapa.ECAPA_TDNN_c512, "ECAPA_TDNN_GLOB_c512": ecapa.ECAPA_TDNN_GLOB_c512,
        "ECAPA_TDNN_c1024": ecapa.ECAPA_TDNN_c1024, "ECAPA_TDNN_GLOB_c1024": ecapa.ECAPA_TDNN_GLOB_c1024,
        "ResNet18": resnet.ResNet18, "ResNet34": resnet.ResNet34, "CAMPPlus": campplus.CAMPPlus,
    }
    # (model, seed, B, T) cases.  T=198 is 2 s; odd/short T exercise the strided / ceil-mode paths.
    cases = [
        ("ECAPA_TDNN_c512", 0, 4, 198), ("ECAPA_TDNN_c512", 0, 2, 61),
        ("ECAPA_TDNN_GLOB_c512", 0, 4, 198), ("ECAPA_TDNN_c1024", 1, 2, 200),
        ("ECAPA_TDNN_GLOB_c1024", 1, 2, 200),
        ("ResNet34", 0, 2, 200), ("ResNet34", 0, 2, 99), ("ResNet18", 0, 2, 57),
        ("CAMPPlus", 0, 2, 198), ("CAMPPlus", 0, 2, 455), ("CAMPPlus", 0, 1, 98),
    ]
    out = {}
    for name, seed, B, T in cases:
        args = syn.DEFAULT_MODEL_ARGS[name]
        model = ctors[name](**args)
        sd_np = syn.make_state_dict(name, seed)
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
        ref_keys = list(model.state_dict().keys())
        assert ref_keys == list(sd.keys()), (name, set(ref_keys) ^ set(sd.keys()))
        model.load_state_dict(sd, strict=True)
        model.eval()
        feats = torch.from_numpy(syn.make_feats(B, T, 80, seed=seed + 17 * T))
        with torch.no_grad():
            o = model(du.apply_cmvn(feats))  # CMN is idempotent on already-normalised feats
            o = o[-1] if isinstance(o, tuple) else o
        key = f"{name}__s{seed}_B{B}_T{T}"
        out[key] = o.numpy().astype(np.float32)
        nparam = sum(p.numel() for p in model.parameters())
        print(f"{key}: emb {tuple(o.shape)} |e|={o.norm(dim=1).mean():.4f} params={nparam/1e6:.2f}M")
    np.savez_compressed(os.path.join(HERE, "models.npz"), **out)

    # ---- fbank (third-party arithmetic: torchaudio.compliance.kaldi.fbank) + CMN
    import torchaudio.compliance.kaldi as kaldi
    fb = {}
    wavs = syn.make_wavs(3, 32000, seed=0)
    for wt in ("hamming", "povey"):
        mats = [kaldi.fbank(torch.from_numpy(w)[None], num_mel_bins=80, frame_length=25, frame_shift=10,
                            dither=0.0, sample_frequency=16000, window_type=wt, use_energy=False)
                for w in wavs]
        fb[f"fbank_{wt}"] = torch.stack(mats).numpy()
    short = syn.make_wavs(1, 16000 + 77, seed=5)[0]
    fb["fbank_hamming_short"] = kaldi.fbank(torch.from_numpy(short)[None], num_mel_bins=80, dither=0.0,
                                            window_type="hamming").numpy()
    fb["cmvn_hamming"] = du.apply_cmvn(torch.from_numpy(fb["fbank_hamming"])).numpy()
    np.savez_compressed(os.path.join(HERE, "fbank.npz"), **fb)
    print("fbank:", {k: v.shape for k, v in fb.items()})

    # ---- PLDA (reference TwoCovPLDA with fields set directly; h5py/kaldiio are stubs)
    pl = {}
    for nl in (True, False):
        pm = syn.make_plda(256, seed=3, normalize_length=nl)
        plda = TwoCovPLDA(embed_dim=256, normalize_length=nl)
        plda.mu, plda.transform, plda.psi, plda.offset = pm["mu"], pm["transform"], pm["psi"], pm["offset"]
        plda.dim = 256
        enroll = syn.make_embeddings(48, 256, seed=3).astype(np.float64)
        test = syn.make_embeddings(40, 256, seed=4).astype(np.float64)
        counts = (np.arange(48) % 5) + 1
        if nl:
            e_t = np.stack([plda.transform_embedding(norm_embeddings(e)) for e in enroll])
            t_t = np.stack([plda.transform_embedding(norm_embeddings(t)) for t in test])
        else:
            e_t = np.stack([plda.transform_embedding(e) for e in enroll])
            t_t = np.stack([plda.transform_embedding(t) for t in test])
        s1 = np.array([[plda.log_likelihood_ratio(e, t, 1) for t in t_t] for e in e_t])
        sn = np.array([[plda.log_likelihood_ratio(e, t, int(c)) for t in t_t] for e, c in zip(e_t, counts)])
        tag = "norm" if nl else "raw"
        pl[f"enroll_t_{tag}"], pl[f"test_t_{tag}"] = e_t, t_t
        pl[f"scores_n1_{tag}"], pl[f"scores_nvar_{tag}"] = s1, sn
    np.savez_compressed(os.path.join(HERE, "plda.npz"), **pl)
    print("plda:", {k: v.shape for k, v in pl.items()})


if __name__ == "__main__":
    main()
