"""Generate tests/golden/plda_train.npz by running the REAL reference PLDA trainer / adapter
(`wespeaker/utils/plda/two_cov_plda.py:39-154` PldaStats + TwoCovPLDA.__init__/train/em_one_iter/get_output, `:258-309`
adapt) on a small synthetic embedding set.  Build container only (reads /root/reference); kaldiio is replaced by an
in-memory table, h5py / tqdm stubbed."""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/wespeaker"
TABLES = {}


def _import_reference():
    for pkg, path in [("wespeaker", REF), ("wespeaker.utils", REF + "/utils"), ("wespeaker.utils.plda", REF + "/utils/plda")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    for name in ("h5py", "kaldi_io", "kaldi_io.kaldi_io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    kk = sys.modules["kaldi_io.kaldi_io"]
    for mod in (kk, sys.modules["kaldi_io"]):
        for a in ("open_or_fd", "BadSampleSize", "UnknownMatrixHeader", "_read_compressed_mat", "_read_mat_ascii"):
            if not hasattr(mod, a):
                setattr(mod, a, None)
    sys.modules["kaldi_io"].kaldi_io = kk
    kio = types.ModuleType("kaldiio")
    kio.load_scp_sequential = lambda path: iter(TABLES[path])
    sys.modules["kaldiio"] = kio
    from wespeaker.utils.plda.two_cov_plda import TwoCovPLDA
    return TwoCovPLDA


def synth(seed, nspk, dim, nmin, nmax, shift=0.0):
    rng = np.random.default_rng(seed)
    basis = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    keys, vecs, spk = [], [], []
    for s in range(nspk):
        mu = basis @ (rng.standard_normal(dim) * np.linspace(2.0, 0.2, dim)) + shift
        for u in range(int(rng.integers(nmin, nmax + 1))):
            keys.append(f"spk{s:03d}-utt{u}")
            vecs.append((mu + 0.6 * rng.standard_normal(dim)).astype(np.float32))
            spk.append(f"spk{s:03d}")
    return keys, np.stack(vecs), spk


def main():
    TwoCovPLDA = _import_reference()
    out = {}
    with tempfile.TemporaryDirectory() as td:
        keys, X, spk = synth(11, 40, 24, 2, 7)
        akeys, AX, _ = synth(12, 30, 24, 3, 5, shift=0.3)
        scp, u2s, ascp = os.path.join(td, "x.scp"), os.path.join(td, "utt2spk"), os.path.join(td, "adapt.scp")
        TABLES[scp] = list(zip(keys, X))
        TABLES[ascp] = list(zip(akeys, AX))
        with open(u2s, "w") as f:
            for k, s in zip(keys, spk):
                f.write(f"{k} {s}\n")
        out.update(X=X, spk=np.array(spk), AX=AX)
        for tag, sub, nl in (("plain", False, False), ("norm", True, True)):
            with contextlib.redirect_stdout(io.StringIO()):
                p = TwoCovPLDA(scp, u2s, embed_dim=24, subtract_train_set_mean=sub, normalize_length=nl)
                out[f"{tag}_scatter"], out[f"{tag}_sum"] = p.stats.offset_scatter.copy(), p.stats.sum_.copy()
                Bs, Ws = [], []
                for _ in range(4):
                    p.em_one_iter()
                    Bs.append(p.B.copy()); Ws.append(p.W.copy())
                p.get_output()
                a = p.adapt(ascp, ac_scale=0.5, wc_scale=0.5)
            out[f"{tag}_B"], out[f"{tag}_W"] = np.stack(Bs), np.stack(Ws)
            out[f"{tag}_mu"], out[f"{tag}_transform"], out[f"{tag}_psi"], out[f"{tag}_offset"] = p.mu, p.transform, p.psi, p.offset
            out[f"{tag}_adapt_mu"], out[f"{tag}_adapt_transform"], out[f"{tag}_adapt_psi"] = a.mu, a.transform, a.psi
            out[f"{tag}_adapt_offset"] = a.offset
    np.savez_compressed(os.path.join(HERE, "plda_train.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
