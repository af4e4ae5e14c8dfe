"""Generate tests/golden/plda_eval.npz by running the REAL reference `TwoCovPLDA.eval_sv`
(`/root/reference/wespeaker/utils/plda/two_cov_plda.py:186-256`) and `read_plda`
(`utils/plda/kaldi_utils.py:24-108`, binary format) on synthetic files.  Build container only.

`kaldiio` is not installed, so the reference's `kaldiio.load_scp_sequential` is served by this repository's byte-level
ark/scp reader (wespeaker_b200.kaldi_io — pinned separately by the header-byte test); `h5py`, `kaldi_io`, `tqdm` are
stubs.  Everything from `get_data_for_plda` to the formatted score lines is the reference's own code.  Inputs AND the
reference's output score files are stored, so the GPU test needs neither the reference nor this script.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/wespeaker"


def _import_reference():
    from wespeaker_b200 import kaldi_io as our_io
    for pkg, path in [("wespeaker", REF), ("wespeaker.utils", REF + "/utils"), ("wespeaker.utils.plda", REF + "/utils/plda")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    kio = types.ModuleType("kaldiio")
    kio.load_scp_sequential = our_io.load_scp_sequential
    sys.modules["kaldiio"] = kio
    sys.modules["h5py"] = types.ModuleType("h5py")
    tq = types.ModuleType("tqdm")
    tq.tqdm = lambda it, *a, **k: it
    sys.modules["tqdm"] = tq
    kk = types.ModuleType("kaldi_io.kaldi_io")
    kp = types.ModuleType("kaldi_io")
    kp.open_or_fd = lambda f: open(f, "rb") if isinstance(f, str) else f
    kp.BadSampleSize = type("BadSampleSize", (Exception,), {})
    kp.UnknownMatrixHeader = type("UnknownMatrixHeader", (Exception,), {})
    kk._read_compressed_mat = None
    kk._read_mat_ascii = None
    kp.kaldi_io = kk
    sys.modules["kaldi_io"], sys.modules["kaldi_io.kaldi_io"] = kp, kk
    from wespeaker.utils.plda.two_cov_plda import TwoCovPLDA
    from wespeaker.utils.plda.kaldi_utils import read_plda
    return TwoCovPLDA, read_plda


def main():
    from wespeaker_b200 import kaldi_io as our_io, synthetic as syn
    TwoCovPLDA, read_plda = _import_reference()
    rng = np.random.default_rng(11)
    D, NSPK, NTEST, NIN = 256, 12, 30, 50
    nsess = rng.integers(1, 5, NSPK)
    spk_c = rng.standard_normal((NSPK, D)).astype(np.float32)
    enroll_keys, enroll_spk, enroll_vecs = [], [], []
    for s in range(NSPK):
        for j in range(int(nsess[s])):
            enroll_keys.append(f"enr{s:02d}_{j}")
            enroll_spk.append(f"spk{s:02d}")
            enroll_vecs.append((spk_c[s] + 0.5 * rng.standard_normal(D)).astype(np.float32))
    enroll_keys.append("orphan_utt")           # not in utt2spk: the reference prints a warning and skips it
    enroll_spk.append(None)
    enroll_vecs.append(rng.standard_normal(D).astype(np.float32))
    test_keys = [f"tst{i:03d}" for i in range(NTEST)]
    test_vecs = (spk_c[rng.integers(0, NSPK, NTEST)] + 0.6 * rng.standard_normal((NTEST, D))).astype(np.float32)
    indom = (0.3 + rng.standard_normal((NIN, D))).astype(np.float32)
    trials = [(f"spk{rng.integers(0, NSPK):02d}", test_keys[rng.integers(0, NTEST)],
               "target" if rng.random() < 0.3 else "nontarget") for _ in range(200)]
    out = {"enroll_keys": np.array(enroll_keys), "enroll_spk": np.array([s or "" for s in enroll_spk]),
           "enroll_vecs": np.stack(enroll_vecs), "test_keys": np.array(test_keys), "test_vecs": test_vecs,
           "indomain_vecs": indom, "trials": np.array(trials)}
    with tempfile.TemporaryDirectory() as td:
        def write(name, keys, vecs):
            with our_io.VectorWriter(os.path.join(td, name + ".ark"), os.path.join(td, name + ".scp")) as w:
                for k, v in zip(keys, vecs):
                    w(k, v)
            return os.path.join(td, name + ".scp")
        e_scp = write("enroll", enroll_keys, enroll_vecs)
        t_scp = write("test", test_keys, test_vecs)
        i_scp = write("indomain", [f"in{i}" for i in range(NIN)], indom)
        u2s = os.path.join(td, "utt2spk")
        with open(u2s, "w") as f:
            for k, s in zip(enroll_keys, enroll_spk):
                if s is not None:
                    f.write(f"{k} {s}\n")
        tr = os.path.join(td, "trials")
        with open(tr, "w") as f:
            for a, b, l in trials:
                f.write(f"{a} {b} {l}\n")
        for nl in (True, False):
            pm = syn.make_plda(D, seed=3, normalize_length=nl)
            p = TwoCovPLDA()
            p.mu, p.transform, p.psi, p.offset = pm["mu"], pm["transform"], pm["psi"], pm["offset"]
            p.normalize_length, p.dim = nl, D
            for avg in (True, False):
                for ind in (True, False):
                    sf = os.path.join(td, "scores")
                    p.eval_sv(e_scp, u2s, t_scp, tr, sf, multisession_avg=avg, indomain_scp=i_scp if ind else None)
                    out[f"scores_nl{int(nl)}_avg{int(avg)}_ind{int(ind)}"] = np.array(open(sf).read())
        # Kaldi <Plda>: binary files (double and float payloads) written by wespeaker_b200.kaldi_io.write_plda and by hand,
        # parsed by the REFERENCE read_plda: pins the byte layout both ways
        pm = syn.make_plda(16, seed=5)
        kp = os.path.join(td, "plda_bin")
        our_io.write_plda(kp, pm["mu"], pm["transform"], pm["psi"], binary=True)
        mu, trm, psi = read_plda(kp)
        assert np.array_equal(mu, pm["mu"]) and np.array_equal(trm, pm["transform"]) and np.array_equal(psi, pm["psi"])
        out["kaldi_plda_bin"] = np.frombuffer(open(kp, "rb").read(), dtype=np.uint8)
        out["kaldi_plda_mu"], out["kaldi_plda_transform"], out["kaldi_plda_psi"] = mu, trm, psi
        import struct
        kf = os.path.join(td, "plda_f32")
        with open(kf, "wb") as f:   # float payloads (FV / FM), as Kaldi writes them when compiled in single precision
            f.write(b"\0B<Plda> ")
            f.write(b"FV \4" + struct.pack("<i", 16) + pm["mu"].astype("<f4").tobytes())
            f.write(b"FM \4" + struct.pack("<i", 16) + b"\4" + struct.pack("<i", 16) + pm["transform"].astype("<f4").tobytes())
            f.write(b"FV \4" + struct.pack("<i", 16) + pm["psi"].astype("<f4").tobytes())
            f.write(b"</Plda> ")
        mu, trm, psi = read_plda(kf)
        out["kaldi_plda_f32"] = np.frombuffer(open(kf, "rb").read(), dtype=np.uint8)
        out["kaldi_plda_f32_mu"], out["kaldi_plda_f32_transform"], out["kaldi_plda_f32_psi"] = mu, trm, psi
    np.savez_compressed(os.path.join(HERE, "plda_eval.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
