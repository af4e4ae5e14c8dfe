"""Generate tests/golden/embproc.npz by running the REAL reference embedding processing chain
(`wespeaker/utils/embedding_processing.py:23-271`: chain_string_to_dict, MeanSubtraction, Length_norm, Lda,
EmbeddingProcessingChain) on synthetic embeddings.  Build container only; kaldiio replaced by an in-memory table."""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_plda_train import synth, TABLES  # noqa: E402

REF = "/root/reference/wespeaker"


def _import_reference():
    for pkg, path in [("wespeaker", REF), ("wespeaker.utils", REF + "/utils"), ("wespeaker.utils.plda", REF + "/utils/plda")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    kio = types.ModuleType("kaldiio")
    kio.load_scp_sequential = lambda path: iter(TABLES[path])
    sys.modules["kaldiio"] = kio
    import wespeaker.utils.embedding_processing as ep
    return ep


def main():
    ep = _import_reference()
    out = {}
    with tempfile.TemporaryDirectory() as td:
        keys, X, spk = synth(21, 30, 20, 1, 6)          # some speakers have a single utterance (skipped by the LDA stats)
        tkeys, TX, _ = synth(22, 8, 20, 2, 3, shift=0.2)
        scp, u2s = os.path.join(td, "x.scp"), os.path.join(td, "utt2spk")
        TABLES[scp] = list(zip(keys, X))
        with open(u2s, "w") as f:
            for k, s in zip(keys, spk):
                f.write(f"{k} {s}\n")
        chain = f"mean-subtract --scp {scp} | length-norm | lda --scp {scp} --utt2spk={u2s} --dim 8 | length-norm"
        out["parsed"] = np.array(repr(ep.chain_string_to_dict("mean-subtract --scp a.scp | length-norm |"
                                                             " lda  --scp b.scp --utt2spk=u2s --dim 100 | length-norm")))
        with contextlib.redirect_stdout(io.StringIO()):
            c = ep.EmbeddingProcessingChain(chain)
            lda = c.chain_of_classes[2]
            out.update(X=X, spk=np.array(spk), TX=TX, mean0=c.chain_of_classes[0].mean, lda_m=lda.m, lda_mat=lda.lda,
                       y=c(TX.copy()), y_train=c(X.copy()))
            m, bc, wc = lda.compute_mean_and_lda_scatter_matrices(scp, u2s, equal_speaker_weight=True,
                                                                  current_chain=lambda e: e)
            out.update(eq_mean=m, eq_bc=bc, eq_wc=wc)
            m, bc, wc = lda.compute_mean_and_lda_scatter_matrices(scp, u2s, equal_speaker_weight=False,
                                                                  current_chain=lambda e: e)
            out.update(w_mean=m, w_bc=bc, w_wc=wc)
    np.savez_compressed(os.path.join(HERE, "embproc.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
