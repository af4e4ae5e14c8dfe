"""Generate tests/golden/score.npz by running the REAL reference scoring code (cosine trials + S-norm / AS-norm).

Runs only in the build container (reads /root/reference).  `wespeaker/bin/score.py:38-72` (trials_cosine_score) and
`wespeaker/bin/score_norm.py:26-37,54-117` (get_mean_std, main) are imported with the package __init__ bypassed;
`fire` is a stub and `kaldiio.load_scp_sequential` is replaced by an in-memory table (kaldiio is not installed here),
so the arithmetic, the trial handling and the output formatting are the reference's own.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/wespeaker"
TABLES = {}   # scp path -> list[(utt, np.float32 vector)]


def _import_reference():
    for pkg, path in [("wespeaker", REF), ("wespeaker.bin", REF + "/bin"), ("wespeaker.utils", REF + "/utils")]:
        m = types.ModuleType(pkg)
        m.__path__ = [path]
        sys.modules[pkg] = m
    sys.modules["fire"] = types.ModuleType("fire")
    kio = types.ModuleType("kaldiio")
    kio.load_scp_sequential = lambda path: iter(TABLES[path])
    sys.modules["kaldiio"] = kio
    import wespeaker.bin.score as score
    import wespeaker.bin.score_norm as score_norm
    return score, score_norm


def main():
    score, score_norm = _import_reference()
    rng = np.random.default_rng(5)
    D, NE, NC = 48, 60, 200
    spk = rng.standard_normal((12, D)).astype(np.float32)
    evals = (spk[rng.integers(0, 12, NE)] + 0.7 * rng.standard_normal((NE, D))).astype(np.float32)
    cohort = (spk[rng.integers(0, 12, NC)] * 0.5 + rng.standard_normal((NC, D))).astype(np.float32)
    cohort[17] = cohort[3]           # exact ties inside the cohort
    cohort[101] = cohort[3]
    mean_vec = (0.1 * rng.standard_normal(D)).astype(np.float32)
    utts = [f"utt{i:03d}" for i in range(NE)]
    coh = [f"spk{i:03d}" for i in range(NC)]
    pairs = [(utts[rng.integers(0, 20)], utts[rng.integers(20, NE)], "target" if rng.random() < 0.3 else "nontarget")
             for _ in range(150)]
    out = {"evals": evals, "cohort": cohort, "mean_vec": mean_vec, "utts": np.array(utts), "coh": np.array(coh),
           "pairs": np.array(pairs)}
    with tempfile.TemporaryDirectory() as td:
        eval_scp, coh_scp = os.path.join(td, "eval.scp"), os.path.join(td, "cohort.scp")
        TABLES[eval_scp] = list(zip(utts, evals))
        TABLES[coh_scp] = list(zip(coh, cohort))
        with open(coh_scp, "w") as f:                      # read_table(cohort_emb_scp) needs 2 columns per line
            for c in coh:
                f.write(f"{c} dummy.ark:0\n")
        mv = os.path.join(td, "mean_vec.npy")
        np.save(mv, mean_vec)
        trials = os.path.join(td, "trials.kaldi")
        with open(trials, "w") as f:
            for a, b, l in pairs:
                f.write(f"{a} {b} {l}\n")
        trials2 = os.path.join(td, "trials_nolabel")
        with open(trials2, "w") as f:
            for a, b, _ in pairs[:20]:
                f.write(f"{a} {b}\n")
        for tag, mvp in (("mean", mv), ("nomean", None)):
            sd = os.path.join(td, "scores_" + tag)
            os.makedirs(sd)
            score.trials_cosine_score(eval_scp, sd, mvp, (trials, trials2))
            out[f"cos_{tag}"] = np.array(open(os.path.join(sd, "trials.kaldi.score")).read())
            out[f"cos_nolabel_{tag}"] = np.array(open(os.path.join(sd, "trials_nolabel.score")).read())
            for method, top_n in (("asnorm", 30), ("snorm", 30)):
                dst = os.path.join(td, f"norm_{method}_{tag}")
                score_norm.main(method, top_n, os.path.join(sd, "trials.kaldi.score"), dst, coh_scp, eval_scp, mvp)
                out[f"{method}_{tag}"] = np.array(open(dst).read())
        m, s = score_norm.get_mean_std(evals - mean_vec, cohort - mean_vec, 30)
        out["topn_mean"], out["topn_std"] = m, s
    np.savez_compressed(os.path.join(HERE, "score.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
