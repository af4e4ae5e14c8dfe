"""Oracle: numpy restatement of cosine trial scoring and S-norm / AS-norm score normalisation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows `wespeaker/bin/score.py:38-72` (trials_cosine_score) and
`wespeaker/bin/score_norm.py:26-37` (get_mean_std), `:54-117` (main: trial handling, normalisation, output columns).
Arithmetic in fp64 (the reference runs fp32 numpy / sklearn on fp32 Kaldi vectors; tests compare at 2e-5, the last
printed digit).  Pinned by tests/golden/score.npz, produced by the reference's own code
(tests/golden/make_golden_score.py).
"""
from __future__ import annotations

import numpy as np


def cosine_trials(embs, enroll_idx, test_idx, mean_vec=None):
    """score.py:50-63: emb - mean_vec, cosine_similarity per listed trial."""
    x = np.asarray(embs, dtype=np.float64) - (0.0 if mean_vec is None else np.asarray(mean_vec, dtype=np.float64))
    u = x / np.linalg.norm(x, axis=1, keepdims=True)
    return np.einsum("kd,kd->k", u[np.asarray(enroll_idx)], u[np.asarray(test_idx)])


def get_mean_std(emb, cohort, top_n):
    """score_norm.py:26-37, literally (fp64)."""
    emb = np.asarray(emb, dtype=np.float64)
    cohort = np.asarray(cohort, dtype=np.float64)
    emb = emb / np.sqrt(np.sum(emb ** 2, axis=1, keepdims=True))
    cohort = cohort / np.sqrt(np.sum(cohort ** 2, axis=1, keepdims=True))
    emb_cohort_score = np.matmul(emb, cohort.T)
    emb_cohort_score = np.sort(emb_cohort_score, axis=1)[:, ::-1]
    emb_cohort_score_topn = emb_cohort_score[:, :top_n]
    return np.mean(emb_cohort_score_topn, axis=1), np.std(emb_cohort_score_topn, axis=1)


def score_norm_lines(method, top_n, trial_rows, eval_table, cohort_table, mean_vec=None):
    """score_norm.py:54-117 on in-memory tables: trial_rows = [(enroll, test, score_str, label)], *_table = {utt: vector}.
    Returns the output file's lines (without newline)."""
    mv = 0.0 if mean_vec is None else np.asarray(mean_vec, dtype=np.float64)
    enroll_list = sorted(set(r[0] for r in trial_rows))
    test_list = sorted(set(r[1] for r in trial_rows))
    e_emb = np.stack([np.asarray(eval_table[u], dtype=np.float64) - mv for u in enroll_list])
    t_emb = np.stack([np.asarray(eval_table[u], dtype=np.float64) - mv for u in test_list])
    c_emb = np.stack([np.asarray(v, dtype=np.float64) - mv for v in cohort_table.values()])
    e_idx = {u: i for i, u in enumerate(enroll_list)}
    t_idx = {u: i for i, u in enumerate(test_list)}
    if method == "asnorm":
        n = top_n
    elif method == "snorm":
        n = c_emb.shape[0]
    else:
        raise ValueError(method)
    em, es = get_mean_std(e_emb, c_emb, n)
    tm, ts = get_mean_std(t_emb, c_emb, n)
    out = []
    for r in trial_rows:
        i, j, s = e_idx[r[0]], t_idx[r[1]], float(r[2])
        normed = 0.5 * ((s - em[i]) / es[i] + (s - tm[j]) / ts[j])
        out.append("{} {} {:.5f} {} {:.4f} {:.4f} {:.4f} {:.4f}".format(
            r[0], r[1], normed, r[3], np.linalg.norm(e_emb[i]), np.linalg.norm(t_emb[j]), em[i], tm[j]))
    return out
