"""Oracle: numpy fp64 restatement of two-covariance PLDA scoring.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
`wespeaker/utils/plda/two_cov_plda.py:156-184` (transform_embedding, log_likelihood_ratio),
`:186-256` (eval_sv preparation) and `wespeaker/utils/plda/plda_utils.py:46-58`
(norm_embeddings).  Pinned by tests/golden/plda_*.npz.
"""
from __future__ import annotations

import math

import numpy as np

M_LOG_2PI = 1.8378770664093454835606594728112


def norm_embeddings(e, kaldi_style=True):
    """plda_utils.py:46-58: sqrt(D) * x / ||x||."""
    e = np.asarray(e, dtype=np.float64)
    scale = math.sqrt(e.shape[-1]) if kaldi_style else 1.0
    return scale * e / np.linalg.norm(e, axis=-1, keepdims=True)


def transform_embedding(plda, x):
    """two_cov_plda.py:156-163 for one (D,) or a batch (N,D) of embeddings."""
    x = np.asarray(x, dtype=np.float64)
    y = x @ plda["transform"].T + plda["offset"]
    if plda["normalize_length"]:
        y = y * (math.sqrt(plda["dim"]) / np.linalg.norm(y, axis=-1, keepdims=True))
    return y


def log_likelihood_ratio(plda, e, t, n):
    """two_cov_plda.py:165-184, literally (one trial)."""
    psi = plda["psi"]
    dim = plda["dim"]
    mean = n * psi / (n * psi + 1.0) * e
    variance = 1.0 + psi / (n * psi + 1.0)
    logdet = np.sum(np.log(variance))
    sqdiff = np.power(t - mean, 2.0)
    variance = 1.0 / variance
    loglike_given_class = -0.5 * (logdet + M_LOG_2PI * dim + np.dot(sqdiff, variance))
    sqdiff = np.power(t, 2.0)
    variance = psi + 1.0
    logdet = np.sum(np.log(variance))
    variance = 1.0 / variance
    loglike_without_class = -0.5 * (logdet + M_LOG_2PI * dim + np.dot(sqdiff, variance))
    return loglike_given_class - loglike_without_class


def llr_matrix(plda, enroll_t, test_t, n):
    """All-pairs scores S[i,j] = log_likelihood_ratio(enroll_t[i], test_t[j], n[i]) in fp64,
    vectorised over j (the per-trial formula above applied row by row)."""
    enroll_t = np.asarray(enroll_t, dtype=np.float64)
    test_t = np.asarray(test_t, dtype=np.float64)
    n = np.broadcast_to(np.asarray(n, dtype=np.float64), (enroll_t.shape[0],))
    psi = plda["psi"]
    out = np.empty((enroll_t.shape[0], test_t.shape[0]), dtype=np.float64)
    wo = -0.5 * (np.sum(np.log(psi + 1.0)) + (test_t ** 2) @ (1.0 / (psi + 1.0)))
    for i in range(enroll_t.shape[0]):
        mean = n[i] * psi / (n[i] * psi + 1.0) * enroll_t[i]
        var = 1.0 + psi / (n[i] * psi + 1.0)
        given = -0.5 * (np.sum(np.log(var)) + ((test_t - mean) ** 2) @ (1.0 / var))
        out[i] = given - wo
    return out


def prepare_enroll(plda, sessions, mean_vec=None, multisession_avg=True):
    """two_cov_plda.py:218-235: per speaker, (sessions - mean_vec).mean(0) -> [norm] -> transform.
    ``sessions`` is a list of (k_i, D) arrays.  Returns (transformed (S,D), counts (S,))."""
    d = plda["dim"]
    mv = np.zeros(d) if mean_vec is None else np.asarray(mean_vec, dtype=np.float64)
    outs, counts = [], []
    for value in sessions:
        value = np.vstack(value).astype(np.float64)
        counts.append(1 if multisession_avg else len(value))
        value = value - mv
        tmp = np.mean(value, 0)
        if plda["normalize_length"]:
            tmp = norm_embeddings(tmp)
        outs.append(transform_embedding(plda, tmp))
    return np.vstack(outs), np.asarray(counts)


def prepare_test(plda, embs, mean_vec=None):
    """two_cov_plda.py:237-244."""
    d = plda["dim"]
    mv = np.zeros(d) if mean_vec is None else np.asarray(mean_vec, dtype=np.float64)
    v = np.asarray(embs, dtype=np.float64) - mv
    if plda["normalize_length"]:
        v = norm_embeddings(v)
    return transform_embedding(plda, v)
