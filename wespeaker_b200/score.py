"""Cosine scoring and S-norm / AS-norm on the B200 — mirror of `wespeaker/bin/score.py` and
`wespeaker/bin/score_norm.py`: same function names, arguments, trial handling and output
formats; the arithmetic (mean subtraction, L2 normalisation, per-trial cosine, embedding x cohort scores, per-row
top-N mean / std, per-trial normalisation) runs in fp64 on the device (ws_score.cu).  numpy here is file parsing and
host<->device staging only; there is no CPU fallback.
"""
from __future__ import annotations

import logging
import os
from pathlib import Path

import numpy as np
import torch

from . import lib as _lib
from .kaldi_io import load_scp_sequential


def read_table(table_file):
    """`wespeaker/utils/file_utils.py:51-65`."""
    table_list = []
    with open(table_file, "r", encoding="utf8") as fin:
        for line in fin:
            table_list.append(line.strip().split())
    return table_list


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.B200Error("wespeaker_b200.score needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device() if device is None else device)


def _unit_rows(embs: np.ndarray, mean_vec, dev):
    """(N,D) fp32 host embeddings -> device fp64 unit rows of (emb - mean_vec) and their norms."""
    L = _lib.load()
    x = torch.from_numpy(np.ascontiguousarray(embs, dtype=np.float32)).to(dev)
    N, D = x.shape
    mv = None
    if mean_vec is not None and not np.isscalar(mean_vec):
        mv = torch.from_numpy(np.ascontiguousarray(mean_vec, dtype=np.float64)).to(dev)
    unit = torch.empty((N, D), dtype=torch.float64, device=dev)
    norms = torch.empty((N,), dtype=torch.float64, device=dev)
    _lib.check(L.ws_score_unit_rows(x.data_ptr(), N, D, mv.data_ptr() if mv is not None else None, unit.data_ptr(),
                                    norms.data_ptr(), _lib.cur_stream_ptr(dev)), "ws_score_unit_rows")
    return unit, norms


def cosine_scores(embs: np.ndarray, enroll_idx, test_idx, mean_vec=None, device=None) -> np.ndarray:
    """Cosine similarity of the listed (enroll_idx[k], test_idx[k]) rows of `embs` after subtracting `mean_vec`."""
    dev = _device(device)
    unit, _ = _unit_rows(embs, mean_vec, dev)
    ei = torch.as_tensor(np.asarray(enroll_idx, dtype=np.int64)).to(dev)
    ti = torch.as_tensor(np.asarray(test_idx, dtype=np.int64)).to(dev)
    out = torch.empty((ei.numel(),), dtype=torch.float64, device=dev)
    _lib.check(_lib.load().ws_score_cosine_trials(unit.data_ptr(), ei.data_ptr(), ti.data_ptr(), ei.numel(), unit.shape[1],
                                                  out.data_ptr(), _lib.cur_stream_ptr(dev)), "ws_score_cosine_trials")
    return out.cpu().numpy()


def _cohort_stats(unit, unit_cohort, top_n, dev, tile_bytes=1 << 30):
    N, D = unit.shape
    M = unit_cohort.shape[0]
    rows = max(1, min(N, tile_bytes // (4 * M)))
    work = torch.empty((rows, M), dtype=torch.float32, device=dev)
    mean = torch.empty((N,), dtype=torch.float64, device=dev)
    std = torch.empty((N,), dtype=torch.float64, device=dev)
    _lib.check(_lib.load().ws_score_cohort_stats(unit.data_ptr(), N, unit_cohort.data_ptr(), M, D, int(top_n), work.data_ptr(),
                                                 rows, mean.data_ptr(), std.data_ptr(), _lib.cur_stream_ptr(dev)),
               "ws_score_cohort_stats")
    return mean, std


def get_mean_std(emb, cohort, top_n, device=None):
    """`score_norm.py:26-37`: mean / std of each embedding's top_n cosine scores against the cohort."""
    dev = _device(device)
    unit, _ = _unit_rows(np.asarray(emb), None, dev)
    ucoh, _ = _unit_rows(np.asarray(cohort), None, dev)
    m, s = _cohort_stats(unit, ucoh, top_n, dev)
    return m.cpu().numpy(), s.cpu().numpy()


# ------------------------------------------------------------------------------------------------- score.py
def calculate_mean_from_kaldi_vec(scp_path):
    """`score.py:25-36`."""
    vec_num = 0
    mean_vec = None
    for _, vec in load_scp_sequential(scp_path):
        if mean_vec is None:
            mean_vec = np.zeros_like(vec)
        mean_vec += vec
        vec_num += 1
    return mean_vec / vec_num


def trials_cosine_score(eval_scp_path="", store_dir="", mean_vec=None, trials=()):
    """`score.py:38-72`: one `<trial basename>.score` file per trial list, `enroll test score [label]`."""
    if mean_vec is None or not os.path.exists(mean_vec):
        mean_vec = None
    else:
        mean_vec = np.load(mean_vec)
    keys, vecs = [], []
    for utt, emb in load_scp_sequential(eval_scp_path):
        keys.append(utt)
        vecs.append(emb)
    index = {k: i for i, k in enumerate(keys)}
    embs = np.stack(vecs)
    for trial in trials:
        store_path = os.path.join(store_dir, os.path.basename(trial) + ".score")
        rows = [ln.strip().split() for ln in open(trial, "r")]
        scores = cosine_scores(embs, [index[r[0]] for r in rows], [index[r[1]] for r in rows], mean_vec)
        with open(store_path, "w") as w_f:
            for segs, cos_score in zip(rows, scores):
                if len(segs) == 3:  # enroll_name test_name target/nontarget
                    w_f.write("{} {} {:.5f} {}\n".format(segs[0], segs[1], cos_score, segs[2]))
                else:               # enroll_name test_name
                    w_f.write("{} {} {:.5f}\n".format(segs[0], segs[1], cos_score))


def main(exp_dir, eval_scp_path, cal_mean, cal_mean_dir, *trials):
    """`score.py:75-92`."""
    if not cal_mean:
        print("Do not do mean normalization for evaluation embeddings.")
        mean_vec_path = None
    else:
        scp_path = os.path.join(cal_mean_dir, "xvector.scp")
        print("Calculate mean statistics from {}.".format(scp_path))
        mean_vec = calculate_mean_from_kaldi_vec(scp_path)
        mean_vec_path = os.path.join(cal_mean_dir, "mean_vec.npy")
        np.save(mean_vec_path, mean_vec)
    store_score_dir = os.path.join(exp_dir, "scores")
    Path(store_score_dir).mkdir(parents=True, exist_ok=True)
    trials_cosine_score(eval_scp_path, store_score_dir, mean_vec_path, trials)


# ------------------------------------------------------------------------------------------------- score_norm.py
def split_embedding(utt_list, emb_scp, mean_vec):
    """`score_norm.py:40-51` (mean subtraction happens on the device, so the raw vectors are returned)."""
    utt2emb = {utt: emb for utt, emb in load_scp_sequential(emb_scp)}
    embs, utt2idx = [], {}
    for utt in utt_list:
        embs.append(utt2emb[utt])
        utt2idx[utt] = len(embs) - 1
    return np.array(embs), utt2idx


def score_norm_main(score_norm_method, top_n, trial_score_file, score_norm_file, cohort_emb_scp, eval_emb_scp,
                    mean_vec_path=None, device=None):
    """`score_norm.py:54-117` (its `main`): writes `enroll test normed label enroll_mag test_mag enroll_mean test_mean`."""
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
    if not mean_vec_path:
        print("Do not do mean normalization for evaluation embeddings.")
        mean_vec = None
    else:
        assert os.path.exists(mean_vec_path), "mean_vec file ({}) does not exist !!!".format(mean_vec_path)
        mean_vec = np.load(mean_vec_path)
    logging.info("get embedding ...")
    rows = read_table(trial_score_file)
    enroll_list = sorted(set(r[0] for r in rows))
    test_list = sorted(set(r[1] for r in rows))
    enroll_emb, enroll_utt2idx = split_embedding(enroll_list, eval_emb_scp, mean_vec)
    test_emb, test_utt2idx = split_embedding(test_list, eval_emb_scp, mean_vec)
    cohort_list = [r[0] for r in read_table(cohort_emb_scp)]
    cohort_emb, _ = split_embedding(cohort_list, cohort_emb_scp, mean_vec)
    logging.info("computing normed score ...")
    if score_norm_method == "asnorm":
        top_n = top_n
    elif score_norm_method == "snorm":
        top_n = cohort_emb.shape[0]
    else:
        raise ValueError(score_norm_method)
    dev = _device(device)
    L = _lib.load()
    e_unit, e_mag = _unit_rows(enroll_emb, mean_vec, dev)
    t_unit, t_mag = _unit_rows(test_emb, mean_vec, dev)
    c_unit, _ = _unit_rows(cohort_emb, mean_vec, dev)
    e_mean, e_std = _cohort_stats(e_unit, c_unit, top_n, dev)
    t_mean, t_std = _cohort_stats(t_unit, c_unit, top_n, dev)
    ei = torch.as_tensor(np.array([enroll_utt2idx[r[0]] for r in rows], dtype=np.int64)).to(dev)
    ti = torch.as_tensor(np.array([test_utt2idx[r[1]] for r in rows], dtype=np.int64)).to(dev)
    sc = torch.as_tensor(np.array([float(r[2]) for r in rows], dtype=np.float64)).to(dev)
    normed = torch.empty_like(sc)
    _lib.check(L.ws_score_asnorm(sc.data_ptr(), ei.data_ptr(), ti.data_ptr(), sc.numel(), e_mean.data_ptr(), e_std.data_ptr(),
                                 t_mean.data_ptr(), t_std.data_ptr(), normed.data_ptr(), _lib.cur_stream_ptr(dev)),
               "ws_score_asnorm")
    normed = normed.cpu().numpy()
    e_mag, t_mag, e_mean, t_mean = (a.cpu().numpy() for a in (e_mag, t_mag, e_mean, t_mean))
    ei, ti = ei.cpu().numpy(), ti.cpu().numpy()
    with open(score_norm_file, "w", encoding="utf-8") as fout:
        for k, line in enumerate(rows):
            fout.write("{} {} {:.5f} {} {:.4f} {:.4f} {:.4f} {:.4f}\n".format(
                line[0], line[1], normed[k], line[3], e_mag[ei[k]], t_mag[ti[k]], e_mean[ei[k]], t_mean[ti[k]]))
