"""Two-covariance PLDA scoring on the B200 — mirror of the scoring half of
`wespeaker/utils/plda/two_cov_plda.py` (seam B5): ``load_model``, ``transform_embedding``,
``log_likelihood_ratio``, ``eval_sv``; plus the batched entry points the GPU makes worthwhile
(``transform_batch`` on (N,D) matrices, ``score_matrix`` all-pairs, ``score_trials``).  All arithmetic is fp64 on
the device (ws_plda.cu); numpy here is only file parsing and host<->device staging.  Training / adaptation
(`two_cov_plda.py:106-154,258-309`) is out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .kaldi_io import read_vec_scp_file


def read_label_file(label_file):
    """`wespeaker/utils/plda/plda_utils.py:32-43`."""
    labels = {}
    with open(label_file) as fin:
        for line in fin:
            tok = line.strip().split()
            if tok:
                labels[tok[0]] = tok[1]
    return labels


class TwoCovPLDA:
    def __init__(self, embed_dim: int = 256, normalize_length: bool = False, subtract_train_set_mean: bool = False,
                 device: int | None = None):
        self.dim = embed_dim
        self.normalize_length = normalize_length
        self.subtract_train_set_mean = subtract_train_set_mean
        self.mu = np.zeros(embed_dim)
        self.transform = np.zeros((embed_dim, embed_dim))
        self.psi = np.zeros(embed_dim)
        self.offset = np.zeros(embed_dim)
        self._device = device
        self._h = None

    # ------------------------------------------------------------------ model IO (two_cov_plda.py:341-363)
    @staticmethod
    def from_arrays(mu, transform, psi, offset=None, normalize_length=False, subtract_train_set_mean=False,
                    device=None, **_):
        p = TwoCovPLDA(len(mu), normalize_length, subtract_train_set_mean, device)
        p.mu = np.asarray(mu, dtype=np.float64)
        p.transform = np.ascontiguousarray(transform, dtype=np.float64)
        p.psi = np.asarray(psi, dtype=np.float64)
        p.offset = (-1.0 * p.transform @ p.mu) if offset is None else np.asarray(offset, dtype=np.float64)
        return p

    @staticmethod
    def load_model(model_name: str, from_kaldi: bool = False, device=None):
        if from_kaldi:
            raise NotImplementedError("Kaldi <Plda> files are read by the reference's kaldi_utils.read_plda; "
                                      "convert to .npz/.h5 (IO glue, out of the hot path)")
        if model_name.endswith(".npz"):
            z = np.load(model_name)
            get = lambda k: z[k]  # noqa: E731
        else:
            import h5py  # not installed in this image; same dataset names as two_cov_plda.py:311-339
            f = h5py.File(model_name, "r")
            get = lambda k: f.get(k)[()]  # noqa: E731
        return TwoCovPLDA.from_arrays(get("mu"), get("transform"), get("psi"), get("offset"),
                                      bool(get("normalize_length")), bool(get("subtract_train_set_mean")), device)

    def save_model(self, path: str):
        np.savez(path, mu=self.mu, transform=self.transform, psi=self.psi, offset=self.offset,
                 normalize_length=int(self.normalize_length), subtract_train_set_mean=int(self.subtract_train_set_mean))

    # ------------------------------------------------------------------ device handle
    def _dev(self) -> int:
        if self._device is None:
            if not torch.cuda.is_available():
                raise _lib.B200Error("TwoCovPLDA scoring needs a CUDA device (no CPU fallback)")
            self._device = torch.cuda.current_device()
        return int(self._device)

    def _handle(self):
        if self._h is None:
            L = _lib.load()
            h = _lib.c_plda_p()
            mu = np.ascontiguousarray(self.mu, dtype=np.float64)
            A = np.ascontiguousarray(self.transform, dtype=np.float64)
            psi = np.ascontiguousarray(self.psi, dtype=np.float64)
            off = np.ascontiguousarray(self.offset, dtype=np.float64)
            _lib.check(L.ws_plda_create(self.dim, mu.ctypes.data, A.ctypes.data, psi.ctypes.data, off.ctypes.data,
                                        int(self.normalize_length), self._dev(), C.byref(h)), "ws_plda_create")
            self._h = h
        return self._h

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().ws_plda_destroy(self._h)
        except Exception:
            pass

    def _cuda(self, x, dtype):
        t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=torch.device("cuda", self._dev()), dtype=dtype).contiguous()

    # ------------------------------------------------------------------ batched GPU API
    def transform_batch(self, embeddings, mean_vec=None, pre_norm: bool | None = None) -> torch.Tensor:
        """(N,D) fp32 embeddings -> fp64 CUDA (N,D): ``transform_embedding(pre(x - mean_vec))`` where pre is the
        sqrt(D) length-norm (default: iff normalize_length, as eval_sv :225-241 does; transform_embedding :156-163)."""
        if pre_norm is None:
            pre_norm = bool(self.normalize_length)
        x = self._cuda(embeddings, torch.float32)
        if x.dim() == 1:
            x = x[None]
        y = torch.empty(x.shape, dtype=torch.float64, device=x.device)
        mv = None
        if mean_vec is not None:
            mv = np.ascontiguousarray(mean_vec, dtype=np.float64)
        with torch.cuda.device(self._dev()):
            _lib.check(_lib.load().ws_plda_transform(self._handle(), x.data_ptr(), x.shape[0],
                                                     mv.ctypes.data if mv is not None else None, int(pre_norm),
                                                     y.data_ptr(),
                                                     _lib.cur_stream_ptr(self._dev())), "ws_plda_transform")
        return y

    def score_matrix(self, enroll_t, test_t, counts=1, out: torch.Tensor | None = None,
                     out_dtype=torch.float32) -> torch.Tensor:
        """All-pairs LLR (N,M).  ``counts`` = int (constant n, K=D GEMM) or (N,) int array (K=2D GEMM)."""
        e = self._cuda(enroll_t, torch.float64)
        t = self._cuda(test_t, torch.float64)
        N, M = e.shape[0], t.shape[0]
        if out is None:
            out = torch.empty((N, M), dtype=out_dtype, device=e.device)
        if N == 0 or M == 0:
            return out
        cnt, const_n = None, 1
        if np.isscalar(counts) or (torch.is_tensor(counts) and counts.dim() == 0):
            const_n = int(counts)
        else:
            cnt = self._cuda(counts, torch.int32)
        with torch.cuda.device(self._dev()):
            _lib.check(_lib.load().ws_plda_score_matrix(
                self._handle(), e.data_ptr(), cnt.data_ptr() if cnt is not None else None, const_n, N, t.data_ptr(),
                M, out.data_ptr(), 1 if out.dtype == torch.float64 else 0, out.stride(0),
                _lib.cur_stream_ptr(self._dev())), "ws_plda_score_matrix")
        return out

    def score_trials(self, enroll_t, test_t, enroll_idx, test_idx, counts=1) -> torch.Tensor:
        e = self._cuda(enroll_t, torch.float64)
        t = self._cuda(test_t, torch.float64)
        ei = self._cuda(enroll_idx, torch.int64)
        ti = self._cuda(test_idx, torch.int64)
        out = torch.empty((ei.shape[0],), dtype=torch.float64, device=e.device)
        cnt, const_n = None, 1
        if np.isscalar(counts):
            const_n = int(counts)
        else:
            cnt = self._cuda(counts, torch.int32)
        with torch.cuda.device(self._dev()):
            _lib.check(_lib.load().ws_plda_score_trials(
                self._handle(), e.data_ptr(), cnt.data_ptr() if cnt is not None else None, const_n, e.shape[0],
                t.data_ptr(), t.shape[0], ei.data_ptr(), ti.data_ptr(), ei.shape[0], out.data_ptr(),
                _lib.cur_stream_ptr(self._dev())), "ws_plda_score_trials")
        return out

    # ------------------------------------------------------------------ reference-shaped scalar API
    def transform_embedding(self, embedding):
        """two_cov_plda.py:156-163 for one (D,) vector: A x + offset (+ length-norm); like the reference, callers
        pre-normalise with norm_embeddings themselves (eval_sv :225-241).  Input is taken as float32."""
        x = np.asarray(embedding, dtype=np.float32)
        return self.transform_batch(x[None], pre_norm=False).cpu().numpy()[0]

    def log_likelihood_ratio(self, transformed_train_embedding, transformed_test_embedding, n):
        """two_cov_plda.py:165-184 for one trial (runs the same device kernel with one trial)."""
        e = np.asarray(transformed_train_embedding, dtype=np.float64)[None]
        t = np.asarray(transformed_test_embedding, dtype=np.float64)[None]
        cnt = np.asarray([int(n)], dtype=np.int32)
        return float(self.score_trials(e, t, np.zeros(1, np.int64), np.zeros(1, np.int64), cnt).cpu()[0])

    def eval_sv(self, enroll_scp, enroll_utt2spk, test_scp, trials, score_file, multisession_avg=True,
                indomain_scp=None):
        """two_cov_plda.py:186-256 with the per-trial Python loop replaced by one device launch."""
        enroll = read_vec_scp_file(enroll_scp)
        labels = read_label_file(enroll_utt2spk)
        test = read_vec_scp_file(test_scp)
        mean_vec = None
        if indomain_scp is not None:
            mean_vec = np.vstack(list(read_vec_scp_file(indomain_scp).values())).astype(np.float64).mean(0)
        spk_sessions = {}
        for key, vec in enroll.items():
            if key in labels:
                spk_sessions.setdefault(labels[key], []).append(vec)
            else:
                print(f"WARNING: {key} not in utt2spk ({enroll_utt2spk}), skipping it.")
        spks = list(spk_sessions)
        # per-speaker mean of (sessions - mean_vec): mean commutes with the shift
        spk_mean = np.stack([np.mean(np.vstack(spk_sessions[s]).astype(np.float64), 0) for s in spks])
        counts = np.array([1 if multisession_avg else len(spk_sessions[s]) for s in spks], dtype=np.int32)
        e_t = self.transform_batch(spk_mean.astype(np.float32), mean_vec)
        tkeys = list(test)
        t_t = self.transform_batch(np.stack([test[k] for k in tkeys]).astype(np.float32), mean_vec)
        sidx = {s: i for i, s in enumerate(spks)}
        tidx = {k: i for i, k in enumerate(tkeys)}
        lines, ei, ti = [], [], []
        with open(trials) as f:
            for line in f:
                seg = line.strip().split()
                if not seg:
                    continue
                lines.append(seg)
                ei.append(sidx[seg[0]])
                ti.append(tidx[seg[1]])
        scores = self.score_trials(e_t, t_t, np.asarray(ei, np.int64), np.asarray(ti, np.int64), counts).cpu().numpy()
        with open(score_file, "w") as w:
            for seg, s in zip(lines, scores):
                w.write("{} {} {:.5f} {}\n".format(seg[0], seg[1], s, seg[2]))
