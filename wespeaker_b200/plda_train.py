"""Two-covariance PLDA training (EM) and unsupervised adaptation — mirror of
`wespeaker/utils/plda/two_cov_plda.py:39-154` (PldaStats, TwoCovPLDA.__init__/train/em_one_iter/get_output) and `:258-309`
(adapt); SURVEY.md §8f rank 3.

These are one-time D x D dense-algebra jobs (D = 256), not part of the extraction/scoring hot path, so they are written as
device-resident torch fp64 linear algebra (cuBLAS / cuSOLVER through torch: plain library GEMMs, inverses and symmetric
eigendecompositions) rather than hand-written kernels.  What changes against the reference is the shape of the work: the
reference loops over classes in Python with one D x D inverse per class; here the per-class statistics are segment sums
over the whole (N, D) embedding matrix, the posterior covariance `inv(B^-1 + n W^-1)` is computed once per DISTINCT
session count n (batched inverse), and the B / W accumulations are two GEMMs over the class-mean matrix.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .kaldi_io import read_vec_scp_file
from .plda import TwoCovPLDA, read_label_file


def _dev(device):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("wespeaker_b200.plda_train defaults to the GPU; pass device='cpu' explicitly to run elsewhere")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def norm_embeddings(x: torch.Tensor, kaldi_style: bool = True) -> torch.Tensor:
    """`plda_utils.py:46-58` on a (N, D) tensor."""
    scale = math.sqrt(x.shape[-1]) if kaldi_style else 1.0
    return scale * x / torch.linalg.norm(x, dim=-1, keepdim=True)


def compute_normalizing_transform(covar: torch.Tensor) -> torch.Tensor:
    """`plda_utils.py:79-85`: inverse Cholesky factor (with the reference's 1e-6 jitter retry)."""
    c, info = torch.linalg.cholesky_ex(covar)
    if int(info) != 0:
        c = torch.linalg.cholesky(covar + torch.eye(covar.shape[0], dtype=covar.dtype, device=covar.device) * 1e-6)
    return torch.linalg.inv(c)


class PldaStats:
    """`two_cov_plda.py:39-64`, built in one pass from an (N, D) matrix and integer class ids (all weights 1)."""

    def __init__(self, x: torch.Tensor, cls: torch.Tensor, num_classes: int):
        n, d = x.shape
        self.dim = d
        self.num_example, self.num_classes = n, num_classes
        self.class_weight, self.example_weight = float(num_classes), float(n)
        self.count = torch.bincount(cls, minlength=num_classes).to(x.dtype)               # n_k
        self.class_mean = torch.zeros((num_classes, d), dtype=x.dtype, device=x.device).index_add_(0, cls, x)
        self.class_mean /= self.count[:, None]
        xc = x - self.class_mean[cls]
        self.offset_scatter = xc.T @ xc                                                   # sum_k (X_k - m_k)^T (X_k - m_k)
        self.sum_ = self.class_mean.sum(dim=0)                                            # sum of class means


class TwoCovPLDATrainer:
    """Same constructor arguments, attributes and methods as the training half of the reference `TwoCovPLDA`."""

    def __init__(self, scp_file=None, utt2spk_file=None, embed_dim=256, subtract_train_set_mean=False,
                 normalize_length=False, device=None, embeddings=None, labels=None):
        self.subtract_train_set_mean = subtract_train_set_mean
        self.normalize_length = normalize_length
        self.dim = embed_dim
        self.device = _dev(device)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.B, self.W = torch.eye(embed_dim, **f64), torch.eye(embed_dim, **f64)
        self.mu = np.zeros(embed_dim)
        self.transform, self.psi, self.offset = np.zeros((embed_dim, embed_dim)), np.zeros(embed_dim), np.zeros(embed_dim)
        self.stats = None
        if scp_file is not None:
            samples = read_vec_scp_file(scp_file)                                         # get_data_for_plda, plda_utils.py:61-76
            lab = read_label_file(utt2spk_file)
            embeddings = np.stack(list(samples.values()))
            labels = [lab.get(k) for k in samples]
            for k, l in zip(samples, labels):
                if l is None:
                    print("WARNING: {} not in utt2spk ({}), skipping it.".format(k, utt2spk_file))
        if embeddings is not None:
            x_all = torch.as_tensor(np.asarray(embeddings), dtype=torch.float64, device=self.device)
            keep = [i for i, l in enumerate(labels) if l is not None]
            names = {}
            cls = torch.tensor([names.setdefault(labels[i], len(names)) for i in keep], dtype=torch.long, device=self.device)
            train_mean = x_all.mean(dim=0) if subtract_train_set_mean else torch.zeros(embed_dim, **f64)   # mean over ALL samples
            x = x_all[torch.tensor(keep, device=self.device)] - train_mean
            if normalize_length:
                x = norm_embeddings(x)
            self.stats = PldaStats(x, cls, len(names))
            self.mu = (self.stats.sum_ / self.stats.class_weight).cpu().numpy()

    def train(self, num_em_iters):
        for i in range(num_em_iters):
            print("Plda estimation %d of %d" % (i, num_em_iters))
            self.em_one_iter()
        self.get_output()

    def em_one_iter(self):
        """`two_cov_plda.py:112-139` with the class loop replaced by per-distinct-n batched algebra."""
        st = self.stats
        b_inv, w_inv = torch.linalg.inv(self.B), torch.linalg.inv(self.W)
        m = st.class_mean - st.sum_ / st.class_weight                                     # (K, D)
        ns, inverse, cnt = torch.unique(st.count, return_inverse=True, return_counts=True)
        mix_var = torch.linalg.inv(b_inv[None] + ns[:, None, None] * w_inv[None])         # (U, D, D), one per distinct n
        wm = m @ w_inv.T                                                                   # rows: W^-1 m_k
        # w_k = mix_var_{n_k} (n_k W^-1 m_k): one GEMM per distinct n (no (K, D, D) gather of the covariances)
        w = torch.empty_like(wm)
        for u in range(ns.shape[0]):
            sel = inverse == u
            w[sel] = ns[u] * (wm[sel] @ mix_var[u].T)
        m_w = m - w
        cntf = cnt.to(mix_var.dtype)
        b_stats = (cntf[:, None, None] * mix_var).sum(dim=0) + w.T @ w
        w_stats = st.offset_scatter + ((cntf * ns)[:, None, None] * mix_var).sum(dim=0) + (m_w * st.count[:, None]).T @ m_w
        b_count = st.class_weight
        w_count = st.example_weight - st.class_weight + st.class_weight
        self.W = w_stats / w_count
        self.B = b_stats / b_count
        self.W = 0.5 * (self.W + self.W.T)
        self.B = 0.5 * (self.B + self.B.T)
        print("W_count:", w_count, "Trace of W:", float(torch.trace(self.W)))
        print("B_count:", b_count, "Trace of B:", float(torch.trace(self.B)))

    def get_output(self):
        """`two_cov_plda.py:141-154`."""
        st = self.stats
        mu = st.sum_ / st.class_weight
        t1 = compute_normalizing_transform(self.W)
        b_proj = t1 @ self.B @ t1.T
        s, u = torch.linalg.eigh(b_proj)
        s = torch.where(s > 0.0, s, torch.zeros_like(s))
        idx = torch.argsort(-s)                                                            # sort_svd, plda_utils.py:88-100
        s, u = s[idx], u[:, idx]
        transform = u.T @ t1
        self.mu, self.transform, self.psi = mu.cpu().numpy(), transform.cpu().numpy(), s.cpu().numpy()
        self.offset = -1.0 * (transform @ mu).cpu().numpy()

    def save_model(self, output_file_name):
        """`two_cov_plda.py:311-339`: same file formats as `plda.TwoCovPLDA.save_model` (HDF5 with the reference's dataset
        names when h5py is present, `.npz` otherwise)."""
        print("saving the trained plda to {}".format(output_file_name))
        self.to_plda().save_model(output_file_name)

    def to_plda(self, device=None) -> TwoCovPLDA:
        """The trained model as the GPU scorer (`plda.TwoCovPLDA`)."""
        return TwoCovPLDA.from_arrays(self.mu, self.transform, self.psi, self.offset, self.normalize_length,
                                      self.subtract_train_set_mean, device=device)


def adapt(mu, transform, psi, adapt_data, normalize_length=False, ac_scale=0.5, wc_scale=0.5, device=None):
    """`two_cov_plda.py:258-309` (BUT unsupervised adaptation): returns (mu, transform, psi, offset) of the adapted model.
    `adapt_data`: (N, D) embeddings or an scp path."""
    dev = _dev(device)
    f64 = dict(dtype=torch.float64, device=dev)
    if isinstance(adapt_data, str):
        adapt_data = np.array(list(read_vec_scp_file(adapt_data).values()))
    x = torch.as_tensor(np.asarray(adapt_data), **f64)
    x = x - x.mean(dim=0)
    if normalize_length:
        x = norm_embeddings(x)
    tr, ps = torch.as_tensor(np.asarray(transform), **f64), torch.as_tensor(np.asarray(psi), **f64)
    w = torch.linalg.inv(tr.T @ tr)
    w = (w + w.T) / 2
    b = torch.linalg.inv((tr.T / ps) @ tr)
    b = (b + b.T) / 2
    t = b + w
    t = (t + t.T) / 2
    xc = x - x.mean(dim=0)
    data_cov = xc.T @ xc / (x.shape[0] - 1)                                                # np.cov(adp_data.T)
    # generalised symmetric eigenproblem data_cov e = v T e with e^T T e = I (scipy.linalg.eigh(a, b))
    l = torch.linalg.cholesky(t)
    li = torch.linalg.inv(l)
    v, y = torch.linalg.eigh(li @ data_cov @ li.T)
    e = li.T @ y
    iet = torch.linalg.inv(e.T)
    sel = v > 1
    excess = iet[:, sel] * torch.sqrt(v[sel] - 1)[None, :]
    v_adp = excess * math.sqrt(ac_scale)
    b_adp = b + v_adp @ v_adp.T
    u_adp = excess * math.sqrt(wc_scale)
    w_adp = w + u_adp @ u_adp.T
    mu_adp = x.mean(dim=0)
    a_m, b_m = (b_adp + b_adp.T) / 2.0, (w_adp + w_adp.T) / 2.0
    d, vv = torch.linalg.eigh(b_m)
    t1 = torch.diag(1.0 / torch.sqrt(d + 1e-9)) @ vv.T
    a1 = t1 @ a_m @ t1.T
    _, t2 = torch.linalg.eigh(a1)
    tt = t2.T @ t1
    a2 = tt @ a_m @ tt.T
    new_psi = torch.diagonal(a2)
    return (mu_adp.cpu().numpy(), tt.cpu().numpy(), new_psi.cpu().numpy(), (-1.0 * (tt @ mu_adp)).cpu().numpy())


def adapt_model(plda: TwoCovPLDA, adapt_scp, ac_scale=0.5, wc_scale=0.5, device=None) -> TwoCovPLDA:
    """`TwoCovPLDA.adapt(adapt_scp, ac_scale, wc_scale)` (`two_cov_plda.py:258-309`): the adapted model as a scorer object
    (which has `save_model`), like the reference method returns a new TwoCovPLDA."""
    mu, tr, psi, off = adapt(plda.mu, plda.transform, plda.psi, adapt_scp, normalize_length=plda.normalize_length,
                             ac_scale=ac_scale, wc_scale=wc_scale, device=device)
    # the reference's adapted object is a fresh TwoCovPLDA(): normalize_length / subtract_train_set_mean at their defaults
    return TwoCovPLDA.from_arrays(mu, tr, psi, off, False, False, device=plda._device)
