"""Embedding processing chain — mirror of `wespeaker/utils/embedding_processing.py`: the same chain
strings (`"mean-subtract --scp X | length-norm | lda --scp X --utt2spk U --dim 100 | length-norm"`), class names and
call contract (numpy (N, D) in, numpy out), with the arithmetic in device-resident torch fp64 (library GEMMs and
symmetric eigendecompositions: one-time D x D estimation, then a GEMM per batch).  The per-speaker Python loop of the LDA
statistics (`:85-127`) becomes segment sums over the whole embedding matrix.
"""
from __future__ import annotations

import pickle
import re

import numpy as np
import torch

from .kaldi_io import load_scp_sequential, read_vec_scp_file
from .plda import read_label_file


def _dev(device=None):
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("wespeaker_b200.embedding_processing defaults to the GPU; pass device='cpu' explicitly")
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def chain_string_to_dict(chain_string=None):
    """`embedding_processing.py:23-66`: "a --x 1 | b" -> [['a', {'x': '1'}], ['b', {}]]."""
    links = chain_string.split("|") if chain_string is not None else []
    a = []
    for l in links:
        x = l.split("--")
        method = x.pop(0).strip(" ")
        args_and_values = {}
        for xx in x:
            xx = re.sub(" +", " ", re.sub("=", " ", xx)).strip(" ").split(" ")
            assert len(xx) == 2
            args_and_values[xx[0]] = xx[1]
        a.append([method, args_and_values])
    return a


def _t(x, dev):
    if torch.is_tensor(x):
        return x.to(device=dev, dtype=torch.float64)
    return torch.as_tensor(np.asarray(x), dtype=torch.float64, device=dev)


class _Link:
    """Common behaviour of the chain links.  State is numpy only (what gets pickled: no device handles, so a chain saved
    on one host loads on another and pickles written by the reference's classes load by attribute); the device copies are
    created lazily.  ``__call__`` keeps a torch tensor on its device (the chain runs end to end on the GPU with ONE host
    round trip) and returns numpy for numpy input, like the reference."""
    _state = ()

    def _device(self):
        d = getattr(self, "device", None)
        return _dev(d) if d is not None or torch.cuda.is_available() else torch.device("cpu")

    def __getstate__(self):
        return {k: np.asarray(getattr(self, k)) for k in self._state}

    def __setstate__(self, st):
        for k, v in st.items():
            if k in self._state:
                setattr(self, k, np.asarray(v))
        self.device = None

    def _dev_state(self, name, dev):
        cache = self.__dict__.setdefault("_cache", {})
        key = (name, str(dev))
        if key not in cache:
            cache[key] = _t(getattr(self, name), dev)
        return cache[key]

    def apply(self, x: torch.Tensor) -> torch.Tensor:   # fp64 tensor on some device -> same
        raise NotImplementedError

    def __call__(self, embd):
        if torch.is_tensor(embd):
            return self.apply(embd.to(torch.float64))
        return self.apply(_t(embd, self._device())).cpu().numpy()


class Lda(_Link):
    _state = ("m", "lda")

    def compute_mean_and_lda_scatter_matrices(self, scp_file, utt2spk_file, equal_speaker_weight=False, current_chain=None):
        """`:70-127`.  Speakers with a single utterance are skipped (count > 1 rule)."""
        samples = read_vec_scp_file(scp_file)
        lab = read_label_file(utt2spk_file)
        keys = [k for k in samples if k in lab]
        x = np.stack([samples[k] for k in keys])
        dev = self._device()
        xt = _t(x, dev)
        if current_chain is not None and not isinstance(current_chain, list):
            xt = _t(current_chain(xt), dev)                       # stays on the device through the links built so far
        names = {}
        cls = np.array([names.setdefault(lab[k], len(names)) for k in keys])
        ct = torch.as_tensor(cls, device=dev)
        nspk = len(names)
        cnt = torch.bincount(ct, minlength=nspk).to(torch.float64)
        means = torch.zeros((nspk, xt.shape[1]), dtype=torch.float64, device=dev).index_add_(0, ct, xt) / cnt[:, None]
        used = cnt > 1
        xc = (xt - means[ct]) * used[ct][:, None]                 # rows of skipped speakers contribute nothing
        cnt_u, means_u = cnt[used], means[used]
        print("  #speakers: {}, #used {}, #skipped {} (only having one utterances)".format(
            nspk, int(used.sum()), int((~used).sum())))
        if equal_speaker_weight:
            mean = means_u.mean(dim=0)
            d = means_u - mean
            bc = d.T @ d / means_u.shape[0]
            wc = ((xc / cnt[ct][:, None].sqrt()).T @ (xc / cnt[ct][:, None].sqrt())) / nspk       # sum_s cov_s / len(speakers)
        else:
            tot = cnt_u.sum()
            mean = (cnt_u[:, None] * means_u).sum(dim=0) / tot
            d = means_u - mean
            bc = (d * cnt_u[:, None]).T @ d / tot
            wc = xc.T @ xc / tot
        return mean, bc, wc

    def __init__(self, args, current_chain=None, device=None):
        print(" LDA")
        self.device = device if device is not None else getattr(current_chain, "device", None)
        dim = int(args["dim"])
        eps = float(args["eps"]) if "eps" in args else 1e-6
        m, bc, wc = self.compute_mean_and_lda_scatter_matrices(args["scp"], args["utt2spk"], current_chain=current_chain)
        e, mm = torch.linalg.eigh(wc)
        e = torch.clamp(e, min=float(e.max()) * eps)              # floor the within-class eigenvalues (as Kaldi does)
        t1 = torch.diag(1.0 / torch.sqrt(e)) @ mm.T
        d, lda = torch.linalg.eigh(t1 @ bc @ t1.T)
        self.m = m.cpu().numpy()
        self.lda = (t1.T @ lda[:, -dim:]).cpu().numpy()
        print("  Input dimension: {}, output dimension: {}, sum of all eigenvalues {:.2f}, sum of kept eigenvalues {:.2f}".format(
            len(d), dim, float(d.sum()), float(d[-dim:].sum())))

    def apply(self, x):
        return (x - self._dev_state("m", x.device)) @ self._dev_state("lda", x.device)


class Length_norm(_Link):
    def __init__(self, args=None, current_chain=None, device=None):
        self.device = device if device is not None else getattr(current_chain, "device", None)

    def apply(self, x):
        return x / torch.sqrt((x ** 2).sum(dim=1, keepdim=True))


class MeanSubtraction(_Link):
    _state = ("mean",)

    def __init__(self, args, current_chain=None, device=None):
        self.device = device if device is not None else getattr(current_chain, "device", None)
        e = _t(np.vstack([vec for _, vec in load_scp_sequential(args["scp"])]), self._device())
        if current_chain is not None and not isinstance(current_chain, list):
            e = _t(current_chain(e), e.device)
        self.mean = e.mean(dim=0).cpu().numpy()

    def apply(self, x):
        return x - self._dev_state("mean", x.device)


class _RefUnpickler(pickle.Unpickler):
    """Chains pickled by the reference (`wespeaker.utils.embedding_processing.<Class>` objects holding numpy arrays) load
    as the classes of this module."""

    def find_class(self, module, name):
        if module.endswith("embedding_processing") and name in ("Lda", "Length_norm", "MeanSubtraction"):
            return globals()[name]
        return super().find_class(module, name)


class EmbeddingProcessingChain:
    """`:221-271`.  Each link is estimated on data passed through the links built so far (the chain hands ITSELF to the
    link constructors, exactly like the reference).  A call moves the batch to the device once, applies every link there
    and comes back once."""
    string2class = {"lda": Lda, "length-norm": Length_norm, "mean-subtract": MeanSubtraction}

    def __init__(self, chain=None, device=None):
        self.device = _dev(device)
        self.chain_of_classes = []
        for m, a in chain_string_to_dict(chain):
            print("Method: {}".format(m))
            print("Argument: {}".format(a))
            self.chain_of_classes.append(self.string2class[m](a, self))

    def __call__(self, embd):
        as_tensor = torch.is_tensor(embd)
        x = _t(embd, embd.device if as_tensor and embd.is_cuda else self.device)
        for c in self.chain_of_classes:
            x = c.apply(x) if isinstance(c, _Link) else _t(c(x.cpu().numpy()), x.device)
        return x if as_tensor else x.cpu().numpy()

    def save(self, path, data_format="pickle"):
        print("Saving embedding processing chain to {}".format(path))
        with open(path, "wb") as f:
            pickle.dump(self.chain_of_classes, f)

    def load(self, path, data_format="pickle"):
        print("Loading embedding processing chain from {}".format(path))
        with open(path, "rb") as f:
            self.chain_of_classes = _RefUnpickler(f).load()
        for c in self.chain_of_classes:
            if isinstance(c, _Link):
                c.device = self.device

    def update_link(self, link_no_to_replace, new_link):
        nl = chain_string_to_dict(new_link)
        assert len(nl) == 1, "Length of new chain must be one."
        m, a = nl[0]
        old, self.chain_of_classes = self.chain_of_classes, []
        for i, ol in enumerate(old):
            if i != link_no_to_replace:
                self.chain_of_classes.append(ol)
            else:
                print("Replacing link number {} ({}) with".format(i, ol))
                self.chain_of_classes.append(self.string2class[m](a, self))
