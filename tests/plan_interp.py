"""Host re-evaluation of a launch plan written by ws_engine_plan_trace (plan-check engine: no device, nothing computed by
the library).  TEST INFRASTRUCTURE: it checks the plan BUILDER's arithmetic - BN folding, channel padding, torch.cat /
torch.split as channel slices, shortcut convs merged as K ranges, strided convs as parity planes, tap offsets, residuals,
activations - against the oracle on the CPU; the CUDA kernels that execute the same plan are checked on the GPU.

Memory model: every placeholder allocation becomes a flat float64 array indexed in ELEMENTS (address offset / element size of
the tensor's dtype in the plan); values are never rounded, so a bf16 plan and an fp32 plan both reproduce the oracle to
floating-point accuracy.  Ops the interpreter does not model (fused ECAPA / CAM++ kernels) raise NotImplementedError."""
import json
import struct

import numpy as np


def load_trace(path):
    with open(path, "rb") as f:
        assert f.read(6) == b"WSPT1\n"
        (n,) = struct.unpack("<Q", f.read(8))
        meta = json.loads(f.read(n).decode())
        blob = np.frombuffer(f.read(), dtype="<f4")
    return meta, blob


class Memory:
    def __init__(self, meta, blob):
        self.allocs = sorted((a, n) for a, n in meta["allocs"])
        self.starts = np.array([a for a, _ in self.allocs], dtype=np.uint64)
        self.bufs = {}      # base address -> (float64 array, element size)
        self.blobs = {a: blob[o:o + n].astype(np.float64) for a, o, n in meta["blobs"]}

    def _find(self, addr):
        i = int(np.searchsorted(self.starts, np.uint64(addr), side="right")) - 1
        assert i >= 0, addr
        base, nbytes = self.allocs[i]
        assert base <= addr < base + max(nbytes, 1), (addr, base, nbytes)
        return base, nbytes

    def array(self, addr, es):
        """(flat array, element offset) of the allocation holding `addr`, elements of `es` bytes."""
        base, nbytes = self._find(addr)
        if base in self.blobs:              # weights: fp32 source values, whatever the plan's storage dtype
            assert (addr - base) % es == 0
            return self.blobs[base], (addr - base) // es
        if base not in self.bufs:
            self.bufs[base] = (np.zeros(nbytes // es, dtype=np.float64), es)
        arr, es0 = self.bufs[base]
        assert es0 == es and (addr - base) % es == 0, (addr, es0, es)
        return arr, (addr - base) // es

    def vec(self, addr, n, es=4):
        if not addr:
            return None
        arr, off = self.array(addr, es)
        return arr[off:off + n]

    def strided(self, addr, es, shape, strides, write=False):
        arr, off = self.array(addr, es)
        need = off + sum((s - 1) * st for s, st in zip(shape, strides)) + 1
        assert need <= arr.size, (addr, shape, strides, arr.size)
        v = np.lib.stride_tricks.as_strided(arr[off:], shape=shape, strides=[st * 8 for st in strides], writeable=write)
        return v

    def view(self, v, es, write=False):
        """channels-last View {p,B,F,T,C,ld} -> array (B,F,T,C)."""
        B, F, T, C, ld = v["B"], v["F"], v["T"], v["C"], v["ld"]
        return self.strided(v["p"], es, (B, F, T, C), (F * T * ld, T * ld, ld, 1), write)


ROUND = None   # None: exact re-evaluation; "bf16" / "fp16": emulate the plan's 16-bit storage (activations and conv weights
               # rounded where the kernels round them) to PREDICT the 16-bit paths' distance from the oracle


def rnd(x, es):
    if ROUND is None or es != 2:
        return x
    import torch
    dt = torch.bfloat16 if ROUND == "bf16" else torch.float16
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dt).to(torch.float32).numpy().astype(np.float64)


def act(x, code):
    if code == 0:
        return x
    if code == 1:
        return np.maximum(x, 0.0)
    if code == 2:
        return np.tanh(x)
    if code == 3:
        return 1.0 / (1.0 + np.exp(-x))
    if code == 4:
        return np.clip(x, 0.0, 20.0)
    if code == 5:
        return x / (1.0 + np.exp(-x))
    raise NotImplementedError(code)


def shifted(src, df, dt, F, T):
    """src (B,Fs,Ts,C) -> (B,F,T,C): out[b,f,t] = src[b,f+df,t+dt], zero outside the source extents."""
    B, Fs, Ts, C = src.shape
    out = np.zeros((B, F, T, C))
    f0, f1 = max(0, -df), min(F, Fs - df)
    t0, t1 = max(0, -dt), min(T, Ts - dt)
    if f1 > f0 and t1 > t0:
        out[:, f0:f1, t0:t1] = src[:, f0 + df:f1 + df, t0 + dt:t1 + dt]
    return out


def run_conv(mem, tr):
    es = tr["es"]
    for k in ("rowbias", "gate", "out2", "add2", "colsum"):
        if tr[k]:
            raise NotImplementedError("conv epilogue feature " + k)
    B, F, T, Cout, K = tr["B"], tr["F"], tr["T"], tr["Cout"], tr["Ktot"]
    W = rnd(mem.vec(tr["W"], Cout * K).reshape(Cout, K), es)
    srcs = [mem.strided(s["p"], es, (s["B"], s["F"], s["T"], s["C"]), (s["sB"], s["sF"], s["sT"], 1)) for s in tr["src"]]
    acc = np.zeros((B, F, T, Cout))
    for si, c0, dt, df, wk, nch in tr["taps"]:
        x = shifted(srcs[si][..., c0:c0 + nch], df, dt, F, T)
        acc += x @ W[:, wk:wk + nch].T
    if tr["bias"]:
        acc += mem.vec(tr["bias"], Cout)
    acc = act(acc, tr["act1"])
    if tr["scale"]:
        acc = acc * mem.vec(tr["scale"], Cout) + mem.vec(tr["shift"], Cout)
    if tr["res"]:
        acc = acc + mem.view(dict(p=tr["res"], B=B, F=F, T=T, C=Cout, ld=tr["res_ld"]), es)
    acc = act(acc, tr["act2"])
    mem.view(dict(p=tr["out"], B=B, F=F, T=T, C=Cout, ld=tr["out_ld"]), es, write=True)[...] = rnd(acc, es)


def run_conv3x3(mem, tr):
    es = tr["es"]
    if tr["lens"]:
        raise NotImplementedError("masked conv3x3")
    x = mem.view(tr["x"], es)
    o = tr["out"]
    Cin, Cout, sf, st = tr["x"]["C"], o["C"], tr["sf"], tr["st"]
    W = rnd(mem.vec(tr["W"], Cout * 9 * Cin).reshape(Cout, 9, Cin), es)
    Fo, To = o["F"], o["T"]
    assert Fo == (x.shape[1] - 1) // sf + 1 and To == (x.shape[2] - 1) // st + 1
    acc = np.zeros((o["B"], Fo, To, Cout))
    for jf in range(3):
        for jt in range(3):
            full = shifted(x, jf - 1, jt - 1, x.shape[1], x.shape[2])      # full[b,f,t] = x[b,f+jf-1,t+jt-1]
            acc += full[:, ::sf, ::st][:, :Fo, :To] @ W[:, jf * 3 + jt].T
    if tr["bias"]:
        acc += mem.vec(tr["bias"], Cout)
    if tr["res"] is not None:
        acc = acc + mem.view(tr["res"], es)
    acc = {0: acc, 1: np.maximum(acc, 0.0), 2: np.clip(acc, 0.0, 20.0)}[tr["relu"]]
    mem.view(o, es, write=True)[...] = rnd(acc, es)


def run_stem(mem, tr, meta):
    if tr["lens"]:
        raise NotImplementedError("masked stem")
    o = tr["out"]
    B, Fd, T, C = o["B"], o["F"], o["T"], o["C"]
    feats = mem.vec(tr["feats"], B * T * Fd).reshape(B, T, Fd)
    w9 = mem.vec(tr["w9"], C * 9).reshape(C, 9)
    x = feats.transpose(0, 2, 1)[..., None]                                # (B,F,T,1)
    acc = np.zeros((B, Fd, T, C))
    for jf in range(3):
        for jt in range(3):
            acc += shifted(x, jf - 1, jt - 1, Fd, T) * w9[:, jf * 3 + jt]
    acc = np.maximum(acc + mem.vec(tr["shift"], C), 0.0)
    mem.view(o, tr["es"], write=True)[...] = rnd(acc, tr["es"])


def run_tstats(mem, tr):
    if tr["lens"]:
        raise NotImplementedError("masked tstats")
    x = mem.view(tr["x"], tr["es"])
    B, F, T, C = x.shape
    if tr["pre_scale"]:
        x = np.maximum(x * mem.vec(tr["pre_scale"], C) + mem.vec(tr["pre_shift"], C), 0.0)
    out_ld, so = tr["out_ld"], tr["std_off"]
    out = mem.strided(tr["out"], 4, (B, out_ld), (out_ld, 1), write=True)
    mean = x.mean(axis=2)                                                  # (B,F,C) -> index c*F + f
    out[:, :C * F] = mean.transpose(0, 2, 1).reshape(B, C * F)
    if so >= 0:
        std = np.sqrt(x.var(axis=2, ddof=1) + 1e-7)
        out[:, so:so + C * F] = std.transpose(0, 2, 1).reshape(B, C * F)


def run_linear(mem, tr):
    R, I, O = tr["R"], tr["I"], tr["O"]
    x = mem.strided(tr["in"], 4, (R, I), (tr["in_ld"], 1)).copy()
    if tr["in2"]:
        x2 = mem.strided(tr["in2"], 4, (R // tr["rows_per_b"], I), (tr["in2_ld"], 1))
        x = x + np.repeat(x2, tr["rows_per_b"], axis=0)
    W = mem.vec(tr["W"], O * I).reshape(O, I)
    y = x @ W.T
    if tr["bias"]:
        y = y + mem.vec(tr["bias"], O)
    mem.strided(tr["out"], 4, (R, O), (tr["out_ld"], 1), write=True)[...] = act(y, tr["act"])


def run_aff_combine(mem, tr):
    es = tr["es"]
    x, y, t = mem.view(tr["x"], es), mem.view(tr["y"], es), mem.view(tr["t"], es)
    att = 1.0 + t
    mem.view(tr["out"], es, write=True)[...] = rnd(x * att + y * (2.0 - att), es)


def run_convert(mem, tr):
    n = tr["n"]
    mem.array(tr["out"], tr["es"])  # materialise with the right element size
    arr, off = mem.array(tr["out"], tr["es"])
    arr[off:off + n] = rnd(mem.vec(tr["in"], n), tr["es"])


def run_plan(path, feats):
    """feats (B,T,feat_dim) float -> embeddings (B,embed_dim) float64 by re-evaluating the traced plan on the host."""
    meta, blob = load_trace(path)
    B, T, Fd, E = meta["B"], meta["T"], meta["feat_dim"], meta["embed_dim"]
    assert feats.shape == (B, T, Fd)
    mem = Memory(meta, blob)
    arr, off = mem.array(meta["feats_in"], 4)
    arr[off:off + B * T * Fd] = np.asarray(feats, np.float64).reshape(-1)
    for op in meta["ops"]:
        tr = op["trace"]
        if tr is None:
            raise NotImplementedError("no trace for op " + op["label"])
        kind = tr["kind"]
        if kind == "conv":
            run_conv(mem, tr)
        elif kind == "conv3x3":
            run_conv3x3(mem, tr)
        elif kind == "stem":
            run_stem(mem, tr, meta)
        elif kind == "tstats":
            run_tstats(mem, tr)
        elif kind == "linear":
            run_linear(mem, tr)
        elif kind == "aff_combine":
            run_aff_combine(mem, tr)
        elif kind == "convert":
            run_convert(mem, tr)
        else:
            raise NotImplementedError(kind)
    return mem.vec(meta["emb"], B * E).reshape(B, E).copy(), meta
