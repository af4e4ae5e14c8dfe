"""Host re-evaluation of a launch plan written by ws_engine_plan_trace (plan-check engine: no device, nothing computed by
the library).  TEST INFRASTRUCTURE: it checks the plan BUILDER's arithmetic - BN folding, channel padding, torch.cat /
torch.split as channel slices, shortcut convs merged as K ranges, strided convs as parity planes, tap offsets, residuals,
activations - against the oracle on the CPU; the CUDA kernels that execute the same plan are checked on the GPU.

Memory model: every placeholder allocation becomes a flat float64 array indexed in ELEMENTS (address offset / element size of
the tensor's dtype in the plan); values are never rounded, so a bf16 plan and an fp32 plan both reproduce the oracle to
floating-point accuracy.  The fused kernels (Res2 chain, ASTP tail, CAM++ dense block, SE gate) are re-evaluated from their
documented contracts (ws_host.h).  Length-masked plans are modelled too: rows behind an utterance's end are conv padding and
every statistic runs over the utterance's own frames."""
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


def _tf32_trunc(a):
    return (np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32x3_matmul(x, Wt):
    """The engine's 3xTF32 product x @ Wt as the tensor core sees it (ROUND = "tf32x3" or "tf32x3_trunc_lo"): operands are
    truncated to tf32; x_lo * W_hi + x_hi * W_lo + x_hi * W_hi with lo = the residual rounded to the nearest tf32 (the engine's
    ws_tf32_lo) or, for comparison, left to the hardware's truncation."""
    def split(a):
        a32 = np.ascontiguousarray(a, dtype=np.float32)
        hi = _tf32_trunc(a32)
        d = a32 - hi
        lo = ((d.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32) if ROUND == "tf32x3" else _tf32_trunc(d)
        return hi.astype(np.float64), lo.astype(np.float64)
    xh, xl = split(x)
    wh, wl = split(Wt)
    return xl @ wh + xh @ wl + xh @ wh


def rnd(x, es):
    if ROUND in ("tf32x3", "tf32x3_trunc_lo"):
        return np.asarray(x, np.float32).astype(np.float64) if es == 4 else x   # fp32 storage
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
    B, F, T, Cout, K = tr["B"], tr["F"], tr["T"], tr["Cout"], tr["Ktot"]   # (colsum, the fused SE squeeze, is a side output)
    W = rnd(mem.vec(tr["W"], Cout * K).reshape(Cout, K), es)
    srcs = [mem.strided(s["p"], es, (s["B"], s["F"], s["T"], s["C"]), (s["sB"], s["sF"], s["sT"], 1)) for s in tr["src"]]
    acc = np.zeros((B, F, T, Cout))
    for si, c0, dt, df, wk, nch in tr["taps"]:
        x = shifted(srcs[si][..., c0:c0 + nch], df, dt, F, T)
        acc += tf32x3_matmul(x, W[:, wk:wk + nch].T) if ROUND in ("tf32x3", "tf32x3_trunc_lo") and es == 4 else x @ W[:, wk:wk + nch].T
    if tr["bias"]:
        acc += mem.vec(tr["bias"], Cout)
    if tr["rowbias"]:      # per-utterance bias row (ECAPA global-context attention)
        acc += mem.strided(tr["rowbias"], 4, (B, Cout), (tr["rowbias_ld"], 1))[:, None, None, :]
    acc = act(acc, tr["act1"])
    if tr["scale"]:
        acc = acc * mem.vec(tr["scale"], Cout) + mem.vec(tr["shift"], Cout)
    if tr["gate"]:         # CAM context mask: gate[b][t / gate_seg][c]
        nseg, gl = tr["gate_nseg"], tr["gate_ld"]
        g = mem.strided(tr["gate"], 4, (B, nseg, Cout), (nseg * gl, gl, 1))
        acc = acc * g[:, np.arange(T) // tr["gate_seg"], :][:, None]
    if tr["res"]:
        acc = acc + mem.view(dict(p=tr["res"], B=B, F=F, T=T, C=Cout, ld=tr["res_ld"]), es)
    acc = act(acc, tr["act2"])
    acc = rnd(acc, es)
    mem.view(dict(p=tr["out"], B=B, F=F, T=T, C=Cout, ld=tr["out_ld"]), es, write=True)[...] = acc
    if tr["out2"]:         # Res2 chain: s_{i+1} = sp_i + x_{i+1}
        add2 = mem.view(dict(p=tr["add2"], B=B, F=F, T=T, C=Cout, ld=tr["add2_ld"]), es)
        mem.view(dict(p=tr["out2"], B=B, F=F, T=T, C=Cout, ld=tr["out2_ld"]), es, write=True)[...] = rnd(acc + add2, es)


def run_conv3x3(mem, tr):
    es = tr["es"]
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
    if tr["lens"]:      # length-masked batch: output rows behind an utterance's end are stored as zeros
        acc = acc * (np.arange(To)[None, :] < mem.vec(tr["lens"], o["B"])[:, None])[:, None, :, None]
    mem.view(o, es, write=True)[...] = rnd(acc, es)


def run_stem(mem, tr, meta):
    o = tr["out"]
    B, Fd, T, C = o["B"], o["F"], o["T"], o["C"]
    feats = mem.vec(tr["feats"], B * T * Fd).reshape(B, T, Fd)
    valid = np.ones((B, T), bool)
    if tr["lens"]:      # frames behind an utterance's end are padding: read as zeros, written as zeros
        valid = np.arange(T)[None, :] < mem.vec(tr["lens"], B)[:, None]
        feats = feats * valid[:, :, None]
    w9 = mem.vec(tr["w9"], C * 9).reshape(C, 9)
    x = feats.transpose(0, 2, 1)[..., None]                                # (B,F,T,1)
    acc = np.zeros((B, Fd, T, C))
    for jf in range(3):
        for jt in range(3):
            acc += shifted(x, jf - 1, jt - 1, Fd, T) * w9[:, jf * 3 + jt]
    acc = np.maximum(acc + mem.vec(tr["shift"], C), 0.0) * valid[:, None, :, None]
    mem.view(o, tr["es"], write=True)[...] = rnd(acc, tr["es"])


def run_tstats(mem, tr):
    x = mem.view(tr["x"], tr["es"])
    B, F, T, C = x.shape
    if tr["pre_scale"]:
        x = np.maximum(x * mem.vec(tr["pre_scale"], C) + mem.vec(tr["pre_shift"], C), 0.0)
    out_ld, so = tr["out_ld"], tr["std_off"]
    out = mem.strided(tr["out"], 4, (B, out_ld), (out_ld, 1), write=True)
    lens = mem.vec(tr["lens"], B).astype(int) if tr["lens"] else np.full(B, T)   # statistics over each utterance's own frames
    for b in range(B):
        xb = x[b, :, :lens[b]]
        out[b, :C * F] = xb.mean(axis=1).T.reshape(C * F)                       # (F,C) -> index c*F + f
        if so >= 0:
            out[b, so:so + C * F] = np.sqrt(xb.var(axis=1, ddof=1) + 1e-7).T.reshape(C * F)


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


def _dil_conv1d(x, W, dil):
    """x (B,T,Cin), W (Cout,3,Cin) tap-major, zero padding `dil` -> (B,T,Cout): sum_j x[t + (j-1)*dil] @ W[:, j].T"""
    B, T, _ = x.shape
    out = np.zeros((B, T, W.shape[0]))
    for j in range(3):
        off = (j - 1) * dil
        xs = shifted(x[:, None], 0, off, 1, T)[:, 0]
        out += xs @ W[:, j].T
    return out


def _lens(mem, tr, B, T):
    """frames per utterance of a length-masked op (clamped to [1, T]); T for every utterance in ordinary plans"""
    if not tr["lens"]:
        return np.full(B, T, dtype=int)
    return np.clip(mem.vec(tr["lens"], B).astype(int), 1, T)


def run_res2_fused(mem, tr):
    """ws_host.h make_res2_op: 7 dependent dilated k=3 convs on w8-channel groups (ecapa_tdnn.py:29-78):
    sp_i = bn(relu(conv(s_i) + bias_i)), s_0 = x_0, s_{i+1} = sp_i + x_{i+1}; out groups 0..6 = sp_0..sp_6."""
    es, w8, dil = tr["es"], tr["w8"], tr["dil"]
    xall = mem.view(tr["x"], es)[:, 0]                   # (B,T,8*w8)
    out = mem.view(tr["out"], es, write=True)
    W7 = rnd(mem.vec(tr["W7"], 7 * w8 * 3 * w8).reshape(7, w8, 3, w8), es)
    bias, scale, shift = (mem.vec(tr[k], 7 * w8).reshape(7, w8) for k in ("bias", "scale", "shift"))
    B, T, _ = xall.shape
    lens = _lens(mem, tr, B, T)
    for b in range(B):       # length-masked: rows behind the utterance's end are conv padding = the utterance alone
        x = xall[b:b + 1, :lens[b]]
        s = x[..., :w8]
        out[b, 0, lens[b]:, :7 * w8] = 0.0
        for i in range(7):
            sp = rnd(np.maximum(_dil_conv1d(s, W7[i], dil) + bias[i], 0.0) * scale[i] + shift[i], es)
            out[b, 0, :lens[b], i * w8:(i + 1) * w8] = sp[0]
            if i < 6:
                s = rnd(sp + x[..., (i + 1) * w8:(i + 2) * w8], es)


def run_se_gate(mem, tr):
    """gate[b][c] = sigmoid(W2 relu(W1 mean_T(x[b]) + b1) + b2); W2t is W2 transposed to [H][C] (ecapa_tdnn.py:113-126)."""
    x = mem.view(tr["x"], tr["es"])[:, 0]
    B, T, C = x.shape
    H = tr["H"]
    W1 = mem.vec(tr["W1"], H * C).reshape(H, C)
    W2t = mem.vec(tr["W2t"], H * C).reshape(H, C)
    lens = _lens(mem, tr, B, T)
    mean = np.stack([x[b, :lens[b]].mean(axis=0) for b in range(B)])
    hid = np.maximum(mean @ W1.T + mem.vec(tr["b1"], H), 0.0)
    mem.strided(tr["gate"], 4, (B, C), (C, 1), write=True)[...] = act(hid @ W2t + mem.vec(tr["b2"], C), 3)


def run_scale_residual(mem, tr):
    es = tr["es"]
    x, res = mem.view(tr["x"], es), mem.view(tr["res"], es)
    B, _, T, C = x.shape
    gate = mem.strided(tr["gate"], 4, (B, C), (C, 1))
    mem.view(tr["out"], es, write=True)[...] = rnd(x * gate[:, None, None, :] + res, es)


def _astp(x, logits):
    """pooling_layers.py:140-144: softmax over T, weighted mean and sqrt(clamp(E[x^2] - mean^2, 1e-7))."""
    a = np.exp(logits - logits.max(axis=1, keepdims=True))
    a /= a.sum(axis=1, keepdims=True)
    mean = (a * x).sum(axis=1)
    var = (a * x * x).sum(axis=1) - mean * mean
    return np.concatenate([mean, np.sqrt(np.maximum(var, 1e-7))], axis=1)


def run_astp_fused(mem, tr):
    """ws_host.h make_astp_op: logits = h @ W2^T (linear2's bias cancels in the softmax over time)."""
    es = tr["es"]
    x, h = mem.view(tr["x"], es)[:, 0], mem.view(tr["h"], es)[:, 0]
    B, T, C = x.shape
    W2 = rnd(mem.vec(tr["W2"], C * 128).reshape(C, 128), es)
    lens = _lens(mem, tr, B, T)
    st = mem.strided(tr["stats"], 4, (B, 2 * C), (2 * C, 1), write=True)
    for b in range(B):
        st[b] = _astp(x[b:b + 1, :lens[b]], h[b:b + 1, :lens[b]] @ W2.T)[0]


def run_astp_stats(mem, tr):
    es = tr["es"]
    x, lg = mem.view(tr["x"], es)[:, 0], mem.view(tr["logits"], es)[:, 0]
    B, T, C = x.shape
    lens = _lens(mem, tr, B, T)
    st = mem.strided(tr["stats"], 4, (B, 2 * C), (2 * C, 1), write=True)
    for b in range(B):
        st[b] = _astp(x[b:b + 1, :lens[b]], lg[b:b + 1, :lens[b]])[0]


def run_bnrelu(mem, tr):
    es, C = tr["es"], tr["C"]
    x = mem.view(tr["x"], es)[..., :C]
    mem.view(tr["out"], es, write=True)[..., :C] = rnd(np.maximum(x * mem.vec(tr["scale"], C) + mem.vec(tr["shift"], C), 0.0), es)


def _cam_context(h, seg_len):
    """campplus.py:108-135: mean over T + per-segment means (avg_pool1d, ceil_mode: the last segment averages its own frames),
    one context vector per segment: (B, nseg, C)."""
    B, T, C = h.shape
    nseg = (T + seg_len - 1) // seg_len
    seg = np.stack([h[:, s * seg_len:min(T, (s + 1) * seg_len)].mean(axis=1) for s in range(nseg)], axis=1)
    return h.mean(axis=1)[:, None, :] + seg


def run_cam_gate(mem, tr):
    es, H, G, L = tr["es"], tr["H"], tr["G"], tr["seg_len"]
    h = mem.view(tr["x"], es)[:, 0]
    B, T, C = h.shape
    ctx = _cam_context(h, L)
    W1 = mem.vec(tr["W1"], H * C).reshape(H, C)
    W2 = mem.vec(tr["W2"], G * H).reshape(G, H)
    hid = np.maximum(ctx @ W1.T + mem.vec(tr["b1"], H), 0.0)
    nseg = ctx.shape[1]
    mem.strided(tr["gate"], 4, (B, nseg, G), (nseg * G, G, 1), write=True)[...] = act(hid @ W2.T + mem.vec(tr["b2"], G), 3)


def run_seg_means(mem, tr):
    es, L = tr["es"], tr["seg_len"]
    h = mem.view(tr["x"], es)[:, 0]
    B, T, C = h.shape
    nseg = (T + L - 1) // L
    mem.strided(tr["mean"], 4, (B, C), (C, 1), write=True)[...] = h.mean(axis=1)
    seg = np.stack([h[:, s * L:min(T, (s + 1) * L)].mean(axis=1) for s in range(nseg)], axis=1)
    mem.strided(tr["segmean"], 4, (B, nseg, C), (nseg * C, C, 1), write=True)[...] = seg


def run_cam_dense(mem, tr):
    """ws_host.h make_cam_dense_op / campplus.py:86-201, per layer: h = relu(W1 relu(bn1(x[:, :cin])) + bias2) (BN2 folded into
    W1), context mask m = sigmoid(W2c relu(W1c (mean_T(h) + segmean(h)) + b1c) + b2c) per 100-frame segment,
    x[:, cin:cin+32] = (k3 dilated conv of h) * m."""
    es, L = tr["es"], tr["seg_len"]
    Xall = mem.view(tr["X"], es, write=True)
    B, _, Tmax, _ = Xall.shape
    lens = _lens(mem, tr, B, Tmax)
    for b in range(B):       # one utterance at a time over its own frames (what the one-CTA-per-utterance kernel does)
        T = lens[b]
        X = Xall[b:b + 1, :, :T]
        for ly in tr["layers"]:
            cin, dil = ly["cin"], ly["dil"]
            xin = rnd(np.maximum(X[:, 0, :, :cin] * mem.vec(ly["bn1_scale"], cin) + mem.vec(ly["bn1_shift"], cin), 0.0), es)
            W1 = rnd(mem.vec(ly["W1"], 128 * cin).reshape(128, cin), es)
            h = rnd(np.maximum(xin @ W1.T + mem.vec(ly["bias2"], 128), 0.0), es)
            ctx = _cam_context(h, L)                                           # (1,nseg,128)
            w1c_t = mem.vec(ly["w1c_t"], 128 * 64).reshape(128, 64)
            w2c_t = mem.vec(ly["w2c_t"], 64 * 32).reshape(64, 32)
            m = act(np.maximum(ctx @ w1c_t + mem.vec(ly["b1c"], 64), 0.0) @ w2c_t + mem.vec(ly["b2c"], 32), 3)   # (1,nseg,32)
            Wl = rnd(mem.vec(ly["Wl"], 32 * 3 * 128).reshape(32, 3, 128), es)
            y = _dil_conv1d(h, Wl, dil)
            seg_of_t = np.arange(T) // L
            X[:, 0, :, cin:cin + 32] = rnd(y * m[:, seg_of_t, :], es)
            Xall[b, 0, T:, cin:cin + 32] = 0.0


def run_lens_derive(mem, tr):
    """frames per utterance behind each stride-2 level: level 0 clamped to [1, T], level k + 1 = (level k - 1) // 2 + 1"""
    B, T, levels = tr["B"], tr["T"], tr["levels"]
    lens = mem.strided(tr["lens"], 4, (levels, B), (B, 1), write=True)
    lens[0] = np.clip(lens[0], 1, T)
    for k in range(1, levels):
        lens[k] = (lens[k - 1] - 1) // 2 + 1


def run_zero_tail(mem, tr):
    x = mem.view(tr["x"], tr["es"], write=True)
    B, F, T, C = x.shape
    x *= (np.arange(T)[None, :] < mem.vec(tr["lens"], B)[:, None])[:, None, :, None]


EXTRA = {"lens_derive": run_lens_derive, "zero_tail": run_zero_tail, "res2_fused": run_res2_fused, "se_gate": run_se_gate, "scale_residual": run_scale_residual, "astp_fused": run_astp_fused,
         "astp_stats": run_astp_stats, "bnrelu": run_bnrelu, "cam_gate": run_cam_gate, "seg_means": run_seg_means,
         "cam_dense": run_cam_dense}


def run_plan(path, feats, n_frames=None):
    """feats (B,T,feat_dim) float -> embeddings (B,embed_dim) float64 by re-evaluating the traced plan on the host.
    n_frames: per-utterance frame counts of a length-masked plan (ws_engine_forward_masked)."""
    meta, blob = load_trace(path)
    B, T, Fd, E = meta["B"], meta["T"], meta["feat_dim"], meta["embed_dim"]
    assert feats.shape == (B, T, Fd)
    mem = Memory(meta, blob)
    arr, off = mem.array(meta["feats_in"], 4)
    arr[off:off + B * T * Fd] = np.asarray(feats, np.float64).reshape(-1)
    assert (n_frames is not None) == bool(meta["lens"]), "n_frames goes with a masked plan"
    if n_frames is not None:
        arr, off = mem.array(meta["lens"], 4)
        arr[off:off + B] = np.asarray(n_frames, np.float64)
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
        elif kind in EXTRA:
            EXTRA[kind](mem, tr)
        else:
            raise NotImplementedError(kind)
    return mem.vec(meta["emb"], B * E).reshape(B, E).copy(), meta
