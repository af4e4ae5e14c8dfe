"""Oracle: functional torch-CPU restatement of the reference model forwards.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function takes the reference
``state_dict`` (key names as in SURVEY.md Appendix C) and evaluates the eval-mode
forward with ``torch.nn.functional`` ops in the dtype of the inputs (fp32 or fp64).
Pinned against the real reference modules by tests/golden/*.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default, used by every BN in the reference


def _t(sd, key, like):
    v = sd[key]
    if not torch.is_tensor(v):
        v = torch.as_tensor(v)
    return v.to(dtype=like.dtype)


def _bn(x, sd, p, affine=True):
    """Eval-mode BatchNorm{1d,2d}: per-channel affine on dim 1."""
    w = _t(sd, p + ".weight", x) if affine else None
    b = _t(sd, p + ".bias", x) if affine else None
    return F.batch_norm(x, _t(sd, p + ".running_mean", x), _t(sd, p + ".running_var", x),
                        w, b, False, 0.0, BN_EPS)


# --------------------------------------------------------------------------- pooling
def tstp(x):
    """`wespeaker/models/pooling_layers.py:78-85`: cat[mean_T, sqrt(var_T(unbiased)+1e-7)]."""
    mean = x.mean(dim=-1).flatten(start_dim=1)
    std = torch.sqrt(torch.var(x, dim=-1) + 1e-7).flatten(start_dim=1)
    return torch.cat((mean, std), 1)


def astp(x, sd, p="pool", global_context_att=False):
    """`wespeaker/models/pooling_layers.py:119-144` (attentive statistics pooling)."""
    if global_context_att:
        cm = torch.mean(x, dim=-1, keepdim=True).expand_as(x)
        cs = torch.sqrt(torch.var(x, dim=-1, keepdim=True) + 1e-7).expand_as(x)
        x_in = torch.cat((x, cm, cs), dim=1)
    else:
        x_in = x
    a = torch.tanh(F.conv1d(x_in, _t(sd, p + ".linear1.weight", x), _t(sd, p + ".linear1.bias", x)))
    a = torch.softmax(F.conv1d(a, _t(sd, p + ".linear2.weight", x), _t(sd, p + ".linear2.bias", x)), dim=2)
    mean = torch.sum(a * x, dim=2)
    var = torch.sum(a * (x ** 2), dim=2) - mean ** 2
    std = torch.sqrt(var.clamp(min=1e-7))
    return torch.cat([mean, std], dim=1)


# --------------------------------------------------------------------------- ECAPA
def _conv1d_relu_bn(x, sd, p, padding=0, dilation=1):
    """`wespeaker/models/ecapa_tdnn.py:85-106`: bn(relu(conv(x))) — conv -> relu -> bn."""
    y = F.conv1d(x, _t(sd, p + ".conv.weight", x), _t(sd, p + ".conv.bias", x),
                 padding=padding, dilation=dilation)
    return _bn(F.relu(y), sd, p + ".bn")


def _res2(x, sd, p, dilation, scale=8):
    """`wespeaker/models/ecapa_tdnn.py:29-78` (Res2Conv1dReluBn)."""
    width = x.shape[1] // scale
    spx = torch.split(x, width, 1)
    out = []
    sp = spx[0]
    for i in range(scale - 1):
        if i >= 1:
            sp = sp + spx[i]
        sp = F.conv1d(sp, _t(sd, f"{p}.convs.{i}.weight", x), _t(sd, f"{p}.convs.{i}.bias", x),
                      padding=dilation, dilation=dilation)
        sp = _bn(F.relu(sp), sd, f"{p}.bns.{i}")
        out.append(sp)
    out.append(spx[scale - 1])
    return torch.cat(out, dim=1)


def _se(x, sd, p):
    """`wespeaker/models/ecapa_tdnn.py:113-126` (SE_Connect)."""
    s = x.mean(dim=2)
    s = F.relu(F.linear(s, _t(sd, p + ".linear1.weight", x), _t(sd, p + ".linear1.bias", x)))
    s = torch.sigmoid(F.linear(s, _t(sd, p + ".linear2.weight", x), _t(sd, p + ".linear2.bias", x)))
    return x * s.unsqueeze(2)


def _se_res2block(x, sd, p, dilation):
    """`wespeaker/models/ecapa_tdnn.py:133-157`."""
    y = _conv1d_relu_bn(x, sd, p + ".0")
    y = _res2(y, sd, p + ".1", dilation)
    y = _conv1d_relu_bn(y, sd, p + ".2")
    y = _se(y, sd, p + ".3")
    return x + y


def ecapa_forward(sd, feats, global_context_att=False, emb_bn=False, return_taps=False):
    """`wespeaker/models/ecapa_tdnn.py:208-234`.  feats (B,T,F) -> (out4 (B,C,T), emb (B,E))."""
    x = feats.permute(0, 2, 1)
    out1 = _conv1d_relu_bn(x, sd, "layer1", padding=2)
    out2 = _se_res2block(out1, sd, "layer2.se_res2block", 2)
    out3 = _se_res2block(out2, sd, "layer3.se_res2block", 3)
    out4 = _se_res2block(out3, sd, "layer4.se_res2block", 4)
    cat = torch.cat([out2, out3, out4], dim=1)
    out = F.conv1d(cat, _t(sd, "conv.weight", x), _t(sd, "conv.bias", x))
    out = F.relu(out)
    stats = astp(out, sd, "pool", global_context_att)
    emb = F.linear(_bn(stats, sd, "bn"), _t(sd, "linear.weight", x), _t(sd, "linear.bias", x))
    if emb_bn:
        emb = _bn(emb, sd, "bn2")
    if return_taps:
        return dict(out1=out1, out2=out2, out3=out3, out4=out4, frame=out, stats=stats, emb=emb)
    return out4, emb


# --------------------------------------------------------------------------- ResNet
def _basic_block(x, sd, p, stride):
    """`wespeaker/models/resnet.py:35-69` / `campplus.py:245-279` (stride may be (sF, sT))."""
    out = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight", x), None, stride=stride, padding=1),
                     sd, p + ".bn1"))
    out = _bn(F.conv2d(out, _t(sd, p + ".conv2.weight", x), None, padding=1), sd, p + ".bn2")
    if (p + ".shortcut.0.weight") in sd:
        sc = _bn(F.conv2d(x, _t(sd, p + ".shortcut.0.weight", x), None, stride=stride),
                 sd, p + ".shortcut.1")
    else:
        sc = x
    return F.relu(out + sc)


def _bottleneck(x, sd, p, stride):
    """`wespeaker/models/resnet.py:72-107`: 1x1 -> 3x3 (stride) -> 1x1 (x4), BN after each, shortcut, ReLU."""
    out = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight", x)), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, _t(sd, p + ".conv2.weight", x), None, stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(F.conv2d(out, _t(sd, p + ".conv3.weight", x)), sd, p + ".bn3")
    if (p + ".shortcut.0.weight") in sd:
        sc = _bn(F.conv2d(x, _t(sd, p + ".shortcut.0.weight", x), None, stride=stride), sd, p + ".shortcut.1")
    else:
        sc = x
    return F.relu(out + sc)


def resnet_forward(sd, feats, num_blocks=(3, 4, 6, 3), two_emb_layer=False, return_taps=False):
    """`wespeaker/models/resnet.py:171-204`.  feats (B,T,F) -> (tensor(0.), embed_a).  The block type (BasicBlock or
    Bottleneck) follows from the checkpoint (`conv3` exists only in Bottlenecks)."""
    x = feats.permute(0, 2, 1).unsqueeze(1)
    out = F.relu(_bn(F.conv2d(x, _t(sd, "conv1.weight", x), None, padding=1), sd, "bn1"))
    taps = dict(stem=out)
    block = _bottleneck if "layer1.0.conv3.weight" in sd else _basic_block
    for li, (nb, stride) in enumerate(zip(num_blocks, (1, 2, 2, 2)), 1):
        for bi in range(nb):
            out = block(out, sd, f"layer{li}.{bi}", stride if bi == 0 else 1)
        taps[f"layer{li}"] = out
    stats = tstp(out)
    emb = F.linear(stats, _t(sd, "seg_1.weight", x), _t(sd, "seg_1.bias", x))
    taps.update(stats=stats, emb=emb)
    if two_emb_layer:
        o = _bn(F.relu(emb), sd, "seg_bn_1", affine=False)
        emb_b = F.linear(o, _t(sd, "seg_2.weight", x), _t(sd, "seg_2.bias", x))
        return emb, emb_b
    if return_taps:
        return taps
    return torch.tensor(0.0), emb


# --------------------------------------------------------------------------- CAM++
def _fcm(x, sd, p="head"):
    """`wespeaker/models/campplus.py:282-330` (FCM: 2-D conv head, frequency stride only)."""
    x = x.unsqueeze(1)
    out = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight", x), None, padding=1), sd, p + ".bn1"))
    for li in (1, 2):
        for bi in range(2):
            out = _basic_block(out, sd, f"{p}.layer{li}.{bi}", (2, 1) if bi == 0 else 1)
    out = F.relu(_bn(F.conv2d(out, _t(sd, p + ".conv2.weight", x), None, stride=(2, 1), padding=1),
                     sd, p + ".bn2"))
    b, c, f, t = out.shape
    return out.reshape(b, c * f, t)


def _seg_pooling(x, seg_len=100):
    """`wespeaker/models/campplus.py:117-135` (avg_pool1d, ceil_mode=True, expand back)."""
    seg = F.avg_pool1d(x, kernel_size=seg_len, stride=seg_len, ceil_mode=True)
    shape = seg.shape
    seg = seg.unsqueeze(-1).expand(shape[0], shape[1], shape[2], seg_len).reshape(shape[0], shape[1], -1)
    return seg[..., :x.shape[-1]]


def _cam_layer(x, sd, p, dilation):
    """`wespeaker/models/campplus.py:86-115` (CAMLayer)."""
    y = F.conv1d(x, _t(sd, p + ".linear_local.weight", x), None, padding=dilation, dilation=dilation)
    context = x.mean(-1, keepdim=True) + _seg_pooling(x)
    context = F.relu(F.conv1d(context, _t(sd, p + ".linear1.weight", x), _t(sd, p + ".linear1.bias", x)))
    m = torch.sigmoid(F.conv1d(context, _t(sd, p + ".linear2.weight", x), _t(sd, p + ".linear2.bias", x)))
    return y * m


def _cam_dense_layer(x, sd, p, dilation):
    """`wespeaker/models/campplus.py:138-170` (BN-ReLU-Conv1x1, BN-ReLU, CAM)."""
    h = F.conv1d(F.relu(_bn(x, sd, p + ".nonlinear1.batchnorm")), _t(sd, p + ".linear1.weight", x))
    h = F.relu(_bn(h, sd, p + ".nonlinear2.batchnorm"))
    return _cam_layer(h, sd, p + ".cam_layer", dilation)


def campplus_forward(sd, feats, return_taps=False):
    """`wespeaker/models/campplus.py:345-413`.  feats (B,T,F) -> emb (B,E) (bare tensor)."""
    x = feats.permute(0, 2, 1)
    x = _fcm(x, sd, "head")
    taps = dict(head=x)
    x = F.conv1d(x, _t(sd, "xvector.tdnn.linear.weight", x), None, stride=2, padding=2)
    x = F.relu(_bn(x, sd, "xvector.tdnn.nonlinear.batchnorm"))
    taps["tdnn"] = x
    for b, (nl, dil) in enumerate(zip((12, 24, 16), (1, 2, 2)), 1):
        for j in range(1, nl + 1):
            x = torch.cat([x, _cam_dense_layer(x, sd, f"xvector.block{b}.tdnnd{j}", dil)], dim=1)
        taps[f"block{b}"] = x
        x = F.conv1d(F.relu(_bn(x, sd, f"xvector.transit{b}.nonlinear.batchnorm")),
                     _t(sd, f"xvector.transit{b}.linear.weight", x))
        taps[f"transit{b}"] = x
    x = F.relu(_bn(x, sd, "xvector.out_nonlinear.batchnorm"))
    stats = tstp(x)
    emb = F.conv1d(stats.unsqueeze(-1), _t(sd, "xvector.dense.linear.weight", x)).squeeze(-1)
    emb = _bn(emb, sd, "xvector.dense.nonlinear.batchnorm", affine=False)
    taps.update(stats=stats, emb=emb)
    return taps if return_taps else emb


# --------------------------------------------------------------------------- XVEC
def xvec_forward(sd, feats):
    """`wespeaker/models/tdnn.py:23-117`: TdnnLayer = BN(ReLU(Conv1d(x))) without padding (T shrinks by 4 + 4 + 6),
    BN affine=False; TSTP; seg_1 -> ReLU -> seg_bn_1 -> seg_2.  Returns (embed_a, embed_b)."""
    x = feats.permute(0, 2, 1)
    for i, dil in enumerate((1, 2, 3, 1, 1), 1):
        p = f"frame_{i}"
        x = F.conv1d(x, _t(sd, p + ".conv_1d.weight", x), _t(sd, p + ".conv_1d.bias", x), dilation=dil)
        x = _bn(F.relu(x), sd, p + ".bn", affine=False)
    stats = tstp(x)
    emb_a = F.linear(stats, _t(sd, "seg_1.weight", x), _t(sd, "seg_1.bias", x))
    o = _bn(F.relu(emb_a), sd, "seg_bn_1", affine=False)
    return emb_a, F.linear(o, _t(sd, "seg_2.weight", x), _t(sd, "seg_2.bias", x))


# --------------------------------------------------------------------------- Res2Net / ERes2Net
def _relu20(x):
    """`wespeaker/models/eres2net.py:43-52`: the families' "ReLU" is Hardtanh(0, 20)."""
    return torch.clamp(x, 0.0, 20.0)


def _aff(x, y, sd, p):
    """`wespeaker/models/eres2net.py:75-102` (AFF): att = 1 + tanh(BN(conv(SiLU(BN(conv([x | y])))))),
    out = x * att + y * (2 - att)."""
    xa = torch.cat((x, y), dim=1)
    h = F.conv2d(xa, _t(sd, p + ".local_att.0.weight", x), _t(sd, p + ".local_att.0.bias", x))
    h = F.silu(_bn(h, sd, p + ".local_att.1"))
    a = F.conv2d(h, _t(sd, p + ".local_att.3.weight", x), _t(sd, p + ".local_att.3.bias", x))
    att = 1.0 + torch.tanh(_bn(a, sd, p + ".local_att.4"))
    return x * att + y * (2.0 - att)


def _res2net_block(x, sd, p, stride, width, scale, kind):
    """kind "res2net": `res2net.py:61-90` (scale - 1 chain convs, the last split passes through);
    "eres2net": `eres2net.py:141-163` (scale chain convs, sp + spx[i] before conv i >= 1);
    "eres2net_aff": `eres2net.py:203-224` (conv2_1 first, then AFF(sp, spx[i]) before conv i)."""
    out = _relu20(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight", x), None, stride=stride), sd, p + ".bn1"))
    spx = torch.split(out, width, 1)
    outs = []
    if kind == "eres2net_aff":
        sp = _relu20(_bn(F.conv2d(spx[0], _t(sd, p + ".conv2_1.weight", x), None, padding=1), sd, p + ".bn2_1"))
        outs.append(sp)
        for i in range(1, scale):
            sp = _aff(sp, spx[i], sd, f"{p}.fuse_models.{i - 1}")
            sp = _relu20(_bn(F.conv2d(sp, _t(sd, f"{p}.convs.{i - 1}.weight", x), None, padding=1), sd, f"{p}.bns.{i - 1}"))
            outs.append(sp)
    else:
        nums = scale if kind == "eres2net" else scale - 1
        sp = spx[0]
        for i in range(nums):
            if i >= 1:
                sp = sp + spx[i]
            sp = _relu20(_bn(F.conv2d(sp, _t(sd, f"{p}.convs.{i}.weight", x), None, padding=1), sd, f"{p}.bns.{i}"))
            outs.append(sp)
        if kind == "res2net":
            outs.append(spx[nums])
    out = _bn(F.conv2d(torch.cat(outs, 1), _t(sd, p + ".conv3.weight", x)), sd, p + ".bn3")
    if (p + ".shortcut.0.weight") in sd:
        sc = _bn(F.conv2d(x, _t(sd, p + ".shortcut.0.weight", x), None, stride=stride), sd, p + ".shortcut.1")
    else:
        sc = x
    return _relu20(out + sc)


def res2net_forward(sd, feats, m_channels=32, num_blocks=(3, 4, 6, 3), base_width=32, scale=2, expansion=2, fuse=False,
                    two_emb_layer=False, return_taps=False):
    """`wespeaker/models/res2net.py:153-199` (fuse=False) and `wespeaker/models/eres2net.py:354-391` (fuse=True: AFF blocks
    in layers 3-4, stride-2 3x3 downsampling of the running fusion and an AFF with the next stage's output)."""
    import math
    x = feats.permute(0, 2, 1).unsqueeze(1)
    out = F.relu(_bn(F.conv2d(x, _t(sd, "conv1.weight", x), None, padding=1), sd, "bn1"))   # plain ReLU here
    taps = dict(stem=out)
    stage = []
    for li, (nb, mult, stride) in enumerate(zip(num_blocks, (1, 2, 4, 8), (1, 2, 2, 2)), 1):
        width = int(math.floor(m_channels * mult * (base_width / 64.0)))
        kind = "res2net" if not fuse else ("eres2net_aff" if li >= 3 else "eres2net")
        for bi in range(nb):
            out = _res2net_block(out, sd, f"layer{li}.{bi}", stride if bi == 0 else 1, width, scale, kind)
        stage.append(out)
        taps[f"layer{li}"] = out
    if fuse:
        f = stage[0]
        for li, name in ((1, "fuse_mode12"), (2, "fuse_mode123"), (3, "fuse_mode1234")):
            d = F.conv2d(f, _t(sd, f"layer{li}_downsample.weight", x), None, stride=2, padding=1)
            f = _aff(stage[li], d, sd, name)
            taps[name] = f
        out = f
    stats = tstp(out)
    emb = F.linear(stats, _t(sd, "seg_1.weight", x), _t(sd, "seg_1.bias", x))
    taps.update(stats=stats, emb=emb)
    if two_emb_layer:
        o = _bn(F.relu(emb), sd, "seg_bn_1", affine=False)
        return emb, F.linear(o, _t(sd, "seg_2.weight", x), _t(sd, "seg_2.bias", x))
    if return_taps:
        return taps
    return torch.tensor(0.0), emb


# --------------------------------------------------------------------------- dispatch
def forward(model_name: str, sd, feats, **kw):
    """Embedding (B,E) for a reference model name; same call convention as the reference
    callers' ``outputs[-1] if isinstance(outputs, tuple) else outputs`` (extract.py:133-134)."""
    from wespeaker_b200.synthetic import ECAPA_NAMES, RES2NET_NAMES, RESNET_NAMES, XVEC_NAMES
    feats = torch.as_tensor(feats)
    with torch.no_grad():
        if model_name in ECAPA_NAMES:
            return ecapa_forward(sd, feats, ECAPA_NAMES[model_name]["global_context_att"], **kw)[-1]
        if model_name in RESNET_NAMES:
            return resnet_forward(sd, feats, RESNET_NAMES[model_name], **kw)[-1]
        if model_name == "CAMPPlus":
            return campplus_forward(sd, feats, **kw)
        if model_name in XVEC_NAMES:
            return xvec_forward(sd, feats)[-1]
        if model_name in RES2NET_NAMES:
            return res2net_forward(sd, feats, **RES2NET_NAMES[model_name], **kw)[-1]
    raise ValueError(model_name)


def apply_cmvn(feats, norm_mean=True, norm_var=False):
    """`wespeaker/dataset/dataset_utils.py:19-26`."""
    if norm_mean:
        feats = feats - torch.mean(feats, dim=1, keepdim=True)
    if norm_var:
        feats = feats / torch.sqrt(torch.var(feats, dim=1, keepdim=True) + 1e-7)
    return feats
