"""`extract(config, **kwargs)` — drop-in for `wespeaker/bin/extract.py:33-139` (seam B2).

Reads the reference YAML keys (``model``, ``model_args``, ``dataset_args``), loads ``model_path`` with
``load_checkpoint``, extracts embeddings and writes Kaldi ``ark,scp`` (scp path = ark path with ``.scp``).
The reference's DataLoader / shard / augmentation pipeline is out of scope (SURVEY.md §2 #3b); the data list
is the reference ``raw`` format (JSON lines ``{key, wav, spk}``, processor.py:147-166) or ``feat`` (a Kaldi
scp of (T,80) features is not needed offline, so ``feat`` takes an .npz of name -> (T,80) arrays).  Unlike
the reference (CPU fbank in worker processes, extract.py:99-103), waveforms go to the GPU and fbank + CMN +
forward run fused on device; equal-length utterances are batched.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import yaml

from .kaldi_io import VectorWriter
from .models import get_speaker_model, load_checkpoint
from .speaker import read_wav


def parse_config_or_kwargs(config_file, **kwargs):
    """`wespeaker/utils/utils.py:37-51`: yaml dict overridden by kwargs."""
    with open(config_file) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    return dict(cfg, **kwargs)


def extract(config="conf/config.yaml", **kwargs):
    configs = parse_config_or_kwargs(config, **kwargs)
    model_path = configs["model_path"]
    embed_ark = os.path.abspath(configs["embed_ark"])
    batch_size = int(configs.get("batch_size", 1))
    test_conf = dict(configs.get("dataset_args", {}))
    if test_conf.get("frontend", "fbank") != "fbank":
        raise NotImplementedError("only the fbank frontend is on the B200 hot path")
    model = get_speaker_model(configs["model"])(precision=configs.get("precision"), **configs["model_args"])
    load_checkpoint(model, model_path)
    device = torch.device("cuda")
    model.to(device).eval()
    cmvn = test_conf.get("cmvn", True)
    if not cmvn:
        raise NotImplementedError("the fused wav path always applies CMN (reference default cmvn: True)")
    os.makedirs(os.path.dirname(embed_ark), exist_ok=True)
    embed_scp = embed_ark[:-3] + "scp"
    data_type = configs.get("data_type", "raw")
    n = 0
    with torch.no_grad(), VectorWriter(embed_ark, embed_scp) as writer:
        if data_type == "raw":
            pending = {}  # nsamples -> [(key, pcm)]

            def flush(group):
                nonlocal n
                wavs = torch.stack([p for _, p in group]).to(device)
                embs = model.extract_from_wav(wavs).cpu().numpy()
                for (k, _), e in zip(group, embs):
                    writer(k, e)
                    n += 1

            with open(configs["data_list"]) as f:
                for line in f:
                    if not line.strip():
                        continue
                    obj = json.loads(line)
                    pcm, sr = read_wav(obj["wav"], normalize=False)
                    if sr != test_conf.get("resample_rate", 16000):
                        raise NotImplementedError("resampling is out of scope; provide 16 kHz audio")
                    g = pending.setdefault(pcm.shape[1], [])
                    g.append((obj["key"], pcm[0]))
                    if len(g) >= batch_size:
                        flush(g)
                        pending[pcm.shape[1]] = []
            for g in pending.values():
                if g:
                    flush(g)
        elif data_type == "feat":
            z = np.load(configs["data_list"])
            for k in z.files:
                feats = torch.from_numpy(z[k].astype(np.float32))[None].to(device)
                feats = feats - feats.mean(dim=1, keepdim=True)
                out = model(feats)
                out = out[-1] if isinstance(out, tuple) else out
                writer(k, out[0].cpu().numpy())
                n += 1
        else:
            raise NotImplementedError(f"data_type {data_type!r}: shard reading is out of scope")
    return n
