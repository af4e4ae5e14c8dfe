"""Oracle: literal Python restatement of the C++ chunk-and-average extraction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows `runtime/core/speaker/speaker_engine.cc:62-75` (ApplyMean),
`:77-139` (ExtractFeature, vector operations restated one for one on Python lists of frames) and `:141-159`
(ExtractEmbedding).  PINNED: the reference's own speaker_engine.cc / feature_pipeline.cc / fft.cc / fbank.h compile with a
glog shim (`make -C oracle` -> oracle/_ref/libref_engine.so; the full CMake build needs onnxruntime by URL and is not
used); tests/golden/ref_engine.npz holds its chunk composition, ApplyMean output and cosine values, and
tests/test_speaker_engine.py checks this restatement (and oracle/fbank_np.py against the native fbank twin) on them.
"""
from __future__ import annotations

import numpy as np


def num_chunk_frames(per_chunk_samples, sample_rate=16000):
    return 1 + ((per_chunk_samples - sample_rate // 1000 * 25) // (sample_rate // 1000 * 10))


def extract_feature(frames, per_chunk_samples, sample_rate=16000):
    """frames: list of per-frame vectors (the queued fbank frames).  Returns the list of chunks (lists of frames)."""
    queue = list(frames)
    chunks_feat = []
    if per_chunk_samples <= 0:
        return [queue]
    n = num_chunk_frames(per_chunk_samples, sample_rate)
    while len(queue) >= n:                       # feature_pipeline_->Read(num_chunk_frames_)
        chunks_feat.append(queue[:n])
        queue = queue[n:]
    last_frames = len(queue)
    if last_frames > 0:
        chunk_feat = list(queue)
        if not chunks_feat:                      # wav_len < chunk_len
            num_pad = int(n / last_frames)
            for _ in range(1, num_pad):
                chunk_feat = chunk_feat + chunk_feat[:last_frames]
            chunk_feat = chunk_feat + chunk_feat[:n - len(chunk_feat)]
        else:
            chunk_feat = chunk_feat + chunks_feat[0][:n - len(chunk_feat)]
        assert len(chunk_feat) == n
        chunks_feat.append(chunk_feat)
    return chunks_feat


def apply_mean(chunk):
    a = np.asarray(chunk, dtype=np.float32)
    return a - a.mean(axis=0, keepdims=True, dtype=np.float32)


def extract_embedding(frames, per_chunk_samples, embed_one):
    """embed_one(feats (T,F) float32) -> (E,) embedding of ONE chunk (batch 1, like SpeakerModel::ExtractEmbedding)."""
    chunks = extract_feature(frames, per_chunk_samples)
    acc = None
    for c in chunks:
        e = np.asarray(embed_one(apply_mean(c)), dtype=np.float32)
        acc = e.copy() if acc is None else acc + e
    return acc / len(chunks)
