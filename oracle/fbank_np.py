"""Oracle: numpy restatement of Kaldi-compatible log-mel fbank + CMN.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The arithmetic is third-party to the reference: `wespeaker/dataset/processor.py:496-526`
and `wespeaker/cli/speaker.py:90-100` call ``torchaudio.compliance.kaldi.fbank``
(torchaudio 2.11.0 here; reference pins ``torchaudio>=2.0.0``, setup.py:35-36).  This file
restates torchaudio ``kaldi.py`` ``_get_strided :44-83``, ``_get_window :154-217``,
``get_mel_banks :436-512`` and ``fbank :514-646`` for the arguments the reference uses
(SURVEY.md Appendix B), cross-checked against the native statement in
`runtime/core/frontend/fbank.h:33-198`.  Pinned by tests/golden/fbank_*.npz.
"""
from __future__ import annotations

import math

import numpy as np

EPS = np.float32(1.1920928955078125e-07)  # torch.finfo(float32).eps, kaldi.py `_get_epsilon`


def mel_scale(freq):
    return 1127.0 * np.log(1.0 + freq / 700.0)


def window(window_type: str, n: int = 400) -> np.ndarray:
    """kaldi.py `_feature_window_function`: hamming = 0.54-0.46cos(2*pi*j/(n-1)) (non-periodic);
    povey = hann(non-periodic)**0.85 (used by the damo CAM++ checkpoints, cli/speaker.py:343-348)."""
    j = np.arange(n, dtype=np.float64)
    if window_type == "hamming":
        w = 0.54 - 0.46 * np.cos(2.0 * math.pi * j / (n - 1))
    elif window_type == "povey":
        w = (0.5 - 0.5 * np.cos(2.0 * math.pi * j / (n - 1))) ** 0.85
    elif window_type == "hanning":
        w = 0.5 - 0.5 * np.cos(2.0 * math.pi * j / (n - 1))
    elif window_type == "rectangular":
        w = np.ones(n)
    else:
        raise ValueError(window_type)
    return w.astype(np.float32)


def mel_banks(num_bins=80, padded=512, sample_freq=16000.0, low_freq=20.0, high_freq=0.0):
    """kaldi.py `get_mel_banks` (vtln_warp=1): (num_bins, padded/2 + 1) float32, Nyquist column 0."""
    num_fft_bins = padded // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float32)[:, None]
    # torchaudio does this arithmetic in float32 tensors
    left = np.float32(mel_low) + b * np.float32(delta)
    center = np.float32(mel_low) + (b + np.float32(1.0)) * np.float32(delta)
    right = np.float32(mel_low) + (b + np.float32(2.0)) * np.float32(delta)
    mel = (np.float32(1127.0) * np.log(np.float32(1.0) + (np.float32(fft_bin_width) *
           np.arange(num_fft_bins, dtype=np.float32)) / np.float32(700.0))).astype(np.float32)[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = np.maximum(np.float32(0.0), np.minimum(up, down)).astype(np.float32)
    return np.pad(bins, ((0, 0), (0, 1)))


def fbank(wav, num_mel_bins=80, frame_length=25.0, frame_shift=10.0, sample_frequency=16000.0,
          window_type="hamming", preemph=0.97, dtype=np.float32):
    """wav: (N,) samples in int16 range (i.e. after ``* (1 << 15)``).  Returns (m, num_mel_bins).

    dither=0 (extract.py:84-85), remove_dc_offset, snip_edges, round_to_power_of_two, use_power,
    use_log_fbank, use_energy=False — the argument set of processor.py:518-525."""
    wav = np.asarray(wav, dtype=dtype).reshape(-1)
    shift = int(sample_frequency * frame_shift * 0.001)
    size = int(sample_frequency * frame_length * 0.001)
    padded = 1 if size == 0 else 2 ** (size - 1).bit_length()
    n = wav.shape[0]
    if n < size:
        return np.zeros((0, num_mel_bins), dtype=dtype)
    m = 1 + (n - size) // shift
    idx = np.arange(m)[:, None] * shift + np.arange(size)[None, :]
    fr = wav[idx]
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=dtype)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = fr - dtype(preemph) * prev
    fr = fr * window(window_type, size).astype(dtype)[None, :]
    fr = np.pad(fr, ((0, 0), (0, padded - size)))
    spec = np.abs(np.fft.rfft(fr, axis=1)).astype(dtype) ** dtype(2.0)
    mel = mel_banks(num_mel_bins, padded, sample_frequency).astype(dtype)
    e = spec @ mel.T
    return np.log(np.maximum(e, dtype(EPS))).astype(dtype)


def cmn(feats):
    """`wespeaker/dataset/dataset_utils.py:19-26` / `cli/speaker.py:98-99`: subtract mean over T."""
    feats = np.asarray(feats)
    return feats - feats.mean(axis=-2, keepdims=True, dtype=feats.dtype)
