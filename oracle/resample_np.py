"""TEST INFRASTRUCTURE ONLY.  numpy restatement of `torchaudio.functional.resample` (torchaudio 2.x
`functional/functional.py::_get_sinc_resample_kernel` + `_apply_sinc_resample_kernel`, defaults sinc_interp_hann,
lowpass_filter_width = 6, rolloff = 0.99), which is what the reference calls through `torchaudio.transforms.Resample`
(`wespeaker/dataset/processor.py:242-262`, `wespeaker/cli/speaker.py:157-159`).  Third-party arithmetic (torchaudio is a
dependency of the reference, `setup.py:35-36`); pinned against torchaudio itself by `tests/golden/resample.npz`
(`tests/golden/make_golden_resample.py`)."""
import math

import numpy as np


def sinc_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    g = math.gcd(int(orig_freq), int(new_freq))
    of, nf = int(orig_freq) // g, int(new_freq) // g
    base = min(of, nf) * rolloff
    width = math.ceil(lowpass_filter_width * of / base)
    idx = np.arange(-width, width + of, dtype=np.float64)[None, :] / of
    t = np.arange(0, -nf, -1, dtype=np.float64)[:, None] / nf + idx
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / of)
    return k.astype(np.float32), width, of, nf


def resample(wav: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """(B, N) -> (B, ceil(new * N / orig)) float32; accumulation in float64 over the float32 taps."""
    if orig_freq == new_freq:
        return wav.astype(np.float32)
    k, width, of, nf = sinc_kernel(orig_freq, new_freq)
    x = np.atleast_2d(wav).astype(np.float64)
    B, N = x.shape
    pad = np.pad(x, ((0, 0), (width, width + of)))
    nj = (pad.shape[1] - k.shape[1]) // of + 1
    out = np.empty((B, nj, nf), dtype=np.float64)
    kd = k.astype(np.float64)
    for j in range(nj):
        out[:, j, :] = pad[:, j * of:j * of + k.shape[1]] @ kd.T
    n_out = -(-nf * N // of)
    return out.reshape(B, -1)[:, :n_out].astype(np.float32)
