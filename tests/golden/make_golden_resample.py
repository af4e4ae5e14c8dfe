"""Golden vectors for resampling: outputs of torchaudio.transforms.Resample itself (the call the reference makes in
`wespeaker/dataset/processor.py:258-259` and `wespeaker/cli/speaker.py:158-159`) on seeded int16-range waveforms.
Run in the build container (torchaudio installed): python tests/golden/make_golden_resample.py"""
import os

import numpy as np
import torch
import torchaudio

out = {}
rng = np.random.default_rng(7)
for tag, (orig, new, n) in {"8k_16k": (8000, 16000, 4001), "44k1_16k": (44100, 16000, 9000), "48k_16k": (48000, 16000, 7777),
                            "22k05_16k": (22050, 16000, 5000), "16k_8k": (16000, 8000, 3333)}.items():
    t = np.arange(n) / orig
    x = (6000 * np.sin(2 * np.pi * 220 * t) + 3000 * np.sin(2 * np.pi * 1800 * t + 0.3) + 800 * rng.standard_normal(n))
    x = np.round(np.stack([x, x[::-1]])).astype(np.float32)
    y = torchaudio.transforms.Resample(orig_freq=orig, new_freq=new)(torch.from_numpy(x)).numpy()
    out[f"{tag}_x"] = x
    out[f"{tag}_y"] = y.astype(np.float32)
    out[f"{tag}_rates"] = np.array([orig, new])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resample.npz"), **out)
print({k: v.shape for k, v in out.items()})
