"""BASELINE.json configs[3]: CAM++ bf16, variable-length 1-10 s utterances bucketed by exact length (the reference has no
masking).  Reports utterances/s and seconds of audio per second on ONE GPU (rank-local shard)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wespeaker_b200 import synthetic as syn
from wespeaker_b200.models import from_synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(2)
durs = rng.integers(1, 11, size=N)
base = torch.from_numpy(syn.make_wavs(1, 160000, seed=9)[0].astype(np.int16)).cuda()
wavs = [base[: int(d) * 16000].clone() for d in durs]
m = from_synthetic("CAMPPlus", 0, precision="bf16").to("cuda:0")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    emb = m.extract_from_wav_list(wavs, max_batch=64, device="cuda:0")
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {N} utts (1-10 s, mean {durs.mean():.2f} s) in {dt*1e3:.1f} ms -> {N/dt:.0f} utt/s, {durs.sum()/dt:.0f} x real time; finite={bool(torch.isfinite(emb).all())}")
