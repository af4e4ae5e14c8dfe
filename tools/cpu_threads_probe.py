"""Probe: oracle (reference CPU path port) throughput vs torch thread count on this host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fbank_np, models_torch
from wespeaker_b200 import synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "ECAPA_TDNN_c1024"
sd = {k: torch.from_numpy(np.asarray(v)) for k, v in syn.make_state_dict(name, 0).items()}
wavs = syn.make_wavs(16, 32320, seed=0)
t0 = time.perf_counter(); feats = np.stack([fbank_np.cmn(fbank_np.fbank(w)) for w in wavs]); tf = time.perf_counter() - t0
print(f"fbank numpy 16 utts: {tf:.3f}s")
x = torch.from_numpy(feats)
for nt in (8, 16, 32, 64, 128):
    if nt > (os.cpu_count() or 1): break
    torch.set_num_threads(nt)
    models_torch.forward(name, sd, x)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 3.0:
        models_torch.forward(name, sd, x); n += 1
    dt = time.perf_counter() - t0
    print(f"threads {nt}: {16*n/dt:.1f} utt/s (forward only)")
