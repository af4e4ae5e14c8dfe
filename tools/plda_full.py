"""BASELINE.json configs[4]: TwoCovPLDA 1M enroll x 100k test, D=256, all pairs = 1e11 scores, tiled over enroll rows into
a reused fp32 score buffer (400 GB of scores do not fit in HBM).  Scores are consumed on device (running checksum)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wespeaker_b200 import synthetic as syn
from wespeaker_b200.plda import TwoCovPLDA
from oracle import plda_np
NE = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
NT, D, TILE = 100_000, 256, 32768
pm = syn.make_plda(D, seed=3, normalize_length=True)
p = TwoCovPLDA.from_arrays(**pm)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
enroll = torch.randn(NE, D, generator=g, device=dev)
test = torch.randn(NT, D, generator=g, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
e_t, t_t = p.transform_batch(enroll), p.transform_batch(test)
torch.cuda.synchronize(); t1 = time.perf_counter()
out = torch.empty((TILE, NT), dtype=torch.float32, device=dev)
chk = torch.zeros((), dtype=torch.float64, device=dev)
for r in range(0, NE, TILE):
    n = min(TILE, NE - r)
    p.score_matrix(e_t[r:r + n], t_t, 1, out=out[:n])
    chk += out[:n, ::997].sum(dtype=torch.float64)
torch.cuda.synchronize(); t2 = time.perf_counter()
ref = plda_np.llr_matrix(pm, e_t[NE - 64:].cpu().numpy(), t_t[:256].cpu().numpy(), 1)
err = float(np.abs(out[n - 64:n, :256].double().cpu().numpy() - ref).max())
print(f"transform {NE}+{NT} rows: {(t1-t0)*1e3:.1f} ms; {NE*NT:.3e} scores in {(t2-t1):.3f} s -> {NE*NT/(t2-t1):.3e} scores/s; checksum {float(chk):.6e}; last-tile max err vs fp64 oracle {err:.2e}")
