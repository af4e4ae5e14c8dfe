import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wespeaker_b200 import lib, synthetic as syn
from wespeaker_b200.models import from_synthetic
m = from_synthetic("ECAPA_TDNN_c1024", 0, precision="bf16").to("cuda:0")
B, N = 256, 32320
base = syn.make_wavs(B, N, seed=1).astype(np.int16)
pins = [torch.from_numpy(np.roll(base, i, axis=0)).pin_memory() for i in range(4)]
for i in range(3): m.extract_from_wav(pins[i % 4])
torch.cuda.synchronize()
L = lib.load(); h = m._engine
realc, reals = L.ws_engine_collect, L.ws_engine_submit_wav_host
log = []
class Wrap:
    def __getattr__(self, k): return getattr(L, k)
    def ws_engine_collect(self, *a):
        t = time.perf_counter(); r = realc(*a); log.append(("collect", round((time.perf_counter() - t) * 1e3, 2))); return r
    def ws_engine_submit_wav_host(self, *a):
        t = time.perf_counter(); r = reals(*a); log.append(("submit", round((time.perf_counter() - t) * 1e3, 2), a[1], hex(a[2]), hex(a[7]))); return r
lib._lib = Wrap()
for rep in range(2):
    log.clear()
    t0 = time.perf_counter(); tl = t0; gaps = []
    for o in m.extract_stream(pins[i % 4] for i in range(10)):
        now = time.perf_counter(); gaps.append(round((now - tl) * 1e3, 2)); tl = now
    print("generator ms/step", (time.perf_counter() - t0) * 100, gaps)
    print(log)
print([p.is_pinned() for p in pins], [hex(p.data_ptr()) for p in pins])
