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
t0 = time.perf_counter()
for i in range(10): m.extract_from_wav(pins[i % 4])
print("sync path ms/step", (time.perf_counter() - t0) * 100)
L = lib.load(); h = m._engine
outs = [torch.empty((B, 192)).pin_memory() for _ in range(2)]
for rep in range(2):
    ts = []
    t0 = time.perf_counter()
    for i in range(10):
        s = i & 1
        a = time.perf_counter()
        if i >= 2: lib.check(L.ws_engine_collect(h, s), "collect")
        b = time.perf_counter()
        lib.check(L.ws_engine_submit_wav_host(h, s, pins[i % 4].data_ptr(), 1, N, B, b"hamming", outs[s].data_ptr()), "submit")
        c = time.perf_counter()
        ts.append((round((b - a) * 1e3, 2), round((c - b) * 1e3, 2)))
    lib.check(L.ws_engine_collect(h, 0), "collect"); lib.check(L.ws_engine_collect(h, 1), "collect")
    print("stream path ms/step", (time.perf_counter() - t0) * 100, ts)
t0 = time.perf_counter()
n = 0
for o in m.extract_stream(pins[i % 4] for i in range(10)): n += 1
print("generator ms/step", (time.perf_counter() - t0) * 100)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for o in m.extract_stream(pins[i % 4] for i in range(10)): n += 1
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
