"""Experiment: can tcgen05.mma read a K-major SW128 operand whose start address is shifted by one 128-B row, and which
base_offset value does it need?  (Needed for conv taps served from one smem-resident activation buffer.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_parity as T
g = torch.Generator().manual_seed(1)
x = torch.randn(1, 1, 512, 128, generator=g); w = torch.randn(128, 128, 1, 1, generator=g) / 11
args = (None, None, None, None, "bf16", 2, 1, 1, (1, 1), (0, 0), (1, 1), 0, 0)
os.environ.pop("WS_TC2_SHIFT_TEST", None)
ref = T.run_conv(x, w, *args)[0, 0]
for bo in (0, 1, 7):
    os.environ["WS_TC2_SHIFT_TEST"] = str(bo)
    out = T.run_conv(x, w, *args)[0, 0]
    # with the shift experiment tile-row i holds global row t0-1+i and the MMA starts at row 1 -> output row i == ref row i,
    # except the last row of each 128-row tile (reads one row past the loaded tile)
    ok = [(out[t] - ref[t]).abs().max().item() for t in range(512) if t % 128 != 127]
    bad = sum(1 for v in ok if v > 1e-2)
    print(f"base_offset={bo}: rows mismatching {bad}/{len(ok)}, max diff {max(ok):.4f}")
os.environ.pop("WS_TC2_SHIFT_TEST", None)
