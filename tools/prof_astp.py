"""WS_ASTP_PROF=1 wait-cycle profile of the fused ASTP kernel inside an ECAPA plan run without CUDA graphs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wespeaker_b200 import synthetic as syn
from wespeaker_b200.models import from_synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "ECAPA_TDNN_c1024"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
m = from_synthetic(name, 0, precision="bf16").to("cuda:0")
m.set_option("cuda_graph", 0)
x = torch.from_numpy(syn.make_feats(B, T, 80, seed=1)).cuda()
for _ in range(3):
    m.embed(x)
torch.cuda.synchronize()
