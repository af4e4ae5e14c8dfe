"""Launch the dominant conv-GEMM (ECAPA 3C->1536 1x1 over B*T positions) a few times for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
model = sys.argv[1] if len(sys.argv) > 1 else "ECAPA_TDNN_c1024"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
r = bench.time_dominant_kernel(model, prec, 256, 200, iters=3, tc_version=3)
print(r)
