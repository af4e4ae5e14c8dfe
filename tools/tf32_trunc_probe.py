"""Does tcgen05 kind::tf32 truncate or round fp32 operands?  Compare a conv on x against the same conv on x with the
13 low mantissa bits cleared: bit-identical outputs <=> hardware truncation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
import test_gpu_parity as T
g = torch.Generator().manual_seed(1)
x = torch.randn(2, 1, 200, 256, generator=g); w = torch.randn(256, 256, 1, 1, generator=g) / 16
xt = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
wt = (w.view(torch.int32) & ~0x1FFF).view(torch.float32)
a = T.run_conv(x, w, None, None, None, None, "tf32", 2, 1, 1, (1, 1), (0, 0), (1, 1), 0, 0)
b = T.run_conv(xt, w, None, None, None, None, "tf32", 2, 1, 1, (1, 1), (0, 0), (1, 1), 0, 0)
c = T.run_conv(xt, wt, None, None, None, None, "tf32", 2, 1, 1, (1, 1), (0, 0), (1, 1), 0, 0)
ref = torch.nn.functional.conv2d(xt.double().permute(0, 3, 1, 2), wt.double()).permute(0, 2, 3, 1)
print("x vs trunc(x):      max diff", (a - b).abs().max().item())
print("W vs trunc(W):      max diff", (b - c).abs().max().item())
print("trunc/trunc vs fp64 product of truncated operands: max err", (c.double() - ref).abs().max().item(), "(fp32 accumulation only)")
