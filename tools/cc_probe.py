"""Why does the 1024->1024 1x1 conv take 174 us inside the model but ~116 us alone?  Vary one factor at a time."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wespeaker_b200 import lib
L = lib.load()
B, T, ci, co = 256, 200, 1024, 1024
dev = torch.device("cuda")
def run(label, ld_mult=1, bn=False, nbuf=2, out_ld_mult=1, iters=10):
    xs = [torch.randn(B, 1, T, ci * ld_mult, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    outs = [torch.empty(B, 1, T, co * out_ld_mult, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    w = (torch.randn(co, ci, device=dev) / 32).to(torch.bfloat16)
    bias = torch.zeros(co, device=dev); sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    descs = []
    for x, o in zip(xs, outs):
        d = lib.ConvDesc()
        d.x, d.B, d.F, d.T, d.Cin, d.x_ld = x.data_ptr(), B, 1, T, ci, ci * ld_mult
        d.w, d.Cout, d.kf, d.kt = w.data_ptr(), co, 1, 1
        d.dil_f = d.dil_t = d.stride_f = d.stride_t = 1
        d.bias, d.act1, d.out, d.out_ld, d.dtype, d.use_tc = bias.data_ptr(), 1, o.data_ptr(), co * out_ld_mult, 1, 3
        if bn: d.scale, d.shift = sc.data_ptr(), sh.data_ptr()
        descs.append(d)
    st = lib.cur_stream_ptr()
    for i in range(3): lib.check(L.ws_conv(C.byref(descs[i % nbuf]), st), "c")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): lib.check(L.ws_conv(C.byref(descs[i % nbuf]), st), "c")
    e1.record(); torch.cuda.synchronize()
    print(f"{label:40s} {e0.elapsed_time(e1) / iters * 1e3:8.1f} us", flush=True)
run("baseline (ld=C, bias+relu, 2 bufs)")
run("with BN scale/shift", bn=True)
run("input ld = 3C", ld_mult=3)
run("output ld = 3C", out_ld_mult=3)
run("4 rotating buffers", nbuf=4)
run("1 buffer (L2-resident W, same x)", nbuf=1)
run("50 iterations (sustained)", iters=50)
