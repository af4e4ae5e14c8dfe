"""Phase timing of the 2-CTA conv-GEMM kernel (needs a build with -DWS_TC3_PROFILE: counters from CTA 0 only).
epilogue thread 0: [0] wait staging free  [1] param stage+sync  [2] wait accumulator  [3] process chunks  [4] sync+store issue
MMA thread:        [8] wait tempty  [9] mainloop issue (incl. waiting for operands)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wespeaker_b200 import lib
L = lib.load()
B, T = 256, 200
dev = torch.device("cuda")
def run(label, ci, co, kt=1, bn=False, iters=5):
    x = torch.randn(B, 1, T, ci, device=dev).to(torch.bfloat16)
    o = torch.empty(B, 1, T, co, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(co, ci * kt, device=dev) / 32).to(torch.bfloat16)
    bias = torch.zeros(co, device=dev); sc = torch.ones(co, device=dev); sh = torch.zeros(co, device=dev)
    d = lib.ConvDesc()
    d.x, d.B, d.F, d.T, d.Cin, d.x_ld = x.data_ptr(), B, 1, T, ci, ci
    d.w, d.Cout, d.kf, d.kt = w.data_ptr(), co, 1, kt
    d.dil_f = d.dil_t = d.stride_f = d.stride_t = 1
    d.pad_t = kt // 2
    d.bias, d.act1, d.out, d.out_ld, d.dtype, d.use_tc = bias.data_ptr(), 1, o.data_ptr(), co, 1, 3
    if bn: d.scale, d.shift = sc.data_ptr(), sh.data_ptr()
    st = lib.cur_stream_ptr()
    for i in range(3): lib.check(L.ws_conv(C.byref(d), st), "c")
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    L.ws_tc3_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): lib.check(L.ws_conv(C.byref(d), st), "c")
    e1.record(); torch.cuda.synchronize()
    L.ws_tc3_prof_read(buf, 1)
    v = list(buf)
    nt = max(v[5], 1); nm = max(v[10], 1)
    print(f"{label:28s} {e0.elapsed_time(e1) / iters * 1e3:7.1f} us/launch  tiles/CTA/launch {nt / iters:.1f} | per tile (cycles): "
          f"stg_wait {v[0] / nt:.0f} par {v[1] / nt:.0f} acc_wait {v[2] / nt:.0f} process {v[3] / nt:.0f} store {v[4] / nt:.0f} | "
          f"mma: tempty_wait {v[8] / nm:.0f} mainloop {v[9] / nm:.0f}", flush=True)
run("1024->1024 k1", 1024, 1024)
run("1024->1024 k1 +BN", 1024, 1024, bn=True)
run("3072->1536 k1", 3072, 1536)
run("512->512 k1 +BN", 512, 512, bn=True)
