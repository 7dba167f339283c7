"""HBM-bound apply passes against a plain copy of the same tensor (16x128 level, B = 1024, fp16): is the 1:1 read:write limit ours or the memory's?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safediffcon_b200 import _lib as L, unet as U
lib = L.lib()
F16 = U.PREC_F16
B, HW, C = 1024, 2048, 128
x = torch.randn(B * HW, C, device="cuda").half()
y = torch.empty_like(x)
res = torch.randn(B * HW, C, device="cuda").half()
stats = torch.zeros(B, 2, dtype=torch.float64, device="cuda")
stats[:, 1] = HW * C
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


nb = x.numel() * 2
gn = lambda xin, r, out: L.check(lib.sdc_gn_silu(F16, L.ptr(xin), 1, L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None, 0, L.ptr(r) if r is not None else None, 1,
                                                 L.ptr(out), B, HW, C, L.stream_ptr()))
for name, fn, byts in (("torch copy_ (1:1)", lambda: y.copy_(x), 2 * nb),
                       ("gn_silu in place, no residual (1:1)", lambda: gn(x, None, x), 2 * nb),
                       ("gn_silu out of place, no residual (1:1)", lambda: gn(x, None, y), 2 * nb),
                       ("gn_silu in place + residual (2:1)", lambda: gn(x, res, x), 3 * nb),
                       ("torch add out= (2:1)", lambda: torch.add(x, res, out=y), 3 * nb)):
    us = timed(fn)
    print(f"{name:44s} {us:7.1f} us  {byts / us / 1e3:7.0f} GB/s")
    x.normal_()
