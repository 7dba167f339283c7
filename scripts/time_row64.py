"""3x3 convolution 128 -> 128 on the 8x64 level (B = 1024): W = 64 row kernel (conv_row64.cu) against the generic implicit GEMM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safediffcon_b200 import _lib as L, unet as U
lib = L.lib()
F16 = U.PREC_F16
for (B, H, W, c0, c1, cout) in ((1024, 8, 64, 128, 0, 128), (1024, 8, 64, 128, 128, 128), (250, 8, 64, 128, 0, 128)):
    x0 = torch.randn(B * H * W, c0, device="cuda").half()
    x1 = torch.randn(B * H * W, c1, device="cuda").half() if c1 else None
    w = torch.randn(cout, c0 + c1, 3, 3, device="cuda") * 0.05
    wp = U.pack_conv_weight(1, w, F16)
    b = torch.randn(cout, device="cuda")
    out = torch.empty(B * H * W, cout, dtype=torch.float16, device="cuda")
    stats = torch.zeros(B, 2, dtype=torch.float64, device="cuda")
    res = {}
    for name, fn in (("row64", lambda: lib.sdc_conv3x3_row(F16, L.ptr(x0), c0, L.ptr(x1), c1, L.ptr(wp), L.ptr(b), None, L.ptr(out), L.ptr(stats), 1, B, H, W, cout, L.stream_ptr())),
                     ("generic", lambda: lib.sdc_conv_gemm(F16, 1, L.ptr(x0), c0, L.ptr(x1), c1, L.ptr(wp), L.ptr(b), None, L.ptr(out), L.ptr(stats), 1, B, H, W, cout, L.stream_ptr()))):
        for _ in range(3):
            rc = fn()
        assert rc == 0, (name, rc)
        torch.cuda.synchronize()
        res[name] = out.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"B={B} {H}x{W} {c0}+{c1}->{cout} {name}: {us:.1f} us, {2.0 * B * H * W * cout * 9 * (c0 + c1) / us / 1e6:.0f} TFLOP/s")
    print("   max |row64 - generic| =", (res["row64"].float() - res["generic"].float()).abs().max().item())
