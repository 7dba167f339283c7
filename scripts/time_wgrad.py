"""Device time of the conv weight gradient, tcgen05 kernel (sdc_conv_wgrad_tc) against the mma.sync kernel (sdc_conv_wgrad), on the
convolution shapes of the dim-128 U-Net at the reference's fine-tuning batch (50) and at 256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safediffcon_b200 import _lib as L, unet as U

lib = L.lib()
SHAPES = [  # kind, H, W, c0, c1, cout
    (1, 16, 128, 128, 0, 128), (1, 16, 128, 128, 128, 128), (1, 8, 64, 256, 128, 256), (1, 4, 32, 512, 256, 512), (1, 2, 16, 1024, 512, 1024),
    (1, 2, 16, 1024, 0, 1024), (0, 16, 128, 128, 0, 384), (0, 2, 16, 1024, 512, 1024)]
for B in [int(a) for a in (sys.argv[1:] or ["50", "256"])]:
    tot = [0.0, 0.0]
    for kind, H, W, c0, c1, cout in SHAPES:
        M = B * H * W
        a0 = torch.randn(M, c0, device="cuda").half()
        a1 = torch.randn(M, c1, device="cuda").half() if c1 else None
        dy = torch.randn(M, cout, device="cuda")
        ks = 3 if kind == 1 else 1
        dw = torch.zeros(cout, c0 + c1, ks, ks, device="cuda")
        nb = int(lib.sdc_conv_wgrad_tc_scratch(1, c0, c1, B, H, W))
        scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")

        def tc():
            assert lib.sdc_conv_wgrad_tc(kind, 1, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, L.ptr(scratch), nb, L.stream_ptr()) == 0

        def old():
            L.check(lib.sdc_conv_wgrad(kind, 1, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, L.stream_ptr()))

        res = []
        for fn in (tc, old):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 5 * 1e3)
        fl = 2.0 * M * cout * (c0 + c1) * ks * ks
        tot[0] += res[0]; tot[1] += res[1]
        print(f"B={B} k{kind} {H}x{W} {c0}+{c1}->{cout}: tcgen05 {res[0]:8.1f} us ({fl/res[0]/1e6:6.1f} TF/s)   mma.sync {res[1]:8.1f} us ({fl/res[1]/1e6:6.1f} TF/s)", flush=True)
    print(f"B={B} sum over these shapes: tcgen05 {tot[0]/1e3:.2f} ms, mma.sync {tot[1]/1e3:.2f} ms")
