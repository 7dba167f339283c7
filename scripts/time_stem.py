import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safediffcon_b200 import _lib as L, unet as U
lib = L.lib()
B, cout, kp = 1024, 128, 320
x = torch.randn(B, 3, 16, 128, device="cuda")
wrep = torch.randn(cout, kp, device="cuda") * 0.1
wp = U.pack_conv_weight(0, wrep.reshape(cout, kp, 1, 1), 1)
b = torch.randn(cout, device="cuda")
out = torch.empty(B * 2048, cout, dtype=torch.float16, device="cuda")
for B_, dbg in ((1024, 0), (50, 0), (1024, 1), (1024, 2), (1024, 4), (1024, 3), (1024, 5), (1024, 6)):
    os.environ['SDC_STEM_DBG'] = str(dbg)
    xb = x[:B_]
    for _ in range(3):
        assert lib.sdc_stem_conv7_tc(L.ptr(xb), L.ptr(wp), L.ptr(b), L.ptr(out), B_, 3, 16, 128, cout, kp, L.stream_ptr()) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.sdc_stem_conv7_tc(L.ptr(xb), L.ptr(wp), L.ptr(b), L.ptr(out), B_, 3, 16, 128, cout, kp, L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"B={B_} dbg={dbg}: {us:.0f} us, {B_ * 2048 * (128 * 2 + 12) / us / 1e3:.0f} GB/s")
