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
for dbg in ("0", "1", "2"):
    os.environ["SDC_STEM_DBG"] = dbg
    for _ in range(3):
        assert lib.sdc_stem_conv7_tc(L.ptr(x), L.ptr(wp), L.ptr(b), L.ptr(out), B, 3, 16, 128, cout, kp, L.stream_ptr()) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.sdc_stem_conv7_tc(L.ptr(x), L.ptr(wp), L.ptr(b), L.ptr(out), B, 3, 16, 128, cout, kp, L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print(f"dbg={dbg}: {e0.elapsed_time(e1)/10*1e3:.0f} us")
