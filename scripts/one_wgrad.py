"""A few launches of the tcgen05 conv weight-gradient kernel (ncu target): 3x3, 8x64 level, 384 -> 256 channels, B = 256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safediffcon_b200 import _lib as L, unet as U
lib = L.lib()
B, H, W, c0, c1, cout = 256, 8, 64, 256, 128, 256
M = B * H * W
a0, a1 = torch.randn(M, c0, device="cuda"), torch.randn(M, c1, device="cuda")
dy = torch.randn(M, cout, device="cuda")
dw = torch.zeros(cout, c0 + c1, 3, 3, device="cuda")
for _ in range(3):
    assert lib.sdc_conv_wgrad_tc(1, 0, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, None, 0, L.stream_ptr()) == 0
torch.cuda.synchronize()
