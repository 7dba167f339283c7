"""Timing of the fused conv + GroupNorm kernel (sdc_conv3x3_row_gn) against conv + separate GroupNorm, with debug variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safediffcon_b200 import _lib as L, unet as U

B, H, W, c0, cout = 1024, 16, 128, 128, 128
M = B * H * W
g = torch.Generator().manual_seed(0)
a0 = (torch.randn(M // 8, c0, generator=g) * 0.8).half().cuda().repeat(8, 1)
w = (torch.randn(cout, c0, 3, 3, generator=g) / 34.0).cuda()
bias = torch.randn(cout, generator=g).cuda()
gamma, beta = torch.ones(cout).cuda(), torch.zeros(cout).cuda()
res = torch.randn(M // 8, cout, generator=g).half().cuda().repeat(8, 1)
cw = dict(w=U.pack_conv_weight(1, w, 1), b=bias, cout=cout)
y = torch.empty(M, cout, dtype=torch.float16).cuda()
s = torch.zeros(B, 2, dtype=torch.float64).cuda()
cnt = torch.full((B, 256), -1, dtype=torch.int32).cuda()
lib = L.lib()


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def unfused(r):
    s.zero_()
    U.conv_gemm(1, a0, c0, None, 0, cw["w"], bias, None, y, s, True, B, H, W, cout, 1)
    L.check(lib.sdc_gn_silu(1, L.ptr(y), 1, L.ptr(s), L.ptr(gamma), L.ptr(beta), None, None, 0, L.ptr(r), 1, L.ptr(y), B, H * W, cout, L.stream_ptr()))


def fused(r):
    s.zero_()
    cnt.fill_(-1)
    assert U.conv_row_gn(a0, c0, None, 0, cw, y, s, cnt, (gamma, beta), None, None, 0, r, B, H, W, cout) == 0


s.zero_()
conv_only = timed(lambda: U.conv_gemm(1, a0, c0, None, 0, cw["w"], bias, None, y, s, True, B, H, W, cout, 1))
print(f"conv_row2 alone: {conv_only:.0f} us")
for r, tag in ((None, "no residual"), (res, "residual")):
    print(f"{tag}: unfused {timed(lambda: unfused(r)):.0f} us", flush=True)
    for dbg in ("0", "4", "8", "12", "16", "28", "32", "60"):   # 4: no partner wait, 8: no pass 1, 16: no SiLU, 32: no stores
        os.environ["SDC_ROW_DBG"] = dbg
        print(f"{tag}: fused dbg={dbg} {timed(lambda: fused(r)):.0f} us", flush=True)
    os.environ["SDC_ROW_DBG"] = "0"
