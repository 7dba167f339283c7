"""Timing experiment: per-item globaltimer trace of the fused conv + GroupNorm kernel (cluster 0): where does a round go?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
lib.sdc_debug_row_trace.argtypes = [ctypes.c_void_p]
for tag, r in (("no residual", None), ("residual", res)):
    for dbg in (128, 128 + 4):
        os.environ["SDC_ROW_DBG"] = str(dbg)
        for _ in range(2):
            cnt.fill_(-1)
            assert U.conv_row_gn(a0, c0, None, 0, cw, y, s, cnt, (gamma, beta), None, None, 0, r, B, H, W, cout) == 0
        torch.cuda.synchronize()
        buf = np.zeros(4096, dtype=np.int64)
        lib.sdc_debug_row_trace(buf.ctypes.data)
        mma = buf[:2 * 57].reshape(57, 2)
        epi = buf[2048:2048 + 5 * 57].reshape(57, 5)
        t0 = mma[0, 0]
        print(f"--- {tag} dbg={dbg}: ns relative to the first MMA start; MMA (start, end) | epilogue (acc ready, pass1, partners, coef, pass2)")
        for it in list(range(0, 8)) + list(range(30, 34)):
            print(it, (mma[it] - t0).tolist(), (epi[it] - t0).tolist(), "durations", np.diff(epi[it]).tolist())
        d = np.diff(epi[5:55], axis=1).mean(axis=0)
        print("mean phase durations (pass1, wait, coef, pass2):", d.round(0).tolist(), " round:", float(np.diff(epi[5:55, 0]).mean()),
              " mma busy:", float((mma[5:55, 1] - mma[5:55, 0]).mean()))
