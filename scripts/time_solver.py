"""Quick device timing of the rollout kernel (CUDA events) at several N; strict and fast modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s
from safediffcon_b200.synthetic import burgers_instances

for n in (8192, 18944, 100000):
    u0, f = burgers_instances(n, seed=1)
    du0, df = torch.from_numpy(u0).cuda(), torch.from_numpy(f).cuda()
    for strict in (True, False):
        for _ in range(2):
            s.burgers_numeric_solve_free(du0, df, 0.01, 1.0, strict=strict)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = s.burgers_numeric_solve_free(du0, df, 0.01, 1.0, strict=strict)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"N={n} strict={strict}: {ms:.2f} ms  {n/ms*1e3:.0f} rollouts/s  {n*17.92e6/ms/1e9:.1f} TFLOP/s(alg)", flush=True)
