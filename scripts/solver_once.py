"""One rollout + scoring launch of N trajectories between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s
from safediffcon_b200.synthetic import burgers_instances

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
u0, f = burgers_instances(n, seed=1)
du0, df = torch.from_numpy(u0).cuda(), torch.from_numpy(f).cuda()
tgt = torch.zeros(n, 128, device="cuda")
out = s.burgers_numeric_solve_free(du0, df, 0.01, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = s.burgers_numeric_solve_free(du0, df, 0.01, 1.0)
sc = s.burgers_score(out, tgt, 0.8)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(out.abs().max()))
