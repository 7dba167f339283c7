"""Kernel-level breakdown (CUPTI via torch.profiler, no replay) of one backward-data pass: total device time per kernel name."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
g = torch.randn(B, 3, 16, 128, device="cuda")
for _ in range(2):
    net.vjp(x, 500, g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net.vjp(x, 500, g)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"B={B}: device time {tot / 1e3:.2f} ms over {sum(e.count for e in rows)} kernels")
for e in rows[:28]:
    print(f"  {e.key[:70]:70s} n={e.count:3d} {e.device_time_total / 1e3:8.3f} ms  max {max(1, e.device_time_total) / e.count / 1e3:7.3f} ms avg")
if len(sys.argv) > 2:   # per-launch durations of the kernels whose name contains argv[2]
    for e in prof.events():
        if sys.argv[2] in e.name:
            print(f"    {e.name[:60]:60s} {e.device_time / 1e3:8.3f} ms")
