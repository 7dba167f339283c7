"""Per-launch CUDA-event profile of one denoiser evaluation through the C++ executor (sdc_unet_profile_*): every launch in schedule
order with its family, ms, algorithmic GB/s and TFLOP/s.  Usage: exec_profile.py [B] [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
for _ in range(3):
    net.denoise_uniform(x, 500)
torch.cuda.synchronize()
plan = net._plan_ready()
runs = []
for _ in range(R):
    plan.profile(True)
    net.denoise_uniform(x, 500)
    torch.cuda.synchronize()
    runs.append(plan.profile_entries())
    plan.profile(False)
n = len(runs[0])
tot = 0.0
fam = {}
for i in range(n):
    name, _, by, fl = runs[0][i]
    ms = min(r[i][1] for r in runs)
    tot += ms
    d = fam.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += ms
    print(f"{i:3d} {name:22s} {ms*1e3:8.1f} us  {by/ms/1e6:7.0f} GB/s  {fl/ms/1e9:7.0f} TF/s")
print(f"sum of launches {tot:.3f} ms (min over {R} runs)")
for k, (c, ms) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:22s} n={c:3d} {ms:7.3f} ms")
