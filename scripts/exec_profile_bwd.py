"""Per-launch CUDA-event profile of ONE backward-data pass (recording forward + reverse walk, sdc_unet_backward_data) through the
C++ executor: every launch family with its count and ms.  Usage: exec_profile_bwd.py [B] [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
g = torch.randn(B, 3, 16, 128, device="cuda")
for _ in range(2):
    net.vjp(x, 500, g)
torch.cuda.synchronize()
plan = net._plan_ready(backward=True)
runs = []
for _ in range(R):
    plan.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    net.vjp(x, 500, g)
    e1.record()
    torch.cuda.synchronize()
    runs.append((plan.profile_entries(), e0.elapsed_time(e1)))
    plan.profile(False)
ents = runs[0][0]
n = len(ents)
fam, tot = {}, 0.0
for i in range(n):
    name, _, by, fl = ents[i]
    ms = min(r[0][i][1] for r in runs)
    tot += ms
    d = fam.setdefault(name, [0, 0.0, 0.0, 0.0])
    d[0] += 1; d[1] += ms; d[2] += by; d[3] += fl
    if "-v" in sys.argv:
        print(f"{i:3d} {name:22s} {ms*1e3:8.1f} us  {by/ms/1e6:7.0f} GB/s  {fl/ms/1e9:7.0f} TF/s")
print(f"B={B}: {n} launches, sum {tot:.3f} ms (min over {R} runs); wall (events around the call, profiling on) {min(r[1] for r in runs):.2f} ms; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for k, (c, ms, by, fl) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:22s} n={c:3d} {ms:8.3f} ms   {by/ms/1e6 if ms else 0:7.0f} GB/s {fl/ms/1e9 if ms else 0:7.0f} TF/s")
