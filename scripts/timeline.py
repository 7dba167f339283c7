"""In-situ kernel timeline of reverse steps (torch.profiler / CUPTI, no replay): per-kernel busy time and the idle gaps between
consecutive kernels -- what ncu's serialised cold-cache launch list cannot show."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import safediffcon_b200 as s

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NSTEP = 3
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, temporal=True, use_conv2d=True, is_condition_u0=True,
                         is_condition_uT=True).cuda()
cfg = type("Cfg", (), dict(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0}))()
gs = s.safety_guidance(cfg, 0.0).struct()
table, times, rows = gd._coef_table(1, None)
img = torch.randn(B, 3, 16, 128, device="cuda")
nxt = torch.empty_like(img)
u0 = torch.zeros(B, 128, device="cuda")


def step(i):
    global img, nxt
    eps = net.denoise_uniform(img, times[i])
    gd._step(1, img, eps, None, nxt, table, i, gs, None, (u0, u0, None), True, 1, 0)
    img, nxt = nxt, img


for i in range(3):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3, 3 + NSTEP):
        step(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
busy = collections.defaultdict(float)
cnt = collections.Counter()
gap_after = collections.defaultdict(float)
total_gap = 0.0
for a, b in zip(ev, ev[1:] + [None]):
    name = a.name.split("(")[0][:70]
    busy[name] += a.time_range.end - a.time_range.start
    cnt[name] += 1
    if b is not None:
        g = max(0.0, b.time_range.start - a.time_range.end)
        gap_after[name] += g
        total_gap += g
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"B={B}: {NSTEP} steps, span {span / NSTEP / 1e3:.3f} ms/step, busy {sum(busy.values()) / NSTEP / 1e3:.3f} ms/step, "
      f"idle gaps {total_gap / NSTEP / 1e3:.3f} ms/step, {len(ev) / NSTEP:.0f} kernels/step")
for k, v in sorted(busy.items(), key=lambda kv: -kv[1]):
    print(f"{k:72s} n={cnt[k] / NSTEP:5.0f}  busy {v / NSTEP / 1e3:7.3f} ms  gap-after {gap_after[k] / NSTEP / 1e3:6.3f} ms")
