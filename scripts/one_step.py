"""One reverse step of the B-sample guided DDPM chain between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s
from safediffcon_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
torch.cuda.profiler.start()
step(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", img.abs().max().item())
