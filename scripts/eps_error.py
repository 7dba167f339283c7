"""eps relative error of the CUDA U-Net against the reference goldens (per fixture, both operand precisions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import safediffcon_b200 as s

torch.set_grad_enabled(False)   # the inference path (fused attention, fused upsample convolutions)

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
g1 = np.load(os.path.join(G, "config1_ddim.npz"))
gu = np.load(os.path.join(G, "unet_dim128.npz"))
for prec in ("f16", "tf32"):
    net.precision = prec
    rows = []
    x = torch.randn(2, 3, 16, 128, generator=torch.Generator().manual_seed(61))   # = oracle.fixtures.unet_inputs(2)
    t = torch.tensor([999, 417])
    e = net(x.cuda(), t.cuda()).cpu()
    ref = torch.from_numpy(gu["eps"])
    rows.append(("unet_dim128", ((e - ref).norm() / ref.norm()).item(), max(((e[i] - ref[i]).norm() / ref[i].norm()).item() for i in range(2))))
    for k in (0, 1, 60, 120, 199):
        x, t, ref = torch.from_numpy(g1[f"x_{k}"]), torch.from_numpy(g1[f"t_{k}"]), torch.from_numpy(g1[f"eps_{k}"])
        e = net(x.cuda(), t.cuda()).cpu()
        rows.append((f"config1 step {k}", ((e - ref).norm() / ref.norm()).item(),
                     max(((e[i] - ref[i]).norm() / ref[i].norm()).item() for i in range(8))))
    for name, r, mx in rows:
        print(f"{prec:5s} {name:18s} rel {r:.3e}  worst sample {mx:.3e}")
