"""Accuracy margins of the backward-data pass against the reference's autograd goldens (tests/golden/unet_dim*_vjp.npz): eps and
d<eps, g>/dx of Unet2D.vjp, and the distance between the recording forward's eps and the inference eps."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import safediffcon_b200 as s
from oracle import fixtures as fx
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm()).item()
for dim, prec in ((32, "tf32"), (64, "f16"), (64, "tf32")):
    torch.manual_seed(42)
    net = s.Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    net.precision = prec
    B = 2
    x, t = fx.unet_inputs(B); g = fx.unet_cotangent(B)
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"unet_dim{dim}_vjp.npz"))
    with torch.no_grad():
        eps, gx = net.vjp(x.cuda(), t.cuda(), g.cuda())
        e_inf = net(x.cuda(), t.cuda())
    print(dim, prec, "eps vs golden", rel(eps.cpu(), torch.from_numpy(gold["eps"])), "gx vs golden", rel(gx.cpu(), torch.from_numpy(gold["grad_x"])), "vjp eps vs inference eps", rel(eps, e_inf))
