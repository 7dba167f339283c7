"""Device time of one backward-data pass (VJP w.r.t. x_t) next to the forward, B control instances, dim-128 U-Net."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
g = torch.randn(B, 3, 16, 128, device="cuda")
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    e[0].record()
    net.denoise_uniform(x, 500)
    e[1].record()
    eps, gx = net.vjp(x, 500, g)
    e[2].record()
    torch.cuda.synchronize()
    print(f"B={B} forward {e[0].elapsed_time(e[1]):.2f} ms, forward(saving)+backward-data {e[1].elapsed_time(e[2]):.2f} ms, "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, |gx| {gx.abs().mean().item():.3e}")
