"""One backward-data pass between cudaProfilerStart/Stop (for ncu --profile-from-start off).  Usage: bwd_once.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
g = torch.randn(B, 3, 16, 128, device="cuda")
for _ in range(2):
    net.vjp(x, 500, g)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.vjp(x, 500, g)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
