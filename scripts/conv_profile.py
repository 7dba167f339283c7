"""Per-conv-launch CUDA-event timing of one U-Net evaluation (no profiler): shape, ms, TFLOP/s."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s
from safediffcon_b200 import unet as U

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
for _ in range(3):
    net.denoise_uniform(x, 500)
torch.cuda.synchronize()
U.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    net.denoise_uniform(x, 500)
e1.record()
torch.cuda.synchronize()
prof, U.PROFILE = U.PROFILE, None
agg = collections.OrderedDict()
for a, b, fl, shape in prof:
    d = agg.setdefault(shape, [0, 0.0, fl])
    d[0] += 1
    d[1] += a.elapsed_time(b)
tot = 0.0
print("kind  B    H   W   Cin  Cout   n   ms/launch  TFLOP/s   ms/eval")
for shape, (n, ms, fl) in agg.items():
    per = ms / n
    tot += ms / 3
    print(f"{shape[0]:3d} {shape[1]:5d} {shape[2]:3d} {shape[3]:4d} {shape[4]:5d} {shape[5]:5d} {n//3:3d}  {per:8.3f}  {fl/per/1e9:8.1f}  {ms/3:7.2f}")
print(f"conv total {tot:.2f} ms/eval; whole eval {e0.elapsed_time(e1)/3:.2f} ms (with event overhead)")
