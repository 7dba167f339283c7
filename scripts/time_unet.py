"""Device timing of one denoiser evaluation (CUDA events), per batch size; prints achieved conv TFLOP/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
for B in [int(a) for a in (sys.argv[1:] or ["8", "64", "256", "1024"])]:
    x = torch.randn(B, 3, 16, 128, device="cuda")
    for _ in range(3):
        net.denoise_uniform(x, 500)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        net.denoise_uniform(x, 500)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B}: {ms:.2f} ms/eval  {B*27.811e9/ms/1e9:.1f} TFLOP/s conv  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
