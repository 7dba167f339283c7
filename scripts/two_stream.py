"""Experiment: one denoiser evaluation of B samples as S sub-batches on S CUDA streams (same executor handle, one workspace per
stream).  The HBM-bound passes of one sub-batch can then run under the tensor-bound convolutions of another.
Usage: two_stream.py [B] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s
from safediffcon_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
R = int(sys.argv[2]) if len(sys.argv) > 2 else 10
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
x = torch.randn(B, 3, 16, 128, device="cuda")
for _ in range(3):
    ref = net.denoise_uniform(x, 500)
torch.cuda.synchronize()
plan = net._plan_ready()


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(R):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / R


print(f"B={B} one stream: {timed(lambda: plan.forward(x, None, 500)):.3f} ms")
for S in (2, 4):
    for skew in (False,):
        nb = B // S
        xs = [x[i * nb:(i + 1) * nb].contiguous() for i in range(S)]
        need = int(L.lib().sdc_unet_workspace_bytes(plan.handle, nb, 16, 128))
        wss = [torch.empty(need, dtype=torch.uint8, device="cuda") for _ in range(S)]
        streams = [torch.cuda.Stream() for _ in range(S)]
        outs = [None] * S

        def run():
            cur = torch.cuda.current_stream()
            for i in range(S):
                streams[i].wait_stream(cur)
                with torch.cuda.stream(streams[i]):
                    outs[i] = plan.forward(xs[i], None, 500, wss[i])
            for i in range(S):
                cur.wait_stream(streams[i])

        ms = timed(run)
        got = torch.cat(outs)
        err = ((got - ref).norm() / ref.norm()).item()
        print(f"B={B} as {S} x {nb} on {S} streams: {ms:.3f} ms  (rel diff to the single pass {err:.2e})")
# sequential sub-batches on ONE stream (what the split alone costs)
nb = B // 2
xs = [x[i * nb:(i + 1) * nb].contiguous() for i in range(2)]
print(f"B={B} as 2 x {nb} on one stream: {timed(lambda: [plan.forward(xi, None, 500) for xi in xs]):.3f} ms")
