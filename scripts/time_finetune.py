"""Inference-time fine-tuning iteration of the reference (InferenceFT.run_epoch inner loop, inference_ft.py:228-238): guided
DDIM-200 chain with the last step under autograd -> finetune_step's hinge loss -> backward (CUDA parameter gradients) ->
AdamW step.  Device-timed per phase at the reference's test batch size (50)."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=200, ddim_sampling_eta=1.0, temporal=True,
                         use_conv2d=True, is_condition_u0=True, is_condition_uT=True).cuda()
cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
opt = torch.optim.AdamW(gd.parameters(), lr=1e-5)
u0 = 0.1 * torch.randn(B, 128, device="cuda")
Q = 0.0


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def iteration():
    t = [ev()]
    out = gd.sample(batch_size=B, clip_denoised=True, u_init=u0, u_final=u0, guidance_u0=True,
                    nablaJ=lambda x: s.get_finetune_guidance(cfg, x, Q), enable_grad=True, seed=1)
    t.append(ev())
    pred = out * 10.0
    sfty = pred[:, 2, :11, :].amax(dim=(-1, -2))
    loss = torch.nn.functional.mse_loss(torch.maximum(sfty + Q - 0.64, torch.zeros_like(sfty)), torch.zeros_like(sfty))
    # a random-init chain saturates the clamp (zero gradient); add a dense term so that the backward does real work
    loss = loss + 1e-3 * (out * out).mean()
    loss.backward()
    t.append(ev())
    opt.step()
    opt.zero_grad()
    t.append(ev())
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in zip(t[:-1], t[1:])]


for _ in range(3):
    iteration()
ms = iteration()
print(f"B={B}: guided DDIM-200 chain + recorded last step {ms[0]:.1f} ms | loss + backward (dgrad + wgrad, all 276 parameters) {ms[1]:.1f} ms | "
      f"AdamW {ms[2]:.1f} ms | peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
