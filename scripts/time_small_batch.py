"""Host-wall vs device time of one guided reverse step per batch size (shows where the chain is launch-bound)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import safediffcon_b200 as s

torch.manual_seed(42)
net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=50, ddim_sampling_eta=1.0, temporal=True,
                         use_conv2d=True, is_condition_u0=True, is_condition_uT=True).cuda()
cfg = type("Cfg", (), dict(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0}))()
for B in [int(a) for a in (sys.argv[1:] or ["8", "50", "250", "1024"])]:
    u0 = torch.zeros(B, 128, device="cuda")
    kw = dict(batch_size=B, u_init=u0, u_final=u0, guidance_u0=True, nablaJ=s.safety_guidance(cfg, 0.0), enable_grad=False, seed=1)
    gd.sample(**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    gd.sample(**kw)
    t_host = time.perf_counter() - t0      # host time to ENQUEUE the chain
    e1.record(); torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"B={B}: device {e0.elapsed_time(e1)/50:.3f} ms/step  host-enqueue {t_host*1e3/50:.3f} ms/step  wall {t_all*1e3/50:.3f} ms/step", flush=True)
