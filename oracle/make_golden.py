"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (build container only).

The reference has no tests or fixtures of its own, so parity is pinned against the
outputs this script records.  Usage:  python -m oracle.make_golden [names...]   (default: all fast sets;
``config1`` runs the real U-Net DDIM-200 chain at B=8 and takes a few minutes.)

Shared deterministic inputs live in oracle/fixtures.py so that tests regenerate them instead of storing them.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle import fixtures as fx  # noqa: E402

ref_import.install()
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


class NoisePatch:
    """Feed the reference's global-RNG draws (torch.randn / torch.randn_like) from a supplied list."""

    def __init__(self, noises):
        self.noises = list(noises)
        self.k = 0

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def take(*a, **k):
            z = self.noises[self.k]
            self.k += 1
            return z.clone()

        torch.randn = take
        torch.randn_like = take
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def gen_schedule():
    from model.diffusion import GaussianDiffusion
    for T in (1000, 20):
        gd = GaussianDiffusion(fx.FakeEps(), seq_length=(16, 128), timesteps=T, temporal=True, use_conv2d=True)
        sd = {k: v for k, v in gd.state_dict().items() if not k.startswith("model.")}
        save(f"schedule_T{T}", **sd)


def gen_solver():
    from data.generate_burgers import burgers_numeric_solve, burgers_numeric_solve_free
    u0, f = fx.solver_inputs(16, seed=0)
    traj = burgers_numeric_solve_free(u0, f, visc=0.01, T=1.0, dt=1e-4, num_t=10)
    save("solver_free", traj=traj)
    # stiff / diverging inputs: NaN and Inf must propagate exactly like the reference
    u0b, fb = fx.solver_inputs_wild(4, seed=3)
    trajb = burgers_numeric_solve_free(u0b, fb, visc=0.01, T=1.0, dt=1e-4, num_t=10)
    save("solver_free_wild", traj=trajb)
    u0c, fc = fx.solver_inputs(3, seed=5)
    trajc = burgers_numeric_solve(u0c, fc[:2], visc=0.01, T=1.0, dt=1e-4, num_t=10)
    save("solver_cartesian", traj=trajc)
    # metrics on the first set, target = rolled trajectories
    from utils.metrics import evaluate_samples
    tgt = torch.roll(traj, 1, dims=0)
    diffused = torch.zeros(16, 3, 16, 128)
    m = evaluate_samples(diffused, traj, tgt, nt=11, u_bound=0.8)
    m2 = evaluate_samples(diffused, traj, tgt, nt=11, u_bound=0.3)
    save("metrics", **{f"b08_{i}": np.asarray(v, dtype=np.float64) for i, v in enumerate(m.values())},
         **{f"b03_{i}": np.asarray(v, dtype=np.float64) for i, v in enumerate(m2.values())},
         keys=np.array(list(m.keys())))


def _diffusion(T, S, eta=1.0, model=None):
    from model.diffusion import GaussianDiffusion
    return GaussianDiffusion(model or fx.FakeEps(), seq_length=(16, 128), timesteps=T, sampling_timesteps=S,
                             ddim_sampling_eta=eta, temporal=True, use_conv2d=True, is_condition_u0=True,
                             is_condition_uT=True, condition_idx=10, train_on_padded_locations=False)


def _guidance_fn(Q, w_score=500.0, use_max_safety=True):
    from utils.guidance import get_finetune_guidance
    cfg = types.SimpleNamespace(use_max_safety=use_max_safety, u_bound=0.8, guidance_weights={"w_score": w_score})
    return lambda x: get_finetune_guidance(cfg, x, Q)


def gen_chains():
    B = 4
    u_init, u_final, w_gt = fx.chain_conditions(B)
    out = {}
    for name, T, S, kw in fx.CHAIN_CASES:
        gd = _diffusion(T, S)
        noises = fx.chain_noise(B, fx.n_draws(T, S, kw["guidance_u0"]), seed=kw["seed"])
        guide = _guidance_fn(kw["Q"], use_max_safety=kw.get("use_max_safety", True)) if kw["guided"] else None
        with NoisePatch(noises) as npch:
            res = gd.sample(batch_size=B, clip_denoised=True, u_init=u_init, u_final=u_final,
                            guidance_u0=kw["guidance_u0"], nablaJ=guide, J_scheduler=None, w_scheduler=None,
                            w_groundtruth=(w_gt if kw["w_gt"] else None), enable_grad=kw["enable_grad"], device="cpu")
            assert npch.k == len(noises), (name, npch.k, len(noises))
        out[name] = res
    save("chains", **out)


def gen_guidance():
    from inference.guidance import get_weight, normalize_weights
    from inference.conformal import ConformalCalculator
    out = {}
    x = fx.guidance_states(6)
    for Q in (0.0, 0.05, -0.5):
        for ums in (True, False):
            out[f"grad_Q{Q}_{int(ums)}"] = _guidance_fn(Q, use_max_safety=ums)(x.clone().requires_grad_())
            cfg = types.SimpleNamespace(use_max_safety=ums, u_bound=0.8, guidance_weights={"w_score": 500.0})
            out[f"weight_Q{Q}_{int(ums)}"] = get_weight(x, Q, cfg)
    for i, w in enumerate(fx.weight_vectors()):
        out[f"norm_{i}"] = normalize_weights(w.clone())
    cc = ConformalCalculator(None, types.SimpleNamespace(device="cpu"))
    for i, (s, alpha) in enumerate(fx.score_vectors()):
        out[f"quant_{i}"] = cc.calculate_quantile(s, None, None, alpha)
    save("guidance", **out)


def gen_conformal():
    from inference.conformal import ConformalCalculator
    B, nb = 6, 2
    gd = _diffusion(1000, 6)
    cfg = types.SimpleNamespace(device="cpu", num_cal_batch=nb, nt=11, InfFT_Q=None, use_max_safety=True, u_bound=0.8,
                                guidance_weights={"w_score": 500.0})
    states = fx.calibration_states(B * nb)
    loader = iter([states[i * B:(i + 1) * B] for i in range(nb)])
    noises = []
    for i in range(nb):
        noises += fx.chain_noise(B, fx.n_draws(1000, 6, False), seed=100 + i)
    with NoisePatch(noises):
        scores, weights, st = ConformalCalculator(gd, cfg).get_conformal_scores(loader, Q=0.02)
    save("conformal", scores=scores, weights=weights)


def _unet(dim):
    from model.unet import Unet2D
    torch.manual_seed(42)
    return Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).eval()


def gen_unet():
    for dim, B in ((32, 3), (128, 2)):
        net = _unet(dim)
        x, t = fx.unet_inputs(B)
        taps = {}
        hooks = []
        for nm in ("init_conv", "downs.0.0", "downs.0.2", "downs.0.3", "mid_attn", "ups.0.3", "final_res_block"):
            mod = dict(net.named_modules())[nm]
            hooks.append(mod.register_forward_hook(lambda m, i, o, nm=nm: taps.__setitem__("tap_" + nm, o)))
        with torch.no_grad():
            eps = net(x, t)
        for h in hooks:
            h.remove()
        if dim == 128:  # keep the fixture small: full eps only, plus a strided view of the taps
            taps = {k: v[:, ::8] for k, v in taps.items()}
        save(f"unet_dim{dim}", eps=eps, **taps)
        sd = net.state_dict()
        save(f"unet_dim{dim}_wsum", **{"sum": np.array([float(v.double().sum()) for v in sd.values()]),
                                       "abssum": np.array([float(v.double().abs().sum()) for v in sd.values()]),
                                       "keys": np.array(list(sd.keys()))})


def gen_unet_vjp():
    """Input gradient of the unmodified reference denoiser: d<eps, g>/dx by autograd."""
    for dim, B in ((32, 2), (64, 2)):
        net = _unet(dim)
        x, t = fx.unet_inputs(B)
        g = fx.unet_cotangent(B)
        x = x.clone().requires_grad_()
        eps = net(x, t)
        (gx,) = torch.autograd.grad(eps, x, g)
        save(f"unet_dim{dim}_vjp", eps=eps.detach(), grad_x=gx)


def gen_unet_pgrad():
    """Parameter gradients of the unmodified reference, stored as digests (fixtures.grad_digest):
    (a) dense cotangent: d<eps, g>/d(theta) for the dim-64 denoiser, per-sample times;
    (b) the inference-time fine-tuning loss back-propagated through sample(enable_grad=True) (DDIM, last step under
        autograd, diffusion.py:524-551 + inference_ft.py:189-203), Q chosen so that the hinge is active."""
    dim, B = 64, 2
    net = _unet(dim).train()
    x, t = fx.unet_inputs(B)
    g = fx.unet_cotangent(B)
    eps = net(x, t)
    net.zero_grad()
    eps.backward(g)
    save(f"unet_dim{dim}_pgrad", eps=eps.detach(), **fx.grad_digest([(n, p.grad) for n, p in net.named_parameters()]))
    # (b) the reference's last-step code path (diffusion.py:524-531): model_predictions under autograd at the last time of
    # the DDIM-200 grid, then finetune_step's loss.  (A free-running chain of a random-init model saturates the clamp of
    # x0, which zeroes every gradient -- hence a dataset-like x_t.)
    net = _unet(dim).train()
    gd = _diffusion(1000, 200, model=net)
    t_last, Q = 4, 0.0
    assert gd_time_pairs(gd)[-1] == (t_last, -1)
    img = fx.last_step_state(B)
    with torch.enable_grad():
        time_cond = torch.full((B,), t_last, dtype=torch.long)
        pred_noise, x_start, *_ = gd.model_predictions(img, time_cond, None, clip_x_start=True, rederive_pred_noise=True,
                                                       nablaJ=_guidance_fn(Q), J_scheduler=None)
    loss = fx.finetune_loss(x_start, Q=Q)
    assert float(loss) > 0 and float(x_start.abs().max()) < 1.0
    net.zero_grad()
    loss.backward()
    save(f"unet_dim{dim}_ft_pgrad", x_start=x_start.detach(), loss=loss.detach(),
         **fx.grad_digest([(n, p.grad) for n, p in net.named_parameters()]))


def gen_ploss():
    """Training loss of the unmodified reference (GaussianDiffusion.p_losses, diffusion.py:638-733) and its parameter
    gradients (digest), on dataset-like states with fixed t and noise."""
    dim, B = 64, 2
    net = _unet(dim).train()
    gd = _diffusion(1000, 200, model=net)
    x0 = fx.calibration_states(B)
    t = torch.tensor([417, 3])
    noise = fx.chain_noise(B, 1, seed=9)[0]
    per_sample = gd.p_losses(x0.clone(), t, noise=noise.clone(), mean=False)
    loss = gd.p_losses(x0.clone(), t, noise=noise.clone(), mean=True)
    net.zero_grad()
    loss.backward()
    save(f"unet_dim{dim}_ploss", per_sample=per_sample.detach(), loss=loss.detach(),
         **fx.grad_digest([(n, p.grad) for n, p in net.named_parameters()]))


def gen_datagen():
    """Reference generator + dataset assembly: make_data_varying_f under np.random.seed(0), the
    'front_rear_quarter' partial-control variant, and the [3,16,128] state of BurgersDataset._process_data (that method is
    executed unbound on a stand-in object: the class itself needs h5py files)."""
    from data.generate_burgers import make_data_varying_f, burgers_numeric_solve_free
    import numpy as np_
    np_.random.seed(0)
    u0, f = make_data_varying_f(16, 12, 128, 10)
    np_.random.seed(3)
    u0p, fp = make_data_varying_f(3, 4, 128, 10, partial_control='front_rear_quarter', alpha=1.7)
    save("datagen", u0=u0, f=f, u0_partial=u0p, f_partial=fp)
    # dataset states from reference rollouts
    import data.burgers as rb
    traj = burgers_numeric_solve_free(torch.tensor(u0[:12], dtype=torch.float32), f, visc=0.01, T=1.0, dt=1e-4, num_t=10)
    outs = {}
    for use_max in (True, False):
        ds = types.SimpleNamespace(safety_transform=lambda u: u.pow(2), use_max_safety=use_max, stack_u_and_f=True,
                                   pad_for_2d_conv=True, pad_size=16, scaler=10.0)
        outs[f"states_max{int(use_max)}"] = torch.stack([rb.BurgersDataset._process_data(ds, (traj[i], f[i])) for i in range(12)])
    save("dataset_states", traj=traj, f=f, **outs)


def gd_time_pairs(gd):
    times = torch.linspace(-1, gd.num_timesteps - 1, steps=gd.sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def _config1_run(sampler, Q, steps):
    """One reference run of BASELINE config 1: real dim-128 U-Net (seed 42), B=8, guided (w_score 500),
    `sampler` = 'ddim' (S=200, eta=1: the repo default) or 'ddpm' (the north star's 1000-step p_sample_loop), then the
    reference solver + metrics.  Records (x_t, t, eps) at `steps` (indices of U-Net evaluations) and the state after step 0."""
    from utils.metrics import control_trajectories, evaluate_samples
    B = 8
    S = 200 if sampler == "ddim" else 1000
    net = _unet(128)
    gd = _diffusion(1000, S, model=net)
    u0, uT, tgt = fx.config1_conditions(B)
    noises = fx.chain_noise(B, fx.n_draws(1000, S, True), seed=1234)
    keep = {}
    cnt = {"k": 0}
    orig = net.forward

    def spy(x, time, *a, **k):
        o = orig(x, time, *a, **k)
        if cnt["k"] in steps:
            keep[f"x_{cnt['k']}"] = x.detach().clone()
            keep[f"eps_{cnt['k']}"] = o.detach().clone()
            keep[f"t_{cnt['k']}"] = time.detach().clone()
        cnt["k"] += 1
        return o

    net.forward = spy
    import time as _time
    t0 = _time.time()
    with NoisePatch(noises):
        res = gd.sample(batch_size=B, clip_denoised=True, u_init=u0, u_final=uT, guidance_u0=True,
                        nablaJ=_guidance_fn(Q), J_scheduler=None, w_scheduler=None, enable_grad=False, device="cpu")
    t_chain = _time.time() - t0
    assert cnt["k"] == S
    pred = res * 10.0
    t0 = _time.time()
    uc = control_trajectories(pred, 11)
    t_solve = _time.time() - t0
    m = evaluate_samples(pred, uc, tgt, nt=11, u_bound=0.8)
    return dict(sample=res, u_controlled=uc, J=m["control_mse_mean (J)"], Rp=m["point_exceed_ratio (R_p)"],
                Rt=m["time_exceed_ratio (R_t)"], Rs=m["sample_exceed_ratio (R_s)"], t_chain=t_chain, t_solve=t_solve,
                threads=torch.get_num_threads(), Q=Q, **keep)


def gen_config1():
    """BASELINE config 1, DDIM-200 eta=1 guided, Q 0 (file name kept from round 1) and Q 0.05."""
    save("config1_ddim", **_config1_run("ddim", 0.0, (0, 1, 60, 120, 199)))
    save("config1_ddim_Q005", **_config1_run("ddim", 0.05, (0, 120, 199)))


def gen_config1_ddpm():
    """BASELINE config 1 with the north star's sampler: the 1000-step DDPM p_sample_loop (diffusion.py:368-449), Q 0 and 0.05.
    ~10 min of CPU per run on 8 cores."""
    save("config1_ddpm", **_config1_run("ddpm", 0.0, (0, 1, 250, 500, 750, 900, 999)))
    save("config1_ddpm_Q005", **_config1_run("ddpm", 0.05, (0, 500, 999)))


ALL = {"schedule": gen_schedule, "solver": gen_solver, "chains": gen_chains, "guidance": gen_guidance,
       "conformal": gen_conformal, "unet": gen_unet, "unet_vjp": gen_unet_vjp, "unet_pgrad": gen_unet_pgrad, "ploss": gen_ploss, "datagen": gen_datagen, "config1": gen_config1,
       "config1_ddpm": gen_config1_ddpm}

if __name__ == "__main__":
    names = sys.argv[1:] or [k for k in ALL if not k.startswith("config1")]
    for n in names:
        ALL[n]()
