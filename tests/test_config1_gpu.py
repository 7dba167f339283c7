"""GPU parity on BASELINE config 1: the reference's guided chains of the real dim-128 U-Net (seed-42
weights, B=8, w_score 500) recorded by oracle/make_golden.py -- DDIM-200 (the repo default) and the north star's DDPM-1000
p_sample_loop, each at Q = 0 and Q = 0.05 -- then solver + metrics."""
import types

import numpy as np
import pytest
import torch

from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
STEPS = (0, 1, 60, 120, 199)


@pytest.fixture(scope="module")
def model():
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
    return s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=200, ddim_sampling_eta=1.0,
                               temporal=True, use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10,
                               train_on_padded_locations=False).cuda()


def test_teacher_forced_eps_within_1e3(model, golden):
    """Feed the reference's own x_t at 5 points of its chain: eps must agree within 1e-3 relative (north star)."""
    g = golden("config1_ddim")
    for k in STEPS:
        x, t, ref = torch.from_numpy(g[f"x_{k}"]), torch.from_numpy(g[f"t_{k}"]), torch.from_numpy(g[f"eps_{k}"])
        eps = model.model(x.cuda(), t.cuda()).cpu()
        r = ((eps - ref).norm() / ref.norm()).item()
        per = [((eps[i] - ref[i]).norm() / ref[i].norm()).item() for i in range(eps.shape[0])]
        assert r < 1e-3 and max(per) < 1e-3, (k, r, per)


def test_teacher_forced_step_matches_reference_next_state(model, golden):
    """x_0 -> x_1 of the reference chain: with the reference's eps the fused step is bit-exact; with our TF32 eps the
    state agrees to the clamp-limited tolerance."""
    import safediffcon_b200 as s
    g = golden("config1_ddim")
    B = 8
    u0, uT, _ = fx.config1_conditions(B)
    noises = fx.chain_noise(B, fx.n_draws(1000, 200, True), seed=1234)
    table, times, rows = model._coef_table(0, None)
    cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
    gs = s.safety_guidance(cfg, 0.0).struct()
    x0 = torch.from_numpy(g["x_0"]).cuda()
    out = torch.empty_like(x0)
    model._step(0, x0, torch.from_numpy(g["eps_0"]).cuda(), noises[1].cuda(), out, table, 0, gs, None,
                (u0.cuda(), uT.cuda(), None), True, 0, 0)
    assert torch.equal(out.cpu(), torch.from_numpy(g["x_1"]))


def test_free_running_chain_metrics_within_1pct(model, golden):
    """Whole drop-in call with the reference's draws, then control_trajectories + evaluate_samples.
    TF32 chains decorrelate sample-by-sample from the fp32 reference, so the comparison is
    on the reported metrics: J and the violation rates within 1% (north star)."""
    import safediffcon_b200 as s
    g = golden("config1_ddim")
    B = 8
    u0, uT, tgt = fx.config1_conditions(B)
    noises = fx.chain_noise(B, fx.n_draws(1000, 200, True), seed=1234)
    cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
    res = model.sample(batch_size=B, clip_denoised=True, u_init=u0.cuda(), u_final=uT.cuda(), guidance_u0=True,
                       nablaJ=s.safety_guidance(cfg, 0.0), J_scheduler=None, w_scheduler=None, enable_grad=False, device="cuda",
                       noise=noises)
    pred = res * 10.0
    uc = s.control_trajectories(pred, 11)
    m = s.evaluate_samples(pred, uc, tgt.cuda(), nt=11, u_bound=0.8)
    ref = dict(J=float(g["J"]), Rp=float(g["Rp"]), Rt=float(g["Rt"]), Rs=float(g["Rs"]))
    got = dict(J=m["control_mse_mean (J)"], Rp=m["point_exceed_ratio (R_p)"], Rt=m["time_exceed_ratio (R_t)"],
               Rs=m["sample_exceed_ratio (R_s)"])
    drift = (res.cpu() - torch.from_numpy(g["sample"])).abs().mean().item()
    print("config1:", got, ref, "mean |sample drift|", drift)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 0.01 * max(abs(ref[k]), 1e-6), (k, got[k], ref[k])


# ---------------------------------------------------------------------------------------------- all four reference runs
RUNS = {   # golden file -> (sampler steps, Q, recorded U-Net evaluations)
    "config1_ddim_Q005": (200, 0.05, (0, 120, 199)),
    "config1_ddpm": (1000, 0.0, (0, 1, 250, 500, 750, 900, 999)),
    "config1_ddpm_Q005": (1000, 0.05, (0, 500, 999)),
}


@pytest.fixture(scope="module")
def ddpm_model(model):
    import safediffcon_b200 as s
    return s.GaussianDiffusion(model.model, seq_length=(16, 128), timesteps=1000, sampling_timesteps=1000, ddim_sampling_eta=1.0,
                               temporal=True, use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10,
                               train_on_padded_locations=False).cuda()


@pytest.mark.parametrize("name", sorted(RUNS))
def test_teacher_forced_eps_all_runs(model, golden, name):
    """eps on the reference's own x_t at the recorded points of each chain (DDPM: t = 999 ... 0): < 1e-3 per sample."""
    g = golden(name)
    for k in RUNS[name][2]:
        x, t, ref = torch.from_numpy(g[f"x_{k}"]), torch.from_numpy(g[f"t_{k}"]), torch.from_numpy(g[f"eps_{k}"])
        eps = model.model(x.cuda(), t.cuda()).cpu()
        per = [((eps[i] - ref[i]).norm() / ref[i].norm()).item() for i in range(eps.shape[0])]
        assert max(per) < 1e-3, (name, k, per)


@pytest.mark.parametrize("name", sorted(RUNS))
def test_free_running_chain_metrics_all_runs(model, ddpm_model, golden, name):
    """The whole drop-in call with the reference's draws (DDPM-1000: 1000 denoiser evaluations), then rollout + metrics:
    J and the violation rates within 1 % of the reference run (north star)."""
    import safediffcon_b200 as s
    S, Q, _ = RUNS[name]
    g = golden(name)
    assert abs(float(g["Q"]) - Q) < 1e-12
    gd = model if S == 200 else ddpm_model
    B = 8
    u0, uT, tgt = fx.config1_conditions(B)
    noises = fx.chain_noise(B, fx.n_draws(1000, S, True), seed=1234)
    cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
    res = gd.sample(batch_size=B, clip_denoised=True, u_init=u0.cuda(), u_final=uT.cuda(), guidance_u0=True,
                    nablaJ=s.safety_guidance(cfg, Q), J_scheduler=None, w_scheduler=None, enable_grad=False, device="cuda", noise=noises)
    pred = res * 10.0
    uc = s.control_trajectories(pred, 11)
    m = s.evaluate_samples(pred, uc, tgt.cuda(), nt=11, u_bound=0.8)
    ref = dict(J=float(g["J"]), Rp=float(g["Rp"]), Rt=float(g["Rt"]), Rs=float(g["Rs"]))
    got = dict(J=m["control_mse_mean (J)"], Rp=m["point_exceed_ratio (R_p)"], Rt=m["time_exceed_ratio (R_t)"],
               Rs=m["sample_exceed_ratio (R_s)"])
    drift = (res.cpu() - torch.from_numpy(g["sample"])).abs().mean().item()
    print(name, got, ref, "mean |sample drift|", drift)
    for k in ref:
        assert abs(got[k] - ref[k]) <= 0.01 * max(abs(ref[k]), 1e-6), (name, k, got[k], ref[k])


def test_ddpm_first_step_is_bit_exact_given_reference_eps(ddpm_model, golden):
    """x_999 -> x_998 of the reference DDPM chain: with the reference's eps and draw the fused step reproduces the state the
    reference fed to its next U-Net evaluation bit for bit (conditions are written at the start of that iteration)."""
    import safediffcon_b200 as s
    g = golden("config1_ddpm")
    B = 8
    u0, uT, _ = fx.config1_conditions(B)
    noises = fx.chain_noise(B, fx.n_draws(1000, 1000, True), seed=1234)
    table, times, rows = ddpm_model._coef_table(1, None)
    cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
    gs = s.safety_guidance(cfg, 0.0).struct()
    x0 = torch.from_numpy(g["x_0"]).cuda()
    out = torch.empty_like(x0)
    ddpm_model._step(1, x0, torch.from_numpy(g["eps_0"]).cuda(), noises[1].cuda(), out, table, 0, gs, None,
                     (u0.cuda(), uT.cuda(), None), True, 0, 0)
    assert torch.equal(out.cpu(), torch.from_numpy(g["x_1"]))
