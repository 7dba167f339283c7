"""Oracle (TEST INFRASTRUCTURE): deterministic inputs shared by oracle/make_golden.py and tests/.

Everything here is regenerated from seeds (numpy Generator / torch CPU Generator are reproducible across
machines), so the committed golden files only need to hold the reference's OUTPUTS.
"""
import math
import numpy as np
import torch
import torch.nn as nn

from safediffcon_b200.synthetic import burgers_instances


def solver_inputs(n, seed=0):
    u0, f = burgers_instances(n, seed=seed)
    return torch.from_numpy(u0), torch.from_numpy(f)


def solver_inputs_wild(n, seed=3):
    """Amplitudes far outside the data range: the explicit scheme blows up (Inf/NaN) for some rows."""
    u0, f = burgers_instances(n, seed=seed)
    scale = np.array([1.0, 6.0, 40.0, 400.0], dtype=np.float32)[:n, None]
    return torch.from_numpy(u0 * scale), torch.from_numpy(f * scale[:, None])


class FakeEps(nn.Module):
    """Cheap deterministic stand-in for the U-Net (has the attributes GaussianDiffusion reads)."""
    channels = 3
    self_condition = False
    out_dim = 3

    def forward(self, x, time, x_self_cond=None, residual=None):
        ph = (time.to(torch.float32) * (2 * math.pi / 1000.0)).reshape(-1, 1, 1, 1)
        return 0.8 * torch.tanh(1.5 * x.flip(-1)) + 0.3 * torch.sin(ph + 3.0 * x) + 0.05


# name, T, S, options.  S == T -> DDPM loop; S < T -> DDIM.
CHAIN_CASES = (
    ("ddim_guided", 1000, 8, dict(guided=True, Q=1.35, guidance_u0=True, w_gt=False, enable_grad=False, seed=11)),
    ("ddim_guided_grad", 1000, 8, dict(guided=True, Q=1.25, guidance_u0=True, w_gt=False, enable_grad=True, seed=12)),
    ("ddim_guided_amax", 1000, 5, dict(guided=True, Q=-0.2, guidance_u0=True, w_gt=False, enable_grad=False, seed=13,
                                       use_max_safety=False)),
    ("ddim_plain", 1000, 8, dict(guided=False, Q=0.0, guidance_u0=True, w_gt=False, enable_grad=True, seed=14)),
    ("ddim_calib", 1000, 6, dict(guided=False, Q=0.0, guidance_u0=False, w_gt=True, enable_grad=False, seed=15)),
    ("ddpm_guided", 20, 20, dict(guided=True, Q=1.0, guidance_u0=True, w_gt=False, enable_grad=False, seed=16)),
    ("ddpm_guided_grad", 20, 20, dict(guided=True, Q=0.95, guidance_u0=True, w_gt=False, enable_grad=True, seed=17)),
    ("ddpm_calib", 20, 20, dict(guided=False, Q=0.0, guidance_u0=False, w_gt=True, enable_grad=False, seed=18)),
    ("ddpm_calib_grad", 20, 20, dict(guided=False, Q=0.0, guidance_u0=False, w_gt=True, enable_grad=True, seed=19)),
)


def n_draws(T, S, guidance_u0, enable_grad=False):
    """Number of randn draws the reference makes (SURVEY.md section 7 'RNG parity')."""
    if S < T:
        return S  # initial + one per non-final pair
    per = 1 if guidance_u0 else 2
    return 1 + per * (T - 1)


def chain_noise(B, n, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 3, 16, 128, generator=g) for _ in range(n)]


def chain_conditions(B, seed=7):
    g = torch.Generator().manual_seed(seed)
    u_init = 0.2 * torch.randn(B, 128, generator=g)
    u_final = 0.1 * torch.randn(B, 128, generator=g)
    w_gt = 0.3 * torch.randn(B, 16, 128, generator=g)
    return u_init, u_final, w_gt


def guidance_states(B, seed=21):
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(B, 3, 16, 128, generator=g)
    x[:, 2] += torch.linspace(0.0, 0.12, B).reshape(B, 1, 1)
    return x


def weight_vectors():
    g = torch.Generator().manual_seed(31)
    w = torch.rand(17, generator=g)
    w_inf = w.clone()
    w_inf[[2, 9]] = float("inf")
    return [w, w_inf, torch.zeros(5), torch.tensor([0.0, 0.0, 3.0, 1.0])]


def score_vectors():
    g = torch.Generator().manual_seed(41)
    a = torch.rand(1000, generator=g)
    b = torch.randint(0, 7, (1000,), generator=g).to(torch.float32) * 0.25  # tie-heavy
    c = torch.rand(37, generator=g)
    d = torch.rand(5000, generator=g) * 1e-3
    return [(a, 0.98), (b, 0.98), (c, 0.9), (c, 0.999), (d, 0.5), (b, 0.5)]


def calibration_states(n, seed=51):
    g = torch.Generator().manual_seed(seed)
    st = 0.1 * torch.randn(n, 3, 16, 128, generator=g)
    st[:, 2] = (0.02 + 0.08 * torch.rand(n, 1, 1, generator=g)).expand(n, 16, 128)
    st[:, 0, 11:] = 0
    st[:, 1, 10:] = 0
    st[:, 2, 11:] = 0
    return st


def unet_inputs(B, seed=61):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, 16, 128, generator=g)
    t = torch.tensor([999, 417, 3, 250, 0, 731][:B], dtype=torch.long)
    return x, t


def unet_cotangent(B, seed=67):
    """Cotangent g for the input-gradient tests: d<eps, g>/dx."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, 16, 128, generator=g)


def config1_conditions(B, seed=0):
    """u0 / target final state / full target trajectory for BASELINE config 1 (model units for u0,uT)."""
    from oracle import solver_ref
    u0, f = burgers_instances(B, seed=seed)
    traj = solver_ref.solve_free_c(u0, f)
    u0_t = torch.from_numpy(u0) / 10.0
    uT_t = torch.from_numpy(traj[:, -1, :].copy()) / 10.0
    return u0_t, uT_t, torch.from_numpy(traj)


PGRAD_SAMPLES = 256


def grad_digest(named_grads):
    """Compact fingerprint of a set of parameter gradients (full tensors would be ~100 MB): per parameter the L2 norm,
    the sum, and PGRAD_SAMPLES evenly spaced entries of the flattened tensor."""
    import numpy as np
    out = {}
    for name, g in named_grads:
        flat = g.detach().double().flatten().cpu()
        idx = torch.linspace(0, flat.numel() - 1, PGRAD_SAMPLES).round().long()
        out[name + "|norm"] = np.float64(flat.norm().item())
        out[name + "|sum"] = np.float64(flat.sum().item())
        out[name + "|samples"] = flat[idx].numpy()
    return out


def finetune_loss(sample, Q=0.0, u_bound=0.8, scaler=10.0):
    """InferenceFT.finetune_step's objective (/root/reference/1D/inference/inference_ft.py:189-203) on a sample() output."""
    pred = sample * scaler
    s = pred[:, 2, :11, :].amax(dim=(-1, -2))
    obj = torch.maximum(s + Q - u_bound ** 2, torch.zeros_like(s))
    return torch.nn.functional.mse_loss(obj, torch.zeros_like(s))


def last_step_state(B, seed=71):
    """x_t entering the last DDIM-200 pair (t = 4): a dataset-like state plus a little noise, so that the predicted x0
    stays inside the clamp and the fine-tuning hinge is active."""
    g = torch.Generator().manual_seed(seed)
    return calibration_states(B) + 0.01 * torch.randn(B, 3, 16, 128, generator=g)
