"""Oracle (TEST INFRASTRUCTURE): diffusion schedule + guided DDIM / DDPM reverse chains on CPU.

Restates, in plain torch-CPU fp32 ops evaluated in the reference's order, the algebra of
  /root/reference/1D/model/model_utils.py:148-158   cosine beta schedule (fp64)
  /root/reference/1D/model/diffusion.py:111-156     13 schedule buffers (fp64 -> fp32)
  /root/reference/1D/model/diffusion.py:193-203     x0 <-> eps conversions
  /root/reference/1D/model/diffusion.py:226-286     model_predictions (guided)
  /root/reference/1D/model/diffusion.py:288-306     DDPM p_mean_variance / p_sample
  /root/reference/1D/model/diffusion.py:336-366     condition writes
  /root/reference/1D/model/diffusion.py:368-449     p_sample_loop
  /root/reference/1D/model/diffusion.py:451-555     ddim_sample
  /root/reference/1D/utils/guidance.py:58-86        safety guidance (closed-form gradient)
Pinned against the unmodified reference by tests/golden/chain_*.npz (see oracle/make_golden.py).
"""
import math
import torch

SCALER = 10.0  # reference: 1D/utils/common.py:17

BUFFER_NAMES = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
    "posterior_mean_coef1", "posterior_mean_coef2", "loss_weight",
)


def cosine_betas(T, s=0.008):
    grid = torch.linspace(0, T, T + 1, dtype=torch.float64)
    abar = torch.cos(((grid / T) + s) / (1 + s) * math.pi * 0.5) ** 2
    abar = abar / abar[0]
    return torch.clip(1 - (abar[1:] / abar[:-1]), 0, 0.999)


def schedule_buffers(T=1000):
    """fp32 buffers keyed like GaussianDiffusion.state_dict() (objective pred_noise)."""
    b = cosine_betas(T)
    a = 1.0 - b
    abar = torch.cumprod(a, dim=0)
    abar_prev = torch.cat([torch.ones(1, dtype=torch.float64), abar[:-1]])
    pv = b * (1.0 - abar_prev) / (1.0 - abar)
    out = {
        "betas": b,
        "alphas_cumprod": abar,
        "alphas_cumprod_prev": abar_prev,
        "sqrt_alphas_cumprod": torch.sqrt(abar),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - abar),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / abar),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / abar - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": torch.log(pv.clamp(min=1e-20)),
        "posterior_mean_coef1": b * torch.sqrt(abar_prev) / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * torch.sqrt(a) / (1.0 - abar),
        "loss_weight": torch.ones_like(abar),
    }
    return {k: v.to(torch.float32) for k, v in out.items()}


def ddim_time_pairs(T=1000, S=200):
    ts = torch.linspace(-1, T - 1, steps=S + 1).int().tolist()
    ts = ts[::-1]
    return list(zip(ts[:-1], ts[1:]))


def safety_guidance_grad(x0, Q, w_score, u_bound, use_max_safety=True, nt=11):
    """Gradient of sum_b w_score*relu(red(10*x0[b,2,:nt,:]) + Q - u_bound^2) wrt x0 (closed form).

    red = mean when use_max_safety (sic, reference utils/guidance.py:68-71) else amax.
    """
    g = torch.zeros_like(x0)
    s = x0[:, 2, :nt, :] * SCALER
    if use_max_safety:
        m = s.mean(dim=(-1, -2))
        on = (m + Q - u_bound ** 2) > 0
        val = torch.tensor(w_score * SCALER / float(nt * x0.shape[-1]), dtype=torch.float32)
        g[:, 2, :nt, :] = on.to(x0.dtype)[:, None, None] * val
    else:
        B = x0.shape[0]
        flat = s.reshape(B, -1)
        m = flat.amax(dim=1)
        on = (m + Q - u_bound ** 2) > 0
        ties = (flat == m[:, None]).to(x0.dtype)
        ties = ties / ties.sum(dim=1, keepdim=True)  # autograd amax splits ties evenly
        g[:, 2, :nt, :] = (ties * on.to(x0.dtype)[:, None] * (w_score * SCALER)).reshape(B, nt, -1)
    return g


def write_conditions(img, u_init, u_final, w_gt, cond_idx=10, pad=True):
    img[:, 0, 0, :] = u_init
    img[:, 0, cond_idx, :] = u_final
    if w_gt is not None:
        img[:, 1, :, :] = w_gt
    if pad:
        img[:, 0, cond_idx + 1:, :] = 0
        img[:, 1, cond_idx:, :] = 0
        img[:, 2, cond_idx:, :] = 0
    return img


def _coef(buf, t):
    return buf[t].reshape(1, 1, 1, 1)


def predictions(bufs, x, t, eps, clip, guide):
    """(pred_noise, x_start) for a uniform integer t.  guide = None or dict(Q,w_score,u_bound,use_max_safety,sched)."""
    c1 = _coef(bufs["sqrt_recip_alphas_cumprod"], t)
    c2 = _coef(bufs["sqrt_recipm1_alphas_cumprod"], t)
    x0 = c1 * x - c2 * eps
    if clip:
        x0 = x0.clamp(-1.0, 1.0)
    if guide is not None:
        g = safety_guidance_grad(x0, guide["Q"], guide["w_score"], guide["u_bound"], guide.get("use_max_safety", True))
        eps = eps + g * guide.get("sched", 1.0)
    x0 = c1 * x - c2 * eps
    if clip:
        x0 = x0.clamp(-1.0, 1.0)
        eps = (c1 * x - x0) / c2
    return eps, x0


def ddim_chain(eps_fn, bufs, noises, u_init, u_final, w_gt=None, guide=None, S=200, eta=1.0, T=1000,
               cond_idx=10, record=None):
    """noises: sequence of S tensors (initial draw + one per non-final pair).  record(step, x_t, t, eps_model)."""
    abar = bufs["alphas_cumprod"]
    img = noises[0].clone()
    write_conditions(img, u_init, u_final, w_gt, cond_idx)
    k = 1
    for step, (t, tn) in enumerate(ddim_time_pairs(T, S)):
        e_model = eps_fn(img, t)
        if record is not None:
            record(step, img, t, e_model)
        eps, x0 = predictions(bufs, img, t, e_model, True, guide)
        if tn < 0:
            img = x0
            continue
        a, an = abar[t], abar[tn]
        sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        c = (1 - an - sigma ** 2).sqrt()
        img = x0 * an.sqrt() + c * eps + sigma * noises[k]
        k += 1
        write_conditions(img, u_init, u_final, w_gt, cond_idx)
    return img


def ddpm_chain(eps_fn, bufs, noises, u_init, u_final, w_gt=None, guide=None, guidance_u0=True, T=1000,
               cond_idx=10, clip_denoised=True, record=None, enable_grad=False):
    """DDPM loop.  guidance_u0=True: one eval + one draw per step (none at t=0).
    guidance_u0=False (calibration call): two p_sample calls per step; the second reuses the first's
    pred_noise (+0 guidance when nablaJ is None) and its draw is the one kept (diffusion.py:416-423)."""
    img = noises[0].clone()
    k = 1

    def p_sample(img, t, eps_in, g):
        nonlocal k
        eps, x0 = predictions(bufs, img, t, eps_in, False, g)
        if clip_denoised:
            x0 = x0.clamp(-1.0, 1.0)
        mean = _coef(bufs["posterior_mean_coef1"], t) * x0 + _coef(bufs["posterior_mean_coef2"], t) * img
        if t > 0:
            z = noises[k]
            k += 1
            out = mean + (0.5 * _coef(bufs["posterior_log_variance_clipped"], t)).exp() * z
        else:
            out = mean + (0.5 * _coef(bufs["posterior_log_variance_clipped"], t)).exp() * 0.0
        return out, x0, eps

    for step, t in enumerate(reversed(range(T))):
        write_conditions(img, u_init, u_final, w_gt, cond_idx)
        e_model = eps_fn(img, t)
        if record is not None:
            record(step, img, t, e_model)
        if guidance_u0:
            img, _, _ = p_sample(img, t, e_model, guide)
        elif t == 0 and enable_grad:
            # reference quirk (diffusion.py:431-447): under enable_grad the t=0 branch only adopts the
            # p_sample result when guidance_u0 is set, so x_1 (conditions written) is returned as is
            pass
        else:
            _, _, eps1 = p_sample(img, t, e_model, None)
            # the reference evaluates the model again but then overrides its output with pred_noise
            img, _, _ = p_sample(img, t, eps1, None)
    return img
