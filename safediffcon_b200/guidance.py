"""Safety guidance + importance weights: drop-in for /root/reference/1D/utils/guidance.py:58-86 and
/root/reference/1D/inference/guidance.py:9-66 (the posttrain copy is identical).

``config`` is any object with ``use_max_safety``, ``u_bound`` and ``guidance_weights['w_score']`` (the reference's
InferenceConfig / PostTrainConfig).  The shipped guidance has a closed-form gradient, so
``get_finetune_guidance`` needs no autograd; ``safety_guidance(config, Q)`` returns a callable that
``GaussianDiffusion.sample`` recognises and fuses into the reverse-step kernel.
"""
import math

import torch

from . import _lib as L

SCALER = 10.0
NT = 11


def _gstruct(config, Q, nt=NT):
    g = L.Guidance()
    g.mode = 1 if config.use_max_safety else 2
    g.Q = float(Q)
    g.u_bound_sq = float(config.u_bound ** 2)
    g.w_score = float(config.guidance_weights["w_score"])
    g.scaler = SCALER
    g.nt = nt
    return g


def safety_stat(state, use_max_safety=True, nt=NT):
    """red(SCALER * state[:, 2, :nt, :]) per sample; red = mean when use_max_safety (sic) else amax."""
    x = L.dev_f32(state, "state")
    B, C, H, W = x.shape
    out = torch.empty(B, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        L.check(L.lib().sdc_safety_stat(L.ptr(x), L.ptr(out), int(bool(use_max_safety)), SCALER, nt, B, H, W, L.stream_ptr()))
    return out


def calculate_guidance(state, Q, config):
    """w_score * max(red(10*state[:,2,:11,:]) + Q - u_bound^2, 0)   (reference inference/guidance.py:9-37)."""
    s = safety_stat(state, config.use_max_safety)
    Qv = Q if isinstance(Q, torch.Tensor) else float(Q)
    return torch.clamp_min(s + Qv - config.u_bound ** 2, 0.0) * config.guidance_weights["w_score"]


def get_weight(state, Q, config):
    """exp(-guidance) per sample (reference inference/guidance.py:39-46)."""
    x = L.dev_f32(state, "state")
    B, C, H, W = x.shape
    w = torch.empty(B, device=x.device, dtype=torch.float32)
    g = _gstruct(config, Q)
    with torch.cuda.device(x.device):
        L.check(L.lib().sdc_conformal_scores(None, L.ptr(x), None, L.ptr(w), g, float("inf"), B, H, W, L.stream_ptr()))
    return w


def normalize_weights(weights):
    """n*w/sum(w); inf entries are replaced IN PLACE by the largest finite weight, an all-zero vector maps to ones
    (reference inference/guidance.py:48-66)."""
    w = weights
    if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
        raise RuntimeError("safediffcon_b200.normalize_weights: expects a contiguous fp32 CUDA vector (no CPU fallback)")
    out = torch.empty_like(w)
    with torch.cuda.device(w.device):
        L.check(L.lib().sdc_normalize_weights(L.ptr(w), L.ptr(out), None, w.shape[0], L.stream_ptr()))
    return out


class SafetyGuidance:
    """Callable nablaJ for the shipped safety guidance; carries (config, Q) so the sampler can fuse it."""

    def __init__(self, config, Q):
        self.config = config
        self.Q = Q

    def struct(self):
        Q = self.Q.item() if isinstance(self.Q, torch.Tensor) else self.Q
        return _gstruct(self.config, Q)

    def __call__(self, x):
        return get_finetune_guidance(self.config, x, self.Q)


def safety_guidance(config, Q):
    return SafetyGuidance(config, Q)


def get_finetune_guidance(config, x, Q):
    """Gradient of sum_b calculate_guidance(x, Q)[b] w.r.t. x, in closed form (reference utils/guidance.py:79-86)."""
    Qf = Q.item() if isinstance(Q, torch.Tensor) else float(Q)
    xs = x.detach()
    B, C, H, W = xs.shape
    w = float(config.guidance_weights["w_score"])
    g = torch.zeros_like(xs, dtype=torch.float32)
    s = safety_stat(xs, config.use_max_safety)
    margin = (s + Qf) - config.u_bound ** 2
    on = (margin > 0).to(torch.float32) + 0.5 * (margin == 0).to(torch.float32)
    if config.use_max_safety:
        g[:, 2, :NT, :] = (on * (w * SCALER / float(NT * W)))[:, None, None]
    else:
        v = xs[:, 2, :NT, :].to(torch.float32) * SCALER
        ties = (v == s[:, None, None]).to(torch.float32)
        g[:, 2, :NT, :] = ties * ((w * on) / ties.sum(dim=(-1, -2)).clamp_min(1.0) * SCALER)[:, None, None]
    return g
