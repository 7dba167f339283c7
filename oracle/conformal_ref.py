"""Oracle (TEST INFRASTRUCTURE): conformal weights, nonconformity scores and quantile on CPU.

Restates /root/reference/1D/inference/guidance.py:9-66 (calculate_guidance, get_weight, normalize_weights),
/root/reference/1D/inference/conformal.py:68-93 (score / weight assembly) and :95-118 (calculate_quantile).
Pinned against the unmodified reference by tests/golden/conformal.npz.
"""
import numpy as np
import torch

SCALER = 10.0


def safety_stat(state, use_max_safety=True, nt=11):
    s = (state * SCALER)[:, 2, :nt, :]
    return s.mean(dim=(-1, -2)) if use_max_safety else s.amax(dim=(-1, -2))


def guidance_value(state, Q, w_score, u_bound, use_max_safety=True):
    s = safety_stat(state, use_max_safety)
    return torch.maximum(s + Q - u_bound ** 2, torch.zeros_like(s)) * w_score


def raw_weight(state, Q, w_score, u_bound, use_max_safety=True):
    return torch.exp(-guidance_value(state, Q, w_score, u_bound, use_max_safety))


def normalize_weights(w):
    """n*w/sum(w) with the reference's guards: inf -> largest finite (in place), zero sum -> ones."""
    inf = torch.isinf(w)
    if inf.any():
        w[inf] = w[~inf].max()
    if w.sum() == 0:
        return torch.ones_like(w)
    return w.shape[0] * w / w.sum()


def nonconformity(pred_scaled, state_scaled, use_max_safety=True):
    """|red(10*pred[:,2,:11]) - red(10*state[:,2,:11])| (inputs in model units, i.e. divided by SCALER)."""
    return (safety_stat(pred_scaled, use_max_safety) - safety_stat(state_scaled, use_max_safety)).abs()


def quantile_rank(n, alpha):
    return min(int(np.ceil(alpha * (n + 1))), n) - 1


def quantile(scores, alpha):
    """rank-th order statistic (value is unique even though torch.sort's tie order is not)."""
    s = torch.as_tensor(scores)
    r = quantile_rank(s.shape[0], alpha)
    return torch.sort(s).values[r]


def quantile_index(scores, alpha):
    """Index of the selected element with the documented tie-break: among equal values, lowest index first
    (i.e. a stable sort).  The reference's torch.sort is unstable, so only the VALUE is contractual."""
    s = np.asarray(scores, dtype=np.float32)
    r = quantile_rank(s.shape[0], alpha)
    return int(np.argsort(s, kind="stable")[r])
