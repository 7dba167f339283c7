"""Scoring of sampled controls: drop-in for /root/reference/1D/utils/metrics.py (same names, keys, shapes)."""
from typing import Dict

import torch

from .solver import burgers_numeric_solve_free, burgers_score


def control_trajectories(diffused: torch.Tensor, nt: int) -> torch.Tensor:
    """diffused: (batch, channels, padded_time, space) unscaled -> solver rollout (batch, nt, space).
    Reference: utils/metrics.py:42-65."""
    return burgers_numeric_solve_free(diffused[:, 0, 0, :], diffused[:, 1, :nt - 1, :], visc=0.01, T=1.0, dt=1e-4, num_t=10)


def _reduce(J, pts, tms, flg, nt1, s) -> Dict[str, float]:
    n = flg.shape[0]
    flag = flg.to(torch.bool)
    m = {}
    if J is not None:
        m['control_mse_mean (J)'] = J.mean().item()
        m['control_mse_std'] = J.std().item()
    # exact integer counts -> the same ratios the reference obtains from mask.float().mean()
    m['point_exceed_ratio (R_p)'] = (pts.sum(dtype=torch.int64).to(torch.float64) / (n * nt1 * s)).to(torch.float32).item()
    m['time_exceed_ratio (R_t)'] = (tms.sum(dtype=torch.int64).to(torch.float64) / (n * nt1)).to(torch.float32).item()
    m['sample_exceed_ratio (R_s)'] = flag.float().mean().item()
    m['sample_excedd_indices'] = flag.nonzero(as_tuple=True)[0].tolist()
    return m


def calculate_safety_metrics(u: torch.Tensor, threshold: float, diffused_s: torch.Tensor,
                             use_max_safety: bool = True) -> Dict[str, float]:
    """Exceed ratios of |u| > threshold by point / time row / sample (reference utils/metrics.py:67-94)."""
    _, pts, tms, flg = burgers_score(u, None, threshold)
    return _reduce(None, pts, tms, flg, u.shape[1], u.shape[2])


def evaluate_samples(diffused: torch.Tensor, u_controlled: torch.Tensor, u_target: torch.Tensor, nt: int, u_bound: float,
                     use_max_safety: bool = True) -> Dict[str, float]:
    """J = MSE of the final controlled state vs the target + safety ratios (reference utils/metrics.py:8-40)."""
    J, pts, tms, flg = burgers_score(u_controlled, u_target[:, -1, :], u_bound)
    return _reduce(J, pts, tms, flg, u_controlled.shape[1], u_controlled.shape[2])


def calculate_safety_score(u: torch.Tensor) -> torch.Tensor:
    return (u.square()).amax((-1, -2))
