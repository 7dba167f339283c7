"""Conformal calibration: drop-in for ConformalCalculator (/root/reference/1D/inference/conformal.py:11-118; the
posttrain copy differs only by the optional InfFT_Q reweight, which is honoured when the config has it)."""
import logging
from typing import Tuple

import numpy as np
import torch

from . import _lib as L
from .guidance import _gstruct, SCALER


def kth_select(scores, rank):
    """(value, index) of the rank-th order statistic of a fp32 CUDA vector, computed on the device.
    value is bit-exact; index is the one a stable ascending sort would select."""
    s = L.dev_f32(scores, "scores")
    n = s.shape[0]
    val = torch.empty((), device=s.device, dtype=torch.float32)
    idx = torch.empty((), device=s.device, dtype=torch.int64)
    ws = torch.empty(int(L.lib().sdc_kth_select_workspace(n)), device=s.device, dtype=torch.uint8)
    with torch.cuda.device(s.device):
        L.check(L.lib().sdc_kth_select(L.ptr(s), n, int(rank), L.ptr(val), L.ptr(idx), L.ptr(ws), L.stream_ptr()))
    return val, idx


def quantile_rank(n, alpha):
    return min(int(np.ceil(alpha * (n + 1))), n) - 1  # 'n-1' to avoid the worst case (reference conformal.py:112)


def scores_and_weights(pred, state, config, Q):
    """Per-sample nonconformity score |red(10 pred_s) - red(10 state_s)| and raw importance weight."""
    p, s = L.dev_f32(pred, "pred"), L.dev_f32(state, "state")
    B, C, H, W = s.shape
    score = torch.empty(B, device=s.device, dtype=torch.float32)
    weight = torch.empty(B, device=s.device, dtype=torch.float32)
    q2 = getattr(config, "InfFT_Q", None)
    q2 = float("inf") if q2 is None else float(q2)
    Qf = Q.item() if isinstance(Q, torch.Tensor) else float(Q)
    with torch.cuda.device(s.device):
        L.check(L.lib().sdc_conformal_scores(L.ptr(p), L.ptr(s), L.ptr(score), L.ptr(weight), _gstruct(config, Qf), q2, B, H, W,
                                             L.stream_ptr()))
    return score, weight


class ConformalCalculator:
    """Class for calculating conformal scores and quantiles"""

    def __init__(self, model, config):
        self.model = model
        self.config = config
        self.device = config.device

    def get_conformal_scores(self, dataloader, Q: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        scores, weights, states = [], [], []
        logging.info("===Start calculating conformal scores...")
        for i in range(self.config.num_cal_batch):
            logging.info(f"====Calculate {i}-th Batch in Calibration set")
            state = next(dataloader)
            states.append(state)
            state = state.to(self.device)
            with torch.no_grad():
                output = self.model.sample(
                    batch_size=state.shape[0], clip_denoised=True, guidance_u0=False, device=self.device,
                    u_init=state[:, 0, 0, :], u_final=state[:, 0, self.config.nt - 1, :], w_groundtruth=state[:, 1, :, :],
                    nablaJ=None, J_scheduler=None, w_scheduler=None, enable_grad=False)
            sc, w = scores_and_weights(output, state, self.config, Q)
            scores.append(sc)
            weights.append(w)
        weights = torch.cat(weights)
        sc = torch.cat(scores)
        out = torch.empty_like(weights)
        with torch.cuda.device(weights.device):
            L.check(L.lib().sdc_normalize_weights(L.ptr(weights), L.ptr(out), L.ptr(sc), weights.shape[0], L.stream_ptr()))
        return sc, out, torch.cat(states)

    def calculate_quantile(self, scores: torch.Tensor, weights: torch.Tensor, states: torch.Tensor, alpha: float) -> torch.Tensor:
        n = scores.shape[0]
        rank = quantile_rank(n, alpha)
        quantile, _ = kth_select(scores, rank)
        logging.info(f"===Calculate {alpha}-th quantile, No.{rank}")
        return quantile
