"""Data-parallel driver of the hot path: shard control instances / calibration trajectories over ranks (one process
per GPU), run the chains and rollouts locally, all-gather the per-sample results.

The path shards over independent units; the ONLY collective is an all-gather of small per-sample vectors
(nonconformity scores + weights after calibration; J and exceed counters after evaluation).  Every rank then
holds the full vectors in global index order and derives identical scalars (quantile Q, metrics), so the
result does not depend on the number of ranks.  Call sites it stands in for: InferenceFT.inference /
evaluate_model / calibrate (/root/reference/1D/inference/inference_ft.py:263-347).
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L
from .conformal import kth_select, quantile_rank, scores_and_weights
from .guidance import SCALER, safety_guidance
from .solver import control_and_score


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous shard [lo, hi) of n units for `rank`; the first n % world_size ranks take one extra unit."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_concat(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather variable-length per-rank vectors (dim 0) into global index order.  Works with NCCL (CUDA
    tensors) and gloo (CPU tensors); with one rank it is the identity."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def metrics_from_vectors(J: Optional[torch.Tensor], pts: torch.Tensor, tms: torch.Tensor, flag: torch.Tensor, nt1: int,
                         s: int) -> Dict[str, float]:
    """The reference's metric dictionary (utils/metrics.py:32-92) from gathered per-sample vectors."""
    n = flag.shape[0]
    f = flag.to(torch.bool)
    m = {}
    if J is not None:
        m['control_mse_mean (J)'] = J.mean().item()
        m['control_mse_std'] = J.std().item() if n > 1 else float('nan')
    m['point_exceed_ratio (R_p)'] = (pts.sum(dtype=torch.int64).double() / (n * nt1 * s)).float().item()
    m['time_exceed_ratio (R_t)'] = (tms.sum(dtype=torch.int64).double() / (n * nt1)).float().item()
    m['sample_exceed_ratio (R_s)'] = f.float().mean().item()
    m['sample_excedd_indices'] = f.nonzero(as_tuple=True)[0].tolist()
    return m


def shared_seed(seed, device):
    """One Philox seed for the whole job: an explicit seed is used as is; otherwise rank 0 draws one from torch's CPU RNG and
    broadcasts it, so ranks that all called torch.manual_seed(s) do not each re-derive the SAME stream with offset 0."""
    rank, ws = world()
    if seed is not None:
        return int(seed)
    if ws == 1:
        return int(torch.randint(0, 2 ** 62, (1,)).item())
    t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
    t = t.to(device) if dist.get_backend() == "nccl" else t
    dist.broadcast(t, src=0)
    return int(t.item())


def default_offset(sample_offset, n_local, n_total):
    """Global index of this rank's first unit: explicit value, else the start of its contiguous shard of n_total (when the
    shard sizes follow shard_range), else rank * n_local.  The in-kernel noise is keyed by (seed, global index, t), which
    makes the samples independent of the number of ranks."""
    if sample_offset is not None:
        return int(sample_offset)
    rank, ws = world()
    if ws == 1:
        return 0
    if n_total is not None:
        lo, hi = shard_range(n_total, rank, ws)
        if hi - lo == n_local:
            return lo
    return rank * n_local


def sample_controls(model, u_init, u_final, config=None, Q=None, n_total=None, sample_offset=None, seed=None, noise=None,
                    w_groundtruth=None, guidance_u0=True):
    """Guided reverse chain for the local shard; returns the UNSCALED prediction (what InferenceFT.inference returns).
    u_init / u_final may be host (pinned) or device tensors in model units."""
    dev = model.betas.device
    seed = shared_seed(seed, dev) if noise is None else 0
    sample_offset = default_offset(sample_offset, u_init.shape[0], n_total)
    u0 = u_init.to(dev, non_blocking=True)
    uT = u_final.to(dev, non_blocking=True)
    nabla = safety_guidance(config, Q) if (config is not None and Q is not None and guidance_u0) else None
    out = model.sample(batch_size=u0.shape[0], clip_denoised=True, u_init=u0, u_final=uT, guidance_u0=guidance_u0, nablaJ=nabla,
                       J_scheduler=None, w_scheduler=None, enable_grad=False, device=dev, seed=seed, sample_offset=sample_offset,
                       noise=noise, w_groundtruth=None if w_groundtruth is None else w_groundtruth.to(dev, non_blocking=True))
    return out * SCALER


def evaluate_controls(pred_unscaled, target_final, u_bound, n_total=None, nt=11, want_traj=False):
    """Rollout + scoring of the local shard in one launch, then ONE all-gather of (J, points, times, flag)."""
    dev = pred_unscaled.device
    tf = target_final.to(dev, non_blocking=True)
    traj, J, pts, tms, flg = control_and_score(pred_unscaled, tf, u_bound, nt=nt, want_traj=want_traj)
    n_total = n_total if n_total is not None else J.shape[0]
    packed = torch.stack([J, pts.float(), tms.float(), flg.float()], dim=1)  # one collective for all four vectors
    g = all_gather_concat(packed, n_total)
    m = metrics_from_vectors(g[:, 0], g[:, 1].long(), g[:, 2].long(), g[:, 3].long(), nt, pred_unscaled.shape[-1])
    return m, traj


def calibrate_quantile(model, states, config, Q, alpha, n_total=None, sample_offset=None, seed=None, noise=None):
    """Nonconformity scores of the local calibration shard (unguided chain clamped to the ground-truth control,
    inference/conformal.py:53-85) -> all-gather of (score, raw weight) -> weights normalised over the FULL
    vector in index order -> rank-th order statistic on the device.  Returns (quantile 0-d tensor, scores, weights)."""
    dev = model.betas.device
    seed = shared_seed(seed, dev) if noise is None else 0
    sample_offset = default_offset(sample_offset, states.shape[0], n_total)
    st = states.to(dev, non_blocking=True)
    pred = model.sample(batch_size=st.shape[0], clip_denoised=True, guidance_u0=False, u_init=st[:, 0, 0, :],
                        u_final=st[:, 0, config.nt - 1, :], w_groundtruth=st[:, 1, :, :], nablaJ=None, J_scheduler=None,
                        w_scheduler=None, enable_grad=False, device=dev, seed=seed, sample_offset=sample_offset, noise=noise)
    score, weight = scores_and_weights(pred, st, config, Q)
    n_total = n_total if n_total is not None else score.shape[0]
    g = all_gather_concat(torch.stack([score, weight], dim=1), n_total)
    sc, w = g[:, 0].contiguous(), g[:, 1].contiguous()
    wn = torch.empty_like(w)
    L.check(L.lib().sdc_normalize_weights(L.ptr(w), L.ptr(wn), L.ptr(sc), n_total, L.stream_ptr()))
    q, _ = kth_select(sc, quantile_rank(n_total, alpha))
    return q, sc, wn
