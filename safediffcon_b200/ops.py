"""``torch.ops.safediffcon_b200.*``: the public CUDA entry points registered as PyTorch custom operators (torch.library).

Thin layer over the ctypes binding of the C ABI (include/*.h): each operator has a CUDA implementation only -- calling one with
CPU tensors raises NotImplementedError from the dispatcher (there is no CPU fallback) -- and a fake (meta) implementation so the
ops trace under torch.compile / FakeTensor without running.  The Python classes (GaussianDiffusion, Unet2D, ...) call the same C
functions directly; these registrations are for callers that want dispatcher-visible ops (profiling by op name, export, compile).

  burgers_solve_free(u0 [N,s], f [N,nt,s], visc, T, dt, strict) -> traj [N,nt+1,s]        /root/reference/1D/data/generate_burgers.py:207-299
  burgers_solve_cartesian(u0 [Nu,s], f [Nf,nt,s], ...) -> [Nu,Nf,nt+1,s]                   generate_burgers.py:113-205
  burgers_control_score(diffused [N,3,pad,s], target_final [N,s], u_bound, nt) -> (traj, J, points, times, flag)   utils/metrics.py:42-94
  kth_select(scores [n], rank) -> (value [], index [])                                     inference/conformal.py:95-118
  unet_forward(x [B,C,H,W], t [B] int32, handle) -> eps                                   model/unet.py:382-426 (handle = UnetPlan)
"""
import torch

from . import _lib as L
from . import conformal as _conformal
from . import solver as _solver

NS = "safediffcon_b200"


@torch.library.custom_op(f"{NS}::burgers_solve_free", mutates_args=(), device_types="cuda")
def burgers_solve_free(u0: torch.Tensor, f: torch.Tensor, visc: float, T: float, dt: float, strict: bool) -> torch.Tensor:
    return _solver.burgers_numeric_solve_free(u0, f, visc, T, dt=dt, num_t=f.shape[1], strict=strict)


@burgers_solve_free.register_fake
def _(u0, f, visc, T, dt, strict):
    return u0.new_empty(u0.shape[0], f.shape[1] + 1, u0.shape[1], dtype=torch.float32)


@torch.library.custom_op(f"{NS}::burgers_solve_cartesian", mutates_args=(), device_types="cuda")
def burgers_solve_cartesian(u0: torch.Tensor, f: torch.Tensor, visc: float, T: float, dt: float, strict: bool) -> torch.Tensor:
    return _solver.burgers_numeric_solve(u0, f, visc, T, dt=dt, num_t=f.shape[1], strict=strict)


@burgers_solve_cartesian.register_fake
def _(u0, f, visc, T, dt, strict):
    return u0.new_empty(u0.shape[0], f.shape[0], f.shape[1] + 1, u0.shape[1], dtype=torch.float32)


@torch.library.custom_op(f"{NS}::burgers_control_score", mutates_args=(), device_types="cuda")
def burgers_control_score(diffused: torch.Tensor, target_final: torch.Tensor, u_bound: float,
                          nt: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    traj, J, pts, tms, flg = _solver.control_and_score(diffused, target_final, u_bound, nt=nt, want_traj=True)
    return traj, J, pts, tms, flg


@burgers_control_score.register_fake
def _(diffused, target_final, u_bound, nt):
    N, s = diffused.shape[0], diffused.shape[3]
    i32 = lambda: diffused.new_empty(N, dtype=torch.int32)  # noqa: E731
    return diffused.new_empty(N, nt, s), diffused.new_empty(N), i32(), i32(), i32()


@torch.library.custom_op(f"{NS}::kth_select", mutates_args=(), device_types="cuda")
def kth_select(scores: torch.Tensor, rank: int) -> tuple[torch.Tensor, torch.Tensor]:
    return _conformal.kth_select(scores, rank)


@kth_select.register_fake
def _(scores, rank):
    return scores.new_empty(()), scores.new_empty((), dtype=torch.int64)


@torch.library.custom_op(f"{NS}::unet_forward", mutates_args=(), device_types="cuda")
def unet_forward(x: torch.Tensor, t: torch.Tensor, handle: int) -> torch.Tensor:
    """eps = Unet2D(x, t) through the C++ executor; `handle` = id of a live UnetPlan registered with register_plan()."""
    plan = _PLANS[handle]
    with torch.cuda.device(x.device):
        return plan.forward(L.dev_f32(x, "x"), t.to(torch.int32).contiguous())


@unet_forward.register_fake
def _(x, t, handle):
    return x.new_empty(x.shape[0], _PLANS[handle].out_dim, x.shape[2], x.shape[3])


_PLANS = {}


def register_plan(net) -> int:
    """Make the packed executor of `net` (a safediffcon_b200.Unet2D on a CUDA device) addressable by torch.ops.safediffcon_b200.unet_forward."""
    plan = net._plan_ready()
    if plan is None:
        raise RuntimeError("safediffcon_b200: this Unet2D configuration has no C++ executor (see Unet2D._plan_ready)")
    _PLANS[id(plan)] = plan
    return id(plan)
