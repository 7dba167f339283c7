"""Burgers rollout: drop-in for the reference solver entry points, executed by one CUDA launch.

Mirrors /root/reference/1D/data/generate_burgers.py:113-205 (``burgers_numeric_solve``, Cartesian Nu0 x Nf)
and :207-299 (``burgers_numeric_solve_free``, paired) -- same argument names, shapes and assertion messages.
"""
import torch

from . import _lib as L


def _check_mode(f, num_t, mode, allow_const):
    if mode != 'const':
        assert f.size()[1] == num_t, 'check number of time interval'
        return f
    if not allow_const:
        raise ValueError
    return f.unsqueeze(1).repeat(1, num_t, 1)


def burgers_numeric_solve_free(u0, f, visc, T, dt=1e-4, num_t=10, mode=None, strict=True):
    """u0: (N, s), f: (N, num_t, s) -> trajectories (N, num_t + 1, s); row 0 is u0."""
    f = _check_mode(f, num_t, mode, allow_const=False)
    assert u0.size(0) == f.size(0)
    u0c, fc = L.dev_f32(u0, "u0"), L.dev_f32(f, "f")
    N, s = u0c.shape
    out = torch.empty(N, num_t + 1, s, device=u0c.device, dtype=torch.float32)
    with torch.cuda.device(u0c.device):
        L.check(L.lib().sdc_burgers_solve_free(L.ptr(u0c), L.ptr(fc), L.ptr(out), N, s, num_t, float(visc), float(T),
                                               float(dt), int(strict), L.stream_ptr()))
    return out


def burgers_numeric_solve(u0, f, visc, T, dt=1e-4, num_t=10, mode=None, strict=True):
    """u0: (Nu0, s), f: (Nf, num_t, s) [or (Nf, s) with mode='const'] -> (Nu0, Nf, num_t + 1, s)."""
    f = _check_mode(f, num_t, mode, allow_const=True)
    u0c, fc = L.dev_f32(u0, "u0"), L.dev_f32(f, "f")
    Nu0, s = u0c.shape
    Nf = fc.shape[0]
    out = torch.empty(Nu0, Nf, num_t + 1, s, device=u0c.device, dtype=torch.float32)
    with torch.cuda.device(u0c.device):
        L.check(L.lib().sdc_burgers_solve_cartesian(L.ptr(u0c), L.ptr(fc), L.ptr(out), Nu0, Nf, s, num_t, float(visc),
                                                    float(T), float(dt), int(strict), L.stream_ptr()))
    return out


def burgers_score(traj, target_final, u_bound):
    """Per-trajectory J and exceed counters of a rollout [N, nt1, s] (device tensors)."""
    tr = L.dev_f32(traj, "traj")
    N, nt1, s = tr.shape
    tf = L.dev_f32(target_final, "target_final") if target_final is not None else None
    dev = tr.device
    J = torch.empty(N, device=dev, dtype=torch.float32) if tf is not None else None
    pts = torch.empty(N, device=dev, dtype=torch.int32)
    tms = torch.empty(N, device=dev, dtype=torch.int32)
    flg = torch.empty(N, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        L.check(L.lib().sdc_burgers_score(L.ptr(tr), L.ptr(tf), float(u_bound), N, nt1, s, L.ptr(J), L.ptr(pts), L.ptr(tms),
                                          L.ptr(flg), L.stream_ptr()))
    return J, pts, tms, flg


def control_and_score(diffused, target_final, u_bound, nt=11, visc=0.01, T=1.0, dt=1e-4, want_traj=True, strict=True):
    """Fused control_trajectories + per-sample scoring of an UNSCALED model output [N, 3, pad, s]."""
    d = L.dev_f32(diffused, "diffused")
    N, C, pad, s = d.shape
    assert C == 3
    dev = d.device
    tf = L.dev_f32(target_final, "target_final") if target_final is not None else None
    out = torch.empty(N, nt, s, device=dev, dtype=torch.float32) if want_traj else None
    J = torch.empty(N, device=dev, dtype=torch.float32) if tf is not None else None
    pts = torch.empty(N, device=dev, dtype=torch.int32)
    tms = torch.empty(N, device=dev, dtype=torch.int32)
    flg = torch.empty(N, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        L.check(L.lib().sdc_burgers_control_score(L.ptr(d), pad, L.ptr(tf), float(u_bound), L.ptr(out), N, s, nt - 1,
                                                  float(visc), float(T), float(dt), int(strict), L.ptr(J), L.ptr(pts),
                                                  L.ptr(tms), L.ptr(flg), L.stream_ptr()))
    return out, J, pts, tms, flg


def nonfinite_rollouts(reset=True):
    """Diagnostics: how many rollouts on the current device ended in a non-finite state since the last reset (synchronises)."""
    torch.cuda.synchronize()
    n = L.c_i64(0)
    import ctypes
    L.check(L.lib().sdc_burgers_nonfinite_rollouts(int(reset), ctypes.byref(n)))
    return int(n.value)
