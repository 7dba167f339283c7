"""Oracle (TEST INFRASTRUCTURE): Burgers rollout + scoring on CPU.

Two restatements of /root/reference/1D/data/generate_burgers.py:113-299:
  * ``solve_free_c`` / ``solve_cartesian_c`` -- ctypes binding of oracle/burgers_ref.c (strict fp32, no FMA);
    the bit-exact checker for the CUDA stencil at any N.
  * ``solve_free_torch`` -- the same loop as batched torch-CPU tensor ops (what the reference executes: 10,000
    python iterations of whole-batch elementwise ops); used as the *timed CPU baseline* in bench.py.
Metrics follow /root/reference/1D/utils/metrics.py:8-94 (same dictionary keys).
Pinned against the unmodified reference by tests/golden/solver_free.npz, solver_cartesian.npz, metrics.npz.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force=False):
    so = os.path.join(_HERE, "liboracle_burgers.so")
    src = os.path.join(_HERE, "burgers_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        _LIB.oracle_burgers_solve_free.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                                          ctypes.c_double, ctypes.c_double, ctypes.c_double]
        _LIB.oracle_burgers_solve_free.restype = None
        _LIB.oracle_burgers_score.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int64,
                                              ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
        _LIB.oracle_burgers_score.restype = None
    return _LIB


def solve_free_c(u0, f, visc=0.01, T=1.0, dt=1e-4, num_t=10):
    u0 = np.ascontiguousarray(np.asarray(u0, dtype=np.float32))
    f = np.ascontiguousarray(np.asarray(f, dtype=np.float32))
    N, s = u0.shape
    assert f.shape == (N, num_t, s), "check number of time interval"
    out = np.empty((N, num_t + 1, s), dtype=np.float32)
    _lib().oracle_burgers_solve_free(u0.ctypes.data, f.ctypes.data, out.ctypes.data, N, s, num_t, visc, T, dt)
    return out


def solve_cartesian_c(u0, f, visc=0.01, T=1.0, dt=1e-4, num_t=10):
    u0 = np.asarray(u0, dtype=np.float32)
    f = np.asarray(f, dtype=np.float32)
    Nu0, s = u0.shape
    Nf = f.shape[0]
    uu = np.repeat(u0[:, None, :], Nf, axis=1).reshape(Nu0 * Nf, s)
    ff = np.repeat(f[None], Nu0, axis=0).reshape(Nu0 * Nf, num_t, s)
    return solve_free_c(uu, ff, visc, T, dt, num_t).reshape(Nu0, Nf, num_t + 1, s)


def score_c(traj, target_final, u_bound):
    traj = np.ascontiguousarray(np.asarray(traj, dtype=np.float32))
    tf = np.ascontiguousarray(np.asarray(target_final, dtype=np.float32))
    N, nt1, s = traj.shape
    J = np.empty(N, np.float32)
    pts = np.empty(N, np.int32)
    tms = np.empty(N, np.int32)
    flg = np.empty(N, np.int32)
    _lib().oracle_burgers_score(traj.ctypes.data, tf.ctypes.data, u_bound, N, nt1, s, J.ctypes.data, pts.ctypes.data,
                                tms.ctypes.data, flg.ctypes.data)
    return J, pts, tms, flg


def solve_free_torch(u0, f, visc=0.01, T=1.0, dt=1e-4, num_t=10):
    """Whole-batch torch-CPU loop (the reference's execution model)."""
    N, s = u0.shape
    dx = 1.0 / (s + 1)
    steps = math.ceil(T / dt)
    rec = steps // num_t
    a = torch.tensor(np.float32(1.0 / (2 * dx)))
    na = torch.tensor(np.float32(-1.0 / (2 * dx)))
    d = torch.tensor(np.float32(visc * 1.0 / dx ** 2))
    d2 = torch.tensor(np.float32(visc * -2.0 / dx ** 2))
    u = torch.zeros(N, s + 2)
    u[:, 1:-1] = u0
    fp = torch.zeros(N, num_t, s + 2)
    fp[:, :, 1:-1] = f
    out = torch.empty(N, num_t + 1, s)
    out[:, 0] = u0
    c = 0
    for j in range(steps):
        us = u * u
        tr = na * us[:, :-2] + a * us[:, 2:]
        di = (d * u[:, :-2] + d2 * u[:, 1:-1]) + d * u[:, 2:]
        k = min(j // rec, num_t - 1)
        new = u[:, 1:-1] + dt * ((-0.5 * tr + di) + fp[:, k, 1:-1])
        u = torch.zeros_like(u)
        u[:, 1:-1] = new
        if (j + 1) % rec == 0 and c < num_t:
            out[:, c + 1] = new
            c += 1
    return out


def control_inputs(diffused, nt=11):
    """(u0, f) slices the reference's control_trajectories feeds the solver (utils/metrics.py:52-53)."""
    return diffused[:, 0, 0, :], diffused[:, 1, : nt - 1, :]


def evaluate(u_controlled, u_target, u_bound):
    """Metric dictionary with the reference's key names (utils/metrics.py:32-92)."""
    uc = torch.as_tensor(u_controlled)
    ut = torch.as_tensor(u_target)
    mse = (ut[:, -1, :] - uc[:, -1, :]).square().mean(-1)
    ex = uc.abs() > u_bound
    per_sample = ex.any(dim=(-1, -2))
    return {
        "control_mse_mean (J)": mse.mean().item(),
        "control_mse_std": mse.std().item(),
        "point_exceed_ratio (R_p)": ex.float().mean().item(),
        "time_exceed_ratio (R_t)": ex.any(dim=-1).float().mean().item(),
        "sample_exceed_ratio (R_s)": per_sample.float().mean().item(),
        "sample_excedd_indices": per_sample.nonzero(as_tuple=True)[0].tolist(),
    }
