"""Oracle (TEST INFRASTRUCTURE): the synthetic Burgers generator and dataset-state assembly on CPU.

Restates /root/reference/1D/data/generate_burgers.py:338-418 (make_data_varying_f: two-Gaussian u0, forcing = one always-on
plus seven coin-flipped space-time Gaussian bumps, float64 evaluation, float32 cast, optional alpha scaling with a +-10
clamp) and /root/reference/1D/data/burgers.py:104-142 (BurgersDataset._process_data: safety = u^2 or its per-sample maximum,
zero padding of the time axis to 16, division by the scaler).  The random scalars come from numpy's GLOBAL RNG in the
reference's order, so np.random.seed(k) selects the same instances.  Pinned against the unmodified reference by
tests/golden/datagen.npz and tests/golden/dataset_states.npz (bit-exact: same numpy, same operation order).
"""
import numpy as np
import torch


def make_data_varying_f(Nu0, Nf, s, t, amp_compensate=2, partial_control=None, alpha=1., tmax=1.):
    """-> (u0 [Nu0, s] float64 ndarray, f [Nf, t, s] float32 tensor)."""
    dx = 1.0 / (s + 1)
    x = torch.linspace(0.0 + dx, 1.0 - dx, s).numpy().astype(np.float64)          # float32 nodes promoted, as np.array(x) - loc does
    dt = (tmax - 0.0) / (t + 1)
    ts = torch.linspace(0.0 + dt, tmax - dt, t).numpy().astype(np.float64)

    def bump(lo, hi):
        loc = np.random.uniform(lo, hi, (Nu0, 1))
        amp = np.random.uniform(0, 2, (Nu0, 1)) if lo < 0.5 else np.random.uniform(-2, 0, (Nu0, 1))
        sig = np.random.uniform(0.05, 0.15, (Nu0, 1))
        return amp * np.exp(-0.5 * (x[None, :] - loc) ** 2 / sig ** 2)

    g1 = bump(0.2, 0.4)
    g2 = bump(0.6, 0.8)
    u0 = g1 + g2

    if partial_control is None:
        mask = np.ones(s)
    elif partial_control == 'front_rear_quarter':
        mask = np.zeros(s)
        mask[:s // 4] = 1.
        mask[3 * s // 4:] = 1.
        amp_compensate *= 2
    else:
        raise ValueError('invalid partial control mode')

    def rand_f(is_rand_amp):
        if is_rand_amp:
            amp = np.random.randint(2, size=(Nf, 1, 1)) * np.random.uniform(-1.5, 1.5, (Nf, 1, 1))
        else:
            amp = np.random.uniform(-1.5, 1.5, (Nf, 1, 1))
        loc = np.random.uniform(0, 1, (Nf, 1, 1))
        sig = np.random.uniform(0.1, 0.4, (Nf, 1, 1)) * 0.5
        exp_space = np.exp(-0.5 * (x[None, None, :] - loc) ** 2 / sig ** 2) * mask[None, None, :]
        loc = np.random.uniform(0, 1, (Nf, 1, 1))
        sig = np.random.uniform(0.1, 0.4, (Nf, 1, 1)) * 0.5
        exp_time = amp_compensate * np.exp(-0.5 * (ts[None, :, None] - loc) ** 2 / sig ** 2)
        return (amp * exp_space) * exp_time

    f = rand_f(False)
    for _ in range(7):
        f = f + rand_f(True)
    f = torch.from_numpy(f).to(torch.float32)
    if alpha != 1.:
        f = (f * alpha).clamp(-10., 10.)
    return u0, f


def dataset_states(u_traj, f, pad=16, scaler=10.0, use_max_safety=True):
    """[N, 3, pad, s] states from rollouts [N, nt+1, s] and controls [N, nt, s] (float32 tensors)."""
    u = u_traj.to(torch.float32)
    ff = f.to(torch.float32)
    N, nt1, s = u.shape
    nt = ff.shape[1]
    safety = u * u
    if use_max_safety:
        safety = safety.amax(dim=(1, 2), keepdim=True).expand_as(u)
    out = torch.zeros(N, 3, pad, s, dtype=torch.float32)
    out[:, 0, :nt1] = u
    out[:, 1, :nt] = ff
    out[:, 2, :nt1] = safety
    return out / scaler
