"""Synthetic Burgers data on the device: drop-in for ``make_data_varying_f``
(/root/reference/1D/data/generate_burgers.py:338-418) and for the tensor assembly of ``BurgersDataset._process_data``
(/root/reference/1D/data/burgers.py:104-142) -- SURVEY.md section 8f row 2.

The random scalars (a few per instance) are drawn on the host from numpy's GLOBAL RNG in exactly the reference's order, so
``np.random.seed(k); make_data_varying_f(...)`` describes the same instances as the reference call; the float64 field
evaluation over [N, t, s] -- where the reference spends its time -- runs in a CUDA kernel (csrc/datagen.cu).
"""
import numpy as np
import torch

from . import _lib as L

TERMS = 8   # one always-on forcing bump + sum_num_f = 7 bumps that are on with probability 1/2


def draw_parameters(Nu0, Nf):
    """(params_u0 [Nu0, 6], params_f [Nf, 8, 5]) float64, consuming np.random exactly like the reference generator."""
    loc1 = np.random.uniform(0.2, 0.4, (Nu0, 1))
    amp1 = np.random.uniform(0, 2, (Nu0, 1))
    sig1 = np.random.uniform(0.05, 0.15, (Nu0, 1))
    loc2 = np.random.uniform(0.6, 0.8, (Nu0, 1))
    amp2 = np.random.uniform(-2, 0, (Nu0, 1))
    sig2 = np.random.uniform(0.05, 0.15, (Nu0, 1))
    pu = np.concatenate([loc1, amp1, sig1, loc2, amp2, sig2], axis=1)
    pf = np.empty((Nf, TERMS, 5), dtype=np.float64)
    for k in range(TERMS):
        if k > 0:
            amp = np.random.randint(2, size=(Nf, 1, 1)) * np.random.uniform(-1.5, 1.5, (Nf, 1, 1))
        else:
            amp = np.random.uniform(-1.5, 1.5, (Nf, 1, 1))
        loc_x = np.random.uniform(0, 1, (Nf, 1, 1))
        sig_x = np.random.uniform(0.1, 0.4, (Nf, 1, 1)) * 0.5
        loc_t = np.random.uniform(0, 1, (Nf, 1, 1))
        sig_t = np.random.uniform(0.1, 0.4, (Nf, 1, 1)) * 0.5
        pf[:, k] = np.concatenate([amp, loc_x, sig_x, loc_t, sig_t], axis=1)[:, :, 0]
    return pu, pf


def make_data_varying_f(Nu0, Nf, s, t, amp_compensate=2, partial_control=None, alpha=1., tmax=1., device="cuda"):
    """-> (u0 [Nu0, s] float64, f [Nf, t, s] float32) CUDA tensors (the reference returns a numpy float64 array and a CPU
    float32 tensor with the same values)."""
    if not torch.cuda.is_available():
        raise RuntimeError("safediffcon_b200.make_data_varying_f: CUDA (sm_100a) only, no CPU fallback")
    delta_x = 1.0 / (s + 1)
    x = torch.linspace(0.0 + delta_x, 1.0 - delta_x, s)
    delta_t = (tmax - 0.0) / (t + 1)
    ts = torch.linspace(0.0 + delta_t, tmax - delta_t, t)
    if partial_control is None:
        mode = 0
    elif partial_control == 'front_rear_quarter':
        mode = 1
        amp_compensate *= 2
        print('Generating in partial control mode:', partial_control)
    else:
        raise ValueError('invalid partial control mode')
    pu, pf = draw_parameters(Nu0, Nf)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        pu_d, pf_d = torch.from_numpy(pu).to(dev), torch.from_numpy(pf).to(dev)
        x_d, t_d = x.to(dev), ts.to(dev)
        u0 = torch.empty(Nu0, s, device=dev, dtype=torch.float64)
        f = torch.empty(Nf, t, s, device=dev, dtype=torch.float32)
        L.check(L.lib().sdc_burgers_fields(L.ptr(pu_d), L.ptr(pf_d), L.ptr(x_d), L.ptr(t_d), L.ptr(u0), None, L.ptr(f), Nu0, Nf, s, t,
                                           TERMS, float(amp_compensate), mode, float(alpha), L.stream_ptr()))
    return u0, f


def dataset_states(u_traj, f, pad=16, scaler=10.0, use_max_safety=True):
    """Model-space states [N, 3, pad, s] = (u, f, safety)/scaler on the device from rollouts u_traj [N, nt+1, s] and controls
    f [N, nt, s] (BurgersDataset._process_data with stack_u_and_f, pad_for_2d_conv, is_normalize)."""
    u = L.dev_f32(u_traj, "u_traj")
    ff = L.dev_f32(f, "f")
    N, nt1, s = u.shape
    assert ff.shape[0] == N and ff.shape[2] == s
    out = torch.empty(N, 3, pad, s, device=u.device, dtype=torch.float32)
    with torch.cuda.device(u.device):
        L.check(L.lib().sdc_dataset_states(L.ptr(u), L.ptr(ff), L.ptr(out), N, nt1, ff.shape[1], pad, s, float(scaler),
                                           int(bool(use_max_safety)), L.stream_ptr()))
    return out
