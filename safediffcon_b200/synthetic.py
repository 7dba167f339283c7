"""Synthetic Burgers control instances (host side, numpy) for benchmarks, smoke tests and fixtures.

Draws (u0, f) from the distribution of the reference's data generator
(/root/reference/1D/data/generate_burgers.py:338-418, summarised in SURVEY.md appendix C):
u0 = sum of one positive and one negative Gaussian bump; f = sum of 8 separable space x time Gaussian
bumps (the first always on, the other seven on with probability 1/2), stepwise constant over 10 intervals.
This is an independent implementation on numpy's Generator API (the draws are NOT bit-compatible with the
reference's global-RNG sequence; only the distribution matters for synthetic workloads).
"""
import numpy as np


def burgers_instances(n, seed=0, s=128, nt=10, amp_compensate=2.0):
    """Returns u0 [n, s] float32 and f [n, nt, s] float32 (unscaled physical units)."""
    rng = np.random.default_rng(seed)
    x = (np.arange(1, s + 1, dtype=np.float64) / (s + 1))[None, :]
    tt = (np.arange(1, nt + 1, dtype=np.float64) / (nt + 1))[None, :, None]

    def bump(center, width, grid):
        return np.exp(-0.5 * (grid - center) ** 2 / width ** 2)

    u0 = rng.uniform(0, 2, (n, 1)) * bump(rng.uniform(0.2, 0.4, (n, 1)), rng.uniform(0.05, 0.15, (n, 1)), x)
    u0 = u0 + rng.uniform(-2, 0, (n, 1)) * bump(rng.uniform(0.6, 0.8, (n, 1)), rng.uniform(0.05, 0.15, (n, 1)), x)

    f = np.zeros((n, nt, s))
    for k in range(8):
        amp = rng.uniform(-1.5, 1.5, (n, 1, 1))
        if k > 0:
            amp = amp * rng.integers(0, 2, (n, 1, 1))
        sp = bump(rng.uniform(0, 1, (n, 1, 1)), rng.uniform(0.05, 0.2, (n, 1, 1)), x[:, None, :])
        tm = bump(rng.uniform(0, 1, (n, 1, 1)), rng.uniform(0.05, 0.2, (n, 1, 1)), tt)
        f += amp * sp * (amp_compensate * tm)
    return u0.astype(np.float32), f.astype(np.float32)
