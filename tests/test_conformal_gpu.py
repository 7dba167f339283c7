"""GPU parity: guidance gradient, importance weights, normalisation, nonconformity scores, k-th selection."""
import types

import numpy as np
import pytest
import torch

from oracle import conformal_ref as cr
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu


def _cfg(ums=True, w=500.0, **kw):
    return types.SimpleNamespace(use_max_safety=ums, u_bound=0.8, guidance_weights={"w_score": w}, **kw)


def test_guidance_gradient_and_weights_vs_reference(golden):
    import safediffcon_b200 as s
    g = golden("guidance")
    x = fx.guidance_states(6).cuda()
    for Q in (0.0, 0.05, -0.5):
        for ums in (True, False):
            grad = s.get_finetune_guidance(_cfg(ums), x, Q)
            assert np.allclose(grad.cpu().numpy(), g[f"grad_Q{Q}_{int(ums)}"], rtol=1e-6, atol=0)
            w = s.get_weight(x, Q, _cfg(ums))
            ref = g[f"weight_Q{Q}_{int(ums)}"]
            assert np.allclose(w.cpu().numpy(), ref, rtol=2e-4, atol=1e-38), (Q, ums, w, ref)  # exp of O(10) args
            gv = s.calculate_guidance(x, Q, _cfg(ums))
            assert np.allclose(np.exp(-gv.cpu().numpy().astype(np.float64)), ref, rtol=2e-4, atol=1e-38)


def test_normalize_weights_vs_reference(golden):
    import safediffcon_b200 as s
    g = golden("guidance")
    for i, w in enumerate(fx.weight_vectors()):
        wd = w.clone().cuda()
        out = s.normalize_weights(wd)
        assert np.allclose(out.cpu().numpy(), g[f"norm_{i}"], rtol=1e-6, atol=0)
        if torch.isinf(w).any():  # in-place replacement of inf by the largest finite weight
            assert not torch.isinf(wd).any() and wd.max().item() == w[~torch.isinf(w)].max().item()


@pytest.mark.parametrize("i", range(6))
def test_kth_select_bit_exact(i, golden):
    from safediffcon_b200.conformal import kth_select, quantile_rank, ConformalCalculator
    s, alpha = fx.score_vectors()[i]
    val, idx = kth_select(s.cuda(), quantile_rank(len(s), alpha))
    assert val.item() == float(golden("guidance")[f"quant_{i}"])        # bit-exact value
    assert idx.item() == cr.quantile_index(s, alpha)                     # stable-sort tie-break
    cc = ConformalCalculator(None, types.SimpleNamespace(device="cuda"))
    assert cc.calculate_quantile(s.cuda(), None, None, alpha).item() == float(golden("guidance")[f"quant_{i}"])


def test_kth_select_edge_cases():
    from safediffcon_b200.conformal import kth_select
    rng = np.random.default_rng(3)
    v = rng.normal(size=50000).astype(np.float32)
    v[::7] = -v[::7]
    v[5] = 0.0
    v[9] = -0.0
    v[11] = np.inf
    v[13] = -np.inf
    v[17] = np.nan
    order = np.argsort(v, kind="stable")  # numpy, like torch.sort, places NaN last
    d = torch.from_numpy(v).cuda()
    for r in (0, 1, 49000, 49998, 49999, 25000):
        val, idx = kth_select(d, r)
        assert idx.item() == order[r], r
        a, b = val.item(), float(v[order[r]])
        assert (np.isnan(a) and np.isnan(b)) or a == b
    one = torch.tensor([3.5]).cuda()
    assert kth_select(one, 0)[0].item() == 3.5
    with pytest.raises(ValueError):
        kth_select(one, 1)
    ties = torch.full((4097,), 2.0).cuda()
    assert kth_select(ties, 4000)[1].item() == 4000


def test_conformal_calculator_end_to_end(golden):
    """get_conformal_scores with the reference's kwargs and draws -> reference weighted scores + weights."""
    import safediffcon_b200 as s
    g = golden("conformal")
    B, nb = 6, 2
    gd = s.GaussianDiffusion(fx.FakeEps(), seq_length=(16, 128), timesteps=1000, sampling_timesteps=6, ddim_sampling_eta=1.0,
                             temporal=True, use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    states = fx.calibration_states(B * nb)
    noises = []
    for i in range(nb):
        noises += fx.chain_noise(B, fx.n_draws(1000, 6, False), seed=100 + i)
    it = iter(noises)
    orig = gd.sample
    gd.sample = lambda **kw: orig(noise=[next(it) for _ in range(6)], **kw)
    cfg = _cfg(device="cuda", num_cal_batch=nb, nt=11, InfFT_Q=None)
    loader = iter([states[i * B:(i + 1) * B] for i in range(nb)])
    scores, weights, st = s.ConformalCalculator(gd, cfg).get_conformal_scores(loader, Q=0.02)
    assert st.shape == (12, 3, 16, 128)
    assert np.allclose(weights.cpu().numpy(), g["weights"], rtol=1e-3, atol=1e-30)
    assert np.allclose(scores.cpu().numpy(), g["scores"], rtol=1e-3, atol=2e-4)
    # InfFT_Q multiplies in a second weight (inference/conformal.py:68-72)
    from safediffcon_b200.conformal import scores_and_weights
    sc1, w1 = scores_and_weights(states.cuda(), states.cuda(), _cfg(InfFT_Q=None), 0.02)
    sc2, w2 = scores_and_weights(states.cuda(), states.cuda(), _cfg(InfFT_Q=0.01), 0.02)
    ref2 = cr.raw_weight(states, 0.02, 500.0, 0.8) * cr.raw_weight(states, 0.01, 500.0, 0.8)
    assert torch.count_nonzero(sc1) == 0
    assert np.allclose(w2.cpu().numpy(), ref2.numpy(), rtol=1e-3, atol=1e-30)
