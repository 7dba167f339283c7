"""CPU: the oracle restatement reproduces the outputs recorded from the UNMODIFIED reference (tests/golden)."""
import numpy as np
import torch

from oracle import conformal_ref as cr
from oracle import diffusion_ref as dr
from oracle import fixtures as fx
from oracle import solver_ref


def test_schedule_bit_exact(golden):
    for T in (1000, 20):
        g = golden(f"schedule_T{T}")
        b = dr.schedule_buffers(T)
        assert set(g.files) == set(dr.BUFFER_NAMES)
        for k in g.files:
            assert np.array_equal(g[k], b[k].numpy()), k


def test_solver_c_oracle_bit_exact(golden):
    u0, f = fx.solver_inputs(16, 0)
    assert np.array_equal(solver_ref.solve_free_c(u0.numpy(), f.numpy()), golden("solver_free")["traj"])


def test_solver_c_oracle_nan_propagation(golden):
    u0, f = fx.solver_inputs_wild(4, 3)
    g = golden("solver_free_wild")["traj"]
    assert np.isnan(g).any() and np.isfinite(g[0]).all()
    assert np.array_equal(solver_ref.solve_free_c(u0.numpy(), f.numpy()), g, equal_nan=True)


def test_solver_cartesian(golden):
    u0, f = fx.solver_inputs(3, 5)
    assert np.array_equal(solver_ref.solve_cartesian_c(u0.numpy(), f.numpy()[:2]), golden("solver_cartesian")["traj"])


def test_solver_torch_port_short():
    # the torch-op port (the timed CPU baseline) equals the C restatement; 400 steps keep this fast
    u0, f = fx.solver_inputs(5, 9)
    a = solver_ref.solve_free_torch(u0, f, T=0.04, dt=1e-4, num_t=10).numpy()
    b = solver_ref.solve_free_c(u0.numpy(), f.numpy(), T=0.04, dt=1e-4, num_t=10)
    assert np.array_equal(a, b)


def test_metrics(golden):
    g = golden("metrics")
    traj = torch.from_numpy(golden("solver_free")["traj"])
    tgt = torch.roll(traj, 1, dims=0)
    for tag, bound in (("b08", 0.8), ("b03", 0.3)):
        m = solver_ref.evaluate(traj, tgt, bound)
        for i, k in enumerate(g["keys"]):
            ref = g[f"{tag}_{i}"]
            if ref.ndim:
                assert list(ref.astype(int)) == m[str(k)]
            else:
                assert abs(float(ref) - m[str(k)]) <= 1e-6 * max(1.0, abs(float(ref))), k
        J, pts, tms, flg = solver_ref.score_c(traj.numpy(), tgt[:, -1].numpy(), bound)
        assert abs(J.mean() - m["control_mse_mean (J)"]) < 1e-6
        assert pts.sum() / traj.numel() == np.float64(m["point_exceed_ratio (R_p)"]) or \
            abs(pts.sum() / traj.numel() - m["point_exceed_ratio (R_p)"]) < 1e-7
        assert flg.nonzero()[0].tolist() == m["sample_excedd_indices"]


def _run_case(name, T, S, kw, B=4):
    u_init, u_final, w_gt = fx.chain_conditions(B)
    fake = fx.FakeEps()
    bufs = dr.schedule_buffers(T)
    noises = fx.chain_noise(B, fx.n_draws(T, S, kw["guidance_u0"]), seed=kw["seed"])
    eps_fn = lambda x, t: fake(x, torch.full((B,), t, dtype=torch.long))  # noqa: E731
    guide = dict(Q=kw["Q"], w_score=500.0, u_bound=0.8, use_max_safety=kw.get("use_max_safety", True)) if kw["guided"] else None
    wg = w_gt if kw["w_gt"] else None
    if S < T:
        return dr.ddim_chain(eps_fn, bufs, noises, u_init, u_final, wg, guide, S=S, eta=1.0, T=T)
    return dr.ddpm_chain(eps_fn, bufs, noises, u_init, u_final, wg, guide, guidance_u0=kw["guidance_u0"], T=T,
                         enable_grad=kw["enable_grad"])


def test_chains_bit_exact(golden):
    g = golden("chains")
    for name, T, S, kw in fx.CHAIN_CASES:
        out = _run_case(name, T, S, kw)
        assert np.array_equal(out.numpy(), g[name]), name


def test_guidance_weights_quantile(golden):
    g = golden("guidance")
    x = fx.guidance_states(6)
    for Q in (0.0, 0.05, -0.5):
        for ums in (True, False):
            assert np.array_equal(dr.safety_guidance_grad(x, Q, 500.0, 0.8, ums).numpy(), g[f"grad_Q{Q}_{int(ums)}"])
            assert np.array_equal(cr.raw_weight(x, Q, 500.0, 0.8, ums).numpy(), g[f"weight_Q{Q}_{int(ums)}"])
    for i, w in enumerate(fx.weight_vectors()):
        assert np.array_equal(cr.normalize_weights(w.clone()).numpy(), g[f"norm_{i}"], equal_nan=True)
    for i, (s, a) in enumerate(fx.score_vectors()):
        assert float(cr.quantile(s, a)) == float(g[f"quant_{i}"])
        assert float(s[cr.quantile_index(s, a)]) == float(g[f"quant_{i}"])
    assert cr.quantile_rank(1000, 0.98) == 980 and cr.quantile_rank(50000, 0.98) == 49000


def test_conformal_scores(golden):
    g = golden("conformal")
    B, nb = 6, 2
    states = fx.calibration_states(B * nb)
    bufs = dr.schedule_buffers(1000)
    fake = fx.FakeEps()
    sc, ws = [], []
    for i in range(nb):
        st = states[i * B:(i + 1) * B]
        noises = fx.chain_noise(B, fx.n_draws(1000, 6, False), seed=100 + i)
        eps_fn = lambda x, t: fake(x, torch.full((B,), t, dtype=torch.long))  # noqa: E731
        out = dr.ddim_chain(eps_fn, bufs, noises, st[:, 0, 0, :], st[:, 0, 10, :], st[:, 1], None, S=6)
        sc.append(cr.nonconformity(out, st))
        ws.append(cr.raw_weight(st, 0.02, 500.0, 0.8))
    w = cr.normalize_weights(torch.cat(ws))
    assert np.array_equal(w.numpy(), g["weights"])
    assert np.array_equal((w * torch.cat(sc)).numpy(), g["scores"])


def test_datagen_oracle_bit_exact(golden):
    from oracle import datagen_ref as dg
    g = golden("datagen")
    np.random.seed(0)
    u0, f = dg.make_data_varying_f(16, 12, 128, 10)
    assert u0.dtype == np.float64 and np.array_equal(u0, g["u0"])
    assert f.dtype == torch.float32 and np.array_equal(f.numpy(), g["f"])
    np.random.seed(3)
    u0p, fp = dg.make_data_varying_f(3, 4, 128, 10, partial_control='front_rear_quarter', alpha=1.7)
    assert np.array_equal(u0p, g["u0_partial"]) and np.array_equal(fp.numpy(), g["f_partial"])
    assert (fp.numpy()[:, :, 32:96] == 0).all() and np.abs(fp.numpy()).max() <= 10.0


def test_dataset_states_oracle_bit_exact(golden):
    from oracle import datagen_ref as dg
    g = golden("dataset_states")
    traj, f = torch.from_numpy(g["traj"]), torch.from_numpy(g["f"])
    for use_max in (True, False):
        out = dg.dataset_states(traj, f, use_max_safety=use_max)
        assert np.array_equal(out.numpy(), g[f"states_max{int(use_max)}"])
