"""CPU, build container only (needs /root/reference): the drop-in claims of INTEGRATION.md checked against the LIVE reference.

1. Every entry point SURVEY.md section 8(b) names has the reference's signature: same parameter names, order, kinds and
   defaults; this implementation may only APPEND parameters (strict=, device=, seed= ...).
2. The reference's own call sites -- eval.py:diffuse_samples, inference/conformal.py:ConformalCalculator.get_conformal_scores,
   inference_ft.py:InferenceFT.inference -- are executed UNMODIFIED against the drop-in GaussianDiffusion: the keyword contract
   of sample() (guidance_u0 override, u_init / u_final asserts, unknown keys ignored, dispatch on the sampler) runs for real;
   only the CUDA chain behind it is replaced by a recorder (there is no GPU here)."""
import inspect
import os
import types

import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/1D"), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_import
    ref_import.install()
    import importlib
    names = ("model.unet", "model.diffusion", "data.generate_burgers", "utils.metrics", "utils.guidance", "inference.guidance",
             "inference.conformal", "utils.common")
    return {n: importlib.import_module(n) for n in names}


def _compatible(ours, theirs, name):
    po, pt = list(inspect.signature(ours).parameters.values()), list(inspect.signature(theirs).parameters.values())
    # a trailing **kwargs of the reference may sit behind our appended keyword parameters
    pt_core = [p for p in pt if p.kind is not p.VAR_KEYWORD]
    assert len(po) >= len(pt_core), (name, po, pt)
    for a, b in zip(po, pt_core):
        assert a.name == b.name and a.kind == b.kind, (name, a, b)
        if b.default is inspect.Parameter.empty:
            assert a.default is inspect.Parameter.empty, (name, a, b)
        else:
            assert a.default == b.default or (a.default is not inspect.Parameter.empty and repr(a.default) == repr(b.default)), (name, a, b)
    if any(p.kind is p.VAR_KEYWORD for p in pt):
        assert any(p.kind is p.VAR_KEYWORD for p in po), (name, "reference accepts **kwargs")
    for extra in po[len(pt_core):]:
        assert extra.default is not inspect.Parameter.empty or extra.kind in (extra.VAR_KEYWORD, extra.VAR_POSITIONAL), (name, extra)


def test_signatures_match_the_reference(ref):
    import safediffcon_b200 as s
    rd, ru = ref["model.diffusion"].GaussianDiffusion, ref["model.unet"].Unet2D
    pairs = [(s.Unet2D.__init__, ru.__init__, "Unet2D.__init__"), (s.Unet2D.forward, ru.forward, "Unet2D.forward"),
             (s.GaussianDiffusion.__init__, rd.__init__, "GaussianDiffusion.__init__")]
    for m in ("sample", "p_sample_loop", "ddim_sample", "p_sample", "p_mean_variance", "model_predictions", "q_posterior",
              "predict_start_from_noise", "predict_noise_from_start", "q_sample", "p_losses", "forward", "get_guidance_options",
              "set_condition", "set_pad_condition"):
        pairs.append((getattr(s.GaussianDiffusion, m), getattr(rd, m), "GaussianDiffusion." + m))
    gb = ref["data.generate_burgers"]
    pairs += [(s.burgers_numeric_solve, gb.burgers_numeric_solve, "burgers_numeric_solve"),
              (s.burgers_numeric_solve_free, gb.burgers_numeric_solve_free, "burgers_numeric_solve_free"),
              (s.make_data_varying_f, gb.make_data_varying_f, "make_data_varying_f")]
    um = ref["utils.metrics"]
    pairs += [(s.control_trajectories, um.control_trajectories, "control_trajectories"),
              (s.evaluate_samples, um.evaluate_samples, "evaluate_samples"),
              (s.calculate_safety_metrics, um.calculate_safety_metrics, "calculate_safety_metrics")]
    pairs += [(s.calculate_guidance, ref["utils.guidance"].calculate_guidance, "calculate_guidance"),
              (s.get_finetune_guidance, ref["utils.guidance"].get_finetune_guidance, "get_finetune_guidance"),
              (s.get_weight, ref["inference.guidance"].get_weight, "get_weight"),
              (s.normalize_weights, ref["inference.guidance"].normalize_weights, "normalize_weights")]
    rc = ref["inference.conformal"].ConformalCalculator
    pairs += [(s.ConformalCalculator.__init__, rc.__init__, "ConformalCalculator.__init__"),
              (s.ConformalCalculator.get_conformal_scores, rc.get_conformal_scores, "get_conformal_scores"),
              (s.ConformalCalculator.calculate_quantile, rc.calculate_quantile, "calculate_quantile")]
    uc = ref["utils.common"]
    pairs += [(s.get_target, uc.get_target, "get_target"), (s.build_model, uc.build_model, "build_model"),
              (s.load_model, uc.load_model, "load_model")]
    for ours, theirs, name in pairs:
        _compatible(ours, theirs, name)
    # schedule buffers / attributes the reference's callers read
    gd_r = rd(ref["model.unet"].Unet2D(dim=32, channels=3, resnet_block_groups=1), seq_length=(16, 128), temporal=True, use_conv2d=True)
    torch.manual_seed(0)
    gd_o = s.GaussianDiffusion(s.Unet2D(dim=32, channels=3, resnet_block_groups=1), seq_length=(16, 128), temporal=True, use_conv2d=True)
    assert list(gd_o.state_dict().keys()) == list(gd_r.state_dict().keys())
    for k, v in gd_r.state_dict().items():
        assert gd_o.state_dict()[k].shape == v.shape and gd_o.state_dict()[k].dtype == v.dtype, k
    for attr in ("channels", "self_condition", "traj_size", "num_timesteps", "sampling_timesteps", "is_ddim_sampling", "ddim_sampling_eta",
                 "guidance_u0", "condition_idx", "is_condition_u0", "is_condition_uT", "train_on_padded_locations"):
        assert getattr(gd_o, attr) == getattr(gd_r, attr), attr


class _Recorder:
    """Stands in for the CUDA chain: records what sample() dispatched, returns zeros of the requested shape."""

    def __init__(self):
        self.calls = []

    def __call__(self, gd, sampler):
        def fn(shape, *a, **k):
            self.calls.append((sampler, tuple(shape), dict(k), bool(gd.guidance_u0)))
            return torch.zeros(shape)
        return fn


def _drop_in(sampling_timesteps):
    import safediffcon_b200 as s
    torch.manual_seed(0)
    gd = s.GaussianDiffusion(s.Unet2D(dim=32, channels=3, resnet_block_groups=1), seq_length=(16, 128), timesteps=1000,
                             sampling_timesteps=sampling_timesteps, ddim_sampling_eta=1.0, temporal=True, use_conv2d=True,
                             is_condition_u0=True, is_condition_uT=True, condition_idx=10)
    rec = _Recorder()
    gd.p_sample_loop, gd.ddim_sample = rec(gd, "ddpm"), rec(gd, "ddim")
    return gd, rec


def test_reference_call_sites_run_against_the_drop_in(ref):
    import importlib
    states = 0.1 * torch.randn(6, 3, 16, 128)
    # ---- eval.py:diffuse_samples (reference lines 21-59) ----
    ev = importlib.import_module("eval")
    gd, rec = _drop_in(200)
    ds = types.SimpleNamespace(nt_total=11)
    cfg = types.SimpleNamespace(batch_size=3, n_test_samples=6, ddim_eta=1.0, using_ddim=True, ddim_sampling_steps=200)
    out = ev.diffuse_samples(gd, ds, [states[:3], states[3:]], cfg, torch.device("cpu"))
    assert out.shape == (6, 3, 16, 128) and len(rec.calls) == 2
    sampler, shape, kw, g_u0 = rec.calls[0]
    assert sampler == "ddim" and shape == (3, 3, 16, 128) and g_u0 is True
    assert torch.equal(kw["u_init"], states[:3, 0, 0]) and torch.equal(kw["u_final"], states[:3, 0, 10]) and kw["clip_denoised"] is True
    assert kw["nablaJ"] is None and "timesteps" in kw and "ddim_sampling_eta" in kw   # unknown keys travel and are ignored
    # ---- inference/conformal.py:ConformalCalculator.get_conformal_scores (reference lines 25-93) with the DDPM sampler ----
    gd, rec = _drop_in(1000)
    ccfg = types.SimpleNamespace(device="cpu", num_cal_batch=2, nt=11, InfFT_Q=None, use_max_safety=True, u_bound=0.8,
                                 guidance_weights={"w_score": 500.0})
    sc, w, st = ref["inference.conformal"].ConformalCalculator(gd, ccfg).get_conformal_scores(iter([states[:3], states[3:]]), Q=0.02)
    assert sc.shape == (6,) and w.shape == (6,) and st.shape == states.shape and len(rec.calls) == 2
    sampler, shape, kw, g_u0 = rec.calls[1]
    assert sampler == "ddpm" and g_u0 is False and kw["enable_grad"] is False        # sample() applied the guidance_u0 override
    assert torch.equal(kw["w_groundtruth"], states[3:, 1]) and kw["device"] == "cpu"
    # ---- inference_ft.py:InferenceFT.inference (reference lines 316-347), executed unbound on a stand-in object ----
    ift = importlib.import_module("inference.inference_ft")
    gd, rec = _drop_in(200)
    guide = lambda x: torch.zeros_like(x)  # noqa: E731
    obj = types.SimpleNamespace(device="cpu", config=types.SimpleNamespace(nt=11), model=gd, guidance_fn=guide, J_scheduler=None,
                                w_scheduler=None, get_model_for_inference=lambda: gd)
    pred = ift.InferenceFT.inference(obj, states[:2])
    assert pred.shape == (2, 3, 16, 128)
    assert rec.calls[0][2]["nablaJ"] is guide and rec.calls[0][2]["enable_grad"] is False
    ift.InferenceFT.inference(obj, states[:2], is_backward=True)
    assert rec.calls[1][2]["enable_grad"] is True
    # missing conditions fail like the reference (diffusion.py:581-586)
    with pytest.raises(AssertionError):
        gd.sample(batch_size=2, u_final=states[:2, 0, 10])
