"""GPU: the C++ whole-network executor (sdc_unet_forward, ONE C call per evaluation) against the Python schedule of the same
kernels and against the reference goldens; in-place repacking, graph capture at large batch, FP16-range guard."""
import warnings

import numpy as np
import pytest
import torch

from oracle import fixtures as fx

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def _net(dim=128, prec="f16"):
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    net.precision = prec
    return net


@pytest.mark.parametrize("prec,dim", [("f16", 128), ("tf32", 128), ("f16", 64), ("tf32", 32)])
def test_executor_matches_python_schedule(prec, dim):
    from safediffcon_b200 import unet as U
    net = _net(dim, prec)
    x, t = fx.unet_inputs(6)
    x, t = x.cuda(), t.cuda()
    with torch.no_grad():
        plan = net._plan_ready()
        assert plan is not None
        # identical inputs for an exact comparison: the executor's default FiLM table comes from a split-TF32 tensor-core GEMM
        # (1e-6 from the fp32 loop of the Python schedule) -- a perturbation that small already re-draws the fp16 / TF32
        # rounding noise of everything downstream (two valid evaluations then sit ~6e-4 apart, each ~6.5e-4 from the oracle)
        plan.set_flag(plan.FILM_TC, 0)
        net.invalidate_packed()
        a = net(x, t)                       # C++ executor
        b = net.denoise_uniform(x, 417)
        U.USE_PLAN = False
        try:
            a_py = net(x, t)                # Python schedule, same kernels
            b_py = net.denoise_uniform(x, 417)
        finally:
            U.USE_PLAN = True
    # same kernels in the same order; only the fp64 atomics of the GroupNorm statistics may land in a different order
    assert rel(a, a_py) < 2e-6 and rel(b, b_py) < 2e-6, (rel(a, a_py), rel(b, b_py))


def test_layernorm_fusion_is_a_pure_reformulation():
    """LayerNorms folded into the qkv / output-projection epilogues (option) against the separate LayerNorm passes: same math,
    different rounding sites (the normalised rows are no longer rounded to fp16) -- a few 1e-4 apart, both within 1e-3 of the oracle."""
    from oracle import unet_ref
    net = _net(128, "f16")
    x, t = fx.unet_inputs(4)
    with torch.no_grad():
        ref = unet_ref.unet_forward({k: v.detach().cpu() for k, v in net.state_dict().items()}, x, t)
        net.fuse_layernorm = True
        a = net(x.cuda(), t.cuda()).cpu()
        net.fuse_layernorm = False
        b = net(x.cuda(), t.cuda()).cpu()
    assert not torch.equal(a, b)
    assert rel(a, b) < 8e-4, rel(a, b)
    pa, pb = [rel(a[i], ref[i]) for i in range(4)], [rel(b[i], ref[i]) for i in range(4)]
    print("eps error fused LN", pa, "separate LN", pb)
    assert max(pa) < 1e-3 and max(pb) < 1e-3


def test_film_table_on_tensor_cores_matches_the_time_mlp():
    """The executor's FiLM table (split-TF32 tcgen05 GEMM) against the time MLP evaluated by torch in fp64."""
    import math
    import torch.nn.functional as F
    net = _net(128, "f16")
    with torch.no_grad():
        tab = net._plan_ready().film_table().cpu()
        sd = {k: v.detach().double().cpu() for k, v in net.state_dict().items()}
        t = torch.arange(1000, dtype=torch.float64)
        freq = torch.exp(torch.arange(64, dtype=torch.float64) * -(math.log(10000.0) / 63))
        emb = torch.cat(((t[:, None] * freq).sin(), (t[:, None] * freq).cos()), dim=-1)            # unet.py:81-95
        h = F.linear(F.gelu(F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
        blocks = [k[:-len(".mlp.1.weight")] for k in net.state_dict() if k.endswith(".mlp.1.weight")]   # module order = execution order
        ref = torch.cat([F.linear(F.silu(h), sd[b + ".mlp.1.weight"], sd[b + ".mlp.1.bias"]) for b in blocks], dim=1)
    assert tab.shape == ref.shape == (1000, 16128)
    # against fp64 (the fp32 sinusoid of t up to 999 alone costs ~1e-5) and against the fp32 CUDA-core loop on the same fp32 inputs
    assert rel(tab.double(), ref) < 5e-5 and (tab.double() - ref).abs().max().item() < 1e-4
    with torch.no_grad():
        loop = net._film_table(net._packed()).cpu()
    assert rel(tab, loop) < 5e-6 and (tab - loop).abs().max().item() < 2e-5, (rel(tab, loop), (tab - loop).abs().max().item())


def test_executor_eps_within_1e3_of_reference_golden(golden):
    net = _net(128, "f16")
    g = golden("unet_dim128")
    x, t = fx.unet_inputs(2)
    with torch.no_grad():
        eps = net(x.cuda(), t.cuda()).cpu()
    ref = torch.from_numpy(g["eps"])
    assert rel(eps, ref) < 1e-3
    assert max(rel(eps[i], ref[i]) for i in range(2)) < 1e-3


def test_repack_in_place_follows_data_writes():
    """ema_pytorch updates the EMA copy through p.data (no version bump): revalidate_packed() must notice, and the executor
    must refresh its packed weights IN PLACE (captured graphs hold their addresses)."""
    from oracle import unet_ref
    net = _net(32, "tf32")
    x, t = fx.unet_inputs(2)
    x, t = x.cuda(), t.cuda()
    with torch.no_grad():
        a = net(x, t)
        plan = net._plan_ready()
        versions = [p._version for p in net.parameters()]
        other = [torch.randn_like(p) * 0.05 for p in net.parameters()]
        for p, o in zip(net.parameters(), other):
            p.data.lerp_(p.data + o, 0.5)                      # what EMA.update() does
        assert [p._version for p in net.parameters()] == versions   # invisible to the version counters ...
        stale = net(x, t)
        assert torch.equal(stale, a)                            # ... so a plain call still uses the old pack
        assert net.revalidate_packed() is False                 # the once-per-chain digest check notices
        b = net(x, t)
        assert net._plan_ready() is plan                        # same handle, refreshed in place
        ref = unet_ref.unet_forward({k: v.cpu() for k, v in net.state_dict().items()}, x.cpu(), t.cpu())
    assert not torch.allclose(a, b)
    assert rel(b.cpu(), ref) < 1e-3
    assert net.revalidate_packed() is True
    net.invalidate_packed()
    with torch.no_grad():
        assert torch.allclose(net(x, t), b, rtol=1e-5, atol=1e-6)


def test_sample_sees_ema_style_updates():
    """GaussianDiffusion.sample revalidates once per chain: weights written through .data between two chains change the samples."""
    import safediffcon_b200 as s
    net = _net(64, "f16")
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0, temporal=True,
                             use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    u0, uT, _ = fx.chain_conditions(3)
    kw = dict(batch_size=3, u_init=u0.cuda(), u_final=uT.cuda(), guidance_u0=False, enable_grad=False, seed=5)
    a = gd.sample(**kw)
    a2 = gd.sample(**kw)
    assert torch.equal(a, a2)
    for p in net.parameters():
        p.data.mul_(1.02)
    b = gd.sample(**kw)
    assert not torch.allclose(a, b)


def test_times_outside_table_raise():
    net = _net(64, "f16")
    x, _ = fx.unet_inputs(2)
    with pytest.raises(ValueError, match="FiLM table"):
        net.denoise_uniform(x.cuda(), 1000)
    with pytest.raises(ValueError, match="integer diffusion times"), torch.no_grad():   # (the training path evaluates the time MLP itself)
        net(x.cuda(), torch.tensor([3, 1000]).cuda())
    import safediffcon_b200 as s
    s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1200, sampling_timesteps=4, temporal=True, use_conv2d=True)
    assert net.table_timesteps == 1200
    with torch.no_grad():
        assert torch.isfinite(net.denoise_uniform(x.cuda(), 1150)).all()      # the table was rebuilt with 1200 rows


def test_graph_chain_at_large_batch_equals_eager():
    """The executor keeps every activation in a workspace, so a reverse step is graph-capturable at any batch size."""
    import safediffcon_b200 as s
    from safediffcon_b200 import diffusion as D
    net = _net(64, "f16")
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=5, ddim_sampling_eta=1.0, temporal=True,
                             use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    B = 300
    g = torch.Generator().manual_seed(1)
    u0, uT = 0.2 * torch.randn(B, 128, generator=g), 0.1 * torch.randn(B, 128, generator=g)
    cfg = type("Cfg", (), dict(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0}))()
    kw = dict(batch_size=B, u_init=u0.cuda(), u_final=uT.cuda(), guidance_u0=True, nablaJ=s.safety_guidance(cfg, 0.0), enable_grad=False,
              seed=11, sample_offset=7)
    a = gd.sample(**kw)
    assert len(gd._graphs.entries) == 1
    old = D.GRAPH_MAX_BATCH
    D.GRAPH_MAX_BATCH = 0
    try:
        b = gd.sample(**kw)
    finally:
        D.GRAPH_MAX_BATCH = old
    assert torch.isfinite(a).all()
    assert rel(a, b) < 1e-5, rel(a, b)


def test_fp16_range_guard_falls_back_to_tf32():
    """Weights scaled so that pre-norm activations leave the fp16 range: the chain must not return Inf/NaN samples silently."""
    import safediffcon_b200 as s
    net = _net(64, "f16")
    with torch.no_grad():
        net.downs[0][0].block1.proj.weight.mul_(3.0e5)     # GroupNorm would absorb the scale -- if the conv output were storable
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0, temporal=True,
                             use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    u0, uT, _ = fx.chain_conditions(2)
    kw = dict(batch_size=2, u_init=u0.cuda(), u_final=uT.cuda(), guidance_u0=False, enable_grad=False, seed=3)
    with torch.no_grad():
        x, _ = fx.unet_inputs(2)
        raw = net.denoise_uniform(x.cuda(), 500)
    assert not torch.isfinite(raw).all()                  # the fp16 path overflows on this checkpoint ...
    assert net.take_nonfinite() > 0                       # ... and the device counter saw it
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = gd.sample(**kw)
    assert any("tf32" in str(m.message) for m in w)
    assert net.precision == "tf32"
    assert torch.isfinite(out).all()
    ref = gd.sample(**kw)                                  # now plain TF32
    assert torch.equal(out, ref)


def test_per_launch_profile_covers_the_step():
    net = _net(128, "f16")
    x, _ = fx.unet_inputs(4)
    x = x.cuda()
    with torch.no_grad():
        plan = net._plan_ready()
        net.denoise_uniform(x, 10)
        plan.profile(True)
        net.denoise_uniform(x, 10)
        torch.cuda.synchronize()
        ent = plan.profile_entries()
        plan.profile(False)
    names = {e[0] for e in ent}
    assert {"conv3x3", "conv1x1", "gn_silu", "layernorm", "conv1x1_qkv", "linattn_context", "conv_upsample", "conv_unshuffle"} <= names
    assert 100 <= len(ent) <= 200 and all(e[1] > 0 for e in ent)
    conv_flops = sum(e[3] for e in ent if e[0].startswith("conv"))
    assert abs(conv_flops / 4 / 1e9 - 25.8) < 1.5      # executed conv GFLOP per sample (fused upsample: 4/9 of the reference's MACs there)


def test_torch_ops_dispatch_to_the_same_kernels():
    """torch.ops.safediffcon_b200.* (torch.library registrations) == the Python API, bit for bit."""
    import safediffcon_b200 as s
    from safediffcon_b200 import ops
    ns = torch.ops.safediffcon_b200
    u0, f = fx.solver_inputs(6, seed=1)
    u0, f = u0.cuda(), f.cuda()
    assert torch.equal(ns.burgers_solve_free(u0, f, 0.01, 1.0, 1e-4, True), s.burgers_numeric_solve_free(u0, f, 0.01, 1.0))
    sc = torch.rand(777, generator=torch.Generator().manual_seed(0)).cuda()
    v, i = ns.kth_select(sc, 700)
    v2, i2 = s.kth_select(sc, 700)
    assert torch.equal(v, v2) and torch.equal(i, i2)
    net = _net(64, "f16")
    x, t = fx.unet_inputs(3)
    h = ops.register_plan(net)
    with torch.no_grad():
        a = ns.unet_forward(x.cuda(), t.cuda().int(), h)
        b = net(x.cuda(), t.cuda())
    assert rel(a, b) < 2e-6
    torch.library.opcheck(ns.kth_select.default, (sc, 700), test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(ns.burgers_solve_free.default, (u0, f, 0.01, 1.0, 1e-4, True), test_utils=("test_schema", "test_faketensor"))


@pytest.mark.parametrize("prec,dim", [("f16", 64), ("tf32", 32)])
def test_backward_data_in_one_c_call(prec, dim, golden):
    """sdc_unet_backward_data (recording forward + reverse walk inside the executor) == the Python schedule of the same kernels,
    and == the reference's autograd (golden, dim 32 / 64)."""
    from safediffcon_b200 import unet as U
    net = _net(dim, prec)
    B = 2
    x, t = fx.unet_inputs(B)
    g = fx.unet_cotangent(B)
    with torch.no_grad():
        plan = net._plan_ready(backward=True)
        plan.set_flag(plan.FILM_TC, 0)          # identical FiLM rows in both paths (see test_executor_matches_python_schedule)
        net.invalidate_packed()
        eps_c, gx_c = net.vjp(x.cuda(), t.cuda(), g.cuda())              # C executor
        U.USE_PLAN = False
        try:
            eps_p, gx_p = net.vjp(x.cuda(), t.cuda(), g.cuda())          # Python schedule
        finally:
            U.USE_PLAN = True
    assert rel(eps_c, eps_p) < 2e-6 and rel(gx_c, gx_p) < 1e-5, (rel(eps_c, eps_p), rel(gx_c, gx_p))
    gold = golden(f"unet_dim{dim}_vjp")
    ref_eps, ref_gx = torch.from_numpy(gold["eps"]), torch.from_numpy(gold["grad_x"])
    assert rel(eps_c.cpu(), ref_eps) < 1.5e-3 and rel(gx_c.cpu(), ref_gx) < 3e-3, (rel(eps_c.cpu(), ref_eps), rel(gx_c.cpu(), ref_gx))
    # uniform integer time through the same entry point
    with torch.no_grad():
        e2, g2 = net.vjp(x.cuda(), 417, g.cuda())
        e3, g3 = net.vjp(x.cuda(), torch.full((B,), 417).cuda(), g.cuda())
    assert rel(e2, e3) < 2e-6 and rel(g2, g3) < 1e-5
