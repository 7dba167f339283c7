"""GPU parity of the parameter-gradient pass: each reduction kernel of csrc/unet_wgrad.cu
against torch autograd of the same op (float64), and the whole denoiser against the UNMODIFIED reference's autograd
(tests/golden/unet_dim64_pgrad.npz: dense cotangent; unet_dim64_ft_pgrad.npz: the inference-time fine-tuning loss
through the last DDIM step).  Operands are TF32 / FP16 (10-bit mantissa), accumulation fp32: tolerances are relative."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures as fx

pytestmark = pytest.mark.gpu


def nhwc_rows(x):
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous()


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _L():
    from safediffcon_b200 import _lib as L
    from safediffcon_b200 import unet as U  # noqa: F401  (registers signatures)
    return L, L.lib(), U


WGRAD_CASES = [
    # kind, half, B, H, W (forward OUTPUT resolution), c0, c1, cout
    (1, 1, 2, 16, 128, 64, 0, 64),
    (1, 1, 3, 8, 64, 128, 64, 128),     # two input segments (skip concatenation)
    (1, 0, 2, 4, 32, 96, 32, 32),       # fp32 activations, ragged channel tiles
    (1, 1, 5, 2, 16, 64, 0, 256),
    (0, 1, 2, 8, 64, 128, 0, 384),
    (0, 0, 3, 4, 32, 256, 0, 96),
    (0, 1, 2, 16, 128, 320, 0, 64),     # stem-like 1x1 over the im2col operand
    (2, 1, 2, 8, 64, 64, 0, 128),       # pixel-unshuffle + 1x1: forward input 16x128 with 64 channels
    (2, 0, 3, 2, 16, 32, 0, 64),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[f"k{c[0]}_h{c[1]}_B{c[2]}_{c[3]}x{c[4]}_{c[5]}+{c[6]}to{c[7]}" for c in WGRAD_CASES])
def test_conv_wgrad_vs_autograd(case):
    L, lib, U = _L()
    kind, half, B, H, W, c0, c1, cout = case
    g = torch.Generator().manual_seed(17 * c0 + cout + H)
    cin = c0 + c1
    ksz = 3 if kind == 1 else 1
    hin, win = (2 * H, 2 * W) if kind == 2 else (H, W)
    adt = torch.float16 if half else torch.float32
    x = torch.randn(B, cin, hin, win, generator=g).to(adt)          # operand precision: exactly representable inputs
    gy = torch.randn(B, cout, H, W, generator=g)
    wshape = (cout, cin * 4 if kind == 2 else cin, ksz, ksz)
    w = torch.zeros(wshape, dtype=torch.float64, requires_grad=True)
    xd = x.double()
    if kind == 2:
        xd = xd.reshape(B, cin, H, 2, W, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, cin * 4, H, W)
    y = F.conv2d(xd, w, padding=1 if kind == 1 else 0)
    (want,) = torch.autograd.grad(y, w, gy.double())
    rows = nhwc_rows(x).cuda()
    a0 = rows[:, :c0].contiguous()
    a1 = rows[:, c0:].contiguous() if c1 else None
    dy = nhwc_rows(gy).cuda()
    dw = torch.zeros(wshape, device="cuda")
    L.check(lib.sdc_conv_wgrad(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, L.stream_ptr()))
    err = rel(dw.cpu(), want)
    assert err < 1e-3, err    # dY is rounded to TF32 (2^-11 relative per element, averaged over the pixel reduction)
    # accumulation semantics: a second call adds
    L.check(lib.sdc_conv_wgrad(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, L.stream_ptr()))
    assert rel(dw.cpu(), 2 * want) < 1e-3


TC_CASES = [
    # kind, half, B, H, W, c0, c1, cout
    (1, 1, 3, 16, 128, 128, 0, 128),     # full-resolution level, 64-pixel half rows, fp16 activations (converted)
    (1, 0, 2, 16, 128, 128, 128, 128),   # concat input (two K... two N segments), fp32 activations
    (1, 1, 5, 8, 64, 256, 128, 256),     # 8x64 level
    (0, 1, 4, 4, 32, 128, 0, 384),       # qkv projection, 4x32: two rows per pixel tile
    (1, 1, 6, 2, 16, 512, 0, 128),       # 2x16 level: a pixel tile spans two images
    (1, 0, 7, 2, 16, 128, 0, 256),       # odd batch: the last tile is half out of range (TMA zero fill)
    (0, 1, 3, 16, 128, 256, 0, 128),     # 1x1 res_conv
]


@pytest.mark.parametrize("case", TC_CASES, ids=[f"k{c[0]}_h{c[1]}_B{c[2]}_{c[3]}x{c[4]}_c{c[5]}+{c[6]}_o{c[7]}" for c in TC_CASES])
def test_conv_wgrad_tcgen05_vs_autograd(case):
    """sdc_conv_wgrad_tc (tcgen05 kind::tf32, MN-major operands, pixel axis = K) against autograd in fp64 and against the mma.sync kernel."""
    L, lib, U = _L()
    kind, half, B, H, W, c0, c1, cout = case
    g = torch.Generator().manual_seed(17 * c0 + cout + H)
    cin = c0 + c1
    ksz = 3 if kind == 1 else 1
    adt = torch.float16 if half else torch.float32
    x = torch.randn(B, cin, H, W, generator=g).to(adt)
    if not half:
        x = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)   # TF32-representable (what an fp32 operand tensor holds)
    gy = torch.randn(B, cout, H, W, generator=g)
    gy = (gy.view(torch.int32) & ~0x1FFF).view(torch.float32)     # dY arrives TF32-rounded from the backward-data pass
    w = torch.zeros((cout, cin, ksz, ksz), dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double(), w, padding=1 if kind == 1 else 0)
    (want,) = torch.autograd.grad(y, w, gy.double())
    rows = nhwc_rows(x).cuda()
    a0 = rows[:, :c0].contiguous()
    a1 = rows[:, c0:].contiguous() if c1 else None
    dy = nhwc_rows(gy).cuda()
    dw = torch.zeros((cout, cin, ksz, ksz), device="cuda")
    nb = int(lib.sdc_conv_wgrad_tc_scratch(half, c0, c1, B, H, W))
    scratch = torch.empty(max(nb, 16), dtype=torch.uint8, device="cuda")
    rc = lib.sdc_conv_wgrad_tc(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, L.ptr(scratch), nb, L.stream_ptr())
    assert rc == 0, rc
    torch.cuda.synchronize()
    err = rel(dw.cpu(), want)
    assert err < 2e-5, err          # exact TF32 products, fp32 accumulation
    old = torch.zeros_like(dw)
    L.check(lib.sdc_conv_wgrad(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(old), B, H, W, cout, L.stream_ptr()))
    assert rel(dw, old) < 2e-5
    rc = lib.sdc_conv_wgrad_tc(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, cout, L.ptr(scratch), nb, L.stream_ptr())
    assert rc == 0 and rel(dw.cpu(), 2 * want) < 2e-5     # accumulation semantics


def test_colsum_and_head_wgrad():
    L, lib, U = _L()
    g = torch.Generator().manual_seed(3)
    for M, C in ((4096, 128), (1000, 96), (77, 1024)):
        x = torch.randn(M, C, generator=g)
        out, xc = torch.zeros(C, device="cuda"), x.cuda()
        L.check(lib.sdc_colsum(L.ptr(xc), L.ptr(out), M, C, L.stream_ptr()))
        assert rel(out.cpu(), x.double().sum(0)) < 1e-5
    B, HW, Cin, Cout = 3, 16 * 128, 128, 3
    for half in (1, 0):
        xa = torch.randn(B * HW, Cin, generator=g).to(torch.float16 if half else torch.float32)
        gg = torch.randn(B, Cout, HW, generator=g)
        dw, db = torch.zeros(Cout, Cin, device="cuda"), torch.zeros(Cout, device="cuda")
        ggc, xac = gg.cuda(), xa.cuda()
        L.check(lib.sdc_head_conv1_wgrad(L.ptr(ggc), L.ptr(xac), half, L.ptr(dw), L.ptr(db), B, HW, Cin, Cout, L.stream_ptr()))
        gr = gg.permute(0, 2, 1).reshape(B * HW, Cout).double()
        assert rel(dw.cpu(), gr.T @ xa.double()) < 1e-5
        assert rel(db.cpu(), gr.sum(0)) < 1e-5


@pytest.mark.parametrize("film", [True, False])
def test_gn_param_grad_vs_autograd(film):
    L, lib, U = _L()
    B, HW, C = 3, 8 * 64, 128
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, HW, C, generator=g).double()
    dy = torch.randn(B, HW, C, generator=g).double()
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).double().requires_grad_()
    beta = (0.1 * torch.randn(C, generator=g)).double().requires_grad_()
    ss = (0.2 * torch.randn(B, 2 * C, generator=g)).double().requires_grad_()
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = x.var(dim=(1, 2), unbiased=False, keepdim=True)
    xh = (x - mean) / torch.sqrt(var + 1e-5)
    z = xh * gamma + beta
    if film:
        z = z * (ss[:, None, :C] + 1) + ss[:, None, C:]
    y = F.silu(z)
    grads = torch.autograd.grad(y, [gamma, beta] + ([ss] if film else []), dy)
    stats = torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=1).cuda()   # double [B, 2]
    P = torch.zeros(B, 2, C, device="cuda")
    E = 2 * C
    t_index = torch.arange(B, dtype=torch.int32).cuda()
    ssd = ss.detach().float().cuda().contiguous()
    dyd, xd_, gd_, bd_ = (dy.float().cuda().reshape(B * HW, C), x.float().cuda().reshape(B * HW, C), gamma.detach().float().cuda(),
                          beta.detach().float().cuda())   # keep the device tensors alive across the launch
    L.check(lib.sdc_gn_param_grad(L.ptr(dyd), L.ptr(xd_), L.ptr(stats), L.ptr(gd_), L.ptr(bd_),
                                  L.ptr(ssd) if film else None, L.ptr(t_index) if film else None, E if film else 0, L.ptr(P),
                                  B, HW, C, L.stream_ptr()))
    p0, p1 = P[:, 0].cpu().double(), P[:, 1].cpu().double()
    sc1 = (ss.detach()[:, :C] + 1) if film else torch.ones(B, C, dtype=torch.float64)
    assert rel((sc1 * p1).sum(0), grads[0]) < 1e-4
    assert rel((sc1 * p0).sum(0), grads[1]) < 1e-4
    if film:
        d_ss = torch.cat([gamma.detach() * p1 + beta.detach() * p0, p0], dim=1)
        assert rel(d_ss, grads[2]) < 1e-4


def test_layernorm_gain_grad_vs_autograd():
    L, lib, U = _L()
    g = torch.Generator().manual_seed(13)
    for M, C, half in ((2048, 128, 1), (300, 256, 0), (64, 1024, 1), (500, 64, 0)):
        x = torch.randn(M, C, generator=g).to(torch.float16 if half else torch.float32)
        dy = torch.randn(M, C, generator=g)
        gain = torch.ones(C, dtype=torch.float64, requires_grad=True)
        xd = x.double()
        xh = (xd - xd.mean(1, keepdim=True)) * torch.rsqrt(xd.var(1, unbiased=False, keepdim=True) + 1e-5)
        (want,) = torch.autograd.grad(xh * gain, gain, dy.double())
        dg, dyc, xc = torch.zeros(C, device="cuda"), dy.cuda(), x.cuda()
        L.check(lib.sdc_channel_layernorm_gain_grad(L.ptr(dyc), L.ptr(xc), half, L.ptr(dg), M, C, L.stream_ptr()))
        assert rel(dg.cpu(), want) < 1e-4, (M, C, half)


# ------------------------------------------------------------------------------------------------ whole denoiser
def _net(precision):
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    net.precision = precision
    return net


def _check_digest(net, gold, tol_samples, tol_norm):
    worst = (0.0, None)
    names = [k[:-5] for k in gold.files if k.endswith("|norm")]
    assert set(names) == {n for n, _ in net.named_parameters()}
    for name, p in net.named_parameters():
        assert p.grad is not None, f"{name}: no gradient"
        ref_norm = float(gold[name + "|norm"])
        flat = p.grad.detach().double().flatten().cpu()
        idx = torch.linspace(0, flat.numel() - 1, fx.PGRAD_SAMPLES).round().long()
        ref_s = torch.from_numpy(gold[name + "|samples"])
        # sample error relative to the RMS entry of the reference gradient
        scale = max(ref_norm / np.sqrt(flat.numel()), 1e-30)
        e_s = ((flat[idx] - ref_s).norm() / np.sqrt(len(idx)) / scale).item()
        e_n = abs(flat.norm().item() / max(ref_norm, 1e-30) - 1.0)
        if max(e_s, e_n) > worst[0]:
            worst = (max(e_s, e_n), name)
        assert e_s < tol_samples and e_n < tol_norm, (name, e_s, e_n)
    return worst


@pytest.mark.parametrize("precision", ["f16", "tf32"])
def test_unet_parameter_gradients_vs_reference(precision, golden):
    """d<eps, g>/d(theta) for every one of the 276 parameters == the reference's autograd."""
    gold = golden("unet_dim64_pgrad")
    net = _net(precision)
    x, t = fx.unet_inputs(2)
    g = fx.unet_cotangent(2)
    eps = net(x.cuda(), t.cuda())
    assert eps.requires_grad
    ref = torch.from_numpy(gold["eps"])
    assert rel(eps.detach().cpu(), ref) < 1.5e-3
    eps.backward(g.cuda())
    worst = _check_digest(net, gold, tol_samples=1e-2, tol_norm=5e-3)
    print("worst parameter-gradient error", worst)


def test_finetune_last_step_gradients_vs_reference(golden):
    """The reference's enable_grad last step (model_predictions under autograd at t = 4) + finetune_step's loss."""
    import safediffcon_b200 as s
    gold = golden("unet_dim64_ft_pgrad")
    net = _net("f16")
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=200, ddim_sampling_eta=1.0, temporal=True,
                             use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
    table, times, rows = gd._coef_table(0, None)
    assert times[-1] == 4
    img = fx.last_step_state(2).cuda()
    # the reference's call style: an opaque lambda around get_finetune_guidance
    x0 = gd._last_step_with_grad(img, times[-1], rows[-1], dict(nablaJ=lambda x: s.get_finetune_guidance(cfg, x, 0.0)))
    assert x0.requires_grad
    assert (x0.detach().cpu() - torch.from_numpy(gold["x_start"])).abs().max() < 1e-3
    loss = fx.finetune_loss(x0, Q=0.0)
    assert abs(loss.item() / float(gold["loss"]) - 1) < 2e-2
    loss.backward()
    worst = _check_digest(net, gold, tol_samples=2e-2, tol_norm=1e-2)
    print("worst fine-tuning gradient error", worst)


def test_sample_enable_grad_feeds_an_optimizer_step(monkeypatch):
    """sample(enable_grad=True) returns an x0 with an autograd graph (graph-replayed and eager chains alike); an AdamW step
    on the denoiser changes the next chain (packed weights follow the parameters)."""
    import safediffcon_b200 as s
    import safediffcon_b200.diffusion as D
    net = _net("f16")
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=4, ddim_sampling_eta=1.0, temporal=True,
                             use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    cfg = types.SimpleNamespace(use_max_safety=True, u_bound=0.8, guidance_weights={"w_score": 500.0})
    u_init, u_final, _ = fx.chain_conditions(2)
    kw = dict(batch_size=2, u_init=u_init.cuda(), u_final=u_final.cuda(), guidance_u0=True, nablaJ=s.safety_guidance(cfg, 0.0), seed=5)
    plain = gd.sample(enable_grad=False, **kw)
    outs = []
    for max_batch in (256, 0):
        monkeypatch.setattr(D, "GRAPH_MAX_BATCH", max_batch)
        out = gd.sample(enable_grad=True, **kw)
        assert out.requires_grad and out.grad_fn is not None
        assert (out.detach() - plain).abs().max() < 2e-3
        outs.append(out)
    assert torch.equal(outs[0].detach(), outs[1].detach())
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    loss = (outs[0] * torch.randn_like(plain)).mean()   # any differentiable objective (the hinge of a random-init chain saturates)
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    assert sum(p.grad.abs().sum().item() for p in net.parameters()) > 0
    opt.step()
    after = gd.sample(enable_grad=False, **kw)
    assert not torch.equal(after, plain)
    # frozen parameters (the EMA copy the reference samples from): no graph
    for p in net.parameters():
        p.requires_grad_(False)
    assert not gd.sample(enable_grad=True, **kw).requires_grad


def test_training_loss_and_gradients_vs_reference(golden):
    """GaussianDiffusion.p_losses (the loss PostTrainPipeline re-weights, post_train.py:206-260) and its parameter gradients."""
    import safediffcon_b200 as s
    gold = golden("unet_dim64_ploss")
    net = _net("f16")
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=200, ddim_sampling_eta=1.0, temporal=True,
                             use_conv2d=True, is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    x0 = fx.calibration_states(2).cuda()
    t = torch.tensor([417, 3]).cuda()
    noise = fx.chain_noise(2, 1, seed=9)[0].cuda()
    per_sample = gd.p_losses(x0.clone(), t, noise=noise.clone(), mean=False)
    assert rel(per_sample.detach().cpu(), torch.from_numpy(gold["per_sample"])) < 1e-3
    loss = gd.p_losses(x0.clone(), t, noise=noise.clone(), mean=True)
    assert abs(loss.item() / float(gold["loss"]) - 1) < 1e-3
    net.zero_grad()
    loss.backward()
    worst = _check_digest(net, gold, tol_samples=1e-2, tol_norm=5e-3)
    print("worst training-loss gradient error", worst)
    # forward(img) draws t itself and returns a scalar with a graph
    out = gd(x0)
    assert out.ndim == 0 and out.requires_grad
