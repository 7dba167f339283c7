"""GPU parity of the backward-data pass (VJP w.r.t. the denoiser input): every backward kernel against torch autograd of
the same op, the dgrad convolutions against autograd of F.conv2d, and the whole U-Net against the unmodified
reference's input gradient (tests/golden/unet_dim*_vjp.npz).  Gradients travel as TF32-rounded fp32 (10-bit mantissa
operands, fp32 accumulation): tolerances are relative to the largest reference magnitude."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
TF32 = 0


def tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def nhwc_rows(x):
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous()


def from_rows(r, b, h, w):
    return r.reshape(b, h, w, -1).permute(0, 3, 1, 2)


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def relmax(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _L():
    from safediffcon_b200 import _lib as L
    from safediffcon_b200 import unet as U  # noqa: F401  (registers signatures)
    return L, L.lib(), U


DGRAD_CASES = [
    # kind, B, H, W (output of the FORWARD conv), cin, cout
    (1, 2, 16, 128, 128, 128),   # row kernel shape
    (1, 3, 8, 64, 96, 64),
    (1, 2, 2, 16, 64, 256),
    (0, 2, 8, 64, 64, 384),
    (0, 3, 4, 32, 256, 128),
    (2, 2, 8, 64, 32, 64),       # unshuffle: forward input 16x128 with 32 channels
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[f"k{c[0]}_B{c[1]}_{c[2]}x{c[3]}_{c[4]}to{c[5]}" for c in DGRAD_CASES])
def test_conv_dgrad_vs_autograd(case):
    L, lib, U = _L()
    kind, B, H, W, cin, cout = case
    g = torch.Generator().manual_seed(cin + cout + H)
    ksz = 3 if kind == 1 else 1
    wcin = cin * 4 if kind == 2 else cin
    w = (torch.randn(cout, wcin, ksz, ksz, generator=g) / np.sqrt(wcin * ksz * ksz)).cuda()
    hin, win = (2 * H, 2 * W) if kind == 2 else (H, W)
    x = torch.randn(B, cin, hin, win, generator=g).cuda().requires_grad_()
    gy = tf32(torch.randn(B, cout, H, W, generator=g)).cuda()
    wq = tf32(w.cpu()).cuda()
    if kind == 2:
        xin = x.reshape(B, cin, H, 2, W, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, cin * 4, H, W)
        y = F.conv2d(xin.double(), wq.double())
    else:
        y = F.conv2d(x.double(), wq.double(), padding=ksz // 2)
    (ref,) = torch.autograd.grad(y, x, gy.double())
    wt = U.pack_conv_weight_dgrad(kind, w)
    gy_rows = nhwc_rows(gy)
    if kind == 2:
        t = torch.empty(B * H * W, 4 * cin).cuda()
        U.conv_gemm(0, gy_rows, cout, None, 0, wt, None, None, t, None, False, B, H, W, 4 * cin, TF32)
        addend = torch.randn(B * 4 * H * W, cin, generator=g).cuda()
        out = torch.empty(B * 4 * H * W, cin).cuda()
        L.check(lib.sdc_pixel_shuffle_bwd(L.ptr(t), L.ptr(addend), L.ptr(out), B, H, W, cin, 0, L.stream_ptr()))
        got = from_rows(out - addend, B, 2 * H, 2 * W)
    else:
        # split the input channels in two segments (what a concat input needs) and add a residual to the first
        half = cin // 2 if (cin // 2) % 32 == 0 else cin
        res = torch.randn(B * H * W, half, generator=g).cuda()
        outs = []
        for lo, hi, r in ((0, half, res), (half, cin, None)):
            if hi == lo:
                continue
            o = torch.empty(B * H * W, hi - lo).cuda()
            U.conv_gemm(kind, gy_rows, cout, None, 0, wt[lo:hi], None, r, o, None, False, B, H, W, hi - lo, TF32)
            outs.append(o - r if r is not None else o)
        got = from_rows(torch.cat(outs, 1), B, H, W)
    assert relmax(got.double(), ref) < 5e-5, relmax(got.double(), ref)


def test_gn_silu_bwd_vs_autograd():
    L, lib, U = _L()
    for B, HW, C, film in ((3, 2048, 128, True), (4, 32, 1024, False), (2, 512, 32, True)):
        g = torch.Generator().manual_seed(C)
        x = (torch.randn(B, C, HW, generator=g) * 2 + 0.7).cuda().requires_grad_()
        gamma, beta = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
        table = torch.randn(5, 3 * C, generator=g).cuda()
        tidx = torch.tensor([4, 0, 2, 1][:B], dtype=torch.int32).cuda()
        dy = torch.randn(B * HW, C, generator=g).cuda()
        y = F.group_norm(x, 1, gamma, beta, eps=1e-5)
        if film:
            ss = table[tidx.long()]
            y = y * (ss[:, :C, None] + 1) + ss[:, C:2 * C, None]
        y = F.silu(y).permute(0, 2, 1).reshape(B * HW, C)
        (ref,) = torch.autograd.grad(y, x, dy)
        ref = ref.permute(0, 2, 1).reshape(B * HW, C)
        xr = x.detach().permute(0, 2, 1).reshape(B * HW, C).contiguous()
        stats = torch.stack([xr.double().reshape(B, -1).sum(1), (xr.double() ** 2).reshape(B, -1).sum(1)], 1).contiguous()
        sums = torch.empty(B, 2, dtype=torch.float64).cuda()
        dx = torch.empty_like(xr)
        L.check(lib.sdc_gn_silu_bwd(L.ptr(dy), L.ptr(xr), L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(table) if film else None,
                                    L.ptr(tidx) if film else None, 3 * C if film else 0, L.ptr(sums), L.ptr(dx), B, HW, C, L.stream_ptr()))
        assert relmax(dx, ref) < 1e-3, relmax(dx, ref)   # output is TF32-rounded (2^-11 relative)
        assert rel(dx, ref) < 5e-4


def test_channel_layernorm_bwd_vs_autograd():
    L, lib, U = _L()
    for M, C in ((1000, 128), (77, 1024), (64, 32), (10, 512), (33, 256)):
        g = torch.Generator().manual_seed(C)
        for half in (0, 1):
            x0 = (torch.randn(M, C, generator=g) * 3 + 1)
            xs = x0.half().cuda() if half else x0.cuda()
            x = xs.float().requires_grad_()
            gain = torch.randn(C, generator=g).cuda()
            dy, add = torch.randn(M, C, generator=g).cuda(), torch.randn(M, C, generator=g).cuda()
            y = (x - x.mean(1, keepdim=True)) * (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * gain
            (ref,) = torch.autograd.grad(y, x, dy)
            dx = torch.empty(M, C).cuda()
            L.check(lib.sdc_channel_layernorm_bwd(L.ptr(dy), L.ptr(xs), half, L.ptr(gain), L.ptr(add), L.ptr(dx), M, C, 0, L.stream_ptr()))
            assert relmax(dx - add, ref) < 2e-5, (M, C, half, relmax(dx - add, ref))
            L.check(lib.sdc_channel_layernorm_bwd(L.ptr(dy), L.ptr(xs), half, L.ptr(gain), None, L.ptr(dx), M, C, 1, L.stream_ptr()))
            assert torch.equal(dx, tf32(dx.cpu()).cuda()) and relmax(dx, ref) < 1e-3


def test_attention_bwd_vs_autograd():
    L, lib, U = _L()
    for B, n in ((3, 2048), (2, 512), (5, 32), (2, 96)):
        g = torch.Generator().manual_seed(n)
        qkv = (torch.randn(B * n, 384, generator=g) * 1.5).cuda().requires_grad_()
        dout = torch.randn(B * n, 128, generator=g).cuda()
        q, k, v = (t.reshape(B, n, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=1))  # b h d n
        ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(-1), v)
        out = torch.einsum("bhde,bhdn->bhen", ctx, q.softmax(-2) * 32 ** -0.5).permute(0, 3, 1, 2).reshape(B * n, 128)
        (ref,) = torch.autograd.grad(out, qkv, dout)
        fwd = torch.empty(B * n, 128).cuda()
        ws = torch.empty(lib.sdc_linear_attention_workspace(B), dtype=torch.uint8).cuda()
        L.check(lib.sdc_linear_attention(TF32, L.ptr(qkv.detach()), L.ptr(fwd), L.ptr(ws), B, n, L.stream_ptr()))
        wsb = torch.empty(lib.sdc_linear_attention_bwd_workspace(B), dtype=torch.uint8).cuda()
        dqkv = torch.empty(B * n, 384).cuda()
        L.check(lib.sdc_linear_attention_bwd(L.ptr(qkv.detach()), L.ptr(dout), L.ptr(ws), L.ptr(wsb), L.ptr(dqkv), B, n, L.stream_ptr()))
        assert relmax(dqkv, ref) < 1e-3 and rel(dqkv, ref) < 5e-4, (n, relmax(dqkv, ref), rel(dqkv, ref))
    for B, n in ((4, 32), (1, 32), (3, 20)):
        g = torch.Generator().manual_seed(n + B)
        qkv = (torch.randn(B * n, 384, generator=g) * 1.5).cuda().requires_grad_()
        dout = torch.randn(B * n, 128, generator=g).cuda()
        q, k, v = (t.reshape(B, n, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=1))
        attn = torch.einsum("bhdi,bhdj->bhij", q * 32 ** -0.5, k).softmax(-1)
        out = torch.einsum("bhij,bhdj->bhid", attn, v).permute(0, 2, 1, 3).reshape(B * n, 128)
        (ref,) = torch.autograd.grad(out, qkv, dout)
        dqkv = torch.empty(B * n, 384).cuda()
        L.check(lib.sdc_attention_bwd(L.ptr(qkv.detach()), L.ptr(dout), L.ptr(dqkv), B, n, L.stream_ptr()))
        assert relmax(dqkv, ref) < 1e-3 and rel(dqkv, ref) < 5e-4, (n, relmax(dqkv, ref), rel(dqkv, ref))


def test_small_bwd_ops():
    L, lib, U = _L()
    g = torch.Generator().manual_seed(5)
    # nearest-upsample backward = 2x2 block sums
    dy = torch.randn(2 * 8 * 16, 64, generator=g).cuda()
    dx = torch.empty(2 * 4 * 8, 64).cuda()
    L.check(lib.sdc_upsample2x_bwd(L.ptr(dy), L.ptr(dx), 2, 4, 8, 64, 0, L.stream_ptr()))
    ref = nhwc_rows(F.avg_pool2d(from_rows(dy, 2, 8, 16), 2) * 4)
    assert torch.allclose(dx, ref, atol=1e-5)
    # add in place
    a, b = torch.randn(4096, generator=g).cuda(), torch.randn(4096, generator=g).cuda()
    want = a + b
    L.check(lib.sdc_add_inplace(L.ptr(a), L.ptr(b), 4096, 0, L.stream_ptr()))
    assert torch.equal(a, want)
    # head backward
    gout = torch.randn(3, 3, 16, 128, generator=g).cuda()
    w = torch.randn(3, 128, generator=g).cuda()
    dxh = torch.empty(3 * 2048, 128).cuda()
    L.check(lib.sdc_head_conv1_bwd(L.ptr(gout), L.ptr(w), L.ptr(dxh), 3, 2048, 128, 3, 0, L.stream_ptr()))
    ref = gout.reshape(3, 3, 2048).permute(0, 2, 1).reshape(-1, 3) @ w
    assert torch.allclose(dxh, ref, atol=1e-5)
    # stem backward: 1x1 GEMM + col2im against autograd of the 7x7 convolution
    for cout in (128, 32):
        x = torch.randn(2, 3, 16, 128, generator=g).cuda().requires_grad_()
        ws = (torch.randn(cout, 3, 7, 7, generator=g) * 0.1).cuda()
        gy = tf32(torch.randn(2, cout, 16, 128, generator=g)).cuda()
        wq = tf32(ws.cpu()).cuda()
        (ref,) = torch.autograd.grad(F.conv2d(x.double(), wq.double(), padding=3), x, gy.double())
        kp = 192
        wt = torch.zeros(kp, cout).cuda()
        L.check(lib.sdc_pack_conv_weight_dgrad(0, L.ptr(ws.reshape(cout, 147).contiguous()), L.ptr(wt), cout, 147, L.stream_ptr()))
        t = torch.empty(2 * 2048, kp).cuda()
        U.conv_gemm(0, nhwc_rows(gy), cout, None, 0, wt, None, None, t, None, False, 2, 16, 128, kp, TF32)
        gx = torch.empty(2, 3, 16, 128).cuda()
        L.check(lib.sdc_stem_col2im(L.ptr(t), L.ptr(gx), 2, 3, 16, 128, kp, L.stream_ptr()))
        assert relmax(gx.double(), ref) < 5e-5


@pytest.mark.parametrize("dim,precision", [(32, "tf32"), (64, "f16"), (64, "tf32")])
def test_unet_input_gradient_vs_reference_golden(dim, precision, golden):
    """d<eps, g>/dx of the whole denoiser against the unmodified reference's autograd (seed-42 weights)."""
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    net.precision = precision
    x, t = fx.unet_inputs(2)
    gct = fx.unet_cotangent(2)
    gold = golden(f"unet_dim{dim}_vjp")
    ref_eps, ref_gx = torch.from_numpy(gold["eps"]), torch.from_numpy(gold["grad_x"])
    # autograd path (what a guidance callable that differentiates through the model uses)
    xg = x.cuda().requires_grad_()
    eps = net(xg, t.cuda())
    assert eps.requires_grad
    (gx,) = torch.autograd.grad(eps, xg, gct.cuda())
    assert rel(eps.detach().cpu(), ref_eps) < 1e-3
    r = rel(gx.cpu(), ref_gx)
    per = [rel(gx[i].cpu(), ref_gx[i]) for i in range(2)]
    assert r < 3e-3 and max(per) < 3e-3, (r, per)
    # trainable parameters route autograd through the full backward (FiLM rows by torch); frozen parameters (an EMA copy) through
    # the input-only VJP (Python schedule), which must be bit-identical to the explicit API -- the ONE-call C executor
    # (sdc_unet_backward_data) once both read the same FiLM rows (the executor's default table comes from a tensor-core GEMM)
    plan = net._plan_ready(backward=True)
    if plan is not None:
        plan.set_flag(plan.FILM_TC, 0)
        net.invalidate_packed()
    eps2, gx2 = net.vjp(x.cuda(), t.cuda(), gct.cuda())
    # (the two paths evaluate exp/sin of the time embedding with different libm's: 1-ulp frequency differences times t <= 999)
    assert rel(gx2, gx) < 2e-3 and rel(eps2, eps.detach()) < 1e-3
    for p in net.parameters():
        p.requires_grad_(False)
    xf = x.cuda().requires_grad_()
    eps_f = net(xf, t.cuda())
    (gx_f,) = torch.autograd.grad(eps_f, xf, gct.cuda())
    # same kernels, same order; only the fp64 atomics of the GroupNorm statistics may land in a different order
    assert rel(gx2, gx_f) < 1e-5 and rel(eps2, eps_f.detach()) < 2e-6, (rel(gx2, gx_f), rel(eps2, eps_f.detach()))
    with torch.no_grad():   # the inference path (fused attention, reused buffers) agrees with the recording path to rounding
        net.compact_intermediates = False
        assert rel(net(x.cuda(), t.cuda()), eps.detach()) < 1e-3
        # fp16 norm inputs (FP16 inference default) add one independent 2^-11 rounding per normalised tensor on this path only
        net.compact_intermediates = True
        assert rel(net(x.cuda(), t.cuda()), eps.detach()) < (1.5e-3 if precision == "f16" else 1e-3)
    # linearity of the VJP in the cotangent (size-independent property)
    _, gx3 = net.vjp(x.cuda(), t.cuda(), -2.0 * gct.cuda())
    assert rel(gx3, -2.0 * gx) < 2e-3


def test_guidance_through_the_model_uses_the_vjp():
    """A user nablaJ that differentiates a loss of eps_theta(x0_hat) w.r.t. its argument (the contract of
    GaussianDiffusion.sample's nablaJ, reference diffusion.py:254-262) gets the input gradient from the CUDA backward."""
    import safediffcon_b200 as s
    from oracle import unet_ref
    torch.manual_seed(42)
    net = s.Unet2D(dim=32, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    x, t = fx.unet_inputs(2)
    t = torch.full((2,), 417, dtype=torch.long)

    def nablaJ(model):
        def fn(x0):
            e = model(x0)
            J = (e[:, 0, :11] ** 2).mean() + 0.3 * e[:, 2, :11].mean()   # smooth: the cotangent 2e/N inherits the 1e-3 of eps
            return torch.autograd.grad(J, x0)[0]
        return fn

    xg = x.cuda().requires_grad_()
    got = nablaJ(lambda v: net(v, t.cuda()))(xg)
    xc = x.clone().requires_grad_()
    want = nablaJ(lambda v: unet_ref.unet_forward(sd, v, t))(xc)
    assert rel(got.cpu(), want) < 6e-3   # forward eps error (1e-3, enters through the cotangent) + backward rounding (3e-3)
