"""GPU parity: tcgen05 convolution and the fused U-Net kernels vs torch fp32, and the whole denoiser vs the
reference's eps (tests/golden/unet_*.npz), in both operand precisions (TF32 containers / FP16: 10-bit mantissa,
round to nearest, fp32 accumulation).  North-star bound for the whole network: eps within 1e-3 relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures as fx
from oracle import unet_ref

pytestmark = pytest.mark.gpu


def tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


TF32, F16 = 0, 1


def quant(x, prec):
    """Round to the operand precision, result kept in fp32."""
    return x.half().float() if prec == F16 else tf32(x)


def as_operand(x, prec):
    """fp32 tensor holding operand-representable values -> tensor of the operand dtype."""
    return x.half() if prec == F16 else x


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(autouse=True)
def _fp32_reference():
    torch.set_grad_enabled(True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


CONV_CASES = [
    # kind, B, H, W, c0, c1, cout, bias, residual, stats, round
    (1, 2, 16, 128, 32, 0, 32, True, False, True, False),
    (1, 3, 16, 128, 128, 0, 128, True, False, True, False),
    (1, 2, 8, 64, 64, 64, 128, True, False, True, False),     # two K segments (skip concat)
    (1, 5, 2, 16, 96, 32, 64, True, False, True, True),        # ragged batch: 5 images, 4 per tile
    (1, 4, 4, 32, 256, 0, 512, True, False, False, True),      # two N tiles
    (0, 2, 16, 128, 128, 0, 384, False, False, False, False),  # qkv projection, no bias, N tile 128
    (0, 4, 2, 16, 128, 0, 256, True, True, False, True),       # 1x1 + residual epilogue
    (0, 2, 8, 64, 64, 32, 96, True, False, False, False),      # 1x1 on a concat, N tile 32
    (2, 2, 8, 64, 32, 0, 64, True, False, False, True),        # pixel-unshuffle + 1x1 (input 16x128)
    (2, 3, 2, 16, 128, 0, 256, True, False, False, True),      # unshuffle to the 2x16 level
    (1, 3, 16, 128, 128, 128, 128, True, False, True, False),  # row kernel: two segments (256 -> 128)
    (1, 2, 5, 128, 64, 0, 96, True, True, True, True),         # row kernel: odd H (half-empty last pair), residual
    (1, 150, 16, 128, 32, 0, 64, True, False, True, False),    # row kernel: more tile pairs than SMs
    (1, 40, 8, 64, 64, 32, 128, True, False, True, False),     # CTA-pair kernel (cta_group::2): 160 M tiles, 2 segments
    (1, 149, 4, 32, 64, 0, 256, True, True, True, True),       # CTA-pair: odd tile count (phantom half), N=256, residual
    (0, 80, 8, 64, 128, 0, 384, False, False, False, False),   # CTA-pair 1x1, three N tiles
    (2, 598, 2, 16, 64, 0, 64, True, False, False, True),      # CTA-pair unshuffle, 4 images per tile, ragged last tile
    (1, 600, 2, 16, 32, 32, 96, True, False, True, False),     # CTA-pair, N tile 32 -> half tiles of 16 rows
    (1, 80, 6, 128, 32, 32, 64, True, True, True, True),       # row kernel, CTA pairs, H=6: phantom rows in the 2nd CTA
    (1, 160, 16, 128, 128, 0, 128, True, False, True, False),  # row kernel, CTA pairs, production shape
    (1, 150, 8, 64, 128, 0, 128, True, False, True, True),     # W = 64 row kernel (conv_row64.cu): production shape, 2+ items per cluster
    (1, 80, 8, 64, 64, 64, 96, True, True, True, True),        # W = 64 row kernel: two segments, residual, N = 96
    (1, 75, 16, 64, 32, 32, 64, False, False, True, False),    # W = 64 row kernel: two 8-row items per image, TF32 chunks of 32
]


CONV_PARAMS = [(c, TF32) for c in CONV_CASES] + [(c, F16) for c in CONV_CASES if c[4] % 64 == 0 and c[5] % 64 == 0]


@pytest.mark.parametrize("case,prec", CONV_PARAMS,
                         ids=[f"{'f16' if pr else 'tf32'}_k{c[0]}_B{c[1]}_{c[2]}x{c[3]}_c{c[4]}+{c[5]}_o{c[6]}" for c, pr in CONV_PARAMS])
def test_conv_gemm_vs_torch(case, prec):
    from safediffcon_b200 import unet as U
    kind, B, H, W, c0, c1, cout, use_bias, use_res, use_stats, rnd = case
    g = torch.Generator().manual_seed(H * W + c0 + cout)
    hin, win = (2 * H, 2 * W) if kind == 2 else (H, W)
    cin = c0 + c1
    x = quant(torch.randn(B, cin, hin, win, generator=g), prec).cuda()
    ksz = {0: 1, 1: 3, 2: 1}[kind]
    w = (torch.randn(cout, cin * (4 if kind == 2 else 1), ksz, ksz, generator=g) / np.sqrt(cin * ksz * ksz)).cuda()
    bias = torch.randn(cout, generator=g).cuda() if use_bias else None
    res = quant(torch.randn(B * H * W, cout, generator=g), prec).cuda() if use_res else None
    a0 = as_operand(nhwc(x[:, :c0]), prec)
    a1 = as_operand(nhwc(x[:, c0:]), prec) if c1 else None
    wp = U.pack_conv_weight(kind, w, prec)
    out = torch.full((B * H * W, cout), float("nan"), dtype=U.operand_dtype(prec) if rnd else torch.float32).cuda()
    stats = torch.zeros(B, 2, dtype=torch.float64).cuda() if use_stats else None
    U.conv_gemm(kind, a0, c0, a1, c1, wp, bias, None if res is None else as_operand(res, prec), out, stats, rnd, B, H, W, cout, prec)
    torch.cuda.synchronize()
    out = out.float()
    wq = quant(w.cpu(), prec).cuda()
    if kind == 2:
        xin = x.reshape(B, cin, H, 2, W, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, cin * 4, H, W)
        ref = F.conv2d(xin.double(), wq.double(), None if bias is None else bias.double())
    else:
        ref = F.conv2d(x.double(), wq.double(), None if bias is None else bias.double(), padding=ksz // 2)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, cout)
    if res is not None:
        ref = ref + res.double()
    assert torch.isfinite(out).all()
    err = (out.double() - ref).abs().max().item()
    tol = 2e-3 if rnd else 2e-5   # fp32 accumulation of exact tf32 products; rounding the output costs 2^-11 relative
    assert err < tol * max(1.0, ref.abs().max().item()), err
    if rnd:
        assert torch.equal(out, quant(out.cpu(), prec).cuda())  # stored values are representable in the operand precision
    if stats is not None:
        o = out.double().reshape(B, -1)
        # statistics are taken on the fp32 values BEFORE the optional TF32 rounding of the stored tensor
        rt, at = (1e-3, 0.5) if rnd else (1e-6, 1e-3)
        assert torch.allclose(stats[:, 0], o.sum(1), rtol=rt, atol=at)
        assert torch.allclose(stats[:, 1], (o * o).sum(1), rtol=rt, atol=at)


def _L():
    from safediffcon_b200 import _lib as L
    return L, L.lib()


@pytest.mark.parametrize("prec", [TF32, F16])
def test_stem_conv7_vs_torch(prec):
    L, lib = _L()
    from safediffcon_b200 import unet as U
    for B, cout in ((2, 128), (3, 32)):
        g = torch.Generator().manual_seed(cout)
        x = torch.randn(B, 3, 16, 128, generator=g).cuda()
        w = (torch.randn(cout, 3, 7, 7, generator=g) * 0.1).cuda()
        b = torch.randn(cout, generator=g).cuda()
        out = torch.empty(B * 16 * 128, cout, dtype=U.operand_dtype(prec)).cuda()
        L.check(lib.sdc_stem_conv7(prec, L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(out), B, 3, 16, 128, cout, L.stream_ptr()))
        out = out.float()
        ref = nhwc(F.conv2d(x, w, b, padding=3)).reshape(-1, cout)
        assert (out - ref).abs().max().item() < 1.5e-3 * ref.abs().max().item()  # output rounded to the operand precision
        assert (out - quant(ref.cpu(), prec).cuda()).abs().max().item() < 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("prec", [TF32, F16])
def test_stem_tensor_core_vs_torch(prec):
    """im2col (high | low split of the fp32 input) + tcgen05 1x1 GEMM == 7x7 pad-3 convolution with rounded weights."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    for B, cout in ((2, 128), (3, 64)):
        g = torch.Generator().manual_seed(cout + 1)
        x = (torch.randn(B, 3, 16, 128, generator=g) * 1.3).cuda()
        w = (torch.randn(cout, 3, 7, 7, generator=g) * 0.1).cuda()
        b = torch.randn(cout, generator=g).cuda()
        kp = 320
        wrep = torch.zeros(cout, kp).cuda()
        wrep[:, :147] = w.reshape(cout, 147)
        wrep[:, 160:307] = w.reshape(cout, 147)
        wp = U.pack_conv_weight(0, wrep.reshape(cout, kp, 1, 1), prec)
        patches = torch.full((B * 2048, kp), float("nan"), dtype=U.operand_dtype(prec)).cuda()
        L.check(lib.sdc_stem_im2col(prec, L.ptr(x), L.ptr(patches), B, 3, 16, 128, kp, L.stream_ptr()))
        pf = patches.float()
        assert torch.isfinite(pf).all() and (pf[:, 147:160] == 0).all() and (pf[:, 307:] == 0).all()
        # high + low reproduces the fp32 patch matrix to ~2^-22
        ref_p = F.unfold(x, 7, padding=3).permute(0, 2, 1).reshape(B * 2048, 147)
        assert ((pf[:, :147] + pf[:, 160:307]) - ref_p).abs().max().item() < 4e-6 * ref_p.abs().max().item()
        out = torch.empty(B * 2048, cout).cuda()
        U.conv_gemm(0, patches, kp, None, 0, wp, b, None, out, None, False, B, 16, 128, cout, prec)
        ref = nhwc(F.conv2d(x.double(), quant(w.cpu(), prec).cuda().double(), b.double(), padding=3)).reshape(-1, cout)
        assert (out.double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("B", [1, 3, 37, 300])
def test_stem_one_kernel_vs_torch(B):
    """sdc_stem_conv7_tc (operand tile built in shared memory by the kernel itself, resident weights) == the 7x7 pad-3 convolution
    with fp16-rounded weights on the fp32 input (high + low split), fp16 NHWC output, and == the im2col + GEMM pair."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    cout, kp = 128, 320
    g = torch.Generator().manual_seed(B)
    x = (torch.randn(B, 3, 16, 128, generator=g) * 1.3).cuda()
    w = (torch.randn(cout, 3, 7, 7, generator=g) * 0.1).cuda()
    b = torch.randn(cout, generator=g).cuda()
    wrep = torch.zeros(cout, kp).cuda()
    wrep[:, :147] = w.reshape(cout, 147)
    wrep[:, 160:307] = w.reshape(cout, 147)
    wp = U.pack_conv_weight(0, wrep.reshape(cout, kp, 1, 1), F16)
    out = torch.full((B * 2048, cout), float("nan"), dtype=torch.float16).cuda()
    assert lib.sdc_stem_conv7_tc(L.ptr(x), L.ptr(wp), L.ptr(b), L.ptr(out), B, 3, 16, 128, cout, kp, L.stream_ptr()) == 0
    torch.cuda.synchronize()
    ref = nhwc(F.conv2d(x.double(), quant(w.cpu(), F16).cuda().double(), b.double(), padding=3)).reshape(-1, cout)
    assert torch.isfinite(out.float()).all()
    assert (out.double() - ref).abs().max().item() < 6e-4 * ref.abs().max().item()      # fp16 rounding of the output only
    patches = torch.empty(B * 2048, kp, dtype=torch.float16).cuda()
    L.check(lib.sdc_stem_im2col(F16, L.ptr(x), L.ptr(patches), B, 3, 16, 128, kp, L.stream_ptr()))
    old = torch.empty(B * 2048, cout, dtype=torch.float16).cuda()
    U.conv_gemm(0, patches, kp, None, 0, wp, b, None, old, None, True, B, 16, 128, cout, F16)
    assert (out.float() - old.float()).abs().max().item() <= 2 ** -9 * ref.abs().max().item()   # same products, fp32 sum order differs


@pytest.mark.parametrize("prec", [TF32, F16])
def test_gn_silu_vs_torch(prec):
    L, lib = _L()
    from safediffcon_b200 import unet as U
    od = U.operand_dtype(prec)
    for B, HW, C in ((3, 2048, 128), (4, 32, 1024), (2, 512, 32)):
        g = torch.Generator().manual_seed(C)
        x = (torch.randn(B, C, HW, generator=g) * 2 + 0.7).cuda()
        gamma, beta = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
        table = torch.randn(5, 3 * C, generator=g).cuda()  # rows wider than 2C: exercises the row stride
        tidx = torch.tensor([4, 0, 2, 1][:B], dtype=torch.int32).cuda()
        res = torch.randn(B * HW, C, generator=g).cuda()
        xr = x.permute(0, 2, 1).reshape(B * HW, C).contiguous()
        stats = torch.stack([xr.double().reshape(B, -1).sum(1), (xr.double() ** 2).reshape(B, -1).sum(1)], 1).contiguous()
        y = torch.empty(B * HW, C, dtype=od).cuda()
        ss = table[tidx.long()]
        ref0 = F.group_norm(x, 1, gamma, beta, eps=1e-5) * (ss[:, :C, None] + 1) + ss[:, C:2 * C, None]
        ref0 = F.silu(ref0).permute(0, 2, 1).reshape(B * HW, C)
        # fp32 residual (the 1x1 res_conv output) and operand-precision residual (the block input)
        for res_operand in (0, 1):
            r_in = as_operand(quant(res.cpu(), prec).cuda(), prec) if res_operand else res
            L.check(lib.sdc_gn_silu(prec, L.ptr(xr), 0, L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(table), L.ptr(tidx), 3 * C,
                                    L.ptr(r_in), res_operand, L.ptr(y), B, HW, C, L.stream_ptr()))
            ref = ref0 + r_in.float()
            assert (y.float() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
            if prec == F16 and res_operand and C % 8 == 0 and 256 % (C // 8) == 0:
                # compact intermediates: fp16 input (statistics from the fp32 values), also in place
                xh = xr.half()
                for dst in (y, xh):
                    L.check(lib.sdc_gn_silu(prec, L.ptr(xh), 1, L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(table), L.ptr(tidx), 3 * C,
                                            L.ptr(r_in), 1, L.ptr(dst), B, HW, C, L.stream_ptr()))
                    assert (dst.float() - ref).abs().max().item() < 3e-3 * ref.abs().max().item()
        # no FiLM, no residual, uniform row 0
        L.check(lib.sdc_gn_silu(prec, L.ptr(xr), 0, L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None, 0, None, 0, L.ptr(y), B, HW, C,
                                L.stream_ptr()))
        ref2 = F.silu(F.group_norm(x, 1, gamma, beta, eps=1e-5)).permute(0, 2, 1).reshape(B * HW, C)
        assert (y.float() - ref2).abs().max().item() < 2e-3 * ref2.abs().max().item()


def test_channel_layernorm_vs_torch():
    L, lib = _L()
    import safediffcon_b200.unet  # noqa: F401
    for M, C in ((1000, 128), (77, 1024), (64, 32), (10, 512)):
        g = torch.Generator().manual_seed(C)
        x = (torch.randn(M, C, generator=g) * 3 + 1).cuda()
        gain = torch.randn(C, generator=g).cuda()
        res = torch.randn(M, C, generator=g).cuda()
        y = torch.empty_like(x)
        L.check(lib.sdc_channel_layernorm(TF32, L.ptr(x), 1, L.ptr(gain), L.ptr(res), L.ptr(y), M, C, 0, L.stream_ptr()))
        ref = (x - x.mean(1, keepdim=True)) * (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * gain + res
        assert (y - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
        L.check(lib.sdc_channel_layernorm(TF32, L.ptr(x), 1, L.ptr(gain), None, L.ptr(y), M, C, 1, L.stream_ptr()))
        assert torch.equal(y, tf32((ref - res).cpu()).cuda()) or (y - (ref - res)).abs().max().item() < 1e-3 * ref.abs().max().item()
        # FP16 mode: fp16 or fp32 input, fp16 residual and output
        xh, rh, yh = x.half(), res.half(), torch.empty(M, C, dtype=torch.float16).cuda()
        for x_in, x_operand in ((xh, 1), (x, 0)):
            xf = x_in.float()
            refh = (xf - xf.mean(1, keepdim=True)) * (xf.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * gain + rh.float()
            L.check(lib.sdc_channel_layernorm(F16, L.ptr(x_in), x_operand, L.ptr(gain), L.ptr(rh), L.ptr(yh), M, C, 1, L.stream_ptr()))
            assert (yh.float() - refh).abs().max().item() < 1e-3 * refh.abs().max().item()


@pytest.mark.parametrize("prec", [TF32, F16])
def test_attention_cores_vs_torch(prec):
    L, lib = _L()
    from safediffcon_b200 import unet as U
    od = U.operand_dtype(prec)
    for B, n in ((3, 2048), (2, 512), (5, 32), (2, 96), (3, 128)):   # n % 32 == 0 (tiles of 64 or 32 pixels)
        g = torch.Generator().manual_seed(n)
        qkv = (torch.randn(B * n, 384, generator=g) * 1.5).cuda()
        out = torch.empty(B * n, 128, dtype=od).cuda()
        ws = torch.empty(lib.sdc_linear_attention_workspace(B), dtype=torch.uint8).cuda()
        L.check(lib.sdc_linear_attention(prec, L.ptr(qkv), L.ptr(out), L.ptr(ws), B, n, L.stream_ptr()))
        out = out.float()
        q, k, v = (t.reshape(B, n, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=1))  # b h d n
        ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(-1), v)
        ref = torch.einsum("bhde,bhdn->bhen", ctx, q.softmax(-2) * 32 ** -0.5)
        ref = ref.permute(0, 3, 1, 2).reshape(B * n, 128)
        assert (out - ref).abs().max().item() < 1.5e-3 * ref.abs().max().item()
    for B, n in ((4, 32), (1, 32), (3, 20)):
        g = torch.Generator().manual_seed(n + B)
        qkv = (torch.randn(B * n, 384, generator=g) * 1.5).cuda()
        out = torch.empty(B * n, 128, dtype=od).cuda()
        L.check(lib.sdc_attention(prec, L.ptr(qkv), L.ptr(out), B, n, L.stream_ptr()))
        out = out.float()
        q, k, v = (t.reshape(B, n, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=1))
        attn = torch.einsum("bhdi,bhdj->bhij", q * 32 ** -0.5, k).softmax(-1)
        ref = torch.einsum("bhij,bhdj->bhid", attn, v).permute(0, 2, 1, 3).reshape(B * n, 128)
        assert (out - ref).abs().max().item() < 1.5e-3 * ref.abs().max().item()


@pytest.mark.parametrize("prec", [TF32, F16])
def test_fused_linear_attention_vs_torch(prec):
    """qkv conv with q-softmax epilogue -> context -> folded per-sample output projection == the reference LinearAttention core
    + to_out conv (unet.py:202-222), for single-CTA tiles, CTA pairs, and the smallest legal level (n = 128)."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    od = U.operand_dtype(prec)
    for B, H, W, c in ((3, 16, 128, 128), (80, 8, 64, 64), (150, 4, 32, 128), (2, 4, 32, 64)):
        g = torch.Generator().manual_seed(B + c)
        n, M = H * W, B * H * W
        x = quant(torch.randn(B, c, H, W, generator=g), prec).cuda()
        wqkv = (torch.randn(384, c, 1, 1, generator=g) * (1.5 / np.sqrt(c))).cuda()
        wout = (torch.randn(c, 128, 1, 1, generator=g) / np.sqrt(128)).cuda()
        bout = torch.randn(c, generator=g).cuda()
        # fused path
        a = as_operand(nhwc(x).reshape(M, c), prec)
        qs, kv = torch.empty(M, 128, dtype=od).cuda(), torch.empty(M, 256, dtype=od).cuda()   # F16 mode: k | v in fp16
        U.conv1x1_qkv(a, c, U.pack_conv_weight(0, wqkv, prec), qs, kv, B, H, W, prec)
        ws = torch.empty(lib.sdc_linear_attention_workspace(B), dtype=torch.uint8).cuda()
        import ctypes
        L.check(lib.sdc_linear_attention_context(L.ptr(kv), ctypes.c_void_p(kv.data_ptr() + 128 * kv.element_size()), 256,
                                                 int(kv.dtype == torch.float16), L.ptr(ws), B, n, L.stream_ptr()))
        wf = torch.empty(B * c, 128, dtype=od).cuda()
        L.check(lib.sdc_linear_attention_fold(prec, L.ptr(ws), L.ptr(wout.reshape(c, 128).contiguous()), L.ptr(wf), B, c, L.stream_ptr()))
        out = torch.empty(M, c).cuda()
        U.conv1x1_per_sample(qs, 128, wf, bout, out, B, H, W, c, prec)
        # torch reference on the same (quantised) input and qkv weights
        qkv = F.conv2d(x.double(), quant(wqkv.cpu(), prec).cuda().double())
        q, k, v = (t.reshape(B, 4, 32, n) for t in qkv.chunk(3, dim=1))
        ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(-1), v)
        att = torch.einsum("bhde,bhdn->bhen", ctx, q.softmax(-2) * 32 ** -0.5).reshape(B, 128, H, W)
        ref = nhwc(F.conv2d(att, wout.double(), bout.double())).reshape(M, c)
        # intermediate checks localise failures: softmaxed q and raw k | v
        qref = nhwc((q.softmax(-2) * 32 ** -0.5).reshape(B, 128, H, W)).reshape(M, 128)
        assert (qs.double() - qref).abs().max().item() < 1e-3 * qref.abs().max().item()
        kvref = nhwc(qkv[:, 128:]).reshape(M, 256)
        assert (kv.double() - kvref).abs().max().item() < (2e-5 if kv.dtype == torch.float32 else 6e-4) * kvref.abs().max().item()
        err = (out.double() - ref).abs().max().item()
        assert err < 2e-3 * ref.abs().max().item(), (B, H, W, c, err, ref.abs().max().item())


def test_small_ops_vs_torch():
    L, lib = _L()
    import safediffcon_b200.unet  # noqa: F401
    g = torch.Generator().manual_seed(3)
    for prec, dt in ((TF32, torch.float32), (F16, torch.float16)):
        x = torch.randn(2 * 4 * 8, 64, generator=g).to(dt).cuda()
        y = torch.empty(2 * 8 * 16, 64, dtype=dt).cuda()
        L.check(lib.sdc_upsample2x(prec, L.ptr(x), L.ptr(y), 2, 4, 8, 64, L.stream_ptr()))
        ref = nhwc(F.interpolate(x.float().reshape(2, 4, 8, 64).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")).reshape(-1, 64)
        assert torch.equal(y.float(), ref)
        xin = torch.randn(3 * 2048, 128, generator=g).to(dt).cuda()
        w, b = torch.randn(3, 128, generator=g).cuda(), torch.randn(3, generator=g).cuda()
        out = torch.empty(3, 3, 16, 128).cuda()
        L.check(lib.sdc_head_conv1(prec, L.ptr(xin), L.ptr(w), L.ptr(b), L.ptr(out), 3, 2048, 128, 3, L.stream_ptr()))
        ref = (xin.float() @ w.t() + b).reshape(3, 2048, 3).permute(0, 2, 1).reshape(3, 3, 16, 128)
        assert (out - ref).abs().max().item() < 1e-4
    for act, fn in ((0, lambda v: v), (1, F.silu), (2, F.gelu)):
        xi = torch.randn(7, 512, generator=g).cuda()
        wl, bl = torch.randn(300, 512, generator=g).cuda() * 0.05, torch.randn(300, generator=g).cuda()
        yo = torch.empty(7, 300).cuda()
        L.check(lib.sdc_linear_rows(L.ptr(xi), L.ptr(wl), L.ptr(bl), L.ptr(yo), 7, 512, 300, act, L.stream_ptr()))
        assert (yo - F.linear(fn(xi), wl, bl)).abs().max().item() < 2e-5
    t = torch.tensor([0., 3., 417., 999.]).cuda()
    emb = torch.empty(4, 128).cuda()
    L.check(lib.sdc_sinusoidal_embedding(L.ptr(t), L.ptr(emb), 4, 128, 10000.0, L.stream_ptr()))
    f = torch.exp(torch.arange(64) * -(np.log(10000) / 63))
    a = t.cpu()[:, None] * f[None, :]
    assert (emb.cpu() - torch.cat((a.sin(), a.cos()), -1)).abs().max().item() < 2e-4


@pytest.mark.parametrize("dim,B,precision", [(32, 3, "tf32"), (128, 2, "f16"), (128, 2, "tf32")])
def test_unet_eps_vs_reference_golden(dim, B, precision, golden):
    """Whole denoiser, seed-42 weights, reference inputs -> reference eps within the north-star 1e-3 relative."""
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    assert net.precision == ("f16" if dim % 64 == 0 else "tf32")   # default: FP16 operands whenever the channel counts allow
    net.precision = precision
    x, t = fx.unet_inputs(B)
    torch.set_grad_enabled(False)   # the inference path (the recording path is covered by tests/test_unet_pgrad_gpu.py)
    eps = net(x.cuda(), t.cuda())
    ref = torch.from_numpy(golden(f"unet_dim{dim}")["eps"])
    assert eps.shape == ref.shape
    r = rel(eps.cpu(), ref)
    per = [rel(eps[i].cpu(), ref[i]) for i in range(B)]
    assert r < 1e-3 and max(per) < 1e-3, (r, per)
    # batch-uniform fast path == per-sample time path
    e0 = net.denoise_uniform(x.cuda(), int(t[0]))
    assert torch.allclose(e0[0], eps[0], rtol=0, atol=1e-5 * ref.abs().max().item())
    # float times take the per-sample FiLM evaluation
    e1 = net(x.cuda(), t.cuda().float())
    assert rel(e1.cpu(), ref) < 1e-3
    torch.set_grad_enabled(True)


def test_unet_intermediates_localise_errors(golden):
    """Per-stage comparison against the oracle's taps (diagnostic: which stage drifts first)."""
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=32, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    x, t = fx.unet_inputs(3)
    taps = {}
    with torch.no_grad():
        ref = unet_ref.unet_forward({k: v.cpu() for k, v in net.state_dict().items()}, x, t, taps=taps)
    g = golden("unet_dim32")
    assert np.array_equal(taps["init"].numpy(), g["tap_init_conv"])
    eps = net(x.cuda(), t.cuda())
    assert rel(eps.cpu(), ref) < 1e-3


def test_weight_cache_follows_parameter_updates():
    import safediffcon_b200 as s
    torch.manual_seed(0)
    net = s.Unet2D(dim=32, channels=3, resnet_block_groups=1).cuda()
    x, t = fx.unet_inputs(2)
    a = net(x.cuda(), t.cuda())
    with torch.no_grad():
        net.downs[0][0].block1.proj.weight.mul_(1.5)    # what an optimiser / EMA step does between chains
        net.time_mlp[1].bias.add_(0.1)
    b = net(x.cuda(), t.cuda())
    assert not torch.allclose(a, b)
    with torch.no_grad():
        ref = unet_ref.unet_forward({k: v.cpu() for k, v in net.state_dict().items()}, x, t)
    assert rel(b.cpu(), ref) < 1e-3


UPCONV_CASES = [
    # B, h, w (INPUT size), cin, cout, operand_out
    (2, 2, 16, 64, 64, True),       # 2x16 -> 4x32: a 32-row chunk spans two image rows
    (3, 4, 32, 128, 64, True),
    (2, 8, 64, 64, 128, False),     # 8x64 -> 16x128, fp32 output
    (5, 2, 16, 64, 32, True),       # ragged batch (4 images per tile)
    (70, 4, 32, 64, 128, True),     # CTA pairs
    (37, 8, 64, 64, 256, True),     # CTA pairs, odd tile count, N tile 256
]
UPCONV_PARAMS = [(c, p) for c in UPCONV_CASES for p in (TF32, F16)]


@pytest.mark.parametrize("case,prec", UPCONV_PARAMS, ids=[f"{'f16' if pr else 'tf32'}_B{c[0]}_{c[1]}x{c[2]}_c{c[3]}_o{c[4]}" for c, pr in UPCONV_PARAMS])
def test_fused_upsample_conv_vs_torch(case, prec):
    """Upsample2d (nearest x2 + 3x3 pad 1, reference unet.py:33-37) as four 2x2 phase convolutions on the low-resolution input."""
    from safediffcon_b200 import unet as U
    B, h, w_, cin, cout, rnd = case
    g = torch.Generator().manual_seed(h * w_ + cin + cout)
    x = quant(torch.randn(B, cin, h, w_, generator=g), prec).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    a0 = as_operand(nhwc(x), prec)
    wp = U.pack_conv_weight(U.KIND_UP2X, w, prec)
    assert wp.shape == (4 * cout, 4 * cin)
    out = torch.full((B * 4 * h * w_, cout), float("nan"), dtype=U.operand_dtype(prec) if rnd else torch.float32).cuda()
    U.conv_gemm(U.KIND_UP2X, a0, cin, None, 0, wp, bias, None, out, None, rnd, B, h, w_, cout, prec)
    torch.cuda.synchronize()
    ref = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), w.double(), bias.double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(B * 4 * h * w_, cout)
    out = out.double()
    assert torch.isfinite(out).all()
    # the phase weights are sums of up to four taps rounded once to the operand precision (2^-11 relative)
    err = (out - ref).abs().max().item()
    assert err < 3e-3 * max(1.0, ref.abs().max().item()), err
    assert ((out - ref).norm() / ref.norm()).item() < 6e-4


@pytest.mark.parametrize("B,c0,c1,film,resid", [(296, 128, 0, True, False), (296, 128, 128, False, True), (300, 64, 0, True, True),
                                                (3, 128, 0, True, True), (19, 128, 128, True, False)])
def test_conv3x3_row_with_fused_groupnorm(B, c0, c1, film, resid):
    """sdc_conv3x3_row_gn (conv + GroupNorm/FiLM/SiLU/residual in one kernel, deferred epilogue) == sdc_conv3x3_row followed by
    sdc_gn_silu, and == torch on a few samples.  B = 296 / 300: 72 clusters, 16-17 rounds, ragged last round; B = 3, 19: fewer
    items than SM pairs / a partial second round."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    H, W, cout = 16, 128, 128
    g = torch.Generator().manual_seed(B + c0 + c1)
    M = B * H * W
    a0 = (torch.randn(M, c0, generator=g) * 0.8).half().cuda()
    a1 = (torch.randn(M, c1, generator=g) * 0.8).half().cuda() if c1 else None
    w = (torch.randn(cout, c0 + c1, 3, 3, generator=g) / np.sqrt(9 * (c0 + c1))).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    gamma, beta = (1 + 0.3 * torch.randn(cout, generator=g)).cuda(), (0.2 * torch.randn(cout, generator=g)).cuda()
    table = (0.3 * torch.randn(7, 3 * cout, generator=g)).cuda() if film else None
    tidx = torch.randint(0, 7, (B,), generator=g).to(torch.int32).cuda() if film else None
    res = torch.randn(M, cout, generator=g).half().cuda() if resid else None
    cw = dict(w=U.pack_conv_weight(1, w, F16), b=bias, cout=cout)
    # unfused: conv (fp16 out + statistics) then the in-place fp16 GroupNorm kernel
    s_ref = torch.zeros(B, 2, dtype=torch.float64).cuda()
    y_ref = torch.empty(M, cout, dtype=torch.float16).cuda()
    U.conv_gemm(1, a0, c0, a1, c1, cw["w"], bias, None, y_ref, s_ref, True, B, H, W, cout, F16)
    L.check(lib.sdc_gn_silu(F16, L.ptr(y_ref), 1, L.ptr(s_ref), L.ptr(gamma), L.ptr(beta), L.ptr(table), L.ptr(tidx), 3 * cout if film else 0,
                            L.ptr(res), 1, L.ptr(y_ref), B, H * W, cout, L.stream_ptr()))
    # fused
    s = torch.zeros(B, 2, dtype=torch.float64).cuda()
    n = torch.full((B, 256), -1, dtype=torch.int32).cuda()   # SDC_GN_SLOT_BYTES per sample, every byte 0xFF
    y = torch.full((M, cout), float("nan"), dtype=torch.float16).cuda()
    rc = U.conv_row_gn(a0, c0, a1, c1, cw, y, s, n, (gamma, beta), table, tidx, 3 * cout if film else 0, res, B, H, W, cout)
    assert rc == 0
    torch.cuda.synchronize()
    assert (n.view(B, 128, 2)[:, :, 0] != -1).all()   # every exchange slot of every sample was filled
    assert torch.isfinite(y.float()).all()
    assert torch.allclose(s, s_ref, rtol=1e-6, atol=1e-3)
    d = (y.float() - y_ref.float()).abs()
    scale = y_ref.float().abs().max().item()
    # the fused kernel normalises the fp32 accumulators, the unfused pair their fp16 rounding: one extra 2^-11 on the reference side
    assert d.max().item() < 3e-3 * scale and d.mean().item() < 2e-4 * scale, (d.max().item(), d.mean().item(), scale)
    # torch on the first, a middle and the last sample
    for b in sorted({0, B // 2, B - 1}):
        sl = slice(b * H * W, (b + 1) * H * W)
        x = a0[sl].float() if a1 is None else torch.cat((a0[sl].float(), a1[sl].float()), dim=1)
        x = x.reshape(1, H, W, c0 + c1).permute(0, 3, 1, 2)
        conv = F.conv2d(x, w.half().float(), bias, padding=1)
        z = F.group_norm(conv, 1, gamma, beta, eps=1e-5)
        if film:
            row = table[tidx[b].long()]
            z = z * (row[:cout, None, None] + 1) + row[cout:2 * cout, None, None]
        z = F.silu(z)
        ref = nhwc(z).reshape(H * W, cout) + (res[sl].float() if resid else 0)
        assert (y[sl].float() - ref).abs().max().item() < 2e-3 * ref.abs().max().item(), b


@pytest.mark.parametrize("B", [2, 300])
def test_conv3x3_row_gn_head_vs_torch(B):
    """Last convolution of the network with GroupNorm + SiLU + residual + 1x1 head convolution in its epilogue (NCHW fp32 out)."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    H, W, c0, cout, co = 16, 128, 128, 128, 3
    g = torch.Generator().manual_seed(B)
    M = B * H * W
    a0 = (torch.randn(M, c0, generator=g) * 0.8).half().cuda()
    w = (torch.randn(cout, c0, 3, 3, generator=g) / np.sqrt(9 * c0)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    gamma, beta = (1 + 0.3 * torch.randn(cout, generator=g)).cuda(), (0.2 * torch.randn(cout, generator=g)).cuda()
    res = torch.randn(M, cout, generator=g).half().cuda()
    hw, hb = (torch.randn(co, cout, generator=g) / np.sqrt(cout)).cuda(), torch.randn(co, generator=g).cuda()
    cw = dict(w=U.pack_conv_weight(1, w, F16), b=bias, cout=cout)
    s = torch.zeros(B, 2, dtype=torch.float64).cuda()
    n = torch.full((B, 256), -1, dtype=torch.int32).cuda()
    out = torch.zeros(B, co, H, W).cuda()   # the kernel accumulates the two channel halves of a pixel into it
    assert U.conv_row_gn(a0, c0, None, 0, cw, None, s, n, (gamma, beta), None, None, 0, res, B, H, W, cout, head=(hw, hb, out)) == 0
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    for b in sorted({0, B // 2, B - 1}):
        sl = slice(b * H * W, (b + 1) * H * W)
        x = a0[sl].float().reshape(1, H, W, c0).permute(0, 3, 1, 2)
        z = F.silu(F.group_norm(F.conv2d(x, w.half().float(), bias, padding=1), 1, gamma, beta, eps=1e-5))
        z = z + res[sl].float().reshape(1, H, W, cout).permute(0, 3, 1, 2)
        ref = F.conv2d(z, hw.reshape(co, cout, 1, 1), hb)[0]
        assert (out[b] - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item()), b


def test_fused_groupnorm_path_matches_unfused_network():
    """Whole denoiser at B = 296 (fused conv+GroupNorm on the 16x128 level) against the same network with the separate kernels."""
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(296, 3, 16, 128, generator=g).cuda()
    t = torch.randint(0, 1000, (296,), generator=g).cuda()
    with torch.no_grad():
        net.fuse_groupnorm = True
        a = net(x, t)
        net.fuse_groupnorm = False
        b = net(x, t)
    assert torch.isfinite(a).all()
    # the fused kernels normalise fp32 accumulators where the separate ones read their fp16 rounding: the two paths differ by
    # about one rounding site per GroupNorm of the level (both are within 7e-4 of the fp32 reference, test_unet_eps_*)
    assert rel(a, b) < 8e-4, rel(a, b)


@pytest.mark.parametrize("res_half", [True, False])
def test_gn_silu_head_vs_torch(res_half):
    """Last GroupNorm + SiLU + residual fused with the 1x1 head convolution == torch (fp32), NCHW output."""
    L, lib = _L()
    import safediffcon_b200.unet  # noqa: F401
    for B, HW in ((3, 2048), (300, 2048), (2, 64)):
        C, Cout = 128, 3
        g = torch.Generator().manual_seed(B + HW)
        x = (torch.randn(B, C, HW, generator=g) * 1.7 + 0.3).cuda()
        gamma, beta = (1 + 0.3 * torch.randn(C, generator=g)).cuda(), (0.2 * torch.randn(C, generator=g)).cuda()
        res = torch.randn(B * HW, C, generator=g).cuda()
        if res_half:
            res = res.half()
        hw_, hb_ = (torch.randn(Cout, C, generator=g) / np.sqrt(C)).cuda(), torch.randn(Cout, generator=g).cuda()
        xr = x.permute(0, 2, 1).reshape(B * HW, C).contiguous()
        stats = torch.stack([xr.double().reshape(B, -1).sum(1), (xr.double() ** 2).reshape(B, -1).sum(1)], 1).contiguous()
        out = torch.full((B, Cout, HW), float("nan")).cuda()
        L.check(lib.sdc_gn_silu_head(L.ptr(xr), L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(res), int(res_half), L.ptr(hw_), L.ptr(hb_),
                                     L.ptr(out), B, HW, C, Cout, L.stream_ptr()))
        y = F.silu(F.group_norm(x, 1, gamma, beta, eps=1e-5)).permute(0, 2, 1).reshape(B * HW, C) + res.float()
        ref = (y.double() @ hw_.double().t() + hb_.double()).reshape(B, HW, Cout).permute(0, 2, 1)
        assert torch.isfinite(out).all()
        assert (out.double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


def test_eps_after_adamw_steps_within_1e3():
    """Parity away from the seed-42 initialisation: a few AdamW steps on the diffusion loss (CUDA backward) move every weight,
    bias and norm gain; eps of the updated dim-128 model must still agree with the fp32 oracle within 1e-3, per sample, on both
    operand precisions (north star; the oracle is pinned bit-exactly to the reference, tests/test_oracle_golden.py)."""
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=128, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1).cuda()
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), timesteps=1000, sampling_timesteps=200, temporal=True, use_conv2d=True,
                             is_condition_u0=True, is_condition_uT=True, condition_idx=10).cuda()
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    opt = torch.optim.AdamW(net.parameters(), lr=2e-3, weight_decay=0.01)
    x0 = fx.calibration_states(4).cuda()
    g = torch.Generator().manual_seed(5)
    for it in range(3):
        t = torch.randint(0, 1000, (4,), generator=g).cuda()
        noise = torch.randn(4, 3, 16, 128, generator=g).cuda()
        loss = gd.p_losses(x0.clone(), t, noise=noise)
        opt.zero_grad()
        loss.backward()
        opt.step()
    moved = [((net.state_dict()[k] - v).norm() / (v.norm() + 1e-12)).item() for k, v in before.items()]
    assert min(moved) > 1e-3 and float(np.median(moved)) > 0.02, (min(moved), float(np.median(moved)))   # every tensor changed, most by > 2 %
    x, t = fx.unet_inputs(4)
    with torch.no_grad():
        ref = unet_ref.unet_forward({k: v.detach().cpu() for k, v in net.state_dict().items()}, x, t)
        for prec in ("f16", "tf32"):
            net.precision = prec
            eps = net(x.cuda(), t.cuda()).cpu()
            per = [rel(eps[i], ref[i]) for i in range(4)]
            assert max(per) < 1e-3, (prec, per)


def test_layernorm_folded_into_attention_projections():
    """The two fused LayerNorm epilogues of the FP16 LinearAttention path against the separate kernels and torch:
    (a) sdc_gn_silu_rowstats = sdc_gn_silu + per-row (mean, rstd) of its output; (b) sdc_conv1x1_qkv_ln on raw rows with folded
    weights == LayerNorm -> sdc_conv1x1_qkv; (c) sdc_conv1x1_per_sample_ln == sdc_conv1x1_per_sample -> LayerNorm + residual."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    for B, H, W, c in ((3, 16, 128, 128), (75, 8, 64, 256), (2, 4, 32, 256)):
        g = torch.Generator().manual_seed(B + c)
        n, M = H * W, B * H * W
        raw = (torch.randn(M, c, generator=g) * 1.3 + 0.2).half().cuda()
        res = torch.randn(M, c, generator=g).half().cuda()
        gamma, beta = (1 + 0.3 * torch.randn(c, generator=g)).cuda(), (0.2 * torch.randn(c, generator=g)).cuda()
        stats = torch.stack([raw.float().reshape(B, -1).sum(1), (raw.float() ** 2).reshape(B, -1).sum(1)], dim=1).double().contiguous()
        # (a)
        y_ref = raw.clone()
        L.check(lib.sdc_gn_silu(F16, L.ptr(y_ref), 1, L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None, 0, L.ptr(res), 1, L.ptr(y_ref), B, n, c,
                                L.stream_ptr()))
        y = raw.clone()
        rs = torch.full((M, 2), float("nan")).cuda()
        L.check(lib.sdc_gn_silu_rowstats(L.ptr(y), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None, 0, L.ptr(res), L.ptr(y), L.ptr(rs), B, n, c,
                                         L.stream_ptr()))
        assert torch.equal(y, y_ref)
        yf = y.float()
        assert torch.allclose(rs[:, 0], yf.mean(1), atol=2e-6, rtol=1e-5)
        assert torch.allclose(rs[:, 1], (yf.var(1, unbiased=False) + 1e-5).rsqrt(), rtol=2e-5)
        # (b)
        wqkv = (torch.randn(384, c, generator=g) * (1.5 / np.sqrt(c))).cuda()
        g_in = (1 + 0.2 * torch.randn(c, generator=g)).cuda()
        xn = torch.empty_like(y)
        L.check(lib.sdc_channel_layernorm(F16, L.ptr(y), 1, L.ptr(g_in), None, L.ptr(xn), M, c, 1, L.stream_ptr()))
        q_ref, kv_ref = torch.empty(M, 128, dtype=torch.float16).cuda(), torch.empty(M, 256, dtype=torch.float16).cuda()
        U.conv1x1_qkv(xn, c, U.pack_conv_weight(0, wqkv.reshape(384, c, 1, 1), F16), q_ref, kv_ref, B, H, W, F16)
        wfold, wsum = torch.empty(384, c, dtype=torch.float16).cuda(), torch.empty(384).cuda()
        L.check(lib.sdc_pack_qkv_ln(L.ptr(wqkv), L.ptr(g_in), L.ptr(wfold), L.ptr(wsum), 384, c, L.stream_ptr()))
        assert torch.equal(wfold, (wqkv * g_in).half()) and torch.allclose(wsum, wfold.float().sum(1), rtol=1e-5, atol=1e-5)
        q, kv = torch.empty_like(q_ref), torch.empty_like(kv_ref)
        L.check(lib.sdc_conv1x1_qkv_ln(L.ptr(y), c, L.ptr(wfold), L.ptr(wsum), L.ptr(rs), L.ptr(q), L.ptr(kv), B, H, W, 128, L.stream_ptr()))
        ln = (yf.double() - yf.double().mean(1, keepdim=True)) * (yf.double().var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * g_in.double()
        qkv_t = ln @ wqkv.double().t()
        q_t = (qkv_t[:, :128].reshape(M, 4, 32).softmax(-1) * 32 ** -0.5).reshape(M, 128)
        for got, old, want in ((q, q_ref, q_t), (kv, kv_ref, qkv_t[:, 128:])):
            sc = want.abs().max().item()
            e_new, e_old = (got.double() - want).abs().max().item() / sc, (old.double() - want).abs().max().item() / sc
            assert e_new < 2e-3 and e_new < 1.5 * e_old + 2e-4, (B, c, e_new, e_old)   # no worse than LayerNorm -> fp16 -> projection
        # (c)
        wf = (torch.randn(B * c, 128, generator=g) / np.sqrt(128)).half().cuda()
        bout, g_out = torch.randn(c, generator=g).cuda(), (1 + 0.2 * torch.randn(c, generator=g)).cuda()
        proj = torch.empty(M, c, dtype=torch.float16).cuda()
        U.conv1x1_per_sample(q_ref, 128, wf, bout, proj, B, H, W, c, F16)
        o_ref = torch.empty(M, c, dtype=torch.float16).cuda()
        L.check(lib.sdc_channel_layernorm(F16, L.ptr(proj), 1, L.ptr(g_out), L.ptr(y), L.ptr(o_ref), M, c, 1, L.stream_ptr()))
        o = torch.full((M, c), float("nan"), dtype=torch.float16).cuda()
        L.check(lib.sdc_conv1x1_per_sample_ln(L.ptr(q_ref), 128, L.ptr(wf), L.ptr(bout), L.ptr(g_out), L.ptr(y), L.ptr(o), B, H, W, c, L.stream_ptr()))
        pj = torch.einsum("bnk,bck->bnc", q_ref.double().reshape(B, n, 128), wf.double().reshape(B, c, 128)).reshape(M, c) + bout.double()
        want = (pj - pj.mean(1, keepdim=True)) * (pj.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt() * g_out.double() + yf.double()
        sc = want.abs().max().item()
        e_new, e_old = (o.double() - want).abs().max().item() / sc, (o_ref.double() - want).abs().max().item() / sc
        assert torch.isfinite(o.float()).all() and e_new < 2e-3 and e_new < 1.5 * e_old + 2e-4, (B, c, e_new, e_old)


def test_fold_on_tensor_cores_matches_fp32():
    """linattn_fold_mma_kernel (mma.sync, hi/lo split operands) == the fp32 product W_out (x) ctx to fp16 rounding."""
    L, lib = _L()
    import safediffcon_b200.unet  # noqa: F401
    for B, c in ((5, 128), (3, 256), (2, 512), (4, 64)):
        g = torch.Generator().manual_seed(c)
        ws = torch.zeros(B, 4, 32 * 32 + 64).cuda()
        ctx = torch.randn(B, 4, 32, 32, generator=g).cuda() * 0.7
        ws[:, :, :1024] = ctx.reshape(B, 4, 1024)
        wout = (torch.randn(c, 128, generator=g) / np.sqrt(128)).cuda()
        wf = torch.full((B * c, 128), float("nan"), dtype=torch.float16).cuda()
        L.check(lib.sdc_linear_attention_fold(F16, L.ptr(ws), L.ptr(wout), L.ptr(wf), B, c, L.stream_ptr()))
        ref = torch.einsum("che,bhde->bchd", wout.double().reshape(c, 4, 32), ctx.double()).reshape(B * c, 128)
        err = (wf.double() - ref).abs()
        assert torch.isfinite(wf.float()).all()
        assert (err <= ref.abs() * 2 ** -10.9 + 1e-6).all(), (c, err.max().item())


def test_row64_kernel_declines_ineligible_shapes():
    """sdc_conv3x3_row with W = 64 hands the problem back (-1, nothing launched) when an image is not a whole number of 8-row items or when
    there are fewer items than SM pairs; U.conv_gemm then runs the generic kernel, whose result the W = 64 kernel reproduces bitwise."""
    L, lib = _L()
    from safediffcon_b200 import unet as U
    g = torch.Generator().manual_seed(64)
    cout, c0 = 128, 128
    w = (torch.randn(cout, c0, 3, 3, generator=g) / 34.0).cuda()
    wp = U.pack_conv_weight(1, w, F16)
    for B, H, want in ((80, 4, -1), (8, 8, -1), (80, 8, 0)):
        x = torch.randn(B * H * 64, c0, generator=g).half().cuda()
        out = torch.zeros(B * H * 64, cout, dtype=torch.float16).cuda()
        stats = torch.zeros(B, 2, dtype=torch.float64).cuda()
        rc = lib.sdc_conv3x3_row(F16, L.ptr(x), c0, None, 0, L.ptr(wp), None, None, L.ptr(out), L.ptr(stats), 1, B, H, 64, cout, L.stream_ptr())
        torch.cuda.synchronize()
        assert rc == want, (B, H, rc)
        if rc == 0:
            ref, rstats = torch.zeros_like(out), torch.zeros_like(stats)
            L.check(lib.sdc_conv_gemm(F16, 1, L.ptr(x), c0, None, 0, L.ptr(wp), None, None, L.ptr(ref), L.ptr(rstats), 1, B, H, 64, cout, L.stream_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(out, ref)
            assert torch.allclose(stats, rstats, rtol=1e-12, atol=1e-9)   # same fp32 partials, double atomics in a different order
        else:
            assert not out.any()
