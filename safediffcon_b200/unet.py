"""Denoiser: drop-in for Unet2D (/root/reference/1D/model/unet.py:263-426).

The module tree below exists to own the parameters under the reference's state_dict keys
(``downs.0.0.block1.proj.weight`` ...) and to consume torch's RNG in the reference's construction order, so a
reference checkpoint loads unchanged and ``torch.manual_seed(s); Unet2D(...)`` yields the reference's initial
weights.  ``forward`` does not run those modules: it drives the CUDA kernels of csrc/conv_gemm.cu (tcgen05
implicit-GEMM convolutions) and csrc/unet_ops.cu (fused GroupNorm/SiLU, LayerNorm, attention) over NHWC
activations.  Packed TF32 weights and the per-timestep FiLM table are cached and rebuilt whenever a parameter's
version counter or storage changes (optimiser / EMA updates between chains).
"""
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _lib as L
from ._lib import c_f, c_i, c_i64, c_p

L.register({
    "sdc_pack_conv_weight": (c_i, [c_i, c_i, c_p, c_p, c_i, c_i, c_p]),
    "sdc_conv_gemm": (c_i, [c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_conv3x3_row": (c_i, [c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_conv3x3_row_gn": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_conv3x3_row_gn_head": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_stem_conv7": (c_i, [c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_stem_im2col": (c_i, [c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_stem_conv7_tc": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_gn_silu": (c_i, [c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_i, c_p, c_i, c_i, c_i, c_p]),
    "sdc_gn_silu_head": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_channel_layernorm": (c_i, [c_i, c_p, c_i, c_p, c_p, c_p, c_i64, c_i, c_i, c_p]),
    "sdc_linear_attention_workspace": (c_i64, [c_i]),
    "sdc_linear_attention": (c_i, [c_i, c_p, c_p, c_p, c_i, c_i, c_p]),
    "sdc_attention": (c_i, [c_i, c_p, c_p, c_i, c_i, c_p]),
    "sdc_conv1x1_qkv": (c_i, [c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_linear_attention_context": (c_i, [c_p, c_p, c_i, c_i, c_p, c_i, c_i, c_p]),
    "sdc_linear_attention_fold": (c_i, [c_i, c_p, c_p, c_p, c_i, c_i, c_p]),
    "sdc_conv1x1_per_sample": (c_i, [c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_upsample2x": (c_i, [c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_head_conv1": (c_i, [c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_linear_rows": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_sinusoidal_embedding": (c_i, [c_p, c_p, c_i, c_i, c_f, c_p]),
    "sdc_zero_f64": (c_i, [c_p, c_i64, c_p]),
    "sdc_pack_conv_weight_dgrad": (c_i, [c_i, c_p, c_p, c_i, c_i, c_p]),
    "sdc_gn_silu_bwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_p, c_i, c_i, c_i, c_p]),
    "sdc_channel_layernorm_bwd": (c_i, [c_p, c_p, c_i, c_p, c_p, c_p, c_i64, c_i, c_i, c_p]),
    "sdc_linear_attention_bwd_workspace": (c_i64, [c_i]),
    "sdc_linear_attention_bwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p]),
    "sdc_attention_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p]),
    "sdc_pixel_shuffle_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_upsample2x_bwd": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_add_inplace": (c_i, [c_p, c_p, c_i64, c_i, c_p]),
    "sdc_head_conv1_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_stem_col2im": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "sdc_conv_wgrad": (c_i, [c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_conv_wgrad_tc_scratch": (c_i64, [c_i, c_i, c_i, c_i, c_i, c_i]),
    "sdc_conv_wgrad_tc": (c_i, [c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_i64, c_p]),
    "sdc_colsum": (c_i, [c_p, c_p, c_i64, c_i, c_p]),
    "sdc_gn_param_grad": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_i, c_i, c_i, c_p]),
    "sdc_channel_layernorm_gain_grad": (c_i, [c_p, c_p, c_i, c_p, c_i64, c_i, c_p]),
    "sdc_head_conv1_wgrad": (c_i, [c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
})

L.register({
    "sdc_unet_create": (c_i, [ctypes.POINTER(c_p), c_i, ctypes.POINTER(c_i), c_i, c_i, c_i, c_i, c_f, c_i]),
    "sdc_unet_destroy": (None, [c_p]),
    "sdc_unet_param_count": (c_i, [c_p]),
    "sdc_unet_param_name": (ctypes.c_char_p, [c_p, c_i]),
    "sdc_unet_param_numel": (c_i64, [c_p, c_i]),
    "sdc_unet_pack_weights": (c_i, [c_p, ctypes.POINTER(c_p), c_i, c_p]),
    "sdc_unet_workspace_bytes": (c_i64, [c_p, c_i, c_i, c_i]),
    "sdc_unet_forward": (c_i, [c_p, c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_p, c_i64, c_p, c_p]),
    "sdc_unet_set_flag": (c_i, [c_p, c_i, c_i]),
    "sdc_unet_backward_workspace_bytes": (c_i64, [c_p, c_i, c_i, c_i]),
    "sdc_unet_backward_data": (c_i, [c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_i64, c_p]),
    "sdc_unet_film_table": (c_i, [c_p, c_p, ctypes.POINTER(c_i), ctypes.POINTER(c_i), c_p]),
    "sdc_gn_silu_rowstats": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_i, c_i, c_i, c_p]),
    "sdc_pack_qkv_ln": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p]),
    "sdc_conv1x1_qkv_ln": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_conv1x1_per_sample_ln": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    "sdc_unet_profile_enable": (c_i, [c_p, c_i]),
    "sdc_unet_profile_count": (c_i, [c_p]),
    "sdc_unet_profile_entry": (c_i, [c_p, c_i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_f), ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]),
})

HEADS, DIM_HEAD = 4, 32
KIND_1x1, KIND_3x3, KIND_UNSHUFFLE, KIND_UP2X = 0, 1, 2, 3
PREC_TF32, PREC_F16 = 0, 1   # operand precision of the tensor-core convolutions (include/safediffcon_b200_unet.h)
PREC_NAMES = {"tf32": PREC_TF32, "f16": PREC_F16}


def operand_dtype(prec):
    return torch.float16 if prec == PREC_F16 else torch.float32


# ----------------------------------------------------------------------------- parameter containers
class _Gain(nn.Module):
    """Channel LayerNorm gain (reference LayerNorm.g)."""

    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))


class _Block(nn.Module):
    def __init__(self, dim, dim_out, groups):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm = nn.GroupNorm(groups, dim_out)
        self.act = nn.SiLU()


class _ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, time_emb_dim, groups):
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2))
        self.block1 = _Block(dim, dim_out, groups)
        self.block2 = _Block(dim_out, dim_out, groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out = dim, dim_out


class _LinearAttention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        hidden = HEADS * DIM_HEAD
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), _Gain(dim))


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        hidden = HEADS * DIM_HEAD
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


class _PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = _Gain(dim)


class _Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


def _attn(dim, full=False):
    inner = _Attention(dim) if full else _LinearAttention(dim)
    return _Residual(_PreNorm(dim, inner))


# ----------------------------------------------------------------------------- op wrappers (thin ctypes calls)
def _st():
    return L.stream_ptr()


USE_ROW_KERNEL = os.environ.get("SDC_NO_ROW_KERNEL", "0") != "1"  # halo-reuse kernel for the 16x128 level
FUSE_UPSAMPLE = os.environ.get("SDC_NO_FUSED_UPSAMPLE", "0") != "1"  # Upsample2d as four 2x2 phase convolutions (inference path)
PROFILE = None  # bench.py sets this to a list: every conv launch is then bracketed by CUDA events on its stream


class _Timed:
    """Brackets one conv launch with CUDA events on its stream when bench.py has set PROFILE."""

    def __init__(self, flops, shape):
        self.flops, self.shape = flops, shape

    def __enter__(self):
        self.prof = PROFILE
        if self.prof is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.prof is not None and exc[0] is None:
            self.e1.record()
            self.prof.append((self.e0, self.e1, self.flops, self.shape))


def conv_row_gn(a0, c0, a1, c1, cw, out, stats, counters, gn, ss, t_index, ss_stride, residual, B, H, W, Cout, head=None):
    """3x3 convolution of the full-resolution level FUSED with the following GroupNorm + FiLM + SiLU (+ residual) (FP16, deferred
    epilogue; include/safediffcon_b200_unet.h: sdc_conv3x3_row_gn).  head = (w [o, Cout], b [o], out [B, o, H, W]): the network's
    last block -- the 1x1 head convolution is applied in the same epilogue and `out` (the activation) is not written.
    Returns 0, or -1 when not eligible (nothing launched)."""
    with _Timed(2.0 * B * H * W * Cout * 9 * (c0 + c1), (KIND_3x3, B, H, W, c0 + c1, Cout)):
        if head is None:
            rc = L.lib().sdc_conv3x3_row_gn(L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(cw["w"]), L.ptr(cw["b"]), L.ptr(out), L.ptr(stats),
                                            L.ptr(counters), L.ptr(gn[0]), L.ptr(gn[1]), L.ptr(ss), L.ptr(t_index), ss_stride,
                                            L.ptr(residual), B, H, W, Cout, _st())
        else:
            rc = L.lib().sdc_conv3x3_row_gn_head(L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(cw["w"]), L.ptr(cw["b"]), L.ptr(stats),
                                                 L.ptr(counters), L.ptr(gn[0]), L.ptr(gn[1]), L.ptr(residual), L.ptr(head[0]),
                                                 L.ptr(head[1]), L.ptr(head[2]), head[2].shape[1], B, H, W, Cout, _st())
    if rc > 0:
        L.check(rc)
    return rc


def conv1x1_qkv(a, c, wp, q_out, kv_out, B, H, W, prec):
    """qkv projection of LinearAttention with the q-softmax epilogue (see include/safediffcon_b200_unet.h)."""
    hid = HEADS * DIM_HEAD
    with _Timed(2.0 * B * H * W * 3 * hid * c, (KIND_1x1, B, H, W, c, 3 * hid)):
        L.check(L.lib().sdc_conv1x1_qkv(prec, L.ptr(a), c, L.ptr(wp), L.ptr(q_out), L.ptr(kv_out), int(kv_out.dtype == torch.float16),
                                        B, H, W, hid, _st()))


def conv1x1_per_sample(a, c, w_folded, bias, out, B, H, W, Cout, prec):
    """1x1 convolution whose [Cout, c] weight differs per sample (folded LinearAttention output projection); `out` fp32 or fp16."""
    with _Timed(2.0 * B * H * W * Cout * c, (KIND_1x1, B, H, W, c, Cout)):
        L.check(L.lib().sdc_conv1x1_per_sample(prec, L.ptr(a), c, L.ptr(w_folded), L.ptr(bias), L.ptr(out), int(out.dtype == torch.float16),
                                               B, H, W, Cout, _st()))


def conv_gemm(kind, a0, c0, a1, c1, wp, bias, residual, out, stats, operand_out, B, H, W, Cout, prec=PREC_TF32, algo_k=None):
    """out[B*H*W, Cout] = conv(a0 | a1) + bias (+ residual).  a0/a1/wp/residual are operand-precision tensors; `out` is an
    operand-precision tensor when operand_out else fp32."""
    od = operand_dtype(prec)
    assert a0.dtype == od and wp.dtype == od and (a1 is None or a1.dtype == od) and (residual is None or residual.dtype == od)
    assert out.dtype == (od if operand_out else torch.float32)
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = -1
    if kind == KIND_3x3 and USE_ROW_KERNEL and W in (128, 64) and Cout <= 128:
        rc = L.lib().sdc_conv3x3_row(prec, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(wp), L.ptr(bias), L.ptr(residual), L.ptr(out),
                                     L.ptr(stats), int(operand_out), B, H, W, Cout, _st())
        if rc > 0:
            L.check(rc)
    if rc != 0:
        L.check(L.lib().sdc_conv_gemm(prec, kind, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(wp), L.ptr(bias), L.ptr(residual),
                                      L.ptr(out), L.ptr(stats), int(operand_out), B, H, W, Cout, _st()))
    if prof is not None:
        e1.record()
        taps = {KIND_1x1: 1, KIND_3x3: 9, KIND_UNSHUFFLE: 4, KIND_UP2X: 4}[kind]
        # algorithmic FLOPs: algo_k overrides the GEMM K when the operand carries padding / split columns (tensor-core stem);
        # the fused upsample convolution is counted with the multiply-adds it EXECUTES (4 phases x 4 taps on H x W input pixels),
        # not the 9 taps x 4HW of the reference formulation
        k_alg = taps * (c0 + c1) if algo_k is None else algo_k
        rows = B * H * W * (4 if kind == KIND_UP2X else 1)
        prof.append((e0, e1, 2.0 * rows * Cout * k_alg, (kind, B, H, W, c0 + c1, Cout)))


USE_STEM_TC = os.environ.get("SDC_STEM_TC", "1") != "0"     # the stem as one tcgen05 kernel (csrc/stem_conv.cu)
USE_WGRAD_TC = os.environ.get("SDC_WGRAD_TC", "1") != "0"   # conv weight gradients on tcgen05 where the shape allows


def conv_wgrad_any(kind, a0, c0, a1, c1, dy, dw, B, H, W, Cout):
    """dw += conv weight gradient (OIHW): tcgen05 kernel (pixel axis = K, MN-major operands) when the shape is eligible, else the
    mma.sync kernel (include/safediffcon_b200_unet.h: sdc_conv_wgrad_tc / sdc_conv_wgrad)."""
    lib = L.lib()
    half = int(a0.dtype == torch.float16)
    if USE_WGRAD_TC:
        nb = int(lib.sdc_conv_wgrad_tc_scratch(half, c0, c1, B, H, W))
        scratch = torch.empty(nb, dtype=torch.uint8, device=dy.device) if nb else None
        rc = lib.sdc_conv_wgrad_tc(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, Cout, L.ptr(scratch), nb, _st())
        if rc == 0:
            return
        if rc > 0:
            L.check(rc)
    L.check(lib.sdc_conv_wgrad(kind, half, L.ptr(a0), c0, L.ptr(a1), c1, L.ptr(dy), L.ptr(dw), B, H, W, Cout, _st()))


def pack_conv_weight(kind, w, prec=PREC_TF32):
    w = w.detach().to(torch.float32).contiguous()
    cout, cin = w.shape[0], w.shape[1]
    shape = (4 * cout, 4 * cin) if kind == KIND_UP2X else (cout, w[0].numel())
    wp = torch.empty(shape, device=w.device, dtype=operand_dtype(prec))
    L.check(L.lib().sdc_pack_conv_weight(prec, kind, L.ptr(w), L.ptr(wp), cout, cin, _st()))
    return wp


def pack_conv_weight_dgrad(kind, w):
    """Weight of the data-gradient convolution: Wt[Cin, taps*Cout], TF32 (gradients are never fp16)."""
    w = w.detach().to(torch.float32).contiguous()
    cout, cin = w.shape[0], w.shape[1]
    wt = torch.empty(cin, w[0].numel() // cin * cout, device=w.device, dtype=torch.float32)
    L.check(L.lib().sdc_pack_conv_weight_dgrad(kind, L.ptr(w), L.ptr(wt), cout, cin, _st()))
    return wt


class _InputVJP(torch.autograd.Function):
    """eps = Unet2D(x, t) with a backward that returns d<eps, g>/dx (backward-data only: no parameter gradients)."""

    @staticmethod
    def forward(ctx, x, net, time, table_row):
        tape = []
        with torch.cuda.device(x.device):
            out = net._run(x.detach(), time, table_row, tape)
        ctx.net, ctx.tape = net, tape
        return out

    @staticmethod
    def backward(ctx, g):
        with torch.cuda.device(g.device):
            gx = ctx.net._vjp(ctx.tape, L.dev_f32(g, "grad_eps"))
        ctx.tape = None
        return gx, None, None, None


def _same_layout(a, b):
    """Two packed-weight trees hold tensors of identical shape / dtype / device at identical places."""
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return (isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor) and a.shape == b.shape and a.dtype == b.dtype
                and a.device == b.device)
    if isinstance(a, dict) and isinstance(b, dict):
        ka = [k for k in a if k not in ("table", "dgrad")]
        kb = [k for k in b if k not in ("table", "dgrad")]
        return ka == kb and all(_same_layout(a[k], b[k]) for k in ka)
    if isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        return len(a) == len(b) and all(_same_layout(x, y) for x, y in zip(a, b))
    return a == b


def _copy_into(dst, src):
    """dst <- src leaf by leaf, keeping dst's storage (captured CUDA graphs hold these pointers)."""
    if isinstance(dst, torch.Tensor):
        dst.copy_(src)
    elif isinstance(dst, dict):
        for k in dst:
            if k not in ("table", "dgrad"):
                _copy_into(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for x, y in zip(dst, src):
            _copy_into(x, y)


class _TrainFn(torch.autograd.Function):
    """eps = Unet2D(x, t) with the full backward: d/dx, d/d(FiLM rows) and every convolution / norm parameter gradient
    (csrc/unet_bwd.cu, csrc/unet_wgrad.cu, tcgen05 dgrad convolutions).  The time-MLP parameters receive their gradients
    through `film`, which the caller computes with differentiable torch ops on [B, 4*dim] matrices."""

    @staticmethod
    def forward(ctx, x, film, net, *params):
        tape = []
        with torch.cuda.device(x.device):
            out = net._run(x.detach(), None, None, tape, film_rows=film.detach().float().contiguous())
        ctx.net, ctx.tape, ctx.params = net, tape, params
        return out

    @staticmethod
    def backward(ctx, g):
        pg = {}
        with torch.cuda.device(g.device):
            gx = ctx.net._vjp(ctx.tape, L.dev_f32(g, "grad_eps"), pgrads=pg, need_gx=ctx.needs_input_grad[0])
        ctx.tape = None
        return (gx, pg["film"], None) + tuple(pg.get(id(p)) for p in ctx.params)


USE_PLAN = os.environ.get("SDC_NO_PLAN", "0") != "1"   # inference through the C++ executor (sdc_unet_forward)
WORKSPACE_ENTRIES = 3


class UnetPlan:
    """Handle of the C++ whole-network executor (include/safediffcon_b200_plan.h): owns the packed weights, the FiLM table and
    the launch schedule; `forward` is ONE C call.  Workspaces are torch uint8 tensors, one per problem size (captured graphs keep
    a reference to theirs)."""

    def __init__(self, dim, dim_mults, channels, out_dim, prec, theta, table_timesteps):
        lib = L.lib()
        h = c_p()
        mults = (c_i * len(dim_mults))(*[int(m) for m in dim_mults])
        L.check(lib.sdc_unet_create(ctypes.byref(h), int(dim), mults, len(dim_mults), int(channels), int(out_dim), int(prec),
                                    float(theta), int(table_timesteps)))
        self.handle, self.prec, self.out_dim, self.table_timesteps = h, prec, out_dim, table_timesteps
        self.names = [lib.sdc_unet_param_name(h, i).decode() for i in range(lib.sdc_unet_param_count(h))]
        self.numels = [int(lib.sdc_unet_param_numel(h, i)) for i in range(len(self.names))]
        self.workspaces = {}
        self.nonfinite = None
        self.key = None
        self.flags = (False, False)   # (fuse_ln, fuse_gn): the handle's defaults
        self.backward = False         # data-gradient weights packed (SDC_UNET_BACKWARD)

    def __del__(self):
        try:
            if self.handle:
                L.lib().sdc_unet_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def pack(self, named_params):
        """named_params: {state_dict key: fp32 CUDA tensor}.  Packs / refreshes IN PLACE (captured graphs stay valid)."""
        keep, ptrs = [], (c_p * len(self.names))()
        for i, (nm, ne) in enumerate(zip(self.names, self.numels)):
            t = named_params[nm].detach()
            if not t.is_cuda:
                raise RuntimeError("safediffcon_b200.Unet2D: parameters are on the CPU; move the module to a CUDA device "
                                   "(there is no CPU fallback)")
            t = t.to(torch.float32).contiguous()
            if t.numel() != ne:
                raise ValueError(f"safediffcon_b200: parameter {nm} has {t.numel()} elements, the executor expects {ne}")
            keep.append(t)
            ptrs[i] = t.data_ptr()
        L.check(L.lib().sdc_unet_pack_weights(self.handle, ptrs, len(self.names), _st()))
        if self.prec == PREC_F16 and self.nonfinite is None:
            self.nonfinite = torch.zeros(1, dtype=torch.int32, device=keep[0].device)

    def workspace(self, B, H, W, device):
        key = (B, H, W)
        ws = self.workspaces.pop(key, None)
        if ws is None:
            need = int(L.lib().sdc_unet_workspace_bytes(self.handle, B, H, W))
            if need <= 0:
                L.check(1)
            ws = torch.empty(need, dtype=torch.uint8, device=device)
        self.workspaces[key] = ws   # most recently used last
        while len(self.workspaces) > WORKSPACE_ENTRIES:
            self.workspaces.pop(next(iter(self.workspaces)))
        return ws

    def forward(self, x, t_index=None, t_uniform=0, workspace=None):
        B, _, H, W = x.shape
        ws = workspace if workspace is not None else self.workspace(B, H, W, x.device)
        out = torch.empty(B, self.out_dim, H, W, device=x.device, dtype=torch.float32)
        L.check(L.lib().sdc_unet_forward(self.handle, L.ptr(x), L.ptr(t_index), int(t_uniform), L.ptr(out), B, H, W, L.ptr(ws), ws.numel(),
                                         L.ptr(self.nonfinite), _st()))
        return out

    FUSE_LN, FUSE_GN, FILM_TC, BACKWARD = 1, 2, 3, 4

    def film_table(self):
        """Copy of the executor's FiLM table [table_timesteps, E] (tests compare it with the time MLP evaluated by torch)."""
        rows, cols = c_i(), c_i()
        L.check(L.lib().sdc_unet_film_table(self.handle, None, ctypes.byref(rows), ctypes.byref(cols), None))
        out = torch.empty(rows.value, cols.value, dtype=torch.float32, device="cuda")
        L.check(L.lib().sdc_unet_film_table(self.handle, L.ptr(out), None, None, _st()))
        return out

    def set_flag(self, flag, value):
        """Schedule switches of include/safediffcon_b200_plan.h (SDC_UNET_FUSE_LN, SDC_UNET_FUSE_GN)."""
        L.check(L.lib().sdc_unet_set_flag(self.handle, int(flag), int(value)))
        self.workspaces.clear()

    def vjp(self, x, grad_eps, t_index=None, t_uniform=0):
        """(eps, d<eps, grad_eps>/dx) through ONE C call (sdc_unet_backward_data); needs the BACKWARD flag set before packing."""
        B, _, H, W = x.shape
        need = int(L.lib().sdc_unet_backward_workspace_bytes(self.handle, B, H, W))
        if need <= 0:
            L.check(1)
        ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        eps = torch.empty(B, self.out_dim, H, W, device=x.device, dtype=torch.float32)
        gx = torch.empty_like(x)
        L.check(L.lib().sdc_unet_backward_data(self.handle, L.ptr(x), L.ptr(t_index), int(t_uniform), L.ptr(grad_eps), L.ptr(eps), L.ptr(gx),
                                               B, H, W, L.ptr(ws), ws.numel(), _st()))
        return eps, gx

    def profile(self, enable):
        L.check(L.lib().sdc_unet_profile_enable(self.handle, int(enable)))

    def profile_entries(self):
        """[(kernel family, ms, algorithmic bytes, flops)] of the forwards since profile(True); synchronise first."""
        lib, out = L.lib(), []
        nm, ms, by, fl = ctypes.c_char_p(), c_f(), ctypes.c_double(), ctypes.c_double()
        for i in range(lib.sdc_unet_profile_count(self.handle)):
            L.check(lib.sdc_unet_profile_entry(self.handle, i, ctypes.byref(nm), ctypes.byref(ms), ctypes.byref(by), ctypes.byref(fl)))
            out.append((nm.value.decode(), ms.value, by.value, fl.value))
        return out


class _PackCache:
    """Packed-weight cache that is never deep-copied (EMA wrappers deepcopy the module; the copy repacks lazily)."""

    def __init__(self):
        self.pack, self.key, self.fingerprint = None, None, None
        self.plan, self.plan_gen = None, 0

    def __deepcopy__(self, memo):
        return _PackCache()


class Unet2D(nn.Module):
    '''
    Estimates the noise given the last diffusion step
    The time dimension and the space dimension are treated equally
    '''

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=2, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False, random_fourier_features=False,
                 learned_sinusoidal_dim=16, sinusoidal_pos_emb_theta=10000, attn_dim_head=32, attn_heads=4,
                 condition_on_residual=None):
        super().__init__()
        if self_condition or learned_variance or learned_sinusoidal_cond or random_fourier_features or condition_on_residual:
            raise NotImplementedError("safediffcon_b200.Unet2D: option outside the 1D hot path (SURVEY.md section 8)")
        if resnet_block_groups != 1:
            raise NotImplementedError("safediffcon_b200.Unet2D implements GroupNorm(1, C) (resnet_block_groups=1), the value of "
                                      "every shipped 1D config")
        if attn_dim_head != DIM_HEAD or attn_heads != HEADS:
            raise NotImplementedError("attention is specialised for 4 heads x 32")
        if dim % 32 != 0:
            raise NotImplementedError("dim must be a multiple of 32 (TF32 K blocks of 32 channels)")
        self.condition_on_residual = None
        self.channels = channels
        self.self_condition = False
        self.dim = dim
        self.theta = sinusoidal_pos_emb_theta
        time_dim = dim * 4
        self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(dim, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))
        init_dim = init_dim if init_dim is not None else dim
        self.init_conv = nn.Conv2d(channels, init_dim, 7, padding=3)
        dims = [init_dim, *map(lambda m: dim * m, dim_mults)]
        in_out = list(zip(dims[:-1], dims[1:]))
        block = lambda a, b: _ResnetBlock(a, b, time_dim, resnet_block_groups)  # noqa: E731

        self.downs = nn.ModuleList([])
        for ind, (d_in, d_out) in enumerate(in_out):
            is_last = ind >= (len(in_out) - 1)
            # construction order = the reference's (it fixes the RNG stream of the default initialisation)
            mods = [block(d_in, d_in), block(d_in, d_in), _attn(d_in)]
            mods.append(nn.Sequential(nn.Identity(), nn.Conv2d(d_in * 4, d_out, 1)) if not is_last
                        else nn.Conv2d(d_in, d_out, 3, padding=1))
            self.downs.append(nn.ModuleList(mods))
        mid = dims[-1]
        self.mid_block1 = block(mid, mid)
        self.mid_attn = _attn(mid, full=True)
        self.mid_block2 = block(mid, mid)
        self.ups = nn.ModuleList([])
        for ind, (d_in, d_out) in enumerate(reversed(in_out)):
            is_last = ind == (len(in_out) - 1)
            mods = [block(d_out + d_in, d_out), block(d_out + d_in, d_out), _attn(d_out)]
            mods.append(nn.Sequential(nn.Identity(), nn.Conv2d(d_out, d_in, 3, padding=1)) if not is_last
                        else nn.Conv2d(d_out, d_in, 3, padding=1))
            self.ups.append(nn.ModuleList(mods))
        self.out_dim = out_dim if out_dim is not None else channels
        self.final_res_block = block(dim * 2, dim)
        self.final_conv = nn.Conv2d(dim, self.out_dim, 1)
        self._cache = _PackCache()
        self.table_timesteps = 1000  # rows of the cached FiLM table for integer diffusion times
        # Operand precision of the tensor-core convolutions.  FP16 and TF32 carry the same 10-bit mantissa (identical
        # rounding error on eps); FP16 runs at twice the tensor rate with half the operand bytes but needs every conv
        # input channel count to be a multiple of 64 and activations within +-65504 (GroupNorm/LayerNorm-bounded here).
        # Override with `net.precision = "tf32"` (or SDC_PRECISION=tf32) for checkpoints with extreme activation ranges.
        default = "f16" if dim % 64 == 0 and init_dim % 64 == 0 else "tf32"
        self.precision = os.environ.get("SDC_PRECISION", default)
        # FP16 inference only: convolution outputs that feed a GroupNorm / LayerNorm (and the 1x1 res_conv output) are stored as
        # fp16 instead of fp32.  The norm statistics still come from the fp32 accumulators; only the stored activations carry one
        # more 2^-11 rounding (eps error 6.3e-4 -> 6.6e-4 relative on the dim-128 model with the last block kept in fp32), and the
        # HBM-bound norm kernels move 4 instead of 6 bytes per element.  `net.compact_intermediates = False` (or SDC_COMPACT=0)
        # keeps fp32.  The recording (backward) path always keeps fp32.
        self.compact_intermediates = os.environ.get("SDC_COMPACT", "1") != "0"
        # EXPERIMENTAL (off): the 3x3 convolutions of the 16x128 level can normalise their own output (sdc_conv3x3_row_gn:
        # deferred epilogue, GroupNorm + FiLM + SiLU (+ residual, + the 1x1 head for the last block) applied to the fp32
        # accumulators in TMEM) instead of a separate GroupNorm pass over HBM.  Correct (tests) and more accurate (one rounding
        # site fewer per norm) but SLOWER on B200: 875 us against 417 us conv + 180 us GroupNorm at B = 1024 -- the per-item
        # epilogue (two TMEM passes, cross-cluster exchange of the statistics, SiLU, stores) does not fit into one MMA period
        # (profiles/r02_row_gn_*.txt).  SDC_FUSE_GN=1 enables it.
        self.fuse_groupnorm = os.environ.get("SDC_FUSE_GN", "0") == "1"
        # EXPERIMENTAL (off): FP16 executor, LinearAttention blocks with <= 256 channels -- both channel LayerNorms folded into the
        # qkv / output-projection epilogues (sdc_conv1x1_qkv_ln, sdc_conv1x1_per_sample_ln): 1.24 ms of LayerNorm passes disappear
        # per step at B = 1024, but the 1x1 convolutions are epilogue bound already and lose 1.7 ms (profiles/r02_ln_fusion_ab.txt).
        # SDC_FUSE_LN=1 enables it.
        self.fuse_layernorm = int(os.environ.get("SDC_FUSE_LN", "0"))   # 0 off, 1 both LayerNorms, 2 PreNorm only
        # FP16-range guard (GaussianDiffusion._run_chain): on non-finite eps the chain is repeated with TF32 operands
        self.overflow_fallback = True

    # ------------------------------------------------------------------ weight packing / FiLM table
    def _resnet_blocks(self):
        for lvl in self.downs:
            yield lvl[0]
            yield lvl[1]
        yield self.mid_block1
        yield self.mid_block2
        for lvl in self.ups:
            yield lvl[0]
            yield lvl[1]
        yield self.final_res_block

    def _prec(self):
        if self.precision not in PREC_NAMES:
            raise ValueError(f"safediffcon_b200.Unet2D.precision must be 'f16' or 'tf32', got {self.precision!r}")
        if self.precision == "f16" and (self.dim % 64 != 0 or self.init_conv.weight.shape[0] % 64 != 0):
            raise ValueError("precision 'f16' needs dim and init_dim to be multiples of 64 (128-byte K blocks of fp16 channels)")
        return PREC_NAMES[self.precision]

    def _key(self):
        return (self.precision,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _fingerprint(self):
        """Device-side digest of the parameter VALUES: (L2 norm, L1 norm) of every parameter, [2, n_params] fp32.  The version
        counters of `_key` miss in-place writes made through `p.data` (ema_pytorch's `ema_p.data.lerp_(...)`,
        `p.data.copy_(...)`), which is exactly how the reference's EMA copy -- the module InferenceFT samples from
        (/root/reference/1D/inference/inference_ft.py:167-171,212-222) -- is updated."""
        ps = [p.detach() for p in self.parameters()]
        with torch.no_grad():
            return torch.stack([torch.stack(torch._foreach_norm(ps, 2)), torch.stack(torch._foreach_norm(ps, 1))])

    def invalidate_packed(self):
        """Force the next evaluation to repack the weights / FiLM table (buffers are refreshed in place, so captured graphs stay valid)."""
        self._cache.key = None
        if self._cache.plan is not None:
            self._cache.plan.key = None

    def revalidate_packed(self):
        """Called once per sampling chain (GaussianDiffusion._run_chain) and by anyone who has written parameters behind
        autograd's back: compares the live parameter digest with the one taken when the pack was built (one small D2H sync) and
        invalidates the pack if any parameter changed.  Returns True when the pack was still valid."""
        c = self._cache
        if c.fingerprint is None or (c.key is None and (c.plan is None or c.plan.key is None)):
            return False
        key = self._key()
        stale = (c.key is not None and c.key != key) or (c.plan is not None and c.plan.key is not None and c.plan.key != key)
        if stale or not torch.equal(c.fingerprint, self._fingerprint()):
            self.invalidate_packed()
            return False
        return True

    # ------------------------------------------------------------------ C++ executor (inference)
    def _plan_ready(self, backward=False):
        """The C++ executor with weights packed for the current parameter values, or None when this configuration runs on the
        Python schedule (SDC_NO_PLAN=1, or FP16 with fp32 intermediates).  backward: also pack the data-gradient weights."""
        if not USE_PLAN or (self.precision == "f16" and not self.compact_intermediates):
            return None
        prec = self._prec()
        c = self._cache
        plan = c.plan
        if plan is None or plan.prec != prec or plan.table_timesteps != self.table_timesteps:
            dims = [self.downs[i][0].dim_out // self.dim for i in range(1, len(self.downs))] + [self.mid_block1.dim_out // self.dim]
            plan = UnetPlan(self.dim, dims, self.channels, self.out_dim, prec, self.theta, self.table_timesteps)
            c.plan, c.plan_gen = plan, c.plan_gen + 1
        if (plan.flags != (self.fuse_layernorm, self.fuse_groupnorm)):
            plan.set_flag(UnetPlan.FUSE_LN, int(self.fuse_layernorm))
            plan.set_flag(UnetPlan.FUSE_GN, self.fuse_groupnorm)
            plan.flags = (self.fuse_layernorm, self.fuse_groupnorm)
        if backward and not plan.backward:
            plan.set_flag(UnetPlan.BACKWARD, 1)
            plan.backward, plan.key = True, None   # repack below
        key = self._key()
        if plan.key != key:
            dev = self.init_conv.weight.device
            if dev.type != "cuda":
                raise RuntimeError("safediffcon_b200.Unet2D: parameters are on the CPU; move the module to a CUDA device "
                                   "(there is no CPU fallback)")
            with torch.no_grad(), torch.cuda.device(dev):
                plan.pack(dict(self.named_parameters()))
                plan.key = key
                c.fingerprint = self._fingerprint()
        return plan

    def take_nonfinite(self):
        """Number of non-finite eps entries the FP16 executor has produced since the last call (one D2H sync); resets the count.
        GaussianDiffusion checks it once per chain: fp16 activations beyond +-65504 (checkpoints with extreme pre-norm ranges)
        surface here instead of silently turning into Inf/NaN samples."""
        plan = self._cache.plan
        if plan is None or plan.nonfinite is None:
            return 0
        n = int(plan.nonfinite.item())
        if n:
            plan.nonfinite.zero_()
        return n

    def _packed(self):
        key = self._key()
        if self._cache.pack is not None and key == self._cache.key:
            return self._cache.pack
        dev = self.init_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("safediffcon_b200.Unet2D: parameters are on the CPU; move the module to a CUDA device "
                               "(there is no CPU fallback)")
        prec = self._prec()
        pk = {"prec": prec}
        # every tensor of the pack owns its storage (.clone()): the in-place refresh below must never write through an alias of
        # a parameter (it would bump the parameter's version counter -- invalidating this very cache and autograd's saved tensors)
        with torch.no_grad(), torch.cuda.device(dev):
            def conv(m, kind):
                return dict(w=pack_conv_weight(kind, m.weight, prec), b=None if m.bias is None else m.bias.detach().float().clone(),
                            cout=m.weight.shape[0], mod=m)

            def rb(m):
                return dict(c1=conv(m.block1.proj, KIND_3x3), c2=conv(m.block2.proj, KIND_3x3),
                            g1=(m.block1.norm.weight.detach().float().clone(), m.block1.norm.bias.detach().float().clone()),
                            g2=(m.block2.norm.weight.detach().float().clone(), m.block2.norm.bias.detach().float().clone()),
                            res=conv(m.res_conv, KIND_1x1) if isinstance(m.res_conv, nn.Conv2d) else None, cout=m.dim_out, mod=m)

            def at(m):
                inner = m.fn.fn
                d = dict(g_in=m.fn.norm.g.detach().float().reshape(-1).clone(), qkv=conv(inner.to_qkv, KIND_1x1), mod=m)
                if isinstance(inner, _LinearAttention):
                    d.update(out=conv(inner.to_out[0], KIND_1x1), g_out=inner.to_out[1].g.detach().float().reshape(-1).clone(), full=False,
                             out_w32=inner.to_out[0].weight.detach().float().reshape(inner.to_out[0].weight.shape[0], -1).clone())
                else:
                    d.update(out=conv(inner.to_out, KIND_1x1), g_out=None, full=True)
                return d

            pk["downs"] = []
            for lvl in self.downs:
                is_unshuffle = isinstance(lvl[3], nn.Sequential)
                pk["downs"].append(dict(b1=rb(lvl[0]), b2=rb(lvl[1]), attn=at(lvl[2]), unshuffle=is_unshuffle,
                                        down=conv(lvl[3][1], KIND_UNSHUFFLE) if is_unshuffle else conv(lvl[3], KIND_3x3)))
            pk["mid1"], pk["mid_attn"], pk["mid2"] = rb(self.mid_block1), at(self.mid_attn), rb(self.mid_block2)
            pk["ups"] = []
            for lvl in self.ups:
                is_up = isinstance(lvl[3], nn.Sequential)
                up = conv(lvl[3][1] if is_up else lvl[3], KIND_3x3)
                if is_up:   # inference: nearest-upsample folded into four 2x2 phase convolutions (conv_gemm.cu kind 3)
                    up["w_up"] = pack_conv_weight(KIND_UP2X, lvl[3][1].weight, prec)
                pk["ups"].append(dict(b1=rb(lvl[0]), b2=rb(lvl[1]), attn=at(lvl[2]), upsample=is_up, up=up))
            pk["final"] = rb(self.final_res_block)
            # stem 7x7 as a tcgen05 GEMM over an im2col operand whose K axis holds the input's high and low parts (see
            # sdc_stem_im2col): W[c, Cin*49] repeated at columns 0 and kp/2 of a [c, kp] matrix
            ws = self.init_conv.weight.detach().float()
            c_stem, k_stem = ws.shape[0], ws[0].numel()
            kh = (k_stem + 31) // 32 * 32
            kp = 2 * kh if (2 * kh) % 64 == 0 else 2 * (kh + 32)
            wrep = torch.zeros(c_stem, kp, device=dev, dtype=torch.float32)
            wrep[:, :k_stem] = ws.reshape(c_stem, k_stem)
            wrep[:, kp // 2:kp // 2 + k_stem] = ws.reshape(c_stem, k_stem)
            pk["stem"] = dict(w=pack_conv_weight(KIND_1x1, wrep.reshape(c_stem, kp, 1, 1), prec),
                              b=self.init_conv.bias.detach().float().clone(), cout=c_stem, kp=kp)
            pk["head"] = (self.final_conv.weight.detach().float().reshape(self.out_dim, -1).clone(),
                          self.final_conv.bias.detach().float().clone())
            # all ResnetBlock FiLM projections stacked: one [E_total, time_dim] matrix, block i owns rows [off, off+2*Cout)
            ws, bs, off = [], [], 0
            for i, m in enumerate(self._resnet_blocks()):
                ws.append(m.mlp[1].weight.detach().float())
                bs.append(m.mlp[1].bias.detach().float())
                m._film_off = off
                off += m.mlp[1].weight.shape[0]
            pk["film_w"], pk["film_b"], pk["film_total"] = torch.cat(ws).contiguous(), torch.cat(bs).contiguous(), off
            pk["time"] = tuple(t.detach().float().clone() for t in (self.time_mlp[1].weight, self.time_mlp[1].bias,
                                                                            self.time_mlp[3].weight, self.time_mlp[3].bias))
            pk["table"] = None
            old = self._cache.pack
            if old is not None and _same_layout(old, pk):
                # parameters changed (optimiser / EMA step) but not their layout: refresh the existing buffers in place so
                # that captured reverse-step graphs (GaussianDiffusion._graph_entry) stay valid
                _copy_into(old, pk)
                old.pop("dgrad", None)
                if old["table"] is not None:
                    tab, old["table"] = old["table"], None
                    tab.copy_(self._film_table(old))
                    old["table"] = tab
                pk = old
        self._cache.pack, self._cache.key, self._cache.fingerprint = pk, key, self._fingerprint()
        return pk

    def _film_rows(self, pk, t_float):
        """FiLM (scale | shift) rows [R, E_total] for R diffusion times (float32 tensor on the device)."""
        R, dim, td = t_float.shape[0], self.dim, self.dim * 4
        dev = t_float.device
        emb = torch.empty(R, dim, device=dev)
        h1 = torch.empty(R, td, device=dev)
        h2 = torch.empty(R, td, device=dev)
        out = torch.empty(R, pk["film_total"], device=dev)
        w1, b1, w2, b2 = pk["time"]
        lib = L.lib()
        L.check(lib.sdc_sinusoidal_embedding(L.ptr(t_float), L.ptr(emb), R, dim, float(self.theta), _st()))
        L.check(lib.sdc_linear_rows(L.ptr(emb), L.ptr(w1), L.ptr(b1), L.ptr(h1), R, dim, td, 0, _st()))
        L.check(lib.sdc_linear_rows(L.ptr(h1), L.ptr(w2), L.ptr(b2), L.ptr(h2), R, td, td, 2, _st()))
        L.check(lib.sdc_linear_rows(L.ptr(h2), L.ptr(pk["film_w"]), L.ptr(pk["film_b"]), L.ptr(out), R, td, pk["film_total"], 1, _st()))
        return out

    def _film_table(self, pk):
        if pk["table"] is not None and pk["table"].shape[0] != self.table_timesteps:
            pk["table"] = None   # table_timesteps was raised (a GaussianDiffusion with more steps adopted this denoiser)
            pk["table_gen"] = pk.get("table_gen", 0) + 1   # part of the captured-graph key: graphs hold the table's address
        if pk["table"] is None:
            dev = pk["film_w"].device
            t = torch.arange(self.table_timesteps, device=dev, dtype=torch.float32)
            pk["table"] = self._film_rows(pk, t)
        return pk["table"]

    # ------------------------------------------------------------------ forward
    def denoise_uniform(self, x, t_int):
        """eps for a batch-uniform integer diffusion time (sampler fast path: no per-sample index tensor).  Inference only:
        never records parameter gradients (`forward` does, like the reference module under autograd)."""
        return self._forward(x, table_row=int(t_int), param_grad=False)

    def forward(self, x, time, x_self_cond=None, residual=None):
        if x_self_cond is not None or residual is not None:
            raise NotImplementedError("self-conditioning / residual conditioning are outside the 1D hot path")
        return self._forward(x, time=time)

    def _conv_norm_params(self):
        """Parameters whose gradients the CUDA backward produces directly (everything except the time-embedding MLPs)."""
        return [p for n, p in self.named_parameters() if not n.startswith("time_mlp.") and ".mlp." not in n]

    def wants_param_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _film_rows_autograd(self, B, time, table_row, device):
        """FiLM (scale | shift) rows [B, E_total] with torch ops, differentiable w.r.t. the time-MLP parameters
        (reference unet.py:81-95,152-155,310-315); column layout = the packed FiLM table's."""
        if table_row is not None:
            t = torch.full((B,), float(table_row), device=device)
        else:
            t = time.to(device=device, dtype=torch.float32)
        half = self.dim // 2
        freq = torch.exp(torch.arange(half, device=device) * -(math.log(self.theta) / (half - 1)))
        emb = t[:, None] * freq[None, :]
        h = self.time_mlp(torch.cat((emb.sin(), emb.cos()), dim=-1))
        return torch.cat([m.mlp(h) for m in self._resnet_blocks()], dim=1)

    def _forward(self, x, time=None, table_row=None, param_grad=True):
        x = L.dev_f32(x, "x")
        if param_grad and self.wants_param_grad():
            # training forward: activations are recorded, backward yields parameter gradients
            self._packed()   # fixes the FiLM column offsets
            film = self._film_rows_autograd(x.shape[0], time, table_row, x.device)
            return _TrainFn.apply(x, film, self, *self._conv_norm_params())
        if torch.is_grad_enabled() and x.requires_grad:
            # autograd w.r.t. the INPUT (guidance callables that differentiate through the denoiser); parameters get no
            # gradient from this path (fine-tuning is SURVEY.md section 8f)
            return _InputVJP.apply(x, self, time, table_row)
        with torch.no_grad(), torch.cuda.device(x.device):
            return self._run(x, time, table_row)

    def vjp(self, x, time, grad_eps):
        """(eps, d<eps, grad_eps>/dx) in one call: forward with saved activations, then the backward-data pass."""
        x = L.dev_f32(x, "x")
        if self.init_conv.weight.shape[0] == self.dim and (isinstance(time, int) or not torch.is_floating_point(time)):
            plan = self._plan_ready(backward=True)
            if plan is not None:   # ONE C call: recording forward + reverse walk inside the executor (sdc_unet_backward_data)
                with torch.no_grad(), torch.cuda.device(x.device):
                    g = L.dev_f32(grad_eps, "grad_eps")
                    if isinstance(time, int):
                        if not 0 <= time < self.table_timesteps:
                            raise ValueError(f"safediffcon_b200.Unet2D: diffusion time {time} outside the FiLM table [0, {self.table_timesteps})")
                        return plan.vjp(x, g, None, time)
                    if bool(((time < 0) | (time >= self.table_timesteps)).any()):
                        raise ValueError(f"safediffcon_b200.Unet2D: integer diffusion times must lie in [0, {self.table_timesteps})")
                    return plan.vjp(x, g, time.to(device=x.device, dtype=torch.int32).reshape(-1).contiguous(), 0)
        tape = []
        with torch.no_grad(), torch.cuda.device(x.device):
            if isinstance(time, int):
                eps = self._run(x, None, time, tape)
            else:
                eps = self._run(x, time, None, tape)
            return eps, self._vjp(tape, L.dev_f32(grad_eps, "grad_eps"))

    def denoise_indexed(self, x, t_index, workspace=None):
        """eps with the integer diffusion times read from a device int32 tensor [B] that the caller updates in place
        (captured-graph chains: the same launches serve every step; see GaussianDiffusion._graph_chain)."""
        assert t_index.dtype == torch.int32 and t_index.is_cuda and t_index.numel() == x.shape[0]
        with torch.no_grad(), torch.cuda.device(x.device):
            x = L.dev_f32(x, "x")
            if workspace is not None:
                return self._run_plan(self._plan_ready(), x, None, None, t_index, workspace)
            return self._run(x, None, None, t_index=t_index)

    def _run_plan(self, plan, x, time, table_row, t_index, workspace=None):
        """Inference through the C++ executor: ONE call, activations in the plan's workspace."""
        B = x.shape[0]
        assert x.shape[1] == self.channels
        if table_row is not None:
            if not 0 <= table_row < self.table_timesteps:
                raise ValueError(f"safediffcon_b200.Unet2D: diffusion time {table_row} outside the FiLM table [0, {self.table_timesteps}); "
                                 "set net.table_timesteps (GaussianDiffusion sizes it from its timesteps)")
            return plan.forward(x, None, int(table_row), workspace)
        if t_index is None:
            if time.numel() != B:
                raise ValueError(f"safediffcon_b200.Unet2D: time must hold one entry per sample ({B}), got {tuple(time.shape)}")
            if bool(((time < 0) | (time >= self.table_timesteps)).any()):
                raise ValueError(f"safediffcon_b200.Unet2D: integer diffusion times must lie in [0, {self.table_timesteps}); "
                                 "pass float times or raise net.table_timesteps")
            t_index = time.to(device=x.device, dtype=torch.int32).reshape(B).contiguous()
        return plan.forward(x, t_index, 0, workspace)

    def _run(self, x, time, table_row, tape=None, t_index=None, film_rows=None):
        """tape: None for inference (buffers are reused); a list to record what the backward-data pass needs (every
        normalisation input is then kept in its own buffer)."""
        keep = tape is not None
        if not keep and film_rows is None and (table_row is not None or t_index is not None or
                                               (time is not None and not torch.is_floating_point(time))):
            plan = self._plan_ready() if self.init_conv.weight.shape[0] == self.dim else None
            if plan is not None:
                return self._run_plan(plan, x, time, table_row, t_index)
        pk = self._packed()
        lib = L.lib()
        dev = x.device
        prec = pk["prec"]
        od = operand_dtype(prec)
        B, Cin, H, W = x.shape
        assert Cin == self.channels
        # FiLM rows: integer times index the cached 1000-row table; anything else is evaluated per sample
        if film_rows is not None:
            film, t_index = film_rows, torch.arange(B, device=dev, dtype=torch.int32)
        elif t_index is not None:
            film = self._film_table(pk)
        elif table_row is not None:
            if not 0 <= table_row < self.table_timesteps:
                raise ValueError(f"safediffcon_b200.Unet2D: diffusion time {table_row} outside the FiLM table [0, {self.table_timesteps}); "
                                 "set net.table_timesteps (GaussianDiffusion sizes it from its timesteps)")
            film, t_index = self._film_table(pk)[table_row:table_row + 1], None
        elif not torch.is_floating_point(time):
            # per-sample integer times (user-level forward / model_predictions; the chains use table_row / t_index): one small
            # host check instead of a silent clamp -- the reference's model_predictions synchronises here too (t[0].item())
            if time.numel() != B:
                raise ValueError(f"safediffcon_b200.Unet2D: time must hold one entry per sample ({B}), got {tuple(time.shape)}")
            if bool(((time < 0) | (time >= self.table_timesteps)).any()):
                raise ValueError(f"safediffcon_b200.Unet2D: integer diffusion times must lie in [0, {self.table_timesteps}); "
                                 "pass float times or raise net.table_timesteps")
            film, t_index = self._film_table(pk), time.to(device=dev, dtype=torch.int32).reshape(B).contiguous()
        else:
            film, t_index = self._film_rows(pk, time.to(device=dev, dtype=torch.float32).contiguous()), \
                torch.arange(B, device=dev, dtype=torch.int32)
        E = pk["film_total"]
        n_gn = 2 * sum(1 for _ in self._resnet_blocks())
        cmp_early = od == torch.float16 and not keep and self.compact_intermediates
        stats = torch.zeros(n_gn, B, 2, device=dev, dtype=torch.float64)
        # exchange slots of the fused conv + GroupNorm kernels (SDC_GN_SLOT_BYTES = 1024 per norm and sample, every byte 0xFF)
        counters = torch.full((n_gn, B, 256), -1, device=dev, dtype=torch.int32) if (cmp_early and self.fuse_groupnorm) else None
        stat_i = [0]
        f32 = lambda rows, c: torch.empty(rows, c, device=dev, dtype=torch.float32)  # noqa: E731  (conv outputs ahead of a norm)
        opd = lambda rows, c: torch.empty(rows, c, device=dev, dtype=od)  # noqa: E731  (tensor-core operands)
        cmp = cmp_early

        def conv(kind, a0, c0, a1, c1, cw, residual, out, st, operand_out, h, w, algo_k=None):
            conv_gemm(kind, a0, c0, a1, c1, cw["w"], cw["b"], residual, out, st, operand_out, B, h, w, cw["cout"], prec, algo_k)

        def resnet(p, m, a0, c0, a1, c1, h, w, head=None):
            """ResnetBlock (unet.py:166-180) on one or two concatenated NHWC inputs -> operand [B*h*w, Cout].
            head = (w, b, out): the block is the last one; its second normalisation is fused with the 1x1 head convolution and the
            block keeps fp32 norm inputs (the last rounding sites in front of the output dominate the eps error: with them exact the
            compact path is as accurate as fp32 intermediates everywhere, 6.3-7.0e-4 instead of 7.4-8.3e-4)."""
            M, cout = B * h * w, p["cout"]
            s1, s2 = stats[stat_i[0]], stats[stat_i[0] + 1]
            stat_i[0] += 2
            if cmp and self.fuse_groupnorm and w == 128 and h % 4 == 0 and cout <= 128 and USE_ROW_KERNEL and (head is None or self.out_dim <= 4):
                # full-resolution level: GroupNorm + FiLM + SiLU (+ residual, + 1x1 head) applied to the fp32 accumulators by the
                # conv kernel itself (sdc_conv3x3_row_gn, deferred epilogue): no separate normalisation pass over HBM
                si = stat_i[0] - 2
                n1, n2 = counters[si], counters[si + 1]
                ss = film[:, m._film_off:]
                h1 = opd(M, cout)
                rc = conv_row_gn(a0, c0, a1, c1, p["c1"], h1, s1, n1, p["g1"], ss, t_index, E, None, B, h, w, cout)
                if rc == 0:
                    if p["res"] is not None:
                        res = opd(M, cout)
                        conv(KIND_1x1, a0, c0, a1, c1, p["res"], None, res, None, True, h, w)
                    else:
                        assert a1 is None
                        res = a0
                    out2 = None if head is not None else opd(M, cout)
                    rc = conv_row_gn(h1, cout, None, 0, p["c2"], out2, s2, n2, p["g2"], None, None, 0, res, B, h, w, cout, head=head)
                    assert rc == 0
                    return out2
                # not eligible (nothing was launched): fall through to the unfused sequence
            if head is not None:
                raw = f32(M, cout)
                conv(KIND_3x3, a0, c0, a1, c1, p["c1"], None, raw, s1, False, h, w)
                ss = film[:, m._film_off:]
                h1 = opd(M, cout)
                L.check(lib.sdc_gn_silu(prec, L.ptr(raw), 0, L.ptr(s1), L.ptr(p["g1"][0]), L.ptr(p["g1"][1]), L.ptr(ss), L.ptr(t_index), E,
                                        None, 0, L.ptr(h1), B, h * w, cout, _st()))
                conv(KIND_3x3, h1, cout, None, 0, p["c2"], None, raw, s2, False, h, w)   # conv1's output is dead: reuse its buffer
                if p["res"] is not None:
                    res = h1   # dead after conv2
                    conv(KIND_1x1, a0, c0, a1, c1, p["res"], None, res, None, True, h, w)
                else:
                    assert a1 is None
                    res = a0
                L.check(lib.sdc_gn_silu_head(L.ptr(raw), L.ptr(s2), L.ptr(p["g2"][0]), L.ptr(p["g2"][1]), L.ptr(res), 1, L.ptr(head[0]),
                                             L.ptr(head[1]), L.ptr(head[2]), B, h * w, cout, head[2].shape[1], _st()))
                return None
            if cmp:
                # compact intermediates (FP16 inference): the conv epilogue takes the GroupNorm sums from its fp32 accumulators and
                # stores fp16; normalisation runs in place; the 1x1 res_conv output lands in conv1's dead buffer, also fp16
                raw = opd(M, cout)
                conv(KIND_3x3, a0, c0, a1, c1, p["c1"], None, raw, s1, True, h, w)
                ss = film[:, m._film_off:]
                L.check(lib.sdc_gn_silu(prec, L.ptr(raw), 1, L.ptr(s1), L.ptr(p["g1"][0]), L.ptr(p["g1"][1]), L.ptr(ss), L.ptr(t_index), E,
                                        None, 0, L.ptr(raw), B, h * w, cout, _st()))
                raw2 = opd(M, cout)
                conv(KIND_3x3, raw, cout, None, 0, p["c2"], None, raw2, s2, True, h, w)
                if p["res"] is not None:
                    res = raw
                    conv(KIND_1x1, a0, c0, a1, c1, p["res"], None, res, None, True, h, w)
                else:
                    assert a1 is None
                    res = a0
                L.check(lib.sdc_gn_silu(prec, L.ptr(raw2), 1, L.ptr(s2), L.ptr(p["g2"][0]), L.ptr(p["g2"][1]), None, None, 0, L.ptr(res),
                                        1, L.ptr(raw2), B, h * w, cout, _st()))
                return raw2
            raw = f32(M, cout)
            conv(KIND_3x3, a0, c0, a1, c1, p["c1"], None, raw, s1, False, h, w)
            ss = film[:, m._film_off:]
            h1 = raw if (od == torch.float32 and not keep) else opd(M, cout)   # TF32 inference: in place
            L.check(lib.sdc_gn_silu(prec, L.ptr(raw), 0, L.ptr(s1), L.ptr(p["g1"][0]), L.ptr(p["g1"][1]), L.ptr(ss), L.ptr(t_index), E,
                                    None, 0, L.ptr(h1), B, h * w, cout, _st()))
            raw2 = f32(M, cout)
            conv(KIND_3x3, h1, cout, None, 0, p["c2"], None, raw2, s2, False, h, w)
            if p["res"] is not None:
                # fp32 scratch: conv1's output is dead once h1 exists (TF32 mode: h1 itself, dead after conv2)
                res, res_operand = (f32(M, cout) if keep else raw), 0
                conv(KIND_1x1, a0, c0, a1, c1, p["res"], None, res, None, False, h, w)
            else:
                assert a1 is None
                res, res_operand = a0, 1
            # TF32 mode: in place; F16 mode: h1's buffer is dead after conv2
            out = opd(M, cout) if keep else (raw2 if od == torch.float32 else h1)
            if keep:
                tape.append(("resnet", dict(p=p, c0=c0, c1=c1, h=h, w=w, raw1=raw, s1=s1, raw2=raw2, s2=s2, ss=ss, a0=a0, a1=a1, h1=h1,
                                            film_off=m._film_off)))
            L.check(lib.sdc_gn_silu(prec, L.ptr(raw2), 0, L.ptr(s2), L.ptr(p["g2"][0]), L.ptr(p["g2"][1]), None, None, 0, L.ptr(res),
                                    res_operand, L.ptr(out), B, h * w, cout, _st()))
            return out

        def attention(p, xin, c, h, w):
            """Residual(PreNorm(LinearAttention | Attention)) (unet.py:16-22,65-76,182-258)."""
            M, n = B * h * w, h * w
            xn = opd(M, c)
            L.check(lib.sdc_channel_layernorm(prec, L.ptr(xin), 1, L.ptr(p["g_in"]), None, L.ptr(xn), M, c, 1, _st()))
            hid = HEADS * DIM_HEAD
            if not p["full"] and not keep and n % 128 == 0:
                # fused inference path: q-softmax in the qkv epilogue, context folded into a per-sample output projection
                qs, kv = opd(M, hid), opd(M, 2 * hid)   # F16 mode: k | v in fp16 (half the bytes through the context pass)
                conv1x1_qkv(xn, c, p["qkv"]["w"], qs, kv, B, h, w, prec)
                ws = torch.empty(lib.sdc_linear_attention_workspace(B), device=dev, dtype=torch.uint8)
                L.check(lib.sdc_linear_attention_context(L.ptr(kv), ctypes.c_void_p(kv.data_ptr() + kv.element_size() * hid), 2 * hid,
                                                         int(kv.dtype == torch.float16), L.ptr(ws), B, n, _st()))
                wf = opd(B * c, hid)
                L.check(lib.sdc_linear_attention_fold(prec, L.ptr(ws), L.ptr(p["out_w32"]), L.ptr(wf), B, c, _st()))
                proj = opd(M, c) if cmp else f32(M, c)
                conv1x1_per_sample(qs, hid, wf, p["out"]["b"], proj, B, h, w, c, prec)
                out = xn  # reuse: LN1's output is dead once q / kv exist
                L.check(lib.sdc_channel_layernorm(prec, L.ptr(proj), int(cmp), L.ptr(p["g_out"]), L.ptr(xin), L.ptr(out), M, c, 1, _st()))
                return out
            qkv = f32(M, 3 * HEADS * DIM_HEAD)
            conv(KIND_1x1, xn, c, None, 0, p["qkv"], None, qkv, None, False, h, w)
            att = opd(M, HEADS * DIM_HEAD)
            if p["full"]:
                L.check(lib.sdc_attention(prec, L.ptr(qkv), L.ptr(att), B, n, _st()))
                out = opd(M, c) if keep else xn  # inference: reuse LN1's buffer
                conv(KIND_1x1, att, HEADS * DIM_HEAD, None, 0, p["out"], xin, out, None, True, h, w)
                if keep:
                    tape.append(("attn", dict(p=p, c=c, h=h, w=w, xin=xin, qkv=qkv, xn=xn, att=att)))
                return out
            ws = torch.empty(lib.sdc_linear_attention_workspace(B), device=dev, dtype=torch.uint8)
            L.check(lib.sdc_linear_attention(prec, L.ptr(qkv), L.ptr(att), L.ptr(ws), B, n, _st()))
            proj = f32(M, c)
            conv(KIND_1x1, att, HEADS * DIM_HEAD, None, 0, p["out"], None, proj, None, False, h, w)
            out = opd(M, c) if keep else xn  # inference: LN1's output is dead once qkv exists
            L.check(lib.sdc_channel_layernorm(prec, L.ptr(proj), 0, L.ptr(p["g_out"]), L.ptr(xin), L.ptr(out), M, c, 1, _st()))
            if keep:
                tape.append(("attn", dict(p=p, c=c, h=h, w=w, xin=xin, qkv=qkv, proj=proj, ws=ws, xn=xn, att=att)))
            return out

        c = self.init_conv.weight.shape[0]
        cur = opd(B * H * W, c)
        rc = -1
        if prec == PREC_F16 and USE_STEM_TC:   # the whole stem in one tcgen05 kernel (no patch matrix in HBM)
            rc = lib.sdc_stem_conv7_tc(L.ptr(x), L.ptr(pk["stem"]["w"]), L.ptr(pk["stem"]["b"]), L.ptr(cur), B, Cin, H, W, c, pk["stem"]["kp"], _st())
            if rc > 0:
                L.check(rc)
        if rc != 0:
            patches = opd(B * H * W, pk["stem"]["kp"])
            L.check(lib.sdc_stem_im2col(prec, L.ptr(x), L.ptr(patches), B, Cin, H, W, pk["stem"]["kp"], _st()))
            conv(KIND_1x1, patches, pk["stem"]["kp"], None, 0, pk["stem"], None, cur, None, True, H, W, algo_k=Cin * 49)
            del patches
        r, r_c = cur, c
        h, w = H, W
        skips = []
        rec = tape.append if keep else (lambda item: None)
        rec(("stem", dict(B=B, Cin=Cin, H=H, W=W, c=c, t_index=t_index, E=E, x=x, film=film)))
        for lvl_m, lvl in zip(self.downs, pk["downs"]):
            cur = resnet(lvl["b1"], lvl_m[0], cur, c, None, 0, h, w)
            skips.append((cur, c))
            rec(("push", None))
            cur = resnet(lvl["b2"], lvl_m[1], cur, c, None, 0, h, w)
            cur = attention(lvl["attn"], cur, c, h, w)
            skips.append((cur, c))
            rec(("push", None))
            cout = lvl["down"]["cout"]
            if lvl["unshuffle"]:
                h, w = h // 2, w // 2
                nxt = opd(B * h * w, cout)
                conv(KIND_UNSHUFFLE, cur, c, None, 0, lvl["down"], None, nxt, None, True, h, w)
            else:
                nxt = opd(B * h * w, cout)
                conv(KIND_3x3, cur, c, None, 0, lvl["down"], None, nxt, None, True, h, w)
            rec(("down", dict(p=lvl["down"], unshuffle=lvl["unshuffle"], c=c, h=h, w=w, inp=cur)))
            cur, c = nxt, cout
        cur = resnet(pk["mid1"], self.mid_block1, cur, c, None, 0, h, w)
        cur = attention(pk["mid_attn"], cur, c, h, w)
        cur = resnet(pk["mid2"], self.mid_block2, cur, c, None, 0, h, w)
        for lvl_m, lvl in zip(self.ups, pk["ups"]):
            s, sc = skips.pop()
            cur = resnet(lvl["b1"], lvl_m[0], cur, c, s, sc, h, w)
            c = lvl["b1"]["cout"]
            s, sc = skips.pop()
            cur = resnet(lvl["b2"], lvl_m[1], cur, c, s, sc, h, w)
            cur = attention(lvl["attn"], cur, c, h, w)
            cout = lvl["up"]["cout"]
            if lvl["upsample"] and not keep and FUSE_UPSAMPLE and (w == 16 or w % 32 == 0):
                nxt = opd(B * 4 * h * w, cout)
                conv_gemm(KIND_UP2X, cur, c, None, 0, lvl["up"]["w_up"], lvl["up"]["b"], None, nxt, None, True, B, h, w, cout, prec)
                h, w = 2 * h, 2 * w
                cur, c = nxt, cout
                continue
            if lvl["upsample"]:
                up = opd(B * 4 * h * w, c)
                L.check(lib.sdc_upsample2x(prec, L.ptr(cur), L.ptr(up), B, h, w, c, _st()))
                h, w = 2 * h, 2 * w
                cur = up
            nxt = opd(B * h * w, cout)
            conv(KIND_3x3, cur, c, None, 0, lvl["up"], None, nxt, None, True, h, w)
            rec(("up", dict(p=lvl["up"], upsample=lvl["upsample"], c=c, h=h, w=w, inp=cur)))
            cur, c = nxt, cout
        # zeros: the fused conv + GroupNorm + head kernel accumulates the two channel halves of a pixel into it
        out = torch.zeros(B, self.out_dim, H, W, device=dev, dtype=torch.float32)
        if cmp and pk["final"]["cout"] == 128 and self.out_dim <= 4:
            resnet(pk["final"], self.final_res_block, cur, c, r, r_c, h, w, head=(pk["head"][0], pk["head"][1], out))
            return out
        cur = resnet(pk["final"], self.final_res_block, cur, c, r, r_c, h, w)
        rec(("head", dict(inp=cur)))
        L.check(lib.sdc_head_conv1(prec, L.ptr(cur), L.ptr(pk["head"][0]), L.ptr(pk["head"][1]), L.ptr(out), B, H * W,
                                   self.final_res_block.dim_out, self.out_dim, _st()))
        return out

    # ------------------------------------------------------------------ backward-data (VJP w.r.t. x)
    def _dgrad_pack(self, pk):
        """Transposed / tap-flipped TF32 weights of every convolution, built on first use per parameter version."""
        if "dgrad" in pk:
            return pk["dgrad"]
        with torch.no_grad():
            d = {}

            def add(cw, m, kind):
                d[id(cw)] = pack_conv_weight_dgrad(kind, m.weight)

            def rb(p, m):
                add(p["c1"], m.block1.proj, KIND_3x3)
                add(p["c2"], m.block2.proj, KIND_3x3)
                if p["res"] is not None:
                    add(p["res"], m.res_conv, KIND_1x1)

            def at(p, m):
                inner = m.fn.fn
                add(p["qkv"], inner.to_qkv, KIND_1x1)
                add(p["out"], inner.to_out[0] if isinstance(inner, _LinearAttention) else inner.to_out, KIND_1x1)

            for lvl_m, lvl in zip(self.downs, pk["downs"]):
                rb(lvl["b1"], lvl_m[0]); rb(lvl["b2"], lvl_m[1]); at(lvl["attn"], lvl_m[2])
                if lvl["unshuffle"]:
                    add(lvl["down"], lvl_m[3][1], KIND_UNSHUFFLE)
                else:
                    add(lvl["down"], lvl_m[3], KIND_3x3)
            rb(pk["mid1"], self.mid_block1); at(pk["mid_attn"], self.mid_attn); rb(pk["mid2"], self.mid_block2)
            for lvl_m, lvl in zip(self.ups, pk["ups"]):
                rb(lvl["b1"], lvl_m[0]); rb(lvl["b2"], lvl_m[1]); at(lvl["attn"], lvl_m[2])
                add(lvl["up"], lvl_m[3][1] if lvl["upsample"] else lvl_m[3], KIND_3x3)
            rb(pk["final"], self.final_res_block)
            # stem: dY[M, c] x W[c, Cin*49] as a 1x1 convolution whose output rows are padded to a multiple of 64 columns
            ws = self.init_conv.weight.detach().float()
            c, k = ws.shape[0], ws[0].numel()
            kp = (k + 63) // 64 * 64
            wt = torch.zeros(kp, c, device=ws.device, dtype=torch.float32)
            L.check(L.lib().sdc_pack_conv_weight_dgrad(KIND_1x1, L.ptr(ws.reshape(c, k).contiguous()), L.ptr(wt), c, k, _st()))
            d["stem"] = (wt, kp)
        pk["dgrad"] = d
        return d

    def _vjp(self, tape, g_eps, pgrads=None, need_gx=True):
        """d<eps, g_eps>/dx for the forward recorded on `tape` (reverse walk).

        pgrads: None, or a dict that receives {id(parameter): fp32 gradient} for every convolution / norm parameter of
        the network (csrc/unet_wgrad.cu) plus "film" -> d/d(FiLM rows) [B, E_total], from which autograd reaches the
        time-MLP parameters."""
        pk = self._packed()
        dg = self._dgrad_pack(pk)
        lib = L.lib()
        dev = g_eps.device
        st0 = tape[0][1]
        B, t_index, E = st0["B"], st0["t_index"], st0["E"]
        f32 = lambda rows, c: torch.empty(rows, c, device=dev, dtype=torch.float32)  # noqa: E731
        sums = torch.empty(B, 2, device=dev, dtype=torch.float64)
        want_p = pgrads is not None
        if want_p:
            film = st0["film"]   # per-sample FiLM rows [B, E] (the training forward never uses the per-timestep table)
            assert film.shape[0] == B and t_index is not None
            pgrads["film"] = torch.zeros(B, E, device=dev, dtype=torch.float32)
            rows_of = t_index.long()

        def acc_grad(param, grad):
            key = id(param)
            pgrads[key] = grad if key not in pgrads else pgrads[key] + grad

        def conv_wgrad(kind, cw, a0, c0, a1, c1, dy, h, w):
            """weight (+ bias) gradient of convolution `cw` from its forward inputs and the gradient of its output"""
            m = cw["mod"]
            dw = torch.zeros(m.weight.shape, device=dev, dtype=torch.float32)
            conv_wgrad_any(kind, a0, c0, a1, c1, dy, dw, B, h, w, cw["cout"])
            acc_grad(m.weight, dw)
            if m.bias is not None:
                db = torch.zeros(cw["cout"], device=dev, dtype=torch.float32)
                L.check(lib.sdc_colsum(L.ptr(dy), L.ptr(db), B * h * w, cw["cout"], _st()))
                acc_grad(m.bias, db)

        def gn_pgrad(norm, dy, raw, stats, gb, ss, film_off, hw, c):
            """GroupNorm affine gradients and, for block1, the FiLM (scale, shift) gradients"""
            P = torch.zeros(B, 2, c, device=dev, dtype=torch.float32)
            L.check(lib.sdc_gn_param_grad(L.ptr(dy), L.ptr(raw), L.ptr(stats), L.ptr(gb[0]), L.ptr(gb[1]), L.ptr(ss),
                                          L.ptr(t_index) if ss is not None else None, E if ss is not None else 0, L.ptr(P),
                                          B, hw, c, _st()))
            p0, p1 = P[:, 0], P[:, 1]
            if ss is not None:
                sc1 = film[rows_of, film_off:film_off + c] + 1.0
                acc_grad(norm.weight, (sc1 * p1).sum(0))
                acc_grad(norm.bias, (sc1 * p0).sum(0))
                pgrads["film"][:, film_off:film_off + 2 * c] += torch.cat([gb[0] * p1 + gb[1] * p0, p0], dim=1)
            else:
                acc_grad(norm.weight, p1.sum(0))
                acc_grad(norm.bias, p0.sum(0))

        def ln_pgrad(gain, dy, x, M, c):
            d = torch.zeros(c, device=dev, dtype=torch.float32)
            L.check(lib.sdc_channel_layernorm_gain_grad(L.ptr(dy), L.ptr(x), int(x.dtype == torch.float16), L.ptr(d), M, c, _st()))
            acc_grad(gain.g, d.reshape(gain.g.shape))

        def dgrad(kind, g, cin_g, cw, rows, residual, out_c, operand_out, h, w):
            """Data gradient of conv `cw` restricted to input channels `rows` = (lo, hi): conv of g with Wt[lo:hi]."""
            wt = dg[id(cw)][rows[0]:rows[1]]
            out = f32(B * h * w, out_c)
            conv_gemm(kind, g, cin_g, None, 0, wt, None, residual, out, None, operand_out, B, h, w, out_c, PREC_TF32)
            return out

        def gn_bwd(dy, raw, stats, gb, ss, hw, c):
            dx = torch.empty_like(raw)
            L.check(lib.sdc_gn_silu_bwd(L.ptr(dy), L.ptr(raw), L.ptr(stats), L.ptr(gb[0]), L.ptr(gb[1]), L.ptr(ss),
                                        L.ptr(t_index) if ss is not None else None, E if ss is not None else 0, L.ptr(sums),
                                        L.ptr(dx), B, hw, c, _st()))
            return dx

        def resnet_bwd(r, g):
            """-> (grad of input segment 0, grad of input segment 1 or None); g: TF32-rounded grad of the block output."""
            p, c0, c1, h, w = r["p"], r["c0"], r["c1"], r["h"], r["w"]
            cout = p["cout"]
            d_raw2 = gn_bwd(g, r["raw2"], r["s2"], p["g2"], None, h * w, cout)
            d_h1 = dgrad(KIND_3x3, d_raw2, cout, p["c2"], (0, cout), None, cout, False, h, w)
            d_raw1 = gn_bwd(d_h1, r["raw1"], r["s1"], p["g1"], r["ss"], h * w, cout)
            if want_p:
                m = p["mod"]
                gn_pgrad(m.block2.norm, g, r["raw2"], r["s2"], p["g2"], None, 0, h * w, cout)
                conv_wgrad(KIND_3x3, p["c2"], r["h1"], cout, None, 0, d_raw2, h, w)
                gn_pgrad(m.block1.norm, d_h1, r["raw1"], r["s1"], p["g1"], r["ss"], r["film_off"], h * w, cout)
                conv_wgrad(KIND_3x3, p["c1"], r["a0"], c0, r["a1"], c1, d_raw1, h, w)
                if p["res"] is not None:
                    conv_wgrad(KIND_1x1, p["res"], r["a0"], c0, r["a1"], c1, g, h, w)
            outs = []
            for lo, hi in ((0, c0), (c0, c0 + c1)):
                if hi == lo:
                    outs.append(None)
                    continue
                if p["res"] is not None:
                    side = dgrad(KIND_1x1, g, cout, p["res"], (lo, hi), None, hi - lo, False, h, w)
                else:
                    side = g   # identity residual (single input, c0 == cout)
                outs.append(dgrad(KIND_3x3, d_raw1, cout, p["c1"], (lo, hi), side, hi - lo, True, h, w))
            return outs[0], outs[1]

        def attn_bwd(r, g):
            p, c, h, w = r["p"], r["c"], r["h"], r["w"]
            M, n, hid = B * h * w, h * w, HEADS * DIM_HEAD
            xin = r["xin"]
            d_qkv = f32(M, 3 * hid)
            inner = p["mod"].fn.fn if want_p else None
            if p["full"]:
                d_att = dgrad(KIND_1x1, g, c, p["out"], (0, hid), None, hid, False, h, w)
                L.check(lib.sdc_attention_bwd(L.ptr(r["qkv"]), L.ptr(d_att), L.ptr(d_qkv), B, n, _st()))
                if want_p:
                    conv_wgrad(KIND_1x1, p["out"], r["att"], hid, None, 0, g, h, w)
            else:
                d_proj = f32(M, c)
                L.check(lib.sdc_channel_layernorm_bwd(L.ptr(g), L.ptr(r["proj"]), 0, L.ptr(p["g_out"]), None, L.ptr(d_proj), M, c, 1, _st()))
                d_att = dgrad(KIND_1x1, d_proj, c, p["out"], (0, hid), None, hid, False, h, w)
                wsb = torch.empty(lib.sdc_linear_attention_bwd_workspace(B), device=dev, dtype=torch.uint8)
                L.check(lib.sdc_linear_attention_bwd(L.ptr(r["qkv"]), L.ptr(d_att), L.ptr(r["ws"]), L.ptr(wsb), L.ptr(d_qkv), B, n, _st()))
                if want_p:
                    ln_pgrad(inner.to_out[1], g, r["proj"], M, c)
                    conv_wgrad(KIND_1x1, p["out"], r["att"], hid, None, 0, d_proj, h, w)
            if want_p:
                conv_wgrad(KIND_1x1, p["qkv"], r["xn"], c, None, 0, d_qkv, h, w)
            d_xn = dgrad(KIND_1x1, d_qkv, 3 * hid, p["qkv"], (0, c), None, c, False, h, w)
            if want_p:
                ln_pgrad(p["mod"].fn.norm, d_xn, xin, M, c)
            d_x = f32(M, c)
            L.check(lib.sdc_channel_layernorm_bwd(L.ptr(d_xn), L.ptr(xin), int(xin.dtype == torch.float16), L.ptr(p["g_in"]), L.ptr(g),
                                                  L.ptr(d_x), M, c, 1, _st()))
            return d_x

        def add_(a, b):
            L.check(lib.sdc_add_inplace(L.ptr(a), L.ptr(b), a.numel(), 1, _st()))
            return a

        # ---- reverse walk ----
        i = len(tape) - 1
        kind, r = tape[i]
        assert kind == "head"
        Hh, Ww = st0["H"], st0["W"]
        cfin = self.final_res_block.dim_out
        if want_p:
            dwh = torch.zeros(self.out_dim, cfin, device=dev, dtype=torch.float32)
            dbh = torch.zeros(self.out_dim, device=dev, dtype=torch.float32)
            L.check(lib.sdc_head_conv1_wgrad(L.ptr(g_eps), L.ptr(r["inp"]), int(r["inp"].dtype == torch.float16), L.ptr(dwh), L.ptr(dbh),
                                             B, Hh * Ww, cfin, self.out_dim, _st()))
            acc_grad(self.final_conv.weight, dwh.reshape(self.final_conv.weight.shape))
            acc_grad(self.final_conv.bias, dbh)
        i -= 1
        kind, r = tape[i]
        assert kind == "resnet"   # final_res_block
        g = f32(B * Hh * Ww, cfin)
        L.check(lib.sdc_head_conv1_bwd(L.ptr(g_eps), L.ptr(pk["head"][0]), L.ptr(g), B, Hh * Ww, cfin, self.out_dim, 1, _st()))
        g, g_r = resnet_bwd(r, g)
        i -= 1
        skip_grads = []   # gradients of the skip tensors: filled by the up path in push order, consumed by the down path from the end
        while i > 0:
            kind, r = tape[i]
            if kind == "up":
                c, h, w = r["c"], r["h"], r["w"]
                if want_p:
                    conv_wgrad(KIND_3x3, r["p"], r["inp"], c, None, 0, g, h, w)
                if r["upsample"]:
                    g_hi = dgrad(KIND_3x3, g, r["p"]["cout"], r["p"], (0, c), None, c, False, h, w)
                    g = f32(B * (h // 2) * (w // 2), c)
                    L.check(lib.sdc_upsample2x_bwd(L.ptr(g_hi), L.ptr(g), B, h // 2, w // 2, c, 1, _st()))
                else:
                    g = dgrad(KIND_3x3, g, r["p"]["cout"], r["p"], (0, c), None, c, True, h, w)
            elif kind == "attn":
                g = attn_bwd(r, g)
            elif kind == "resnet":
                g, g_skip = resnet_bwd(r, g)
                if g_skip is not None:
                    skip_grads.append(g_skip)
            elif kind == "down":
                # the tensor entering the downsample was also pushed as a skip: its up-path gradient is the last entry
                c, h, w = r["c"], r["h"], r["w"]
                sg = skip_grads.pop()
                if want_p:
                    conv_wgrad(KIND_UNSHUFFLE if r["unshuffle"] else KIND_3x3, r["p"], r["inp"], c, None, 0, g, h, w)
                if r["unshuffle"]:
                    t = dgrad(KIND_1x1, g, r["p"]["cout"], r["p"], (0, 4 * c), None, 4 * c, False, h, w)
                    g = f32(B * 4 * h * w, c)
                    L.check(lib.sdc_pixel_shuffle_bwd(L.ptr(t), L.ptr(sg), L.ptr(g), B, h, w, c, 1, _st()))
                else:
                    g = dgrad(KIND_3x3, g, r["p"]["cout"], r["p"], (0, c), sg, c, True, h, w)
            elif kind == "push":
                # a "push" directly after an attention block is consumed by the following "down" record; a push after the
                # first ResnetBlock of a level adds its skip gradient here
                if tape[i + 1][0] != "down":
                    g = add_(g, skip_grads.pop())
            i -= 1
        assert not skip_grads
        g = add_(g, g_r)   # the stem output also feeds final_res_block (r = x.clone(), unet.py:393)
        c = st0["c"]
        if want_p:
            # stem 7x7: weight gradient of the GEMM over the (high | low) im2col operand, the two halves folded back
            prec, kps = pk["prec"], pk["stem"]["kp"]
            patches = torch.empty(B * Hh * Ww, kps, device=dev, dtype=operand_dtype(prec))
            L.check(lib.sdc_stem_im2col(prec, L.ptr(st0["x"]), L.ptr(patches), B, st0["Cin"], Hh, Ww, kps, _st()))
            dwp = torch.zeros(c, kps, device=dev, dtype=torch.float32)
            L.check(lib.sdc_conv_wgrad(KIND_1x1, int(patches.dtype == torch.float16), L.ptr(patches), kps, None, 0, L.ptr(g), L.ptr(dwp),
                                       B, Hh, Ww, c, _st()))
            k = st0["Cin"] * 49
            acc_grad(self.init_conv.weight, (dwp[:, :k] + dwp[:, kps // 2:kps // 2 + k]).reshape(self.init_conv.weight.shape))
            dbs = torch.zeros(c, device=dev, dtype=torch.float32)
            L.check(lib.sdc_colsum(L.ptr(g), L.ptr(dbs), B * Hh * Ww, c, _st()))
            acc_grad(self.init_conv.bias, dbs)
            del patches
        if not need_gx:
            return None
        wt, kp = dg["stem"]
        t = f32(B * Hh * Ww, kp)
        conv_gemm(KIND_1x1, g, c, None, 0, wt, None, None, t, None, False, B, Hh, Ww, kp, PREC_TF32)
        gx = torch.empty(B, st0["Cin"], Hh, Ww, device=dev, dtype=torch.float32)
        L.check(lib.sdc_stem_col2im(L.ptr(t), L.ptr(gx), B, st0["Cin"], Hh, Ww, kp, _st()))
        return gx
