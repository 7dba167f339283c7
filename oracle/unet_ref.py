"""Oracle (TEST INFRASTRUCTURE): functional fp32 restatement of the reference denoiser.

Evaluates Unet2D.forward (/root/reference/1D/model/unet.py:382-426) from a plain ``state_dict`` with
torch.nn.functional ops: ResnetBlock (:128-180), LinearAttention (:182-222), Attention (:224-258), channel
LayerNorm (:53-63), Downsample2d / Upsample2d (:33-43), sinusoidal time MLP (:81-107, :310-315).
Floating point => this IS the "plain torch fp32 reference" for the CUDA U-Net; pinned against the
unmodified reference module by tests/golden/unet_*.npz.
"""
import math
import torch
import torch.nn.functional as F

HEADS = 4
DIM_HEAD = 32


def _conv(sd, key, x, pad=0):
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), padding=pad)


def _chan_ln(x, g):
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) * (var + 1e-5).rsqrt() * g


def _block(sd, p, x, groups, scale_shift=None):
    x = _conv(sd, p + ".proj", x, 1)
    x = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    if scale_shift is not None:
        sc, sh = scale_shift
        x = x * (sc + 1) + sh
    return F.silu(x)


def _resnet(sd, p, x, temb, groups):
    te = F.linear(F.silu(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
    sc, sh = te[:, :, None, None].chunk(2, dim=1)
    h = _block(sd, p + ".block1", x, groups, (sc, sh))
    h = _block(sd, p + ".block2", h, groups)
    res = _conv(sd, p + ".res_conv", x) if (p + ".res_conv.weight") in sd else x
    return h + res


def _split_heads(t):
    b, c, h, w = t.shape
    return t.reshape(b, HEADS, c // HEADS, h * w)


def _linear_attention(sd, p, x):
    b, c, h, w = x.shape
    xn = _chan_ln(x, sd[p + ".fn.norm.g"])
    q, k, v = (_split_heads(t) for t in _conv(sd, p + ".fn.fn.to_qkv", xn).chunk(3, dim=1))
    q = q.softmax(dim=-2) * DIM_HEAD ** -0.5
    k = k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(b, HEADS * DIM_HEAD, h, w)
    out = _conv(sd, p + ".fn.fn.to_out.0", out)
    return _chan_ln(out, sd[p + ".fn.fn.to_out.1.g"]) + x


def _attention(sd, p, x):
    b, c, h, w = x.shape
    xn = _chan_ln(x, sd[p + ".fn.norm.g"])
    q, k, v = (_split_heads(t) for t in _conv(sd, p + ".fn.fn.to_qkv", xn).chunk(3, dim=1))
    sim = torch.einsum("bhdi,bhdj->bhij", q * DIM_HEAD ** -0.5, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)  # [b, heads, n, d]
    out = out.permute(0, 1, 3, 2).reshape(b, HEADS * DIM_HEAD, h, w)
    return _conv(sd, p + ".fn.fn.to_out", out) + x


def _time_embedding(sd, t, dim):
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * freq[None, :]
    emb = torch.cat((arg.sin(), arg.cos()), dim=-1)
    emb = F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    return F.linear(F.gelu(emb), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def _pixel_unshuffle(x):
    b, c, h, w = x.shape
    x = x.reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
    return x.reshape(b, c * 4, h // 2, w // 2)


def unet_forward(sd, x, t, groups=1, taps=None):
    """sd: Unet2D state_dict (keys without the 'model.' prefix).  taps: optional dict filled with intermediates."""
    dim = sd["init_conv.weight"].shape[0]
    n_levels = len({k.split(".")[1] for k in sd if k.startswith("downs.")})
    x = _conv(sd, "init_conv", x, 3)
    r = x
    temb = _time_embedding(sd, t, dim)
    if taps is not None:
        taps["init"] = x
        taps["temb"] = temb
    skips = []
    for i in range(n_levels):
        p = f"downs.{i}"
        x = _resnet(sd, p + ".0", x, temb, groups)
        skips.append(x)
        x = _resnet(sd, p + ".1", x, temb, groups)
        x = _linear_attention(sd, p + ".2", x)
        skips.append(x)
        if (p + ".3.1.weight") in sd:
            x = _conv(sd, p + ".3.1", _pixel_unshuffle(x))
        else:
            x = _conv(sd, p + ".3", x, 1)
        if taps is not None:
            taps[f"down{i}"] = x
    x = _resnet(sd, "mid_block1", x, temb, groups)
    x = _attention(sd, "mid_attn", x)
    x = _resnet(sd, "mid_block2", x, temb, groups)
    if taps is not None:
        taps["mid"] = x
    for i in range(n_levels):
        p = f"ups.{i}"
        x = _resnet(sd, p + ".0", torch.cat((x, skips.pop()), dim=1), temb, groups)
        x = _resnet(sd, p + ".1", torch.cat((x, skips.pop()), dim=1), temb, groups)
        x = _linear_attention(sd, p + ".2", x)
        if (p + ".3.1.weight") in sd:
            x = _conv(sd, p + ".3.1", F.interpolate(x, scale_factor=2, mode="nearest"), 1)
        else:
            x = _conv(sd, p + ".3", x, 1)
        if taps is not None:
            taps[f"up{i}"] = x
    x = _resnet(sd, "final_res_block", torch.cat((x, r), dim=1), temb, groups)
    return _conv(sd, "final_conv", x)
