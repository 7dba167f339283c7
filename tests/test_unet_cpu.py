"""CPU: the drop-in Unet2D owns the reference's parameters (keys, shapes, seed-42 initial values) and the oracle
U-Net restatement reproduces the reference's eps."""
import copy

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import unet_ref


@pytest.mark.parametrize("dim,B", [(32, 3), (128, 2)])
def test_state_dict_and_oracle_match_reference(dim, B, golden):
    import safediffcon_b200 as s
    torch.manual_seed(42)
    net = s.Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
    g = golden(f"unet_dim{dim}_wsum")
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert np.array_equal(np.array([float(v.double().sum()) for v in sd.values()]), g["sum"])
    assert np.array_equal(np.array([float(v.double().abs().sum()) for v in sd.values()]), g["abssum"])
    x, t = fx.unet_inputs(B)
    with torch.no_grad():
        eps = unet_ref.unet_forward(sd, x, t)
    assert np.array_equal(eps.numpy(), golden(f"unet_dim{dim}")["eps"])


def test_module_semantics():
    import safediffcon_b200 as s
    net = s.Unet2D(dim=32, channels=3, resnet_block_groups=1)
    assert net.channels == 3 and net.self_condition is False and net.out_dim == 3
    n2 = copy.deepcopy(net)
    assert n2._cache.pack is None
    n2.load_state_dict(net.state_dict())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 16, 128), torch.zeros(1, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        s.Unet2D(dim=32, channels=3, resnet_block_groups=8)
    gd = s.GaussianDiffusion(net, seq_length=(16, 128), temporal=True, use_conv2d=True)
    assert len(gd.state_dict()) == 276 + 13


def test_oracle_input_gradient_matches_reference_golden(golden):
    """The oracle's autograd VJP w.r.t. x (used by the GPU backward-data tests) equals the unmodified reference's."""
    import safediffcon_b200 as s
    from oracle import fixtures as fx, unet_ref
    for dim in (32, 64):
        torch.manual_seed(42)
        net = s.Unet2D(dim=dim, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
        sd = {k: v.detach() for k, v in net.state_dict().items()}
        x, t = fx.unet_inputs(2)
        x = x.clone().requires_grad_()
        eps = unet_ref.unet_forward(sd, x, t)
        (gx,) = torch.autograd.grad(eps, x, fx.unet_cotangent(2))
        g = golden(f"unet_dim{dim}_vjp")
        assert np.allclose(eps.detach().numpy(), g["eps"], rtol=0, atol=1e-5 * np.abs(g["eps"]).max())
        assert np.allclose(gx.numpy(), g["grad_x"], rtol=0, atol=1e-5 * np.abs(g["grad_x"]).max())


def test_oracle_parameter_gradients_match_reference_golden(golden):
    """Oracle restatement under autograd reproduces the reference's parameter-gradient digests: the
    dense-cotangent set and the drop-in module's key layout (all 276 parameters)."""
    import safediffcon_b200 as s
    from oracle import fixtures as fx, unet_ref
    torch.manual_seed(42)
    net = s.Unet2D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, resnet_block_groups=1)
    sd = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items()}
    x, t = fx.unet_inputs(2)
    eps = unet_ref.unet_forward(sd, x, t)
    grads = torch.autograd.grad(eps, list(sd.values()), fx.unet_cotangent(2))
    got = fx.grad_digest(zip(sd.keys(), grads))
    g = golden("unet_dim64_pgrad")
    assert {k[:-5] for k in g.files if k.endswith("|norm")} == set(sd)
    for k in sd:
        assert abs(got[k + "|norm"] / float(g[k + "|norm"]) - 1) < 1e-4, k
        scale = float(g[k + "|norm"]) / np.sqrt(sd[k].numel())
        assert np.abs(got[k + "|samples"] - g[k + "|samples"]).max() < 2e-3 * scale + 1e-12, k
