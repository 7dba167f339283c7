"""CPU: the whole-network executor handle (include/safediffcon_b200_plan.h) describes the reference's parameter layout and
sizes its workspace without touching CUDA."""
import pytest
import torch


@pytest.mark.parametrize("dim,mults", [(64, (1, 2, 4, 8)), (128, (1, 2, 4, 8)), (64, (1, 2))])
def test_parameter_list_is_the_state_dict_layout(dim, mults):
    import safediffcon_b200 as s
    from safediffcon_b200 import unet as U
    net = s.Unet2D(dim=dim, dim_mults=mults, channels=3, resnet_block_groups=1)
    plan = U.UnetPlan(dim, mults, 3, 3, U.PREC_F16, 10000.0, 1000)
    named = list(net.named_parameters())
    assert plan.names == [n for n, _ in named]                 # the reference's state_dict keys, in module order
    assert plan.numels == [p.numel() for _, p in named]
    assert set(plan.names) <= set(net.state_dict().keys())


def test_workspace_grows_with_batch_and_needs_no_gpu():
    from safediffcon_b200 import _lib as L, unet as U
    plan = U.UnetPlan(128, (1, 2, 4, 8), 3, 3, U.PREC_F16, 10000.0, 1000)
    lib = L.lib()
    w8, w64 = (int(lib.sdc_unet_workspace_bytes(plan.handle, b, 16, 128)) for b in (8, 64))
    assert 0 < w8 < w64 and abs(w64 / w8 - 8.0) < 0.5
    assert int(lib.sdc_unet_workspace_bytes(plan.handle, 1024, 16, 128)) < 6 * 2 ** 30   # 180 GB HBM: 1024 samples per GPU fit easily
    tf = U.UnetPlan(128, (1, 2, 4, 8), 3, 3, U.PREC_TF32, 10000.0, 1000)
    assert int(lib.sdc_unet_workspace_bytes(tf.handle, 8, 16, 128)) > w8              # fp32 intermediates


def test_create_rejects_unsupported_architectures():
    from safediffcon_b200 import unet as U
    with pytest.raises(ValueError, match="multiple of 64"):
        U.UnetPlan(96, (1, 2), 3, 3, U.PREC_F16, 10000.0, 1000)
    with pytest.raises(ValueError):
        U.UnetPlan(64, (1, 2), 3, 3, 7, 10000.0, 1000)


def test_forward_before_pack_is_a_state_error():
    from safediffcon_b200 import _lib as L, unet as U
    plan = U.UnetPlan(64, (1, 2), 3, 3, U.PREC_F16, 10000.0, 1000)
    x = torch.zeros(16)
    rc = L.lib().sdc_unet_forward(plan.handle, L.ptr(x), None, 0, L.ptr(x), 1, 16, 128, L.ptr(x), 0, None, None)
    assert rc == 3 and b"pack_weights" in L.lib().sdc_last_error()
