"""CPU: the public entry points are registered as torch custom operators (torch.ops.safediffcon_b200.*), have shape-only fake
implementations, and have NO CPU kernel -- a CPU call fails in the dispatcher instead of computing something else."""
import pytest
import torch


def test_ops_are_registered_cuda_only():
    import safediffcon_b200  # noqa: F401
    ns = torch.ops.safediffcon_b200
    for name in ("burgers_solve_free", "burgers_solve_cartesian", "burgers_control_score", "kth_select", "unet_forward"):
        assert hasattr(ns, name), name
    with pytest.raises(NotImplementedError, match="CPU"):
        ns.burgers_solve_free(torch.zeros(2, 128), torch.zeros(2, 10, 128), 0.01, 1.0, 1e-4, True)
    with pytest.raises(NotImplementedError, match="CPU"):
        ns.kth_select(torch.zeros(8), 3)


def test_fake_implementations_give_the_reference_shapes():
    import safediffcon_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    ns = torch.ops.safediffcon_b200
    with FakeTensorMode():
        u0, f = torch.zeros(5, 128, device="cuda"), torch.zeros(5, 10, 128, device="cuda")
        assert ns.burgers_solve_free(u0, f, 0.01, 1.0, 1e-4, True).shape == (5, 11, 128)
        assert ns.burgers_solve_cartesian(u0, f[:3], 0.01, 1.0, 1e-4, True).shape == (5, 3, 11, 128)
        traj, J, pts, tms, flg = ns.burgers_control_score(torch.zeros(5, 3, 16, 128, device="cuda"), u0, 0.8, 11)
        assert traj.shape == (5, 11, 128) and J.shape == (5,) and pts.dtype == torch.int32 and flg.shape == (5,)
        v, i = ns.kth_select(torch.zeros(100, device="cuda"), 97)
        assert v.shape == () and i.dtype == torch.int64
