"""CPU: checkpoint dictionary format of the reference Trainer (trainer.py:111-148) and build_model (utils/common.py:110-139)."""
import os
import types

import pytest
import torch


def _config(**kw):
    base = dict(dim=32, dim_mults=(1, 2), resnet_block_groups=1, train_on_padded_locations=True, is_condition_u0=True,
                is_condition_uT=True, is_condition_u0_zero_pred_noise=True, is_condition_uT_zero_pred_noise=True, using_ddim=True,
                ddim_sampling_steps=7, ddim_eta=1.0, device="cpu", checkpoints_dir=None, exp_id="exp", checkpoint=3)
    base.update(kw)
    return types.SimpleNamespace(**base)


class _DS:   # what build_model reads from a dataset
    pad_size, nx, nt_total = 16, 128, 11

    def __getitem__(self, i):
        return torch.zeros(3, 16, 128)


def test_build_model_matches_reference_wiring():
    import safediffcon_b200 as s
    gd = s.build_model(_config(), _DS())
    assert isinstance(gd, s.GaussianDiffusion) and isinstance(gd.model, s.Unet2D)
    assert gd.model.channels == 3 and gd.sampling_timesteps == 7 and gd.ddim_sampling_eta == 1.0 and gd.num_timesteps == 1000
    assert gd.condition_idx == 10 and tuple(gd.traj_size) == (16, 128) if hasattr(gd, "traj_size") else True
    gd2 = s.build_model(_config(using_ddim=False), _DS())
    assert gd2.sampling_timesteps == 1000


def test_checkpoint_round_trip_in_trainer_layout(tmp_path):
    import safediffcon_b200 as s
    cfg = _config(checkpoints_dir=str(tmp_path))
    torch.manual_seed(1)
    a = s.build_model(cfg, _DS())
    opt = torch.optim.Adam(a.parameters(), lr=1e-4)
    # an EMA wrapper's state dict carries the averaged copy under 'ema_model.*' (ema_pytorch layout)
    ema_sd = {"initted": torch.tensor(True), "step": torch.tensor(5)}
    ema_sd.update({"online_model." + k: v for k, v in a.state_dict().items()})
    ema_sd.update({"ema_model." + k: v * 0.5 if v.is_floating_point() else v for k, v in a.state_dict().items()})
    ema = types.SimpleNamespace(state_dict=lambda: ema_sd)
    path = s.save_checkpoint(a, os.path.join(str(tmp_path), "exp", "model-3.pt"), step=12, opt=opt, ema=ema, loss=0.25)
    data = torch.load(path, weights_only=False)
    assert set(data) == {"step", "model", "opt", "ema", "scaler", "loss"} and data["step"] == 12   # Trainer.save's keys
    torch.manual_seed(2)
    b, model_path = s.load_model(cfg, _DS())
    assert model_path == os.path.join(str(tmp_path), "exp")
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    c = s.build_model(cfg, _DS())
    s.load_checkpoint(c, path, use_ema=True)
    w = "model.init_conv.weight"
    assert torch.equal(c.state_dict()[w], a.state_dict()[w] * 0.5)
    with pytest.raises(KeyError):
        torch.save({"weights": {}}, os.path.join(str(tmp_path), "bad.pt"))
        s.load_checkpoint(c, os.path.join(str(tmp_path), "bad.pt"))


def test_dataset_needs_files_or_tensors(tmp_path):
    import safediffcon_b200 as s
    with pytest.raises(FileNotFoundError, match="Dataset not found"):
        s.BurgersDataset(root_path=str(tmp_path), split="test")
