"""Data / checkpoint formats either side of the hot path: drop-ins for ``BurgersDataset``
(/root/reference/1D/data/burgers.py:14-156), ``HDF5Dataset`` (/root/reference/1D/data/load_hdf5.py:6-60), ``build_model`` /
``load_model`` / ``get_target`` (/root/reference/1D/utils/common.py:77-160) and the checkpoint dictionary written by
``Trainer.save`` (/root/reference/1D/model/trainer.py:111-148).

B200-first difference: the reference assembles one [3, 16, 128] state per ``__getitem__`` on the host; here the whole split is
assembled ONCE on the device by ``sdc_dataset_states`` (one launch) and ``__getitem__`` is a view into that tensor, so a
``DataLoader`` over it (num_workers=0, pin_memory=False) yields device batches with no host work and no H2D copy.  For the
reference's multi-worker / pinned loaders pass ``host_resident=True`` (items are then CPU tensors, assembled by the same kernel).
"""
import os
from typing import Callable, List, Optional, Union

import torch
from torch.utils.data import Dataset

from .datagen import dataset_states, make_data_varying_f
from .diffusion import GaussianDiffusion
from .solver import burgers_numeric_solve_free
from .unet import Unet2D


def read_burgers_hdf5(path: str, mode: str, nt: int = 11, nx: int = 128):
    """(u [N, nt, nx], f [N, nt-1, nx]) CPU tensors from the reference's file layout: group `mode`, datasets
    ``pde_{nt}-{nx}`` and ``pde_{nt}-{nx}_f`` (load_hdf5.py:28-34; written by generate_burgers.py:539-559)."""
    try:
        import h5py
    except ImportError as e:   # h5py is an optional dependency of the data format, not of the compute path
        raise ImportError("safediffcon_b200.read_burgers_hdf5 needs h5py to read the reference's .h5 datasets; "
                          "use BurgersDataset.from_tensors / BurgersDataset.synthetic without it") from e
    with h5py.File(path, "r") as fh:
        grp = fh[mode]
        u = torch.from_numpy(grp[f"pde_{nt}-{nx}"][:])
        f = torch.from_numpy(grp[f"pde_{nt}-{nx}_f"][:])
    return u, f


class BurgersDataset(Dataset):
    """Same constructor, attributes (``nt_total, nx, pad_size, scaler, use_max_safety``) and item layout as the reference class;
    items are views into a device-resident [N, 3, pad, nx] tensor."""

    def __init__(self, dataset: str = "free_u_f_1e5", split: str = "train", root_path: str = None, nt_total: int = 11, nx: int = 128,
                 is_normalize: bool = True, stack_u_and_f: bool = True, pad_for_2d_conv: bool = True, pad_size: int = 16,
                 safety_transform: Optional[Callable] = None, is_need_idx: bool = False, is_subset: bool = False, config=None,
                 device="cuda", host_resident: bool = False, _tensors=None):
        self.root = root_path or "./datasets"
        self.split = split
        self.nt_total = nt_total
        self.nx = nx
        self.data_folder = dataset
        self.pad_size = pad_size
        self.use_max_safety = config.use_max_safety if config is not None else True
        self.scaler = 10.0 if is_normalize else None
        self.stack_u_and_f = stack_u_and_f
        self.pad_for_2d_conv = pad_for_2d_conv
        self.safety_transform = safety_transform   # None = the default u^2, fused in the assembly kernel
        self.is_need_idx = is_need_idx
        self.is_subset = is_subset
        self.finetune_subset_size = getattr(config, "finetune_subset_size", False) if type(config).__name__ == "PostTrainConfig" else False
        self.device = torch.device(device)
        if _tensors is None:
            path = os.path.join(self.root, self.data_folder, f"burgers_{self.split}.h5")
            if not os.path.exists(path):
                raise FileNotFoundError(f"Dataset not found at {path}")
            _tensors = read_burgers_hdf5(path, self.split, nt_total, nx)
        u, f = _tensors
        assert u.dim() == 3 and f.dim() == 3 and u.shape[0] == f.shape[0] and u.shape[2] == f.shape[2], "u [N, nt, nx], f [N, nt-1, nx]"
        self.u_data, self.f_data = u, f
        if self.split == "train" and self.is_subset:
            self.indices = list(range(min(self.finetune_subset_size, len(u))))
        else:
            self.indices = None
        self._states = self._assemble(u, f)
        # host_resident=True: the split is still assembled once by the CUDA kernel, then kept as a CPU tensor (shared memory),
        # so the reference's DataLoader(num_workers=4/16, pin_memory=True) call sites (inference_ft.py:129-130,148-149,
        # post_train.py:133-134) work unchanged: forked workers must not touch CUDA tensors and pinning needs host memory.
        self.host_resident = bool(host_resident)
        if self.host_resident:
            self._states = self._states.cpu().share_memory_()

    # ---- alternative constructors (no HDF5 needed)
    @classmethod
    def from_tensors(cls, u, f, **kw):
        """Dataset over rollouts u [N, nt, nx] and controls f [N, nt-1, nx] already in memory (host or device)."""
        return cls(_tensors=(u, f), **kw)

    @classmethod
    def synthetic(cls, n, seed=0, **kw):
        """n instances from the reference generator (np.random.seed(seed); make_data_varying_f(n, n, nx, nt-1)) rolled out by the
        CUDA solver -- generator, solver and assembly all on the device."""
        import numpy as np
        nt_total, nx = kw.get("nt_total", 11), kw.get("nx", 128)
        np.random.seed(seed)
        u0, f = make_data_varying_f(n, n, nx, nt_total - 1, device=kw.get("device", "cuda"))
        u = burgers_numeric_solve_free(u0.float(), f, visc=0.01, T=1.0, dt=1e-4, num_t=nt_total - 1)
        return cls(_tensors=(u, f), **kw)

    def _assemble(self, u, f):
        u = u.to(self.device, torch.float32).contiguous()
        f = f.to(self.device, torch.float32).contiguous()
        scaler = self.scaler if self.scaler is not None else 1.0   # x / 1.0 is exact
        if self.safety_transform is None:
            st = dataset_states(u, f, pad=max(self.pad_size, u.shape[1]), scaler=scaler, use_max_safety=self.use_max_safety)
        else:
            # user-supplied safety score: torch evaluates it, the layout rules stay those of _process_data
            s = self.safety_transform(u)
            if self.use_max_safety:
                s = s.amax(dim=(1, 2), keepdim=True).expand_as(s)
            st = torch.zeros(u.shape[0], 3, max(self.pad_size, u.shape[1]), u.shape[2], device=self.device)
            st[:, 0, :u.shape[1]], st[:, 1, :f.shape[1]], st[:, 2, :u.shape[1]] = u, f, s
            st = st / scaler
        if self.stack_u_and_f and self.pad_for_2d_conv:
            return st[:, :, :self.pad_size] if st.shape[2] != self.pad_size else st
        nt1, nt = u.shape[1], f.shape[1]
        return torch.cat((st[:, 0, :nt1], st[:, 1, :nt], st[:, 2, :nt1]), dim=1)   # burgers.py:131-134

    def __len__(self):
        return len(self.indices) if self.indices is not None else self._states.shape[0]

    def _process_data(self, data):
        """Single-item form of the assembly (kept for callers that use it directly): data = (u [nt, nx], f [nt-1, nx])."""
        return self._assemble(data[0][None], data[1][None])[0]

    def __getitem__(self, idx):
        if self._states.is_cuda and torch.utils.data.get_worker_info() is not None:
            raise RuntimeError("safediffcon_b200.BurgersDataset: items are views of a CUDA tensor, which DataLoader worker processes "
                               "cannot read; use num_workers=0 and pin_memory=False, or construct the dataset with host_resident=True")
        i = self.indices[idx] if self.indices is not None else idx
        data = self._states[i]
        return (data, idx) if self.is_need_idx else data

    @property
    def states(self):
        """The whole split as one device tensor [N, 3, pad, nx] (what a sharded calibration run slices per rank)."""
        return self._states if self.indices is None else self._states[:len(self.indices)]


def get_target(target_i: Union[int, List[int]], dataset="free_u_f_1e5", device=None, is_normalize=False, split="test"):
    """Target trajectories [n, nt_total, nx] (utils/common.py:77-108).  `dataset` may be a dataset name (HDF5 under ./datasets) or an
    already constructed BurgersDataset."""
    ds = dataset if isinstance(dataset, BurgersDataset) else BurgersDataset(split=split, root_path="datasets", dataset=dataset,
                                                                            is_normalize=is_normalize)
    if isinstance(target_i, int):
        target = ds[target_i].unsqueeze(0)
    else:
        target = torch.stack([ds[i] for i in target_i], dim=0)
    target = target[:, 0, :ds.nt_total, :]
    return target.to(device) if device is not None else target


def build_model(config, dataset) -> GaussianDiffusion:
    """utils/common.py:110-139: the denoiser + diffusion wrapper for `config` (Eval / Inference / PostTrain config objects)."""
    channels = dataset[0].shape[0]
    unet = Unet2D(dim=config.dim, dim_mults=config.dim_mults, channels=channels, resnet_block_groups=config.resnet_block_groups)
    return GaussianDiffusion(
        unet, seq_length=(dataset.pad_size, dataset.nx), use_conv2d=True, temporal=True,
        train_on_padded_locations=config.train_on_padded_locations, is_condition_u0=config.is_condition_u0,
        is_condition_uT=config.is_condition_uT, condition_idx=dataset.nt_total - 1,
        is_condition_u0_zero_pred_noise=config.is_condition_u0_zero_pred_noise,
        is_condition_uT_zero_pred_noise=config.is_condition_uT_zero_pred_noise,
        sampling_timesteps=config.ddim_sampling_steps if config.using_ddim else 1000, ddim_sampling_eta=config.ddim_eta,
    ).to(config.device)


def checkpoint_path(results_folder, milestone):
    return os.path.join(str(results_folder), f"model-{milestone}.pt" if isinstance(milestone, int) else str(milestone))


def ema_state_dict(ema_sd):
    """The averaged model's parameters out of an ``ema_pytorch.EMA`` state dict (keys ``ema_model.*``; trainer.py:124)."""
    out = {k[len("ema_model."):]: v for k, v in ema_sd.items() if k.startswith("ema_model.")}
    if not out:
        raise KeyError("no 'ema_model.*' entries in the EMA state dict")
    return out


def load_checkpoint(model: GaussianDiffusion, path: str, use_ema: bool = False, map_location=None):
    """Load the reference's checkpoint dictionary {'step','model','opt','ema','scaler','loss'} (trainer.py:115-125) into the drop-in
    module; returns the dictionary (optimizer / EMA states stay available to the caller)."""
    data = torch.load(path, map_location=map_location or next(model.parameters()).device, weights_only=False)
    if "model" not in data:
        raise KeyError(f"{path}: not a Trainer checkpoint (no 'model' entry; keys {sorted(data)})")
    model.load_state_dict(ema_state_dict(data["ema"]) if use_ema else data["model"])
    return data


def save_checkpoint(model: GaussianDiffusion, path: str, step: int = 0, opt=None, ema=None, loss=None):
    """Write a checkpoint the reference's ``Trainer.load`` accepts (same keys; 'ema' = an EMA wrapper's state dict if given)."""
    data = {"step": step, "model": model.state_dict(), "opt": opt.state_dict() if opt is not None else None,
            "ema": ema.state_dict() if ema is not None else None, "scaler": None, "loss": loss}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(data, path)
    return path


def load_model(config, dataset):
    """utils/common.py:141-160 without the Trainer detour: build, then load ``checkpoints_dir/exp_id/model-{checkpoint}.pt``.
    Returns (model, model_path) like the reference."""
    model = build_model(config, dataset)
    model_path = os.path.join(config.checkpoints_dir, config.exp_id)
    load_checkpoint(model, checkpoint_path(model_path, config.checkpoint))
    return model, model_path
