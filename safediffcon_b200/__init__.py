"""safediffcon_b200 -- B200-native (sm_100a) implementation of SafeDiffCon's 1D Burgers hot path.

Public surface mirrors the reference modules it replaces:
  solver    burgers_numeric_solve, burgers_numeric_solve_free      (1D/data/generate_burgers.py)
  metrics   control_trajectories, evaluate_samples, ...            (1D/utils/metrics.py)
  guidance  calculate_guidance, get_finetune_guidance, get_weight, normalize_weights
  conformal ConformalCalculator                                    (1D/inference/conformal.py)
  diffusion GaussianDiffusion                                      (1D/model/diffusion.py)
  unet      Unet2D                                                 (1D/model/unet.py)
  datagen   make_data_varying_f, dataset_states                    (1D/data/generate_burgers.py, 1D/data/burgers.py)
  common    BurgersDataset, build_model, load_model, get_target,   (1D/data/burgers.py, 1D/utils/common.py,
            load_checkpoint, save_checkpoint                        1D/model/trainer.py:111-148)
All compute runs in hand-written CUDA kernels behind the C ABI of include/safediffcon_b200.h; there is no CPU
fallback (calls raise when the library or a CUDA device is missing).
"""
from .solver import burgers_numeric_solve, burgers_numeric_solve_free, burgers_score, control_and_score  # noqa: F401
from .metrics import control_trajectories, evaluate_samples, calculate_safety_metrics, calculate_safety_score  # noqa: F401
from .guidance import (calculate_guidance, get_finetune_guidance, get_weight, normalize_weights, safety_guidance,  # noqa: F401
                       SafetyGuidance, SCALER)
from . import conformal, ops, runner  # noqa: F401  (ops: torch.ops.safediffcon_b200.* registrations)
from .conformal import ConformalCalculator, kth_select  # noqa: F401
from .diffusion import GaussianDiffusion, ModelPrediction  # noqa: F401
from .unet import Unet2D  # noqa: F401
from .datagen import make_data_varying_f, dataset_states  # noqa: F401
from .common import (BurgersDataset, build_model, load_model, get_target, load_checkpoint, save_checkpoint,  # noqa: F401
                     ema_state_dict, read_burgers_hdf5)

__version__ = "0.1.0"
