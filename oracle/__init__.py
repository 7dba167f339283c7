"""CPU oracle for the SafeDiffCon 1D Burgers hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm for the path named by
BASELINE.json's north_star.  It exists so that the CUDA path can be checked on a
GPU box where ``/root/reference`` is absent.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``safediffcon_b200`` never does and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships NO tests, golden vectors or fixtures, so the
oracle is pinned against *outputs of the unmodified reference executed in the build container*
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; the script and its outputs are committed).

Modules
  diffusion_ref   schedule buffers, x0/eps algebra, guidance closed form, DDIM + DDPM chains
                  (reference: 1D/model/diffusion.py, 1D/model/model_utils.py, 1D/utils/guidance.py)
  solver_ref      Burgers explicit solver, torch port + ctypes binding of burgers_ref.c, metrics
                  (reference: 1D/data/generate_burgers.py:113-299, 1D/utils/metrics.py)
  unet_ref        functional fp32 restatement of Unet2D.forward from a state_dict
                  (reference: 1D/model/unet.py)
  conformal_ref   importance weights, nonconformity scores, quantile rank/selection
                  (reference: 1D/inference/conformal.py, 1D/inference/guidance.py)
  ref_import      import harness for the reference itself (build container only)
"""
