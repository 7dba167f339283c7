"""CPU: the C-ABI library loads and exports every symbol the headers declare; host logic fails loudly off-GPU."""
import ctypes
import glob
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(sdc_[a-z0-9_]+)\s*\(", src))
    return names


def test_exports_match_headers():
    from safediffcon_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 16
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the python binding table covers the same set
    missing = decl - set(_lib._SIGS)
    assert not missing, missing


def test_struct_layouts():
    from safediffcon_b200 import _lib
    assert ctypes.sizeof(_lib.StepCoef) == 32
    assert ctypes.sizeof(_lib.Guidance) == 24
    assert ctypes.sizeof(_lib.ChainState) == 24


def test_version_and_error_string():
    from safediffcon_b200 import _lib
    assert _lib.lib().sdc_version() >= 100
    assert isinstance(_lib.lib().sdc_last_error(), bytes)


def test_no_cpu_fallback():
    import safediffcon_b200 as s
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.burgers_numeric_solve_free(torch.zeros(2, 128), torch.zeros(2, 10, 128), 0.01, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.get_weight(torch.zeros(2, 3, 16, 128), 0.0, type("C", (), dict(use_max_safety=True, u_bound=0.8,
                                                                          guidance_weights={"w_score": 1.0}))())


def test_solver_argument_errors_match_reference():
    import safediffcon_b200 as s
    with pytest.raises(AssertionError, match="check number of time interval"):
        s.burgers_numeric_solve_free(torch.zeros(2, 128), torch.zeros(2, 9, 128), 0.01, 1.0)
    with pytest.raises(ValueError):
        s.burgers_numeric_solve_free(torch.zeros(2, 128), torch.zeros(2, 128), 0.01, 1.0, mode='const')


def test_product_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "safediffcon_b200", "*.py")):
        src = open(path).read()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), path
