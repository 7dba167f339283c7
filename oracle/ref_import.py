"""Import harness for the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Only the golden-generation scripts (oracle/make_golden.py) and local validation use this; it needs
/root/reference, which does not exist on the GPU box.  Recipe from SURVEY.md section 8(c): put
/root/reference/1D on sys.path and register empty stub modules for the packages the reference imports
at module top level but the hot path never touches.
"""
import sys
import types

REF_ROOT = "/root/reference/1D"


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Make `import model.unet`, `data.generate_burgers`, `utils.metrics`, ... resolve to the reference."""
    import os
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present (expected on the build container only)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class _Any:  # permissive placeholder class
        def __init__(self, *a, **k):
            pass

    for mod in ("h5py", "tensorboardX", "ema_pytorch", "accelerate", "accelerate.state", "matplotlib",
                "matplotlib.pyplot", "IPython"):
        try:
            __import__(mod)
        except Exception:
            _stub(mod)
    sys.modules["IPython"].embed = getattr(sys.modules["IPython"], "embed", lambda *a, **k: None)
    for mod, name in (("tensorboardX", "SummaryWriter"), ("ema_pytorch", "EMA"), ("accelerate", "Accelerator"),
                      ("accelerate.state", "AcceleratorState"), ("h5py", "File")):
        if not hasattr(sys.modules[mod], name):
            setattr(sys.modules[mod], name, _Any)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
