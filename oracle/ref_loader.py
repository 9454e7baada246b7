"""Import the reference's own modules: from /root/reference in the build container, else from the
unmodified copies oracle/vendor_ref.py placed under oracle/_ref/ (git-ignored, travels to the GPU box).

Used by oracle/make_golden.py (fixture generation), tests/test_oracle_vs_reference.py and the CPU arm of
bench.py (`--impl reference`, `cpu_baseline`).  Missing plotting / model-zoo dependencies of the reference
(matplotlib, seaborn, tensorboardX, pretrainedmodels, efficientnet_pytorch) are stubbed in
sys.modules; none of them is touched by the hot path.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root() -> str:
    env = os.environ.get("FEDMLP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "utils", "FedAvg.py")):
        return "/root/reference"
    return _VENDORED


REFERENCE_ROOT = _pick_root()
_STUBS = ["matplotlib", "matplotlib.pyplot", "seaborn", "tensorboardX", "pretrainedmodels",
          "efficientnet_pytorch"]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "local_training.py"))


def source() -> str:
    """Where the reference modules come from: 'mounted' (/root/reference) or 'vendored' (oracle/_ref)."""
    return "vendored" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_VENDORED) else "mounted"


def _install_stubs() -> None:
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        mod.__dict__["__stub__"] = True
        if name == "tensorboardX":
            mod.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None,
                                                           "add_scalar": lambda self, *a, **k: None})
        if name == "efficientnet_pytorch":
            mod.EfficientNet = type("EfficientNet", (), {})
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)


def load():
    """Returns a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.FedAvg = importlib.import_module("utils.FedAvg")
    ns.FedNoRo = importlib.import_module("utils.FedNoRo")
    ns.utils = importlib.import_module("utils.utils")
    ns.local_training = importlib.import_module("utils.local_training")
    return ns
