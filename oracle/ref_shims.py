"""ORACLE helper (this container only): import the UNMODIFIED reference modules from /root/reference.

The reference needs third-party packages that are not installed here (lightning, termcolor, fvcore, timm, omegaconf);
the stubs below satisfy its imports without changing any of its code. Used by oracle/make_golden.py and by the
non-GPU tests to pin oracle/scalekd_ref.py. /root/reference does not exist on the GPU box: nothing under `-m gpu`,
smoke() or bench.py may call this.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "losses", "scalekd.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs() -> None:
    import torch.nn as nn

    try:
        import termcolor  # noqa: F401
    except ImportError:
        _stub("termcolor", colored=lambda s, *a, **k: s)
    try:
        import lightning  # noqa: F401
    except ImportError:
        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            def optimizers(self):
                return types.SimpleNamespace(param_groups=[{"lr": 0.0}])

        class LightningDataModule:
            pass

        _stub("lightning", LightningModule=LightningModule, LightningDataModule=LightningDataModule)


def import_reference():
    """Returns (scalekd_module, distillation_module_module) of the reference, imported unchanged."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    scalekd = importlib.import_module("losses.scalekd")
    try:
        dm = importlib.import_module("train.distillation_module")
    except Exception:  # utils.logger or other heavy imports may be unavailable
        dm = None
    return scalekd, dm
