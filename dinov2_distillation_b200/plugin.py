"""Swap the B200 implementations into an importable copy of the reference (no reference file is modified).

    import dinov2_distillation_b200.plugin as plugin; plugin.install()
    # then run the reference's train.py / DistillationModule as usual

It rebinds exactly the two plug-in points the reference has:
  * LOSS_REGISTRY['scalekd']                      train/distillation_module.py:11-13
  * DINOv2ViT in models.backbones / models / train   models/backbones/dinov2.py:5, train.py:180
and removes the reference's NCCL_P2P_DISABLE=1 (train.py:23) so the gradient allreduce uses NVLink.
"""
from __future__ import annotations

import importlib
import os
import sys

from .scalekd import ScaleKD
from .teacher import DINOv2ViT


def install(verbose: bool = False) -> dict:
    touched = {}
    for modname, attr, obj in (
        ("models.backbones.dinov2", "DINOv2ViT", DINOv2ViT),
        ("models.backbones", "DINOv2ViT", DINOv2ViT),
        ("models", "DINOv2ViT", DINOv2ViT),
        ("losses.scalekd", "ScaleKD", ScaleKD),
        ("losses", "ScaleKD", ScaleKD),
        ("train.distillation_module", "ScaleKD", ScaleKD),
    ):
        mod = sys.modules.get(modname)
        if mod is None:
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
        if hasattr(mod, attr):
            setattr(mod, attr, obj)
            touched[f"{modname}.{attr}"] = obj
    dm = sys.modules.get("train.distillation_module")
    if dm is not None and hasattr(dm, "LOSS_REGISTRY"):
        dm.LOSS_REGISTRY["scalekd"] = ScaleKD
        touched["train.distillation_module.LOSS_REGISTRY['scalekd']"] = ScaleKD
    main = sys.modules.get("__main__")
    if main is not None and getattr(main, "DINOv2ViT", None) is not None:
        main.DINOv2ViT = DINOv2ViT
    os.environ.pop("NCCL_P2P_DISABLE", None)
    if verbose:
        for k in touched:
            print("b200 plugin: rebound", k)
    return touched
