"""dinov2_distillation_b200 -- B200-native (sm_100a) distillation hot path: frozen DINOv2 ViT teacher forward and the
ScaleKD loss forward/backward, behind the reference's Python plug-in surface (DINOv2ViT, ScaleKD, LOSS_REGISTRY)."""

__version__ = "0.1.0"

from . import _lib  # noqa: F401


def __getattr__(name):
    # lazy: importing torch-dependent shells only when asked for
    if name in ("DINOv2ViT", "TEACHER_CONFIGS"):
        from . import teacher
        return getattr(teacher, name)
    if name in ("ScaleKD", "AttentionProjector", "LOSS_REGISTRY"):
        from . import scalekd
        return getattr(scalekd, name)
    if name == "install":
        from .plugin import install
        return install
    raise AttributeError(name)
