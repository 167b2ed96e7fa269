"""Public API of honerf_b200: the reference's class names for the NeuS hot path."""
from . import ops  # noqa: F401
from ._lib import HonerfError, launch_count  # noqa: F401
from .fields import (Embedding, RenderingNetwork_OBJ, SDFNetwork_OBJ,  # noqa: F401
                     SingleVarianceNetwork)
from .ops import set_default_precision  # noqa: F401
from .renderer import NeuSRenderer  # noqa: F401

__all__ = ["Embedding", "SDFNetwork_OBJ", "RenderingNetwork_OBJ", "SingleVarianceNetwork",
           "NeuSRenderer", "ops", "HonerfError", "launch_count", "set_default_precision"]
