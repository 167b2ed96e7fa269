"""Public API of honerf_b200: the reference's class names for the NeuS hot path."""
from . import ops  # noqa: F401
from ._lib import HonerfError, launch_count  # noqa: F401
from .fields import (Embedding, RenderingNetwork, RenderingNetwork_OBJ, SDFNetwork,  # noqa: F401
                     SDFNetwork_OBJ, SingleVarianceNetwork)
from .ops import set_default_precision  # noqa: F401
from . import losses, rays, renderer, renderer_batch  # noqa: F401
from .renderer import NeuSRenderer, NeuSRenderer_fitting  # noqa: F401

__all__ = ["Embedding", "SDFNetwork", "RenderingNetwork", "SDFNetwork_OBJ", "RenderingNetwork_OBJ", "SingleVarianceNetwork",
           "NeuSRenderer", "NeuSRenderer_fitting", "renderer", "renderer_batch", "losses", "rays", "ops", "HonerfError", "launch_count", "set_default_precision"]
