"""Import alias: the product package lives in the directory ``ho-nerf_b200/`` (not an importable
name), so ``honerf_b200`` extends its ``__path__`` to that directory and re-exports its API."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ho-nerf_b200")
__path__.append(_impl)

from honerf_b200.api import *  # noqa: E402,F401,F403
from honerf_b200.api import __all__  # noqa: E402,F401
