"""One-dimensional tensor networks (reference: onedim/__init__.py)."""
from .onedim_core import *   # noqa: F401,F403
from .onedim_utils import *  # noqa: F401,F403
