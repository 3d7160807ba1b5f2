"""Two-dimensional tensor networks (reference: twodim/__init__.py)."""
from .square_lattice import *  # noqa: F401,F403
