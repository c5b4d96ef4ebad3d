"""B200 mirror of ``pysteps.extrapolation`` (interface + semi-Lagrangian scheme)."""
from . import semilagrangian  # noqa: F401
from .interface import get_method, eulerian_persistence  # noqa: F401
