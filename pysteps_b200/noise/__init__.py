"""B200 mirror of the velocity part of ``pysteps.noise`` (the BPS motion perturbator)."""
from . import motion  # noqa: F401
from .interface import get_method  # noqa: F401
