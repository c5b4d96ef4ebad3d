"""B200 mirror of ``pysteps.motion`` for the methods on the advection hot path."""
from . import lucaskanade  # noqa: F401
from .interface import get_method  # noqa: F401
