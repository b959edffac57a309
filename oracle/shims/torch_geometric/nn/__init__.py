from .norm import LayerNorm  # noqa: F401
from . import conv, dense  # noqa: F401
