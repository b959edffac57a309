from .norm import LayerNorm  # noqa: F401
from . import dense  # noqa: F401
