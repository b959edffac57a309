"""TEST INFRASTRUCTURE ONLY (oracle/): minimal restatement of the torch-geometric
2.1.0 names the reference imports."""
from . import typing, utils, nn, data, loader  # noqa: F401
