"""torch_geometric 2.1.0 names imported by the unchanged X2-GNN callers (see ../README.md)."""
from . import data, loader, nn, typing, utils  # noqa: F401

__version__ = "2.1.0+x2gnn_b200.compat"
