"""Drop-in for the reference's `radial_basis_layer.py`: put this directory on sys.path ahead of the
reference tree (`x2gnn_b200.install()`) and the unchanged callers (model.py, xgnn.py, ...) pick up
the sm_100a implementation."""
from x2gnn_b200.radial_basis_layer import *  # noqa: F401,F403
from x2gnn_b200 import radial_basis_layer as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
