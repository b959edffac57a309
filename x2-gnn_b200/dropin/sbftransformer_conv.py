"""Drop-in for the reference's `sbftransformer_conv.py`: put this directory on sys.path ahead of the
reference tree (`x2gnn_b200.install()`) and the unchanged callers (model.py, xgnn.py, ...) pick up
the sm_100a implementation."""
from x2gnn_b200.sbftransformer_conv import *  # noqa: F401,F403
from x2gnn_b200 import sbftransformer_conv as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
