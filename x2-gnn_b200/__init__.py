"""x2gnn_b200 -- B200-native (sm_100a) implementation of the X2-GNN message-passing hot path:
SBFTransformerConv, the radial / spherical-Bessel / spherical-harmonic bases and the radius-graph /
triplet index construction, behind the reference's own Python API.  See DESIGN.md.

    import x2gnn_b200; x2gnn_b200.install()     # reference-named modules now resolve here
    from sbftransformer_conv import SBFTransformerConv
"""
import os as _os
import sys as _sys

__version__ = "0.1.0"

DROPIN_DIR = _os.path.join(_os.path.abspath(list(__path__)[0]), "dropin")
DROPIN_MODULES = ("sbftransformer_conv", "radial_basis_layer", "angular_basis_layer", "basis_func",
                  "envelop", "edge_graph", "atom_graph")


def install():
    """Make `import sbftransformer_conv` (etc.) resolve to this package's drop-in modules."""
    for m in DROPIN_MODULES:
        _sys.modules.pop(m, None)
    if DROPIN_DIR in _sys.path:
        _sys.path.remove(DROPIN_DIR)
    _sys.path.insert(0, DROPIN_DIR)
