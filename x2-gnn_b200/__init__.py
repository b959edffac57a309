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


COMPAT_DIR = _os.path.join(_os.path.abspath(list(__path__)[0]), "compat")
COMPAT_PACKAGES = ("torch_geometric", "torch_scatter", "torch_sparse")


def install(compat=True):
    """Make `import sbftransformer_conv` (etc.) resolve to this package's drop-in modules, so that the
    reference's model.py / xgnn.py / trainer.py run unchanged on top of them.  With `compat` (default), the
    third-party names those callers import (torch_geometric, torch_scatter, torch_sparse) resolve to the
    plain-PyTorch stand-ins under compat/ -- but only if the real packages are not installed.
    Returns the list of compat packages that were put in place."""
    import importlib.util
    for m in DROPIN_MODULES:
        _sys.modules.pop(m, None)
    if DROPIN_DIR in _sys.path:
        _sys.path.remove(DROPIN_DIR)
    _sys.path.insert(0, DROPIN_DIR)
    used = []
    if compat:
        missing = [m for m in COMPAT_PACKAGES if m not in _sys.modules and importlib.util.find_spec(m) is None]
        if missing and COMPAT_DIR not in _sys.path:
            _sys.path.append(COMPAT_DIR)          # last: an installed package always wins
        used = missing
    return used


def uninstall():
    """Undo install(): drop the path entries and the modules imported through them."""
    for d in (DROPIN_DIR, COMPAT_DIR):
        while d in _sys.path:
            _sys.path.remove(d)
    for name, mod in list(_sys.modules.items()):
        f = getattr(mod, "__file__", None) or ""
        if f.startswith(DROPIN_DIR) or f.startswith(COMPAT_DIR):
            _sys.modules.pop(name, None)
