"""GPU radius graph.  Drop-in for the reference's atom_graph.py:
`calculate_Dij(atom_pos)` (:32-35) and `gen_bonds_mini(Dij, cutoff)` (:42-45), plus a batched
`radius_graph(pos, batch, cutoff)` that goes straight from positions to `edge_index`.

Inputs may be CPU tensors (the reference builds graphs offline on CPU); they are uploaded to
the current CUDA device and results are returned as CUDA tensors.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib


def _to_cuda(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise _lib.X2Error("x2gnn_b200 needs a CUDA device (no CPU path)")
    return t.cuda()


def calculate_Dij(atom_pos: torch.Tensor) -> torch.Tensor:
    """[n,3] positions -> [n,n] fp32 distances, Gram form relu(sqrt(|a|^2+|b|^2-2ab))."""
    pos = _lib.f32(_to_cuda(atom_pos), "calculate_Dij")
    dev = _lib.require_cuda(pos, what="calculate_Dij")
    n = int(pos.size(0))
    out = torch.empty((n, n), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().x2_dij(_lib.ptr(pos), n, _lib.ptr(out), _lib.stream()), "x2_dij")
    return out


def gen_bonds_mini(Dij: torch.Tensor, cutoff: float = 5.0) -> torch.Tensor:
    """edge_index[2,E] int64 = argwhere((Dij < cutoff) & (Dij != 0)), sorted by (i, j)."""
    D = _lib.f32(_to_cuda(Dij), "gen_bonds_mini")
    dev = _lib.require_cuda(D, what="gen_bonds_mini")
    n = int(D.size(0))
    L = _lib.lib()
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ws = _lib.workspace((n + 1) * 4 + 512 + L.x2_scan_workspace_bytes(n), dev)
    _lib.check(L.x2_bonds_count(_lib.ptr(D), n, float(cutoff), _lib.ptr(rowptr), _lib.ptr(ws),
                                ws.numel(), _lib.stream()), "x2_bonds_count")
    E = int(rowptr[n].item())
    ei = torch.empty((2, E), dtype=torch.int64, device=dev)
    _lib.check(L.x2_bonds_fill(_lib.ptr(D), n, float(cutoff), _lib.ptr(rowptr), _lib.ptr(ei), E,
                               _lib.stream()), "x2_bonds_fill")
    return ei


def radius_graph(pos: torch.Tensor, batch: torch.Tensor | None = None, cutoff: float = 5.0):
    """Batched radius graph: all ordered pairs of atoms of the same graph with 0 < d < cutoff
    (same fp32 Gram arithmetic as calculate_Dij).  Returns (edge_index[2,E] int64 sorted by
    (i, j) with global atom ids, edge_num[B] int64 bonds per graph)."""
    pos = _lib.f32(_to_cuda(pos), "radius_graph")
    dev = _lib.require_cuda(pos, what="radius_graph")
    n = int(pos.size(0))
    if batch is None:
        batch = torch.zeros(n, dtype=torch.int64, device=dev)
    batch = _to_cuda(batch).long().contiguous()
    B = int(batch.max().item()) + 1 if n else 0
    counts = torch.bincount(batch, minlength=B)
    ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(counts, 0)
    L = _lib.lib()
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ws = _lib.workspace((n + 1) * 4 + 512 + L.x2_scan_workspace_bytes(n), dev)
    _lib.check(L.x2_radius_graph_count(_lib.ptr(pos), _lib.ptr(batch), _lib.ptr(ptr), n, float(cutoff),
                                       _lib.ptr(rowptr), _lib.ptr(ws), ws.numel(), _lib.stream()),
               "x2_radius_graph_count")
    E = int(rowptr[n].item())
    ei = torch.empty((2, E), dtype=torch.int64, device=dev)
    _lib.check(L.x2_radius_graph_fill(_lib.ptr(pos), _lib.ptr(batch), _lib.ptr(ptr), n, float(cutoff),
                                      _lib.ptr(rowptr), _lib.ptr(ei), E, _lib.stream()),
               "x2_radius_graph_fill")
    rp = rowptr.long()
    edge_num = rp[ptr[1:]] - rp[ptr[:-1]]
    return ei, edge_num
