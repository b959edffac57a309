"""GPU triplet (atom graph -> line graph) construction.  Drop-in for the reference's
edge_graph.py:12-30 `vertex_to_edge_2(edge_index, num_nodes)`, which runs scipy.sparse on the
CPU every forward (xgnn.py:52-53: D2H copy, CSR build, H2D copy).

The caller passes a CPU tensor (xgnn.py:52 `.to('cpu')`); it is uploaded, enumerated by integer
kernels and returned as CUDA int64 tensors, which index CUDA operands directly
(xgnn.py:53,58,61-62).  Ordering is bit-identical to the reference.
"""
from __future__ import annotations

import torch

from . import _lib
from .atom_graph import _to_cuda


def vertex_to_edge_2(edge_index: torch.Tensor, num_nodes: int):
    """-> (triplets_index[2,T] = [jk_idx; ij_idx], edge_j, edge_i, edge_k), all int64.
    NOTE the (j, i, k) return order of the reference (edge_graph.py:30)."""
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
    ei = _to_cuda(edge_index).long().contiguous()
    dev = _lib.require_cuda(ei, what="vertex_to_edge_2")
    E, N = int(ei.size(1)), int(num_nodes)
    L = _lib.lib()
    rowptr = torch.empty(E + 1, dtype=torch.int32, device=dev)
    flags = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = _lib.workspace(L.x2_triplets_workspace_bytes(E, N), dev)
    _lib.check(L.x2_triplets_count(_lib.ptr(ei), E, N, _lib.ptr(rowptr), _lib.ptr(flags), _lib.ptr(ws),
                                   ws.numel(), _lib.stream()), "x2_triplets_count")
    T = int(rowptr[E].item())            # the one host sync (output size)
    if int(flags[1].item()) != 0:
        raise IndexError(f"edge_index has atom ids outside [0, {N})")
    i64 = dict(dtype=torch.int64, device=dev)
    tri = torch.empty((2, T), **i64)
    ej, ei_, ek = (torch.empty(T, **i64) for _ in range(3))
    _lib.check(L.x2_triplets_fill(_lib.ptr(ei), E, N, _lib.ptr(rowptr), T, _lib.ptr(tri), _lib.ptr(ej),
                                  _lib.ptr(ei_), _lib.ptr(ek), _lib.ptr(ws), ws.numel(), _lib.stream()),
               "x2_triplets_fill")
    return tri, ej, ei_, ek
