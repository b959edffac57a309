"""Bond lengths and triplet angles of the line graph (xgnn.py:46,60-66) as two sm_100a kernels.

The reference computes them with ~15 torch launches (five [T,3] / [E,3] position gathers, differences, a dot product,
a cross product, two norms, atan2).  Same fp32 operation order here, no intermediates.  Positions that require grad
(a force loss) take the torch expressions, which autograd differentiates as in the reference.
"""
from __future__ import annotations

import torch

from . import _lib


def _fast(pos, *idx):
    return (pos.is_cuda and pos.dtype == torch.float32 and pos.dim() == 2 and pos.size(1) == 3
            and not (torch.is_grad_enabled() and pos.requires_grad)
            and all(i.is_cuda and i.dtype == torch.int64 and i.dim() == 1 for i in idx))


def bond_lengths(pos: torch.Tensor, a0: torch.Tensor, a1: torch.Tensor) -> torch.Tensor:
    """|pos[a0] - pos[a1]| per bond (xgnn.py:46)."""
    if not _fast(pos, a0, a1):
        return torch.norm(pos[a0] - pos[a1], dim=1)
    p, a0, a1 = pos.contiguous(), a0.contiguous(), a1.contiguous()
    dev = _lib.require_cuda(p, a0, a1, what="bond_lengths")
    if a0.numel() != a1.numel():
        raise ValueError("bond_lengths: index tensors of different lengths")
    d = torch.empty(a0.numel(), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().x2_bond_lengths(_lib.ptr(p), _lib.ptr(a0), _lib.ptr(a1), a0.numel(), _lib.ptr(d),
                                          _lib.stream()), "x2_bond_lengths")
    return d


def triplet_angles(pos: torch.Tensor, a_i: torch.Tensor, a_j: torch.Tensor, a_k: torch.Tensor) -> torch.Tensor:
    """atan2(|ji x jk|, <ji, jk>) per triplet, j the central atom (xgnn.py:60-66)."""
    if not _fast(pos, a_i, a_j, a_k):
        ji, jk = pos[a_i] - pos[a_j], pos[a_k] - pos[a_j]
        return torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    p, a_i, a_j, a_k = pos.contiguous(), a_i.contiguous(), a_j.contiguous(), a_k.contiguous()
    dev = _lib.require_cuda(p, a_i, a_j, a_k, what="triplet_angles")
    if not (a_i.numel() == a_j.numel() == a_k.numel()):
        raise ValueError("triplet_angles: index tensors of different lengths")
    ang = torch.empty(a_j.numel(), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().x2_triplet_angles(_lib.ptr(p), _lib.ptr(a_i), _lib.ptr(a_j), _lib.ptr(a_k), a_j.numel(),
                                            _lib.ptr(ang), _lib.stream()), "x2_triplet_angles")
    return ang
