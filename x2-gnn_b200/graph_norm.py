"""Graph-wise LayerNorm (no affine parameters) on the sm_100a kernels of libx2gnn.

Replaces the call `LayerNorm(in_channels, eps=1e-8, affine=False)(x=out, batch=data.batch)` of the
reference (model.py:24,46; torch_geometric.nn.LayerNorm 2.1.0 with a batch vector): statistics over all
rows and channels of a molecule.  The reference runs it as two scatter-adds, two gathers and ~8
elementwise passes over [E, D] forward and about twice that backward (~0.4 ms per layer at the QM9
batch-128 shape); here it is one kernel each way, one CTA per molecule (x2_graph_layernorm_fwd / _bwd).

The molecules' rows must be contiguous and in graph order -- what PyG collation produces -- and are
described by `rowptr` [B+1] int32 (cumulative rows per graph), built once per batch.
"""
from __future__ import annotations

import torch

from . import _lib


def rowptr_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """[B] rows per graph -> [B+1] int32 offsets (device tensor; no host sync)."""
    rp = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=counts.device)
    rp[1:] = torch.cumsum(counts, 0)
    return rp


class _GraphLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rowptr, eps):
        x_ = _lib.f32(x, "graph_layer_norm.x")
        dev = _lib.require_cuda(x_, rowptr, what="graph_layer_norm")
        if x_.dim() != 2 or x_.size(1) % 4:
            raise ValueError(f"graph_layer_norm: x must be [rows, D] with D % 4 == 0, got {tuple(x_.shape)}")
        if rowptr.dtype != torch.int32 or rowptr.dim() != 1 or rowptr.numel() < 1 or not rowptr.is_contiguous():
            raise ValueError("graph_layer_norm: rowptr must be a contiguous int32 [B+1] tensor")
        B, D = rowptr.numel() - 1, x_.size(1)
        y = torch.empty_like(x_)
        stats = torch.empty((B, 2), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().x2_graph_layernorm_fwd(_lib.ptr(x_), _lib.ptr(rowptr), B, D, float(eps),
                                                     _lib.ptr(y), _lib.ptr(stats), _lib.stream()),
                   "x2_graph_layernorm_fwd")
        ctx.save_for_backward(y, stats, rowptr)
        return y

    @staticmethod
    def backward(ctx, gy):
        y, stats, rowptr = ctx.saved_tensors
        gy_ = _lib.f32(gy, "graph_layer_norm.grad")
        gx = torch.empty_like(y)
        _lib.check(_lib.lib().x2_graph_layernorm_bwd(_lib.ptr(y), _lib.ptr(gy_), _lib.ptr(rowptr),
                                                     rowptr.numel() - 1, y.size(1), _lib.ptr(stats),
                                                     _lib.ptr(gx), _lib.stream()), "x2_graph_layernorm_bwd")
        return gx, None, None


def graph_layer_norm_rows(x: torch.Tensor, rowptr: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """y = (x - mean_g) / sqrt(var_g + eps) per molecule g = rows rowptr[g]..rowptr[g+1] of x, all channels.
    `rowptr[-1]` must equal x.size(0) (checked by the caller that builds it: a check here would synchronise)."""
    return _GraphLayerNormFn.apply(x, rowptr, eps)
