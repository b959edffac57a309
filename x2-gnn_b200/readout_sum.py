"""rbf-gated bond -> atom readout sum on the sm_100a kernels of libx2gnn.

Replaces the first two lines of the reference's `AtomWise.forward` (readout.py:34-43):
`out = scatter(lin_rbf(rbf) * x, edge_index[0], dim=0)` -- a Linear, an elementwise product and an atomic
scatter-add there (and a gather, two products and two GEMMs in the backward) -- with one kernel each way
(x2_rbf_readout_fwd / _bwd): warp per atom over its contiguous bonds, deterministic, no [E, D] intermediate.

Bonds must be sorted by their first atom (the reference's edge_index is lexicographic, atom_graph.py:42-45);
`rowptr` [N+1] int32 = cumulative bonds per atom, built once per batch.
"""
from __future__ import annotations

import torch

from . import _lib


def supported(D: int, R: int) -> bool:
    return D in (128, 256) and 1 <= R <= 16


class _RbfReadoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rbf, weight, bias, rowptr):
        x_, r_ = _lib.f32(x, "rbf_readout.x"), _lib.f32(rbf, "rbf_readout.rbf")
        w_, b_ = _lib.f32(weight, "rbf_readout.weight"), _lib.f32(bias, "rbf_readout.bias")
        dev = _lib.require_cuda(x_, r_, w_, b_, rowptr, what="rbf_readout")
        E, D = x_.shape
        R = r_.size(1)
        if r_.size(0) != E or tuple(w_.shape) != (D, R) or not supported(D, R):
            raise ValueError(f"rbf_readout: x [E,D], rbf [E,R], weight [D,R] with D in (128, 256), R <= 16; got "
                             f"{tuple(x_.shape)}, {tuple(r_.shape)}, {tuple(w_.shape)}")
        if rowptr.dtype != torch.int32 or rowptr.dim() != 1 or rowptr.numel() < 1 or not rowptr.is_contiguous():
            raise ValueError("rbf_readout: rowptr must be a contiguous int32 [N+1] tensor")
        N = rowptr.numel() - 1
        out = torch.empty((N, D), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().x2_rbf_readout_fwd(_lib.ptr(x_), _lib.ptr(r_), _lib.ptr(w_), _lib.ptr(b_),
                                                 _lib.ptr(rowptr), N, E, D, R, _lib.ptr(out), _lib.stream()),
                   "x2_rbf_readout_fwd")
        ctx.save_for_backward(x_, r_, w_, b_, rowptr)
        return out

    @staticmethod
    def backward(ctx, g):
        x_, r_, w_, b_, rowptr = ctx.saved_tensors
        g_ = _lib.f32(g, "rbf_readout.grad")
        dev = g_.device
        (E, D), R, N = x_.shape, r_.size(1), rowptr.numel() - 1
        f32 = dict(dtype=torch.float32, device=dev)
        dx, drbf = torch.empty((E, D), **f32), torch.empty((E, R), **f32)
        dw = torch.empty((D, R), **f32)
        db = torch.empty(D, **f32) if b_ is not None else None
        L = _lib.lib()
        ws = _lib.workspace(L.x2_rbf_readout_bwd_workspace_bytes(N, E, D, R), dev)
        _lib.check(L.x2_rbf_readout_bwd(_lib.ptr(x_), _lib.ptr(r_), _lib.ptr(w_), _lib.ptr(b_), _lib.ptr(rowptr),
                                        _lib.ptr(g_), N, E, D, R, _lib.ptr(dx), _lib.ptr(drbf), _lib.ptr(dw),
                                        _lib.ptr(db), _lib.ptr(ws), ws.numel(), _lib.stream()),
                   "x2_rbf_readout_bwd")
        return dx, drbf, dw, db, None


def rbf_readout(x, rbf, weight, bias, rowptr):
    """out[n] = sum_{e in rowptr[n]..rowptr[n+1]} (weight @ rbf[e] + bias) * x[e]  ->  [N, D]."""
    return _RbfReadoutFn.apply(x, rbf, weight, bias, rowptr)
