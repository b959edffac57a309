"""Radial basis sin(freq_n d / cutoff) with trainable frequencies.  Drop-in for the reference's
radial_basis_layer.py (RadialBasis :26-40, RadialBasis_func :19-24, radialbasis :6-17).
state_dict key: `frequencies` [R]."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import _lib


class _RadialFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, freq, inv_cutoff):
        d_ = _lib.f32(d, "RadialBasis")
        f_ = _lib.f32(freq, "RadialBasis")
        dev = _lib.require_cuda(d_, f_, what="RadialBasis")
        n, R = d_.numel(), f_.numel()
        out = torch.empty((n, R), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().x2_radial_fwd(_lib.ptr(d_), _lib.ptr(f_), None, n, R, inv_cutoff,
                                            _lib.ptr(out), _lib.stream()), "x2_radial_fwd")
        ctx.save_for_backward(d_, f_)
        ctx.inv_cutoff = inv_cutoff
        ctx.d_shape = d.shape
        return out.view(*d.shape, R)

    @staticmethod
    def backward(ctx, go):
        d_, f_ = ctx.saved_tensors
        n, R = d_.numel(), f_.numel()
        go = _lib.f32(go.reshape(n, R), "RadialBasis.backward")
        L = _lib.lib()
        gfreq = torch.empty(R, dtype=torch.float32, device=go.device)
        gd = torch.empty(n, dtype=torch.float32, device=go.device) if ctx.needs_input_grad[0] else None
        ws = _lib.workspace(L.x2_radial_bwd_workspace_bytes(n, R), go.device)
        _lib.check(L.x2_radial_bwd(_lib.ptr(d_), _lib.ptr(f_), None, _lib.ptr(go), n, R, ctx.inv_cutoff,
                                   _lib.ptr(gfreq), _lib.ptr(gd), _lib.ptr(ws), ws.numel(), _lib.stream()),
                   "x2_radial_bwd")
        return (gd.view(ctx.d_shape) if gd is not None else None), gfreq, None


def RadialBasis_func(bond_distances, cutoff=5.0, embedding_size=16):
    freq = (np.pi * torch.arange(1, embedding_size + 1, dtype=torch.float32)).to(bond_distances.device)
    return _RadialFn.apply(bond_distances, freq, 1 / cutoff)


def radialbasis(r, cutoff, embedding_size):
    """DimeNet-style sqrt(2/c) sin(n pi r / c) / r (unused by the model; reference :6-17)."""
    # the reference broadcasts r * n with n [1, embedding_size]: r is one distance ([1]) or a column ([num, 1])
    rr = r.reshape(-1)
    out = RadialBasis_func(rr, cutoff, embedding_size)
    return (2 / cutoff) ** 0.5 * out / rr.unsqueeze(-1)


class RadialBasis(nn.Module):
    def __init__(self, embedding_size, cutoff, Trainable=True, **kwargs):
        super().__init__(**kwargs)
        self.num_radial = embedding_size
        self.inv_cutoff = 1 / cutoff
        freq = np.pi * torch.arange(1, embedding_size + 1, dtype=torch.float32)
        if Trainable:
            self.frequencies = nn.Parameter(freq)
        else:   # the reference keeps a plain tensor attribute (no state_dict key)
            self.frequencies = freq

    def forward(self, bond_distances):
        freq = self.frequencies
        if freq.device != bond_distances.device:
            freq = freq.to(bond_distances.device)
        return _RadialFn.apply(bond_distances, freq, self.inv_cutoff)
