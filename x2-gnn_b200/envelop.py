"""Polynomial cutoff envelope.  Drop-in for the reference's envelop.py (poly_envelop :5-21,
poly_envelop_func :23-32): u(x) = 1/x + a x^(p-1) + b x^p + c x^(p+1), x = d / cutoff, no clamp."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


def _coeffs(exponent: int):
    p = exponent + 1
    return p, -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2


def _envelope_fwd(x, inv_cutoff, p, a, b, c):
    dev = _lib.require_cuda(x, what="poly_envelop")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().x2_envelope_fwd(_lib.ptr(x), x.numel(), inv_cutoff, p, a, b, c, _lib.ptr(out),
                                          _lib.stream()), "x2_envelope_fwd")
    return out


class _EnvelopeFn(torch.autograd.Function):
    """Differentiable w.r.t. the distances like the reference's torch expression (envelop.py:16-21): the U0 training
    graph never asks for it, force training (dE/dpos) does."""

    @staticmethod
    def forward(ctx, d, inv_cutoff, p, a, b, c):
        x = _lib.f32(d, "poly_envelop")
        ctx.save_for_backward(x)
        ctx.args = (inv_cutoff, p, a, b, c)
        return _envelope_fwd(x, inv_cutoff, p, a, b, c)

    @staticmethod
    def backward(ctx, go):
        (x,) = ctx.saved_tensors
        inv_cutoff, p, a, b, c = ctx.args
        go = _lib.f32(go, "poly_envelop.backward")
        _lib.require_cuda(x, go, what="poly_envelop.backward")
        gd = torch.empty_like(x)
        _lib.check(_lib.lib().x2_envelope_bwd(_lib.ptr(x), _lib.ptr(go), x.numel(), inv_cutoff, p, a, b, c,
                                              _lib.ptr(gd), _lib.stream()), "x2_envelope_bwd")
        return gd, None, None, None, None, None


def _envelope(d: torch.Tensor, inv_cutoff: float, p: int, a: float, b: float, c: float):
    if d.requires_grad and torch.is_grad_enabled():
        return _EnvelopeFn.apply(d, inv_cutoff, p, a, b, c)
    return _envelope_fwd(_lib.f32(d, "poly_envelop"), inv_cutoff, p, a, b, c)


class poly_envelop(nn.Module):
    def __init__(self, cutoff, exponent):
        super().__init__()
        self.inv_cutoff = 1 / cutoff
        self.exponent = exponent
        self.p, self.a, self.b, self.c = _coeffs(exponent)

    def forward(self, distances):
        return _envelope(distances, self.inv_cutoff, self.p, self.a, self.b, self.c)


def poly_envelop_func(distances, cutoff=5.0, exponent=5):
    p, a, b, c = _coeffs(exponent)
    return _envelope(distances, 1 / cutoff, p, a, b, c)
