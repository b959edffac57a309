"""2-D Fourier-Bessel basis.  Drop-in for the reference's angular_basis_layer.py
(F_B_2D :51-93, AngularBasisLayer :12-32, AngularBasisLayer_func :34-48).

sbf[t, l*R+n] = env(d_s) * N_ln * j_l(z_ln * d_s / cutoff) * Y_l0(theta_t),  s = edge_index_1[t]

The reference builds 42 sympy-lambdified closures (~15 s construction, ~150 tiny launches and a
fp32-unstable closed form per forward).  Here construction is table generation (scipy, cached)
and forward is two kernels.  No parameters/buffers => no state_dict keys, as in the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .basis_func import bessel_tables
from .envelop import _coeffs


def _no_grad_inputs(name, *ts):
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        raise NotImplementedError(f"{name}: gradients w.r.t. distances/angles are not implemented "
                                  "(the reference training graph never needs them)")


def _angular(angles: torch.Tensor, L: int) -> torch.Tensor:
    _no_grad_inputs("AngularBasisLayer", angles)
    a = _lib.f32(angles, "AngularBasisLayer")
    dev = _lib.require_cuda(a, what="AngularBasisLayer")
    out = torch.empty((a.numel(), L), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().x2_angular_fwd(_lib.ptr(a), a.numel(), L, _lib.ptr(out), _lib.stream()),
               "x2_angular_fwd")
    return out


class AngularBasisLayer(nn.Module):
    def __init__(self, num_sph):
        super().__init__()
        self.num_sph = num_sph

    def forward(self, Angles):
        return _angular(Angles, self.num_sph)


def AngularBasisLayer_func(Angles, num_sph=16):
    return _angular(Angles, num_sph)


_table_cache = {}   # (L, R, device) -> (zeros, norm) device tensors; module-global, not on the Module


def _device_tables(L, R, device):
    key = (L, R, str(device))
    if key not in _table_cache:
        z, n = bessel_tables(L, R)
        _table_cache[key] = (torch.from_numpy(z).to(device), torch.from_numpy(n).to(device))
    return _table_cache[key]


class SbfFactors:
    """What F_B_2D.forward multiplied out: sbf[t, l R + n] = table[idx[t], l R + n] * Y_l0(angles[t]).
    Attached to the returned tensor as `_x2_factors`; SBFTransformerConv uses it to evaluate lin_sbf(sbf)
    inside its attention kernels instead of streaming the [T, L R] tensor (the tensor itself is still a
    plain, fully materialised torch.Tensor for every other consumer).  X2GNN_SGF=0 disables the shortcut."""
    __slots__ = ("table", "angles", "idx", "L", "R", "version", "_checked")

    def __init__(self, table, angles, idx, L, R, version):
        self.table, self.angles, self.idx, self.L, self.R, self.version = table, angles, idx, L, R, version
        self._checked = {}

    def describes(self, sbf: torch.Tensor, edge_index: torch.Tensor) -> bool:
        """True iff `sbf` is unmodified since F_B_2D produced it and idx == edge_index[0] (the source
        line-node of every triplet, xgnn.py:66)."""
        if sbf._version != self.version or edge_index.dim() != 2 or edge_index.size(1) != self.idx.numel():
            return False
        if edge_index.device != self.idx.device:
            return False
        if (self.idx.data_ptr() == edge_index.data_ptr() and edge_index.stride(1) == 1 and self.idx.stride(0) == 1
                and self.idx.dtype == edge_index.dtype):
            return True                       # idx IS row 0 of edge_index (a view): xgnn.py:66
        key = (id(edge_index), edge_index._version)
        if key not in self._checked:          # one comparison (and host sync) per (sbf, edge_index) pair
            self._checked = {key: bool(torch.equal(self.idx.long(), edge_index[0].long()))}
        return self._checked[key]


class F_B_2D(nn.Module):
    def __init__(self, num_spherical, num_radial, cutoff, envelope_exponent=5):
        super().__init__()
        assert num_radial <= 64
        self.num_radial = num_radial
        self.num_spherical = num_spherical
        self.cutoff = cutoff
        self.envelope_exponent = envelope_exponent
        self.envelope_cutoff = 5.0        # hard-coded in the reference (angular_basis_layer.py:60)
        bessel_tables(num_spherical, num_radial)   # warm the host-side table cache

    def radial_table(self, d: torch.Tensor) -> torch.Tensor:
        """[E, L*R] per-bond factor env(d) N_ln j_l(z_ln d / c) (E-scale, fp64 inside)."""
        d_ = _lib.f32(d, "F_B_2D")
        dev = _lib.require_cuda(d_, what="F_B_2D")
        L, R = self.num_spherical, self.num_radial
        zeros, norm = _device_tables(L, R, dev)
        p, a, b, c = _coeffs(self.envelope_exponent)
        table = torch.empty((d_.numel(), L * R), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().x2_sbf_table(_lib.ptr(d_), d_.numel(), L, R, _lib.ptr(zeros), _lib.ptr(norm),
                                           float(self.cutoff), float(self.envelope_cutoff), p, a, b, c,
                                           _lib.ptr(table), _lib.stream()), "x2_sbf_table")
        return table

    def forward(self, d, Angles, edge_index_1):
        _no_grad_inputs("F_B_2D", d, Angles)
        table = self.radial_table(d)
        ang = _lib.f32(Angles, "F_B_2D")
        idx = edge_index_1.long().contiguous()
        dev = _lib.require_cuda(ang, idx, table, what="F_B_2D")
        T = ang.numel()
        if idx.numel() != T:
            raise ValueError("F_B_2D: Angles and edge_index_1 must have the same length")
        E = table.size(0)
        # no host-side range check (it would force a device sync every forward); the kernel
        # clamps out-of-range rows, and edge_index_1 comes from vertex_to_edge_2
        L, R = self.num_spherical, self.num_radial
        out = torch.empty((T, L * R), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().x2_sbf_fwd(_lib.ptr(table), _lib.ptr(ang), _lib.ptr(idx), T, E, L, R,
                                         _lib.ptr(out), _lib.stream()), "x2_sbf_fwd")
        out._x2_factors = SbfFactors(table, ang, edge_index_1, L, R, out._version)
        return out
