"""torch.nn.Linear on the tcgen05 3xTF32 GEMMs of libx2gnn (x2_tc_gemm / x2_tc_wgrad).

Used by the harness model for the dense layers AROUND the hot path (ResidualLayer, readout MLPs,
mat_trans / emb_trans, edgenn -- SURVEY.md §8f rows 1 and 3), which otherwise run as cuBLAS SIMT fp32
GEMMs (fp32 accuracy is required, TF32 is off) and take ~45 % of a training step.  Same parameters and
state_dict keys as torch.nn.Linear; fp32-accurate (3-term split).  Shapes the kernels do not take
(out_features not a multiple of 128) fall back to F.linear on the GPU.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib

_BLK = 128

# Scratch for the GEMM kernels, one buffer per (device, stream), grown on demand.  The kernels of one
# call only use it between their own launches and calls on a stream are ordered, so it can be shared
# by every TCLinear on that stream -- saves two allocator round trips per call (the harness model makes
# ~290 of these calls per training step and is bound by host overhead).
_ws_cache: dict = {}


def _scratch(nbytes: int, dev):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws[0].numel() < nbytes:
        t = _lib.workspace(nbytes, dev)
        ws = (t, _lib.ptr(t), t.numel())
        _ws_cache[key] = ws
    return ws


class _TCLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = _lib.f32(x.reshape(-1, x.size(-1)), "TCLinear.x")
        w = _lib.f32(weight, "TCLinear.weight")
        b = _lib.f32(bias, "TCLinear.bias") if bias is not None else None
        dev = _lib.require_cuda(x2, w, b, what="TCLinear")
        M, K = x2.shape
        N = w.size(0)
        L = _lib.lib()
        y = torch.empty((M, N), dtype=torch.float32, device=dev)
        _, ws_p, ws_n = _scratch(L.x2_tc_gemm_workspace_bytes(_BLK, _BLK), dev)
        st = _lib.stream()
        esz = 4
        for n0 in range(0, N, _BLK):
            nb = min(_BLK, N - n0)
            for k0 in range(0, K, _BLK):
                kb = min(_BLK, K - k0)
                _lib.check(L.x2_tc_gemm(
                    x2.data_ptr() + k0 * esz, K, M, kb, w.data_ptr() + (n0 * K + k0) * esz, 1, K, nb,
                    (b.data_ptr() + n0 * esz) if (b is not None and k0 == 0) else None,
                    y.data_ptr() + n0 * esz, N, 1 if k0 > 0 else 0, ws_p, ws_n, st), "x2_tc_gemm")
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.size(0)
        gy2 = _lib.f32(gy.reshape(M, N), "TCLinear.grad")
        dev = gy2.device
        L = _lib.lib()
        st = _lib.stream()
        esz = 4
        gx = gw = gb = None
        _, ws_p, ws_n = _scratch(max(L.x2_tc_gemm_workspace_bytes(_BLK, _BLK), L.x2_tc_wgrad_workspace_bytes(M, _BLK)), dev)
        if ctx.needs_input_grad[0]:
            gx = torch.empty((M, K), dtype=torch.float32, device=dev)
            for k0 in range(0, K, _BLK):          # output columns (in_features)
                kb = min(_BLK, K - k0)
                for n0 in range(0, N, _BLK):      # reduction over out_features
                    nb = min(_BLK, N - n0)
                    _lib.check(L.x2_tc_gemm(
                        gy2.data_ptr() + n0 * esz, N, M, nb, w.data_ptr() + (n0 * K + k0) * esz, K, 1, kb, None,
                        gx.data_ptr() + k0 * esz, K, 1 if n0 > 0 else 0, ws_p, ws_n, st), "x2_tc_gemm")
            gx = gx.view(ctx.x_shape)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty((N, K), dtype=torch.float32, device=dev)
            gb = torch.empty(N, dtype=torch.float32, device=dev) if ctx.has_bias else None
            for n0 in range(0, N, _BLK):
                for k0 in range(0, K, _BLK):
                    kb = min(_BLK, K - k0)
                    _lib.check(L.x2_tc_wgrad(
                        gy2.data_ptr() + n0 * esz, N, x2.data_ptr() + k0 * esz, K, M, kb,
                        gw.data_ptr() + (n0 * K + k0) * esz, K,
                        (gb.data_ptr() + n0 * esz) if (gb is not None and k0 == 0) else None,
                        ws_p, ws_n, st), "x2_tc_wgrad")
        return gx, gw, gb


def tc_linear(x, weight, bias=None):
    """F.linear on the tensor cores when the shape allows, else F.linear."""
    if (x.is_cuda and x.dtype == torch.float32 and weight.size(0) % _BLK == 0 and x.numel() > 0
            and x.size(-1) == weight.size(1)):
        return _TCLinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class TCLinear(nn.Linear):
    """Drop-in nn.Linear (same parameters / state_dict keys) that runs on x2_tc_gemm / x2_tc_wgrad."""

    def forward(self, x):
        return tc_linear(x, self.weight, self.bias)
