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


def _scratch(nbytes: int, dev, stream_id: int):
    key = (dev.index, stream_id)
    ws = _ws_cache.get(key)
    if ws is None or ws[2] < nbytes:
        t = _lib.workspace(nbytes, dev)
        ws = (t, t.data_ptr(), t.numel())
        _ws_cache[key] = ws
    return ws


_ws_need: dict = {}          # workspace sizes are functions of the (fixed) block shape / row count: ask once


def _need_fwd(L):
    n = _ws_need.get("g")
    if n is None:
        n = _ws_need["g"] = int(L.x2_tc_gemm_workspace_bytes(_BLK, _BLK))
    return n


def _need_bwd(L, M):
    n = _ws_need.get(M)
    if n is None:
        n = _ws_need[M] = max(_need_fwd(L), int(L.x2_tc_wgrad_workspace_bytes(M, _BLK)))
    return n


def _f32c(t, what):
    if t.dtype is not torch.float32:
        raise TypeError(f"{what}: expected float32, got {t.dtype} (the sm_100a kernels are fp32-I/O)")
    return t if t.is_contiguous() else t.contiguous()


class WgradSlots:
    """What a TCLinear needs to defer its weight gradient: the sink (optim_tail.FusedTail) and the blocks of the flat
    gradient buffer that belong to its weight / bias.  Never copied or pickled with the module (copy.deepcopy of a
    model -- the EMA AveragedModel, train_ema.py:45-47 -- must not drag the optimizer state along)."""
    __slots__ = ("sink", "gw", "gb")

    def __init__(self, sink, gw, gb):
        self.sink, self.gw, self.gb = sink, gw, gb

    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (type(None), ())


class DeferredWgrads:
    """Weight gradients of the TCLinear layers of one backward pass, computed together at the end.

    Nothing downstream of a Linear's backward needs dW / db before the optimizer, while each of them, launched on its
    own, is a 25-us kernel + a 4-us reduction for ~7 us of memory traffic (51 of them per training step of the harness
    model: 1.5 ms of 8.9).  With a sink attached (optim_tail.FusedTail.defer_wgrads), `_TCLinearFn.backward` only
    records (grad_out, input, destination block of the FLAT gradient buffer) and returns no weight gradient to
    autograd; `flush()` then runs all of them through x2_tc_wgrad_batch (16 problems per launch, the CTAs dealt over
    the problems, long row ranges per CTA) and the results land directly in the optimizer's gradient buffer.
    One backward per flush: a slot recorded twice (shared weights, two backward passes) raises."""

    def __init__(self):
        self.jobs = {}          # (rows, kb, aligned) -> list of (gy, x, job fields)
        self.pending = set()    # data_ptr of the destination blocks recorded since the last flush
        self.keep = []          # tensors kept alive until the flush has been enqueued

    def add(self, gy2, x2, M, N, K, gw_view, gb_view):
        key_w = gw_view.data_ptr()
        if key_w in self.pending:
            raise RuntimeError("TCLinear: a weight's gradient was deferred twice before the flush (shared weights or "
                               "two backward passes per step are not supported with deferred weight gradients)")
        self.pending.add(key_w)
        self.keep += [gy2, x2]
        gp, xp, gwp = gy2.data_ptr(), x2.data_ptr(), gw_view.data_ptr()
        gbp = gb_view.data_ptr() if gb_view is not None else 0
        for n0 in range(0, N, _BLK):
            for k0 in range(0, K, _BLK):
                kb = min(_BLK, K - k0)
                xa = xp + k0 * 4
                aligned = (xa % 16 == 0) and (K % 4 == 0)
                self.jobs.setdefault((M, kb, aligned), []).append(
                    (gp + n0 * 4, N, xa, K, gwp + (n0 * K + k0) * 4, K, (gbp + n0 * 4) if (gbp and k0 == 0) else 0))

    def flush(self, dev):
        if not self.jobs:
            return 0
        L = _lib.lib()
        cur = torch.cuda.current_stream(dev)
        st = cur.cuda_stream
        # operands recorded by a backward node that ran on another stream (the model's side-stream readouts): the
        # backward has been joined into this stream by autograd, but the allocator must also know they are read here
        for t in self.keep:
            t.record_stream(cur)
        n = 0
        for (M, kb, _), lst in self.jobs.items():
            arr = (_lib.WgradJob * len(lst))(*[_lib.WgradJob(*j) for j in lst])
            _, ws_p, ws_n = _scratch(_need_bwd(L, M), dev, st)
            rc = L.x2_tc_wgrad_batch(arr, len(lst), M, kb, ws_p, ws_n, st)
            if rc:
                _lib.check(rc, "x2_tc_wgrad_batch")
            n += len(lst)
        self.jobs, self.pending, self.keep = {}, set(), []
        return n


class _TCLinearFn(torch.autograd.Function):
    # The harness model makes ~100 of these calls per training step and the step is bound by host time, so
    # the Python around the two C calls is kept minimal: one current_stream() lookup, cached workspace
    # sizes, raw integer pointers, the error path only when a call fails.
    @staticmethod
    def forward(ctx, x, weight, bias, slots=None):
        x2 = _f32c(x.reshape(-1, x.size(-1)), "TCLinear.x")
        w = _f32c(weight, "TCLinear.weight")
        b = _f32c(bias, "TCLinear.bias") if bias is not None else None
        dev = x2.device
        if (dev.index not in _lib._checked_devices or not w.is_cuda or w.device != dev
                or (b is not None and b.device != dev)):
            _lib.require_cuda(x2, w, b, what="TCLinear")      # full check (incl. sm_100) on first use / mismatch
        M, K = x2.shape
        N = w.size(0)
        L = _lib.lib()
        y = torch.empty((M, N), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _, ws_p, ws_n = _scratch(_need_fwd(L), dev, st)
        xp, wp, yp = x2.data_ptr(), w.data_ptr(), y.data_ptr()
        bp = b.data_ptr() if b is not None else 0
        for n0 in range(0, N, _BLK):
            nb = min(_BLK, N - n0)
            for k0 in range(0, K, _BLK):
                kb = min(_BLK, K - k0)
                rc = L.x2_tc_gemm(xp + k0 * 4, K, M, kb, wp + (n0 * K + k0) * 4, 1, K, nb,
                                  (bp + n0 * 4) if (bp and k0 == 0) else None,
                                  yp + n0 * 4, N, 1 if k0 > 0 else 0, ws_p, ws_n, st)
                if rc:
                    _lib.check(rc, "x2_tc_gemm")
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.x_shape = x.shape
        ctx.slots = slots          # WgradSlots or None
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        x2, w = ctx.saved_tensors
        M, K = x2.shape
        N = w.size(0)
        gy2 = _f32c(gy.reshape(M, N), "TCLinear.grad")
        dev = gy2.device
        L = _lib.lib()
        st = torch.cuda.current_stream(dev).cuda_stream
        gx = gw = gb = None
        _, ws_p, ws_n = _scratch(_need_bwd(L, M), dev, st)
        gp, wp, xp = gy2.data_ptr(), w.data_ptr(), x2.data_ptr()
        need = ctx.needs_input_grad
        if need[0]:
            gx = torch.empty((M, K), dtype=torch.float32, device=dev)
            gxp = gx.data_ptr()
            for k0 in range(0, K, _BLK):          # output columns (in_features)
                kb = min(_BLK, K - k0)
                for n0 in range(0, N, _BLK):      # reduction over out_features
                    nb = min(_BLK, N - n0)
                    rc = L.x2_tc_gemm(gp + n0 * 4, N, M, nb, wp + (n0 * K + k0) * 4, K, 1, kb, None,
                                      gxp + k0 * 4, K, 1 if n0 > 0 else 0, ws_p, ws_n, st)
                    if rc:
                        _lib.check(rc, "x2_tc_gemm")
            gx = gx.view(ctx.x_shape)
        sl = ctx.slots
        if sl is not None and sl.sink.active and need[1] and M > 0:
            # dW / db go straight to the optimizer's flat gradient buffer, with all the others, at the end
            sl.sink.queue.add(gy2, x2, M, N, K, sl.gw, sl.gb if ctx.has_bias else None)
            return gx, None, None, None
        if need[1] or (ctx.has_bias and need[2]):
            gw = torch.empty((N, K), dtype=torch.float32, device=dev)
            gb = torch.empty(N, dtype=torch.float32, device=dev) if ctx.has_bias else None
            gwp = gw.data_ptr()
            gbp = gb.data_ptr() if gb is not None else 0
            for n0 in range(0, N, _BLK):
                for k0 in range(0, K, _BLK):
                    kb = min(_BLK, K - k0)
                    rc = L.x2_tc_wgrad(gp + n0 * 4, N, xp + k0 * 4, K, M, kb, gwp + (n0 * K + k0) * 4, K,
                                       (gbp + n0 * 4) if (gbp and k0 == 0) else None, ws_p, ws_n, st)
                    if rc:
                        _lib.check(rc, "x2_tc_wgrad")
        return gx, gw, gb, None


def tc_linear(x, weight, bias=None, slots=None):
    """F.linear on the tensor cores when the shape allows, else F.linear."""
    if (x.is_cuda and x.dtype == torch.float32 and weight.size(0) % _BLK == 0 and x.numel() > 0
            and x.size(-1) == weight.size(1)):
        return _TCLinearFn.apply(x, weight, bias, slots)
    return F.linear(x, weight, bias)


class TCLinear(nn.Linear):
    """Drop-in nn.Linear (same parameters / state_dict keys) that runs on x2_tc_gemm / x2_tc_wgrad."""

    _x2_slots = None       # a WgradSlots, set by optim_tail.FusedTail.defer_wgrads

    def forward(self, x):
        return tc_linear(x, self.weight, self.bias, self._x2_slots)
