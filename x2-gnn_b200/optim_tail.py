"""The parameter-update tail of a training step as two kernel launches.

trainer.py:43-48 runs, per step: clip_grad_norm_ (a norm per tensor, a stack, a norm, a clamp, a multiply per
tensor), Adam (the foreach groups) and the EMA average (train_ema.py:45-47: one lerp per parameter) -- ~300
launches for the 158 tensors of the model.  Here the parameters, their gradients, the Adam moments and the EMA
copy live in FLAT fp32 buffers (every `p.data` is a view into them, so the model sees ordinary tensors; the gradients autograd produces
are gathered by ONE multi-tensor copy) and `x2_optim_tail` reads and writes every element once.  With data
parallelism the flat gradient is the all-reduce buffer.

    tail = FusedTail(model.parameters(), lr=1e-3, max_norm=100.0, ema_decay=0.95)
    tail.zero_grad(); tail.backward(loss); tail.allreduce(); tail.step()      # (loss.backward() works too)

Sync-free and allocation-free after construction: capturable in a CUDA graph (the step count lives on the device).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


class FusedTail:
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 100.0,
                 ema_decay: float | None = None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedTail: no trainable parameters")
        dev = _lib.require_cuda(*self.params, what="FusedTail")
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("FusedTail: parameters must be float32")
        # every tensor starts on a 256-byte boundary of the flat buffers (the GEMM kernels take their fast paths on
        # 16-byte aligned weights; the padding holds zeros in every buffer and never moves)
        ALIGN = 64
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        f32 = dict(dtype=torch.float32, device=dev)
        self.n = n
        self.flat_p = torch.zeros(n, **f32)
        self.flat_g = torch.zeros(n, **f32)
        self.exp_avg = torch.zeros(n, **f32)
        self.exp_avg_sq = torch.zeros(n, **f32)
        self.grad_views, self.ema_views = [], []
        with torch.no_grad():
            for p, off in zip(self.params, offs):
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)        # the module's tensor is now a view
                self.grad_views.append(self.flat_g[off:off + k].view(p.shape))
        self.ema = self.flat_p.clone() if ema_decay is not None else None
        if self.ema is not None:
            for p, off in zip(self.params, offs):
                self.ema_views.append(self.ema[off:off + p.numel()].view(p.shape))
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.max_norm = float(max_norm) if max_norm else 0.0
        self.ema_decay = float(ema_decay) if ema_decay is not None else 0.0
        self.step_count = torch.zeros(1, **f32)
        self.grad_norm = torch.zeros(1, **f32)
        self._ws = _lib.workspace(_lib.lib().x2_optim_workspace_bytes(n), dev)
        self._world = 1
        self._packed = False
        self.active, self.defer, self.queue, self._deferred_ids, self._bias_of = False, False, None, set(), {}

    def defer_wgrads(self, model, on: bool = True):
        """Route the weight / bias gradients of the model's TCLinear layers around autograd: their backward records the
        operands, `_pack` computes all of them in a few batched launches (tc_linear.DeferredWgrads) straight into the
        flat gradient buffer.  Only layers whose out_features the tensor-core kernels take (a multiple of 128) defer;
        everything else keeps the ordinary path."""
        from .tc_linear import DeferredWgrads, TCLinear, WgradSlots
        view = {id(p): v for p, v in zip(self.params, self.grad_views)}
        self.queue = DeferredWgrads()
        self.defer = bool(on)          # `active` is only raised around a backward pass (backward_scope)
        self._deferred_ids = set()
        for m in model.modules():
            if isinstance(m, TCLinear):
                gw = view.get(id(m.weight))
                gb = view.get(id(m.bias)) if m.bias is not None else None
                ok = on and gw is not None and m.out_features % 128 == 0 and (m.bias is None or gb is not None)
                m._x2_slots = WgradSlots(self, gw, gb) if ok else None
                if ok:
                    self._deferred_ids.add(id(m.weight))
                    if m.bias is not None:
                        self._deferred_ids.add(id(m.bias))
                        self._bias_of[id(m.bias)] = gw.data_ptr()
        return len(self._deferred_ids)

    def backward(self, loss):
        """loss.backward() with the TCLinear weight gradients deferred to _pack (when defer_wgrads attached them).
        Only a backward pass run through here defers: any other use of the model's autograd graph is untouched."""
        self.active = self.defer
        try:
            loss.backward()
        finally:
            self.active = False

    def zero_grad(self):
        """Gradients are produced by autograd as usual (fresh tensors, no accumulation kernels) and gathered into
        the flat buffer by step() / allreduce() with one multi-tensor copy."""
        for p in self.params:
            p.grad = None
        self._packed = False

    def _pack(self):
        if self._packed:
            return
        have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
        pend = self.queue.pending if self.queue is not None else ()
        n_def = sum(1 for v, p in zip(self.grad_views, self.params)
                    if p.grad is None and id(p) in self._deferred_ids and self._covered(p, v, pend)) if pend else 0
        if len(have) + n_def != len(self.params):
            self.flat_g.zero_()                    # parameters without a gradient this step
        if pend:
            self.queue.flush(self.flat_g.device)   # deferred TCLinear weight gradients -> their blocks of flat_g
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        self._packed = True

    def _covered(self, p, v, pend):
        """True iff the deferred queue holds this parameter's gradient (a bias rides with its weight)."""
        if v.data_ptr() in pend:
            return True
        w = self._bias_of.get(id(p))
        return w is not None and w in pend

    def allreduce(self):
        """Sum over the ranks; the division by the world size is folded into step()."""
        self._pack()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            self._world = dist.get_world_size()
        return self.flat_g

    def step(self):
        self._pack()
        L = _lib.lib()
        _lib.require_cuda(self.flat_p, what="FusedTail.step")
        _lib.check(L.x2_optim_tail(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.exp_avg),
                                   _lib.ptr(self.exp_avg_sq), _lib.ptr(self.ema), self.n, 1.0 / self._world,
                                   self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps, self.ema_decay,
                                   _lib.ptr(self.step_count), _lib.ptr(self.grad_norm), _lib.ptr(self._ws),
                                   self._ws.numel(), _lib.stream()), "x2_optim_tail")
        self._packed = False
