"""A training step of the harness model replayed as one CUDA graph.

The eager step (trainer.py:37-48: forward, SmoothL1, backward, clip, Adam, EMA) issues ~800 kernel
launches from Python and is bound by the host (17 ms per step for 11.8 ms of GPU work at batch 128).
Everything in it that depends on the parameters is shape-static for a given batch layout, so it is
captured once and replayed; what depends on the batch only -- the integer kernels that build the
triplets and CSR metadata, which read their output sizes back -- stays outside (`XGNNPoly.prepare`).

    step = GraphedTrainStep(model, data, y, lr=1e-3)      # warm-up + capture
    step.replay()                                         # one full update; step.loss is the loss tensor

A graph is tied to the device buffers and the (N, E, T) of the batch it was captured with: new values
are copied INTO `step.data[...]` / `step.y`, and a batch with a different layout needs its own capture
(a loader would keep one per padded shape bucket).  Multi-GPU: pass a `ddp.FlatGradBucket`; its NCCL
all-reduce is captured with the rest.
"""
from __future__ import annotations

import torch

from . import ddp


def fused_dense_step(model, tail, data, prep, y):
    """The same update with the tail (all-reduce buffer, clip, Adam, EMA) in flat buffers and two launches
    (optim_tail.FusedTail / x2_optim_tail) instead of ~300."""
    tail.zero_grad()
    loss = torch.nn.functional.smooth_l1_loss(model(data, prep), y)
    tail.backward(loss)
    tail.allreduce()
    tail.step()
    return loss


def dense_step(model, opt, params, data, prep, y, ema_params=None, ema_decay=0.95, bucket=None,
               max_norm=100.0):
    """The parameter-dependent part of one update, written with sync-free ops only (eager or under capture)."""
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.smooth_l1_loss(model(data, prep), y)
    loss.backward()
    if bucket is not None:
        bucket.pack()
        bucket.allreduce(average=True)
        bucket.unpack()
    torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm, foreach=True)
    opt.step()
    if ema_params is not None:        # swa_utils.get_ema_multi_avg_fn(decay): ema += (p - ema) * (1 - decay)
        torch._foreach_lerp_(ema_params, params, 1.0 - ema_decay)
    return loss


class GraphedTrainStep:
    def __init__(self, model, data: dict, y: torch.Tensor, lr: float = 1e-3, ema_decay: float = 0.95,
                 bucket: "ddp.FlatGradBucket | None" = None, max_norm: float = 100.0, warmup: int = 3,
                 fused_tail: bool = True, defer_wgrads: bool = True):
        """fused_tail: parameters / gradients / Adam moments / EMA in flat buffers, updated by x2_optim_tail (two
        launches; the flat gradient is also the all-reduce buffer, `bucket` is then only a flag that the step is
        data-parallel).  False: torch's capturable fused Adam + foreach clip / lerp, as in round 1."""
        dev = y.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (x2gnn_b200 has no CPU path)")
        self.model, self.data, self.y = model, data, y
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.prep = model.prepare(data)          # index tensors + CSR metadata the graph will point at
        self.stream = torch.cuda.Stream(device=dev)
        if fused_tail:
            from .optim_tail import FusedTail
            self.tail = FusedTail(self.params, lr=lr, max_norm=max_norm, ema_decay=ema_decay)
            # the weight gradients of the TCLinear layers: recorded during the backward, computed in a few batched
            # launches by the tail (tc_linear.DeferredWgrads), written straight into the flat gradient buffer
            self.deferred = self.tail.defer_wgrads(model, defer_wgrads)
            self.opt = None
            self.ema_params = self.tail.ema_views
            run = lambda: fused_dense_step(model, self.tail, data, self.prep, y)
        else:
            self.tail = None
            # capturable: the step counters live on the device, so Adam's bias correction replays correctly
            self.opt = torch.optim.Adam(self.params, lr=lr, fused=True, capturable=True)
            self.ema_params = [p.detach().clone() for p in self.params]
            args = (model, self.opt, self.params, data, self.prep, y, self.ema_params, ema_decay, bucket, max_norm)
            run = lambda: dense_step(*args)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):     # scratch buffers and lazy tables are created per stream
            for _ in range(warmup):
                run()
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.warmup_updates = warmup
        self.graph = torch.cuda.CUDAGraph()
        if self.opt is not None:
            self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.loss = run()

    def replay(self):
        self.graph.replay()
        return self.loss
