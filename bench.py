#!/usr/bin/env python
"""bench.py -- headline benchmark of the X2-GNN hot path on B200.

Metric (BASELINE.json): SBF-conv edge-messages/s, forward + backward, fp32.  One "step" = one
SBFTransformerConv layer (config.json dims D=128 H=16 S=42 R=6 A=128) forward + backward over one
synthetic QM9-shaped batch of 128 molecules (BASELINE.json configs[1]); one edge-message = one
triplet through the layer.  Inputs are resident in HBM for `value`; `e2e` repeats the measurement
through the public module call with pinned HOST buffers (H2D of every input, line-graph metadata
build, forward, backward, D2H of the results inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

N > 1: graph batches are sharded per GPU (rank r draws its own 128-molecule batch, weak scaling);
the only collective is the NCCL all-reduce of the layer's parameter gradients, inside the step.
`--impl reference` times the reference's CPU path (the oracle port of the reference's composite
PyTorch ops -- the reference has no native code to compile) on the host cores.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sbfconv_edge_messages_per_sec_fwd_bwd"
UNIT = "edge-messages/s"
DIMS = dict(D=128, H=16, S=42, R=6, A=128)       # reference config.json:2-7
NMOL = 128                                       # BASELINE.json configs[1]: batch 128


def algorithmic_bytes(E, T, D, S, R, A):
    """every interface tensor touched once, fp32, int64 indices."""
    fwd = 4 * (E * (D + R) + T * (S + A) + E * D) + 16 * T
    bwd = 4 * (E * D + E * (D + R) + T * (S + A)) + 16 * T + 4 * (E * (D + R) + T * A)
    return fwd, bwd


def host_workload(rank: int = 0, world: int = 1, nmol: int = NMOL):
    """The rank's batch of nmol molecules: rank 0 (and world == 1) is the seed-0 batch, every other rank draws its
    own molecules with the same triplet count to within 0.5 % (x2gnn_b200.synth.qm9_shard): weak scaling with
    identical per-rank work at every N."""
    from x2gnn_b200 import synth
    b, mine = synth.qm9_shard(nmol, world, rank, seed=0)
    tri = synth.triplets_host(b["edge_index"], len(b["x"]))[0]
    E = b["edge_index"].shape[1]
    ci = synth.conv_inputs(E, tri, DIMS["D"], DIMS["S"], DIMS["R"], DIMS["A"], seed=rank)
    # central atom j of every directed bond (i -> j): all triplets of a target bond share it (xgnn.py:57-58)
    return dict(N=len(b["x"]), E=E, T=tri.shape[1], center=b["edge_index"][1].copy(), nmol=len(mine), **ci)


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
        self.period = period_s

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------- reference / CPU arm
def oracle_step_fn(w, nthreads):
    """The reference's CPU path for this workload: composite PyTorch ops of the oracle port."""
    import torch
    from oracle import conv as oconv
    torch.set_num_threads(nthreads)
    torch.manual_seed(0)
    layer = oconv.OracleSBFTransformerConv(DIMS["D"], DIMS["D"] // DIMS["H"], heads=DIMS["H"],
                                           sbf_dim=DIMS["S"], rbf_dim=DIMS["R"], edge_dim=DIMS["A"])
    x = torch.from_numpy(w["x"]).requires_grad_(True)
    rbf = torch.from_numpy(w["rbf"]).requires_grad_(True)
    ea = torch.from_numpy(w["edge_attr"]).requires_grad_(True)
    sbf = torch.from_numpy(w["sbf"])
    ei = torch.from_numpy(w["edge_index"])
    gout = torch.randn(w["E"], DIMS["D"], generator=torch.Generator().manual_seed(1))
    params = list(layer.parameters())

    def step():
        out = layer(sbf, rbf, x=x, edge_index=ei, edge_attr=ea)
        torch.autograd.grad(out, [x, rbf, ea] + params, gout)
    return step


def time_cpu(step, iters, warmup=1):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    return (time.perf_counter() - t0) / iters


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    # bounded sample: pick the number of molecules so that (steps + warmup) steps fit ~150 s
    probe = host_workload(0, 1, 16)
    t_probe = time_cpu(oracle_step_fn(probe, ncores), 1, warmup=1)
    rate = probe["T"] / t_probe
    full = host_workload(0, 1, NMOL)
    budget_T = rate * 150.0 / max(steps + warmup, 1)
    nmol = NMOL if budget_T >= full["T"] else max(4, min(NMOL, int(NMOL * budget_T / full["T"])))
    w = full if nmol == NMOL else host_workload(0, 1, nmol)
    sec = time_cpu(oracle_step_fn(w, ncores), steps, warmup=warmup)
    value = w["T"] / sec
    sample = (f"{nmol} of {NMOL} molecules of the seed-0 batch (E={w['E']}, T={w['T']}), "
              f"{warmup} warm-up + {steps} timed fwd+bwd steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "qm9_b128_sbfconv_layer_fwd_bwd (BASELINE.json configs[1])", **DIMS,
                   "E": w["E"], "T": w["T"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- full training step
HPARAMS = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)  # config.json


def train_step_bench(dev, world, rank, steps, warmup, with_cpu):
    """QM9 molecules/s of a full U0 training step (BASELINE.json configs[1]/[4]): forward of the
    xgnn_poly-equivalent harness model on a 128-molecule batch per GPU, SmoothL1 loss, backward,
    flat-bucket gradient all-reduce (N > 1), global-norm clip, Adam, EMA -- trainer.py:37-48."""
    import torch
    import torch.distributed as dist
    from x2gnn_b200 import ddp, synth
    from x2gnn_b200.xgnn_model import XGNNPoly

    torch.manual_seed(0)
    model = XGNNPoly(**HPARAMS).to(dev)
    ema = torch.optim.swa_utils.AveragedModel(model, multi_avg_fn=torch.optim.swa_utils.get_ema_multi_avg_fn(0.95))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)    # same update rule, one kernel
    b, mine = synth.qm9_shard(NMOL, world, rank, seed=0)     # balanced share of the global batch of NMOL * world
    nmine = len(mine)
    data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in b.items()}
    y = torch.zeros(nmine, device=dev)
    bucket = ddp.FlatGradBucket(model.parameters())

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.smooth_l1_loss(model(data), y)
        loss.backward()
        if world > 1:
            bucket.pack()
            bucket.allreduce(average=True)
            bucket.unpack()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=100.0)
        opt.step()
        ema.update_parameters(model)
        return loss

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # The first iterations of a loop that starts from an idle device run slow on every box seen (10-100 ms
    # outliers at fixed positions 1 and 3, whatever the warm-up before the synchronize): LEAD more untimed
    # iterations run inside the loop itself and the K timed steps follow them without a pause (ranks stay in
    # step through the all-reduce of every iteration; barrier + synchronize bracket the whole loop).
    LEAD = 4
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(LEAD + steps + 1)]
    gc.collect()
    gc.disable()
    for i in range(LEAD + steps):
        marks[i].record()
        loss = step()
    marks[-1].record()
    gc.enable()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    marks = marks[LEAD:]
    ms = marks[0].elapsed_time(marks[-1]) / steps
    in_order = [round(marks[i].elapsed_time(marks[i + 1]), 3) for i in range(steps)]
    per_step = sorted(in_order)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    # the step is bound by host-side launch overhead (~800 launches), so host jitter shows: mean and median
    res = {"molecules_per_sec": world * NMOL / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
           "ms_per_step_median": per_step[len(per_step) // 2], "ms_per_step_min": per_step[0],
           "ms_each_step": in_order,
           "molecules_per_gpu": NMOL, "molecules_this_rank": nmine, "loss": float(loss.detach()), "N": len(b["x"]),
           "E": int(b["edge_index"].shape[1]),
           "triplets_this_rank": int(synth.triplets_host(b["edge_index"], len(b["x"]))[0].shape[1]),
           "sharding": "128 molecules per rank; every rank's batch has the N = 1 batch's triplet count to within 0.5 % "
                       "(synth.qm9_shard), so molecules/s and triplets/s scale alike"}
    # ---- the same step with its dense part replayed as ONE CUDA graph (the eager step above is bound by the
    # host: ~800 launches from Python).  Every replayed step still runs the whole path: `prepare` (the integer
    # kernels building triplets + CSR metadata, eager, with their size read-backs) is inside the timed region,
    # then forward, loss, backward, gradient all-reduce, clip, Adam and EMA replay from the graph.  A graph is
    # tied to the batch's (N, E, T): a loader would keep one per padded shape bucket; the bench batch is fixed.
    res["mode"] = "eager"
    try:
        g = _train_step_graph(dev, world, rank, steps, data, y, nmine)
        # the graphed step is the implementation's training step: it leads, the eager numbers stay beside it
        eager = {k: res.pop(k) if k == "ms_each_step" else res[k]
                 for k in ("molecules_per_sec", "ms_per_step", "ms_per_step_median", "ms_per_step_min", "steps",
                           "loss", "ms_each_step")}
        eager["what"] = "the same step issued launch by launch from Python (~560 kernels; bound by the host)"
        res.update({"mode": "cuda_graph", "molecules_per_sec": g.pop("molecules_per_sec"),
                    "ms_per_step": g.pop("ms_per_step"), "ms_per_step_median": g.pop("ms_per_step_median"),
                    "steps": g.pop("steps"), "ms_each_step": g.pop("ms_each_step")})
        res.pop("ms_per_step_min", None)
        res["loss"] = g.get("graph_loss")
        # the host of a shared box stalls single iterations (prepare() runs on it): the median step is the
        # steady state, the mean over the K steps stays the reported ms_per_step
        res["molecules_per_sec_at_median_step"] = world * NMOL / (res["ms_per_step_median"] * 1e-3)
        res["cuda_graph"] = g
        res["eager"] = eager
    except Exception as exc:
        res["cuda_graph"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
    if with_cpu and rank == 0:
        from oracle import model as omodel
        ncores = os.cpu_count() or 1
        torch.set_num_threads(ncores)
        nm = 8
        bc = synth.qm9_batch(nm, seed=0)
        dc = {k: (torch.from_numpy(v) if hasattr(v, "shape") else v) for k, v in bc.items()}
        torch.manual_seed(0)
        ref = omodel.XGNNPoly(**HPARAMS)
        ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)

        def cstep():
            ropt.zero_grad(set_to_none=True)
            torch.nn.functional.smooth_l1_loss(ref(dc), torch.zeros(nm)).backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 100.0)
            ropt.step()
        sec = time_cpu(cstep, 2, warmup=1)
        res["cpu_baseline"] = {"molecules_per_sec": nm / sec, "cores": ncores, "kind": "port",
                               "sample": f"{nm}-molecule batch, 1 warm-up + 2 timed training steps of the oracle model"}
    return res


def _train_step_graph(dev, world, rank, steps, data, y, nmine):
    """The training step with forward + loss + backward + all-reduce + clip + Adam + EMA replayed from one
    CUDA graph (x2gnn_b200.train_graph).  Checks the replayed loss against an identical model stepped eagerly."""
    import torch
    import torch.distributed as dist
    from x2gnn_b200 import ddp
    from x2gnn_b200.train_graph import GraphedTrainStep, dense_step
    from x2gnn_b200.xgnn_model import XGNNPoly

    torch.manual_seed(0)
    model = XGNNPoly(**HPARAMS).to(dev)
    bucket = ddp.FlatGradBucket(model.parameters()) if world > 1 else None
    gs = GraphedTrainStep(model, data, y, lr=1e-3, bucket=bucket)
    # high priority: the few small integer kernels (each followed by a size read-back on the host) are scheduled ahead
    # of the replay's queued CTAs instead of behind a whole persistent kernel -- prepare() then takes ~2 ms of host
    # time beside a replay in flight instead of ~8, and the loop is paced by the GPU, not by the read-backs
    prep_stream = torch.cuda.Stream(device=dev, priority=-1)
    prep_stream.wait_stream(torch.cuda.current_stream(dev))

    def step():
        # the per-batch integer work, every step (results identical to gs.prep), on its own stream so that its
        # size read-backs do not drain the replay in flight -- as a loader thread preparing the next batch does
        with torch.cuda.stream(prep_stream):
            p2 = model.prepare(data)
        gs.replay()
        return p2

    for _ in range(3):
        p2 = step()
    torch.cuda.synchronize(dev)
    same_idx = bool(torch.equal(p2["tri"], gs.prep["tri"]))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    # untimed iterations inside the loop (pipeline fill from an idle device), see the eager loop; with NCCL inside
    # the replayed graph the first replays after the barrier are slow for longer (107 / 22 ms at timed positions
    # 0 / 1 with 4 lead-in replays on a 2-GPU box: gpurun_out/r2y_bench_2gpu.json)
    # (N = 1 as well: with 4 lead-in replays the first ~8 TIMED steps still ran 7.6-7.9 ms against 7.07 afterwards,
    # profiles/r2_bench_1gpu.json of build c8e195d0 -- the device idles through the host-side setup before the loop)
    LEAD = 12
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(LEAD + steps + 1)]
    gc.collect()
    gc.disable()
    for i in range(LEAD + steps):
        marks[i].record()
        step()
    gc.enable()
    torch.cuda.current_stream(dev).wait_stream(prep_stream)    # the last step's index work is inside the timing
    marks[-1].record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    marks = marks[LEAD:]
    ms = marks[0].elapsed_time(marks[-1]) / steps
    in_order = [round(marks[i].elapsed_time(marks[i + 1]), 3) for i in range(steps)]
    per_step = sorted(in_order)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    out = {"molecules_per_sec": world * NMOL / (ms * 1e-3), "ms_per_step": ms,
           "ms_per_step_median": per_step[len(per_step) // 2], "steps": steps, "ms_each_step": in_order,
           "graph_loss": float(gs.loss.detach()), "indices_rebuilt_each_step_identical": same_idx,
           "what": "per step: prepare() (triplet / CSR integer kernels, eager, own stream) + one graph replay of "
                   "forward, loss, backward, all-reduce, clip, Adam, EMA"}
    if world == 1:
        # parity of the replayed arithmetic: an identical model stepped eagerly as many times lands on the same
        # loss (the k-th loss is computed before the k-th update)
        torch.manual_seed(0)
        m2 = XGNNPoly(**HPARAMS).to(dev)
        ps2 = [p for p in m2.parameters() if p.requires_grad]
        o2 = torch.optim.Adam(ps2, lr=1e-3, fused=True)
        pr2 = m2.prepare(data)
        n_updates = gs.warmup_updates + 3 + LEAD + steps
        for _ in range(n_updates):
            l2 = dense_step(m2, o2, ps2, data, pr2, y)
        out["eager_loss_after_same_updates"] = float(l2.detach())
        out["updates"] = n_updates
        out["loss_note"] = ("same arithmetic; the remaining PyTorch index ops use float atomics, so two runs of either "
                            "kind agree to the last bit for ~20 updates and to ~1e-3 after 40")
    return out


# ---------------------------------------------------------------------------------- secondary legs (rank 0, N = 1)
def _timed(fn, iters, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def reference_on_gpu_leg(dev, w, iters=5):
    """SURVEY.md 8(d) last row: the reference's composite PyTorch path (the oracle port: gathers, scatter
    softmax, elementwise passes, index_add -- ~35 launches forward, ~70 backward) on the SAME B200 and the same
    tensors as the headline step.  A baseline leg: the checker runs beside the product, never inside it."""
    import torch
    from oracle import conv as oconv
    D, H, S, R, A = (DIMS[k] for k in "DHSRA")
    torch.manual_seed(0)
    ref = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).to(dev)
    t = {k: torch.from_numpy(w[k]).to(dev) for k in ("x", "rbf", "edge_attr", "sbf", "edge_index")}
    for k in ("x", "rbf", "edge_attr"):
        t[k].requires_grad_(True)
    gout = torch.randn(w["E"], D, device=dev, generator=torch.Generator(dev).manual_seed(1))
    params = list(ref.parameters())

    def step():
        out = ref(t["sbf"], t["rbf"], x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"])
        torch.autograd.grad(out, [t["x"], t["rbf"], t["edge_attr"]] + params, gout)
    ms = _timed(step, iters, warmup=2)
    return {"ms_per_step": ms, "value": w["T"] / (ms * 1e-3), "unit": UNIT,
            "what": "composite PyTorch path of the reference (oracle port) on this GPU, same tensors, fp32"}


def parity_full_size_leg(dev, w, layer):
    """Parity AT THE BENCH BATCH'S SIZE (not only on the small test graphs): one forward + backward of the timed
    layer against the fp64 oracle run on this GPU with the same weights and tensors -- output, input gradients
    and every parameter gradient, max |a - b| / max |b|.  (Sums over 8e5 triplets are where a biased accumulation
    shows: the T-row weight gradients were 4e-5 off here before csrc/tc_gemm.cuh cut the tensor-core accumulation
    into periods.)  The checker runs beside the product, never inside it."""
    import torch
    from oracle import conv as oconv
    D, H, S, R, A = (DIMS[k] for k in "DHSRA")
    ref = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).double().to(dev)
    ref.load_state_dict({k: v.detach().double() for k, v in layer.state_dict().items()})
    names = ("x", "rbf", "edge_attr")
    t = {k: torch.from_numpy(w[k]).to(dev) for k in names + ("sbf", "edge_index")}
    gout = torch.randn(w["E"], D, device=dev, generator=torch.Generator(dev).manual_seed(1))
    t64 = {k: t[k].double().requires_grad_(True) for k in names}
    out64 = ref(t["sbf"].double(), t64["rbf"], x=t64["x"], edge_index=t["edge_index"], edge_attr=t64["edge_attr"])
    rp = dict(ref.named_parameters())
    g64 = torch.autograd.grad(out64, [t64[k] for k in names] + list(rp.values()), gout.double())
    want = dict(zip(names + tuple(rp.keys()), g64))
    want["out"] = out64.detach()
    del out64, g64
    t32 = {k: t[k].requires_grad_(True) for k in names}
    mp = dict(layer.named_parameters())
    out = layer(t["sbf"], t32["rbf"], x=t32["x"], edge_index=t["edge_index"], edge_attr=t32["edge_attr"])
    g32 = torch.autograd.grad(out, [t32[k] for k in names] + [mp[k] for k in rp.keys()], gout)
    got = dict(zip(names + tuple(rp.keys()), g32))
    got["out"] = out.detach()
    errs = {}
    for k, b in want.items():
        if k == "lin_key.bias":          # identically zero in exact arithmetic
            continue
        errs[k] = float((got[k].double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    worst = max(errs, key=errs.get)
    return {"max_rel_err": errs[worst], "worst": worst, "tolerance": 1e-5, "within_tolerance": errs[worst] < 1e-5,
            "per_tensor": {k: float(f"{v:.3e}") for k, v in errs.items()},
            "what": "layer fwd+bwd on the bench batch vs the fp64 oracle on this GPU, same weights"}


def factorised_sbf_leg(dev, iters=20):
    """SURVEY.md 8f row 2 (opt-in, sbftransformer_conv.USE_FACTORS / X2GNN_SGF=1): sbf produced by F_B_2D from the
    bench batch's geometry carries its factors (per-bond radial table x Y_l0(angle)); the conv layer then evaluates
    lin_sbf inside its block-centric attention kernels (csrc/blk_attn.cuh) and neither reads the [T, S] tensor nor
    writes lin_sbf(sbf) / its gradient.  Layer forward + backward, CUDA events, same seed-0 batch; with edge_attr
    [T, A] as in the reference and with the segment-constant table.  The dense path on the same F_B_2D tensor is
    timed beside it (the factors are dropped by cloning the tensor)."""
    import torch
    import x2gnn_b200.sbftransformer_conv as sc
    from x2gnn_b200 import synth
    from x2gnn_b200.angular_basis_layer import F_B_2D
    from x2gnn_b200.edge_graph import vertex_to_edge_2
    D, H, S, R, A = (DIMS[k] for k in "DHSRA")
    b = synth.qm9_batch(NMOL, seed=0)
    ei = torch.from_numpy(b["edge_index"]).to(dev)
    N = len(b["x"])
    tri, aj, ai, ak = vertex_to_edge_2(ei, N)
    pos = torch.from_numpy(b["atom_pos"]).to(dev)
    d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
    ji, jk = pos[ai] - pos[aj], pos[ak] - pos[aj]
    ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    sbf = F_B_2D(7, R, 5.0)(d, ang, tri[0])
    E, T = int(ei.size(1)), int(tri.size(1))
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(E, D, device=dev, generator=g).requires_grad_(True)
    rbf = (torch.rand(E, R, device=dev, generator=g) * 2 - 1).requires_grad_(True)
    ea = torch.randn(T, A, device=dev, generator=g).requires_grad_(True)
    tab = torch.randn(N, A, device=dev, generator=g).requires_grad_(True)
    gout = torch.randn(E, D, device=dev, generator=g)
    torch.manual_seed(0)
    layer = sc.SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).to(dev)
    params = list(layer.parameters())
    res = {"E": E, "T": T, "what": "layer fwd+bwd ms per step on the seed-0 batch with sbf = F_B_2D(geometry)"}
    old = sc.USE_FACTORS
    try:
        for name, kw, inp in (("edge_attr_TA", dict(edge_attr=ea), ea),
                              ("edge_attr_table", dict(edge_attr=tab, edge_attr_index=ei[1].contiguous()), tab)):
            for fact in (True, False):
                sc.USE_FACTORS = fact
                s_in = sbf if fact else sbf.clone()
                before = sc.PLAN_COUNTS["factorised"]

                def step():
                    out = layer(s_in, rbf, x=x, edge_index=tri, **kw)
                    torch.autograd.grad(out, [x, rbf, inp] + params, gout)
                ms = _timed(step, iters, warmup=3)
                took = sc.PLAN_COUNTS["factorised"] > before
                res[f"{name}_{'factorised' if fact else 'dense'}_ms"] = round(ms, 4)
                if fact and not took:
                    res[f"{name}_factorised_ms"] = None
    finally:
        sc.USE_FACTORS = old
    return res


def unchanged_reference_callers_leg(dev, iters=5):
    """SURVEY.md 8(d) metric (ii) on the reference's OWN model code: the unmodified xgnn.py / model.py / readout.py /
    residual_layer.py / atom_embedding.py (staged under baseline/_ref by __graft_entry__.build()) imported after
    `x2gnn_b200.install()`, so that only their hot-path imports (SBFTransformerConv, F_B_2D, RadialBasis, poly_envelop,
    vertex_to_edge_2) resolve to the sm_100a drop-in modules and their torch_geometric / torch_scatter imports to
    compat/ (plain PyTorch on the GPU).  One training step as trainer.py:37-48 runs it (forward, SmoothL1, backward,
    clip, Adam), batch 128, eager: edge_attr is the [T, A] gather of xgnn.py:57-58 and edgenn runs on T rows, as in
    the reference."""
    import importlib
    import torch
    import x2gnn_b200
    from x2gnn_b200 import synth
    root = os.path.dirname(os.path.abspath(__file__))
    d = os.path.join(root, "baseline", "_ref")
    if not os.path.exists(os.path.join(d, "xgnn.py")):
        return {"unavailable": "baseline/_ref/xgnn.py not staged"}
    callers = ("xgnn", "model", "readout", "residual_layer", "atom_embedding", "initializer")
    x2gnn_b200.install()
    sys.path.append(d)
    for m in callers:
        sys.modules.pop(m, None)
    try:
        xgnn = importlib.import_module("xgnn")
        from torch_geometric.data import Data
        torch.manual_seed(0)
        net = xgnn.xgnn_poly(**HPARAMS, device=str(dev)).to(dev)
        b = synth.qm9_batch(NMOL, seed=0)
        data = Data(**{k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in b.items()})
        y = torch.zeros(NMOL, device=dev)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.smooth_l1_loss(net(data), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=100.0)
            opt.step()
        ms = _timed(step, iters, warmup=3)
        return {"ms_per_step": ms, "molecules_per_sec": NMOL / (ms * 1e-3),
                "what": "unmodified reference xgnn_poly (baseline/_ref) over the drop-in modules + compat/, eager "
                        "training step (forward, loss, backward, clip, Adam), batch 128, edge_attr [T, A] as in xgnn.py:57-58"}
    finally:
        sys.path.remove(d)
        for m in callers:
            sys.modules.pop(m, None)
        x2gnn_b200.uninstall()


def ocelot_inference_leg(dev, iters=5):
    """BASELINE.json configs[2]: inference throughput on OCELOT-sized molecules -- here the real 60-146-atom
    geometries the reference ships (raw/AID_kcal.xyz, numeric fixture tests/golden/aid_geometries.npz), batches
    of 12, full harness model forward (radius graph is part of the dataset in the reference; triplets, bases
    and the 4 conv layers run per batch)."""
    import torch
    from x2gnn_b200 import edge_graph, synth
    from x2gnn_b200.xgnn_model import XGNNPoly
    torch.manual_seed(0)
    net = XGNNPoly(**HPARAMS).to(dev).eval()
    out = []
    for seg, name in ((False, "edge_attr [T,A] as in the reference"), (True, "segment-constant edge_attr table (opt-in)")):
        net.segment_edge_attr = seg
        mols = tot_ms = tot_T = 0
        for b0 in (0, 12, 24):
            ob = synth.aid_batch(range(b0, b0 + 12), seed=b0)
            data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in ob.items()}
            with torch.no_grad():
                tri = edge_graph.vertex_to_edge_2(data["edge_index"], data["x"].size(0))[0]
                ms = _timed(lambda: net(data), iters, warmup=2)
            mols += 12
            tot_ms += ms
            tot_T += int(tri.size(1))
        out.append({"edge_attr": name, "molecules_per_sec": mols / (tot_ms * 1e-3),
                    "edge_messages_per_sec_4_layers": 4 * tot_T / (tot_ms * 1e-3), "ms_per_batch_of_12": tot_ms / 3,
                    "triplets_per_batch": tot_T // 3})
    return {"geometries": "raw/AID_kcal.xyz molecules 0..35 (60-146 atoms), 3 batches of 12", "runs": out}


def ball500_sweep_leg(dev, peak, iters=8):
    """BASELINE.json configs[3]: conv-layer fwd+bwd on ball-packed 500-atom graphs (radius-cutoff edges, segments of
    13-62 triplets), two sizes; HBM fraction in algorithmic bytes against the measured peak."""
    import torch
    from x2gnn_b200 import edge_graph, synth
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    D, H, S, R, A = (DIMS[k] for k in "DHSRA")
    torch.manual_seed(0)
    layer = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).to(dev)
    params = list(layer.parameters())
    rows = []
    for natoms, ngraphs in ((500, 1), (500, 4)):
        b = synth.ball_batch(ngraphs, n_atoms=natoms, seed=0)
        ei = torch.from_numpy(b["edge_index"]).to(dev)
        tri = edge_graph.vertex_to_edge_2(ei, len(b["x"]))[0]
        E, T = ei.size(1), tri.size(1)
        g = torch.Generator(dev).manual_seed(1)
        x = torch.randn(E, D, device=dev, generator=g).requires_grad_(True)
        rbf = (torch.rand(E, R, device=dev, generator=g) * 2 - 1).requires_grad_(True)
        sbf = torch.randn(T, S, device=dev, generator=g)
        ea = torch.randn(T, A, device=dev, generator=g).requires_grad_(True)
        gout = torch.randn(E, D, device=dev, generator=g)

        def step():
            out = layer(sbf, rbf, x=x, edge_index=tri, edge_attr=ea)
            torch.autograd.grad(out, [x, rbf, ea] + params, gout)
        ms = _timed(step, iters)
        fb, bb = algorithmic_bytes(E, T, D, S, R, A)
        rows.append({"atoms": natoms, "graphs": ngraphs, "E": int(E), "T": int(T), "ms_per_step": ms,
                     "edge_messages_per_sec": T / (ms * 1e-3), "hbm_frac": (fb + bb) / (ms * 1e-3) / 1e9 / peak})
        del x, rbf, sbf, ea, gout
    return rows


def e2e_from_atoms_leg(dev, world, rank, steps):
    """The path the north star names, end to end from host memory: pinned host atoms (positions, atomic numbers,
    molecule ids) and pair features edge_attr[E, 338] -> device -> radius graph -> triplets -> bases -> embeddings
    -> 4 SBFTransformerConv layers (+ the model's LayerNorm / residual / readout blocks) forward, loss, backward
    -> loss and predictions back on the host.  One training-style step of the harness model per batch of 128
    molecules per GPU; ~60 MB cross PCIe per step instead of the 588 MB of the conv-boundary `e2e` (sbf[T,42]
    and edge_attr[T,128] are produced on the device here, as in the real model).  Copies of step i+1 overlap
    step i on a copy stream (two buffer sets)."""
    import torch
    import torch.distributed as dist
    from x2gnn_b200 import atom_graph, synth
    from x2gnn_b200.xgnn_model import XGNNPoly
    b, _ = synth.qm9_shard(NMOL, world, rank, seed=0)
    pin = {k: torch.from_numpy(b[k]).pin_memory() for k in ("atom_pos", "x", "batch", "edge_attr", "edge_num", "y")}
    B = int(b["num_graphs"])
    res = {}
    for seg in (False, True):
        torch.manual_seed(0)
        net = XGNNPoly(**HPARAMS).to(dev)
        net.segment_edge_attr = seg
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in pin.items()} for _ in range(2)]
        h_pred, h_loss = torch.empty(B).pin_memory(), torch.empty(()).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        cnt = [0]
        stat = {}

        def stage(slot):
            with torch.cuda.stream(copy_stream), torch.no_grad():
                copy_stream.wait_event(consumed[slot])
                for k in pin:
                    bufs[slot][k].copy_(pin[k], non_blocking=True)
                ready[slot].record(copy_stream)

        def step():
            i = cnt[0]
            cnt[0] += 1
            slot = i & 1
            stage(slot ^ 1)
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[slot])
            t = bufs[slot]
            ei, _ = atom_graph.radius_graph(t["atom_pos"], t["batch"], 5.0)        # atom_graph.py:32-45 on the device
            data = {"x": t["x"], "atom_pos": t["atom_pos"], "edge_index": ei, "edge_attr": t["edge_attr"],
                    "edge_num": t["edge_num"], "batch": t["batch"], "num_graphs": B}
            net.zero_grad(set_to_none=True)
            pred = net(data)
            loss = torch.nn.functional.smooth_l1_loss(pred, t["y"])
            loss.backward()
            consumed[slot].record(cur)
            h_pred.copy_(pred.detach(), non_blocking=True)
            h_loss.copy_(loss.detach(), non_blocking=True)
            cur.synchronize()
            stat["E"] = int(ei.size(1))

        for ev in consumed:
            ev.record()
        stage(0)
        for _ in range(3):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        T = int(synth.triplets_host(b["edge_index"], len(b["x"]))[0].shape[1])
        tt = torch.tensor([ms, float(T)], device=dev, dtype=torch.float64)
        if world > 1:
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            ms, T_all = float(mx[0]), float(tt[1])
        else:
            T_all = float(T)
        key = "segment_constant_edge_attr_table" if seg else "edge_attr_TA_as_in_the_reference"
        res[key] = {"ms_per_step": ms, "molecules_per_sec": world * NMOL / (ms * 1e-3),
                    "edge_messages_per_sec_4_layers": HPARAMS["conv_layers"] * T_all / (ms * 1e-3),
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pin.values()),
                    "d2h_bytes_per_step": h_pred.numel() * 4 + 4}
        del net, bufs
    # ---- the same path through the repo's graphed training step (train_graph.GraphedTrainStep: forward, loss,
    # backward, all-reduce, clip, Adam, EMA replayed as one CUDA graph).  Per step: pinned host -> the graph's
    # static device buffers, radius graph + triplets + CSR metadata rebuilt eagerly from the copied atoms, one
    # replay, predictions-free loss back to the host.  One buffer set: copy and compute do not overlap here.
    try:
        from x2gnn_b200.train_graph import GraphedTrainStep
        from x2gnn_b200 import ddp
        torch.manual_seed(0)
        net = XGNNPoly(**HPARAMS).to(dev)
        sbuf = {k: v.to(dev) for k, v in pin.items()}
        ei0, _ = atom_graph.radius_graph(sbuf["atom_pos"], sbuf["batch"], 5.0)
        data = {"x": sbuf["x"], "atom_pos": sbuf["atom_pos"], "edge_index": ei0, "edge_attr": sbuf["edge_attr"],
                "edge_num": sbuf["edge_num"], "batch": sbuf["batch"], "num_graphs": B}
        bucket = ddp.FlatGradBucket(net.parameters()) if world > 1 else None
        gs = GraphedTrainStep(net, data, sbuf["y"], lr=1e-3, bucket=bucket)
        h_loss = torch.empty(()).pin_memory()
        same = [True]

        def gstep():
            with torch.no_grad():
                for k in pin:
                    sbuf[k].copy_(pin[k], non_blocking=True)
                ei, _ = atom_graph.radius_graph(sbuf["atom_pos"], sbuf["batch"], 5.0)
                p2 = net.prepare(dict(data, edge_index=ei))          # the integer work of this batch, every step
                same[0] = same[0] and ei.shape == ei0.shape and p2["tri"].shape == gs.prep["tri"].shape
                ei0.copy_(ei)
            gs.replay()
            h_loss.copy_(gs.loss.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(3):
            gstep()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            gstep()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            mx = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            ms = float(mx[0])
        res["graphed_training_step"] = {
            "ms_per_step": ms, "molecules_per_sec": world * NMOL / (ms * 1e-3), "loss": float(h_loss),
            "same_layout_every_step": same[0],
            "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pin.values()), "d2h_bytes_per_step": 4,
            "what": "full training step (segment-constant edge_attr table, the harness model's default): H2D into the "
                    "graph's static buffers, radius graph + triplets + metadata eager, one CUDA-graph replay incl. "
                    "optimizer, loss to the host"}
        # ---- the same step fed by the device-resident dataset (collate.DeviceDataset: the whole dataset uploaded
        # once, a batch is a gather on the device -- x2_collate_sizes / x2_collate_fill -- instead of a host
        # collation + a 58 MB copy); the batch is the same 128 molecules, so the captured graph applies
        try:
            import numpy as np
            from x2gnn_b200.collate import DeviceDataset
            an = np.concatenate([[0], np.cumsum(np.bincount(b["batch"], minlength=B))])
            en = np.concatenate([[0], np.cumsum(b["edge_num"])])
            mols = [{"x": b["x"][an[g]:an[g + 1]], "atom_pos": b["atom_pos"][an[g]:an[g + 1]],
                     "edge_index": b["edge_index"][:, en[g]:en[g + 1]] - an[g],
                     "edge_attr": b["edge_attr"][en[g]:en[g + 1]], "y": 0.0} for g in range(B)]
            ds = DeviceDataset.from_molecules(mols, dev)
            ids = torch.arange(B, device=dev)

            def dstep():
                with torch.no_grad():
                    rec = ds.collate(ids)
                    for k in ("x", "atom_pos", "batch", "edge_attr", "edge_num"):
                        sbuf[k].copy_(rec[k])
                    ei0.copy_(rec["edge_index"])
                    net.prepare(dict(data, edge_index=rec["edge_index"]))
                gs.replay()
                h_loss.copy_(gs.loss.detach(), non_blocking=True)
                torch.cuda.current_stream().synchronize()

            for _ in range(3):
                dstep()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                dstep()
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            if world > 1:
                mx = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                ms = float(mx[0])
            res["graphed_training_step_device_dataset"] = {
                "ms_per_step": ms, "molecules_per_sec": world * NMOL / (ms * 1e-3), "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 4,
                "what": "as graphed_training_step, the batch gathered on the device from the resident dataset "
                        "(device-side collation) instead of collated on the host and copied"}
        except Exception as exc:
            res["graphed_training_step_device_dataset"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        del net, gs
    except Exception as exc:
        res["graphed_training_step"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
    res["what"] = ("pinned host atoms + pair features -> device -> radius graph -> triplets -> bases -> 4 conv layers "
                   "(full harness model) forward + loss + backward -> predictions and loss on the host, per step")
    res["value"] = res["edge_attr_TA_as_in_the_reference"]["edge_messages_per_sec_4_layers"]
    res["unit"] = UNIT
    return res


def dp_gradient_check(dev, world, rank, layer, params):
    """SURVEY.md 4 item 5 on the hardware: every rank's shard gradients, all-reduced (sum), against the SAME global
    batch run unsharded on rank 0 (every rank can regenerate every shard: the shards are seeded).  Returns the
    largest relative error over the parameter tensors (max |a - b| / max |b|)."""
    import torch
    import torch.distributed as dist
    D = DIMS["D"]

    def grads_of(w, seed):
        t = {k: torch.from_numpy(w[k]).to(dev) for k in ("x", "rbf", "edge_attr", "sbf", "edge_index")}
        gout = torch.randn(w["E"], D, device=dev, generator=torch.Generator(dev).manual_seed(100 + seed))
        out = layer(t["sbf"], t["rbf"], x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"])
        return torch.autograd.grad(out, params, gout), t, gout

    g_mine, _, _ = grads_of(host_workload(rank, world), rank)
    flat = torch.cat([g.reshape(-1) for g in g_mine])
    dist.all_reduce(flat)
    err = None
    if rank == 0:
        ws = [host_workload(r, world) for r in range(world)]
        import numpy as np
        eoff = np.cumsum([0] + [w["E"] for w in ws])
        cat = {k: np.concatenate([w[k] for w in ws], axis=0) for k in ("x", "rbf", "edge_attr", "sbf")}
        cat["edge_index"] = np.concatenate([w["edge_index"] + eoff[r] for r, w in enumerate(ws)], axis=1)
        t = {k: torch.from_numpy(v).to(dev) for k, v in cat.items()}
        gout = torch.cat([torch.randn(w["E"], D, device=dev, generator=torch.Generator(dev).manual_seed(100 + r))
                          for r, w in enumerate(ws)])
        out = layer(t["sbf"], t["rbf"], x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"])
        g_all = torch.autograd.grad(out, params, gout)
        err, off = 0.0, 0
        top = max(float(g.abs().max()) for g in g_all)
        for g in g_all:
            n = g.numel()
            a, bref = flat[off:off + n].double(), g.reshape(-1).double()
            off += n
            scale = float(bref.abs().max())
            if scale > 1e-6 * top:        # lin_key.bias: its gradient is identically zero (App. A), only rounding noise
                err = max(err, float((a - bref).abs().max()) / scale)
        del t, gout, out, g_all
    dist.barrier()
    torch.cuda.empty_cache()
    return err


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from x2gnn_b200 import _lib, graph_meta
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = args.steps, max(args.warmup, 3)

    w = host_workload(rank, world)
    E, T = w["E"], w["T"]
    D, H, S, R, A = (DIMS[k] for k in "DHSRA")
    torch.manual_seed(0)
    layer = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, dropout=0, edge_dim=A).to(dev)
    layer.precision = {"fp32": 0, "tf32x3": 1, "tf32x3_fused": 2, "tf32": 3}[args.mode]
    params = list(layer.parameters())
    pin = {k: torch.from_numpy(w[k]).pin_memory() for k in ("x", "rbf", "sbf", "edge_attr", "edge_index")}
    x = pin["x"].to(dev).requires_grad_(True)
    rbf = pin["rbf"].to(dev).requires_grad_(True)
    ea = pin["edge_attr"].to(dev).requires_grad_(True)
    sbf = pin["sbf"].to(dev)
    ei = pin["edge_index"].to(dev)
    gout = torch.randn(E, D, device=dev, generator=torch.Generator(dev).manual_seed(1))
    flat = torch.empty(sum(p.numel() for p in params), device=dev)

    def step():
        out = layer(sbf, rbf, x=x, edge_index=ei, edge_attr=ea)
        grads = torch.autograd.grad(out, [x, rbf, ea] + params, gout)
        if world > 1:      # data-parallel training: one flat all-reduce of the parameter gradients
            torch.cat([g.reshape(-1) for g in grads[3:]], out=flat)
            dist.all_reduce(flat)
        return out, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------ N > 1: sharded gradient == unsharded gradient, on the hardware
    dp_err = None
    if world > 1 and not args.layer_only:
        dp_err = dp_gradient_check(dev, world, rank, layer, params)
        if rank == 0 and not (dp_err <= 1e-5):
            raise SystemExit(f"bench.py: all-reduced shard gradients differ from the unsharded global batch: {dp_err}")

    # ------------------------------------------------ resident-input throughput (`value`)
    for _ in range(warmup):
        step()
    barrier()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        barrier()
    launches = (_lib.launch_count() - n0) // max(steps, 1)
    ms = ev0.elapsed_time(ev1)
    local_ms_per_step = ms / steps       # this rank's device time per step (CUDA events around the timed region)
    # per-phase CUDA-event times of the library in a SEPARATE pass: the event records sit between the
    # launches (they would cost the timed region ~2 us each and keep a launch from overlapping the tail
    # of its predecessor)
    _lib.timing_read()
    _lib.timing_enable(rank == 0)
    for _ in range(steps):           # every rank steps (the step holds a collective when N > 1)
        step()
    barrier()
    _lib.timing_enable(False)
    phases = _lib.timing_read() if rank == 0 else {}
    t_all = torch.tensor([ms, float(T)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = t_all.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_all, op=dist.ReduceOp.SUM)
        ms, total_T = float(tmax[0]), float(t_all[1])
    else:
        total_T = float(T)
    ms_per_step = ms / steps
    value = total_T / (ms_per_step * 1e-3)

    # ------------------------------------------------ opt-in fast path: segment-constant edge_attr table
    # (SURVEY.md §8f row 1).  Same layer, same graph; edge_attr is a per-atom table [N, A] indexed by the
    # central atom of the target bond instead of a [T, A] stream.  Reported beside the headline, never as it.
    seg = None
    if world == 1 and not args.layer_only:
        N = w["N"]
        tab = torch.randn(N, A, device=dev, generator=torch.Generator(dev).manual_seed(2)).requires_grad_(True)
        idx = torch.from_numpy(w["center"]).to(dev)

        def seg_step():
            out = layer(sbf, rbf, x=x, edge_index=ei, edge_attr=tab, edge_attr_index=idx)
            return torch.autograd.grad(out, [x, rbf, tab] + params, gout)

        for _ in range(warmup):
            seg_step()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            seg_step()
        ev1.record()
        torch.cuda.synchronize()
        seg_ms = ev0.elapsed_time(ev1) / steps
        _lib.timing_read()
        _lib.timing_enable(True)             # phases in a separate pass, as for the headline
        for _ in range(steps):
            seg_step()
        torch.cuda.synchronize()
        _lib.timing_enable(False)
        seg_phases = {k: round(v[0] / steps, 4) for k, v in _lib.timing_read().items()}
        seg_bytes = (4 * (E * (D + R) + T * S + N * A + E * D) + 16 * T + 8 * E            # fwd
                     + 4 * (E * D + E * (D + R) + T * S + N * A) + 16 * T + 8 * E           # bwd reads
                     + 4 * (E * (D + R) + N * A))                                           # bwd writes
        seg = {"value": T / (seg_ms * 1e-3), "unit": UNIT, "ms_per_step": seg_ms,
               "edge_attr": f"table [{N}, {A}] + edge_attr_index [{E}] (central atom of the target bond)",
               "algorithmic_bytes_per_step": seg_bytes, "phase_ms_per_step": seg_phases}

    if args.layer_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                              "ms_per_step": ms_per_step, "mode": args.mode, "layer_only": True,
                              "gpu_launches_per_step": int(launches), "config": {"E": E, "T": T}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------ end to end through the module call (`e2e`)
    # Every step copies ITS inputs host -> device and its results device -> host.  Like a training
    # input pipeline, the copies of step i+1 run on a copy stream while step i computes (two device
    # buffer sets); nothing is reused between steps.
    names = ("x", "rbf", "edge_attr", "sbf", "edge_index")
    bufs = [{k: torch.empty_like(pin[k], device=dev) for k in names} for _ in range(2)]
    for bset in bufs:
        for k in ("x", "rbf", "edge_attr"):
            bset[k].requires_grad_(True)
    h_out = torch.empty(E, D).pin_memory()
    h_dx = torch.empty(E, D).pin_memory()
    h_drbf = torch.empty(E, R).pin_memory()
    h_flat = torch.empty(flat.numel()).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]       # inputs of the set have landed
    consumed = [torch.cuda.Event(), torch.cuda.Event()]    # compute is done with the set
    e2e_i = [0]

    def stage_inputs(slot):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(consumed[slot])
            for k in names:                                 # edge_index copy bumps _version => metadata is rebuilt
                bufs[slot][k].copy_(pin[k], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        i = e2e_i[0]
        e2e_i[0] += 1
        slot = i & 1
        stage_inputs(slot ^ 1)               # next step's inputs, overlapping this step's compute
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[slot])
        bset = bufs[slot]
        out = layer(bset["sbf"], bset["rbf"], x=bset["x"], edge_index=bset["edge_index"], edge_attr=bset["edge_attr"])
        grads = torch.autograd.grad(out, [bset["x"], bset["rbf"], bset["edge_attr"]] + params, gout)
        torch.cat([g.reshape(-1) for g in grads[3:]], out=flat)
        if world > 1:
            dist.all_reduce(flat)
        consumed[slot].record(cur)
        h_out.copy_(out.detach(), non_blocking=True)
        h_dx.copy_(grads[0], non_blocking=True)
        h_drbf.copy_(grads[1], non_blocking=True)
        h_flat.copy_(flat, non_blocking=True)
        cur.synchronize()                    # the host owns this step's results before the next step

    for ev in consumed:
        ev.record()
    stage_inputs(0)                          # prologue: the first step's inputs (outside the timed region;
                                             # the last timed step stages one extra set in exchange)

    h2d = sum(pin[k].numel() * pin[k].element_size() for k in pin)
    d2h = sum(t.numel() * t.element_size() for t in (h_out, h_dx, h_drbf, h_flat))
    e2e_steps = max(3, min(steps, 30))
    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    barrier()
    e_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t[0])
    e2e_value = total_T / (e_ms / e2e_steps * 1e-3)

    # ------------------------------------------------ end to end from atoms (the path the north star names), all ranks
    try:
        from_atoms = e2e_from_atoms_leg(dev, world, rank, max(3, min(steps, 10)))
    except Exception as exc:
        from_atoms = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}

    # ------------------------------------------------ full training step (molecules/s), all ranks
    train = None
    if not args.no_train_step:
        try:
            train = train_step_bench(dev, world, rank, max(3, min(steps, 30)), 5,
                                     with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as exc:  # keep the headline line even if the secondary metric fails
            train = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------ roofline of the layer's kernels (rank 0)
    fwd_b, bwd_b = algorithmic_bytes(E, T, D, S, R, A)
    # the step's kernels: CUDA-event time of the timed region on this rank (N > 1: includes packing + all-reduce);
    # the phase split comes from the separate pass, whose events between launches cost a little overlap
    kernel_ms = local_ms_per_step
    phase_ms = {k: round(v[0] / max(steps, 1), 4) for k, v in phases.items()}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = (fwd_b + bwd_b) / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    # DRAM traffic of the same step: ncu launch list of THIS library build (tools/step_traffic.py writes the file
    # with the digest of the sources the .so was built from; a file of another build is refused, not reused)
    traffic, traffic_src, phase_rates = None, None, None
    tpath = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
    # (the stamp written next to the binary by build.py, libx2gnn.so.digest, wins over the tracked libx2gnn.sha256:
    # a `git checkout` can restore the latter without rebuilding the former)
    lib_digest = None
    for dpath in (os.path.join(ROOT, "x2-gnn_b200", "lib", "libx2gnn.so.digest"),
                  os.path.join(ROOT, "x2-gnn_b200", "lib", "libx2gnn.sha256")):
        if os.path.exists(dpath):
            lib_digest = open(dpath).read().strip()
            break
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(args.mode)
        if not tj:
            traffic_src = f"no ncu capture for mode {args.mode} in profiles/r2_step_traffic.json"
        elif tj.get("lib_sha256") != lib_digest:
            traffic_src = ("stale: profiles/r2_step_traffic.json was captured with another build of libx2gnn.so "
                           "(re-run tools/step_traffic.py)")
        elif tj.get("E") != E or tj.get("T") != T:
            traffic_src = "profiles/r2_step_traffic.json is for another workload"
        else:
            traffic, traffic_src = tj["dram_bytes_per_step"], tj["source"]
            if "phase_dram_MB" in tj:
                phase_rates = {k: {"dram_MB": round(mb, 1), "ms": phase_ms[k],
                                   "dram_GBs": round(mb / phase_ms[k], 1) if phase_ms.get(k) else None,
                                   "frac_of_peak": round(mb / phase_ms[k] / peak, 3) if phase_ms.get(k) else None}
                               for k, mb in tj["phase_dram_MB"].items() if k in phase_ms}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "all launches of x2_sbfconv_fwd + x2_sbfconv_bwd (one layer step; CUDA events around the timed steps)",
                "algorithmic_bytes_per_step": fwd_b + bwd_b, "kernel_ms_per_step": kernel_ms,
                "phase_ms_per_step": phase_ms,
                # how close each phase runs to the memory system with the bytes it actually moves (the gap
                # between `achieved` and these rates is traffic that is not algorithmic: materialised
                # per-triplet intermediates)
                "phase_dram_rates": phase_rates}

    # ------------------------------------------------ CPU baseline (N = 1 only, bounded)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        sec = time_cpu(oracle_step_fn(w, ncores), 3, warmup=1)
        cpu = {"value": T / sec, "unit": UNIT, "cores": ncores, "kind": "port",
               "sample": f"full seed-0 batch (E={E}, T={T}), 1 warm-up + 3 timed fwd+bwd steps, "
                         f"oracle port of the reference's composite PyTorch path, fp32"}

    extra = {}
    if world == 1:
        for key, fn in (("parity_full_size", lambda: parity_full_size_leg(dev, w, layer)),
                        ("reference_on_gpu", lambda: reference_on_gpu_leg(dev, w)),
                        ("factorised_sbf", lambda: factorised_sbf_leg(dev)),
                        ("unchanged_reference_callers", lambda: unchanged_reference_callers_leg(dev)),
                        ("ocelot_inference", lambda: ocelot_inference_leg(dev)),
                        ("ball500_sweep", lambda: ball500_sweep_leg(dev, peak))):
            try:
                torch.cuda.empty_cache()
                extra[key] = fn()
            except Exception as exc:
                extra[key] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        if isinstance(extra.get("reference_on_gpu"), dict) and "ms_per_step" in extra["reference_on_gpu"]:
            extra["reference_on_gpu"]["x2gnn_b200_ms_per_step"] = ms_per_step
            extra["reference_on_gpu"]["speedup"] = extra["reference_on_gpu"]["ms_per_step"] / ms_per_step

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if args.mode == "tf32" else "fp32", "data": "synthetic",
        "config": {"workload": "qm9_b128_sbfconv_layer_fwd_bwd (BASELINE.json configs[1])", **DIMS,
                   "molecules_per_gpu": NMOL, "E": E, "T": T, "precision_mode": ("fp32 SIMT (1e-5 parity)" if args.mode == "fp32" else
                                      "REDUCED PRECISION: one tf32 pass in the Linear layers (2e-2 class)" if args.mode == "tf32" else
                                      "fp32 I/O, Linear layers on tcgen05 in 3xTF32 split precision (1e-5 parity)"),
                   "parallelism": f"dp{world} ({NMOL} whole molecules per rank, every rank's batch within 0.5 % of the "
                                  f"N = 1 batch's triplet count; NCCL all-reduce of the flat parameter gradient)",
                   "l2": f"no flush: per-step T-row inputs {680 * T / 1e6:.0f} MB > 126 MB L2"},
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e_ms / e2e_steps},
        "gpu_launches": int(launches) * steps, "gpu_launches_per_step": int(launches),
        "clocks": clocks.summary(),
        "train_step": train,
        "segment_constant_edge_attr": seg,
        "e2e_from_atoms": from_atoms,
        "dp_grad_max_rel_err": dp_err,
        "triplets_per_sec": value,
        **extra,
    }
    if seg is not None:
        seg["roofline_frac"] = seg["algorithmic_bytes_per_step"] / (seg["ms_per_step"] * 1e-3) / 1e9 / peak
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the secondary molecules/s measurement")
    ap.add_argument("--layer-only", action="store_true",
                    help="only the headline layer step (for ncu launch lists: tools/step_traffic.py)")
    ap.add_argument("--mode", default="tf32x3", choices=["fp32", "tf32x3", "tf32x3_fused", "tf32"],
                    help="fp32: SIMT GEMMs; tf32x3: tcgen05 3xTF32 GEMMs (both meet the 1e-5 parity bar); "
                         "tf32: one tf32 pass (reduced precision, 2e-2 class -- never the headline)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
