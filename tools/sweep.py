"""Roofline sweep of the SBFTransformerConv layer (BASELINE.json configs[3]): ball-packed 500-atom graphs,
T from ~0.1 M to ~4 M triplets, fwd+bwd, CUDA-event timing, HBM fraction against MEASURED_PEAKS.json.

    python tools/sweep.py            # prints one JSON line per size
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x2gnn_b200 import edge_graph, synth                                  # noqa: E402
from x2gnn_b200.sbftransformer_conv import SBFTransformerConv              # noqa: E402

D, H, S, R, A = 128, 16, 42, 6, 128
peak = 6650.0
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])
dev = torch.device("cuda")
torch.manual_seed(0)
layer = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).to(dev)
params = list(layer.parameters())
for natoms, ngraphs in ((120, 1), (200, 1), (320, 1), (500, 1), (500, 3), (500, 5)):
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
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    alg = 4 * (E * (D + R) + T * (S + A) + E * D) + 16 * T + 4 * (E * D + E * (D + R) + T * (S + A)) + 16 * T + 4 * (E * (D + R) + T * A)
    print(json.dumps({"atoms": natoms, "graphs": ngraphs, "E": E, "T": T, "max_segment": int(torch.bincount(tri[1]).max()),
                      "ms_per_step": round(ms, 4), "edge_messages_per_sec": T / (ms * 1e-3),
                      "hbm_frac_of_measured": alg / (ms * 1e-3) / 1e9 / peak}), flush=True)
