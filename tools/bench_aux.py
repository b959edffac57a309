"""Throughput of the basis / index kernels (SURVEY.md §8d "Basis / index kernels") and of the
OCELOT-sized inference configuration (BASELINE.json configs[2]) on one B200.

    python tools/bench_aux.py        # prints one JSON line per measurement

GB/s = algorithmic bytes of §8d ÷ CUDA-event time of the public call (module / function of the
drop-in API), inputs resident in HBM.  The triplet / radius-graph builders contain one host sync each
(the output size), so their figures are end-to-end call times, not kernel times.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from x2gnn_b200 import atom_graph, edge_graph, graph_meta, synth            # noqa: E402
from x2gnn_b200.angular_basis_layer import F_B_2D                             # noqa: E402
from x2gnn_b200.radial_basis_layer import RadialBasis                         # noqa: E402
from x2gnn_b200.xgnn_model import XGNNPoly                                    # noqa: E402

dev = torch.device("cuda")
peak = 6650.0
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])


def timed(fn, iters=20, warmup=3):
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


def emit(name, ms, nbytes, **kw):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "GB_per_s": round(gbs, 1),
                      "hbm_frac_of_measured": round(gbs / peak, 4), **kw}), flush=True)


def main():
    # ---- QM9 batch 128 (the bench workload): graph + basis kernels
    b = synth.qm9_batch(128, seed=0)
    pos = torch.from_numpy(b["atom_pos"]).to(dev)
    batch = torch.from_numpy(b["batch"]).to(dev)
    N = pos.size(0)
    ei, _ = atom_graph.radius_graph(pos, batch, 5.0)
    E = ei.size(1)
    tri, aj, ai, ak = edge_graph.vertex_to_edge_2(ei, N)
    T = tri.size(1)
    d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
    ji, jk = pos[ai] - pos[aj], pos[ak] - pos[aj]
    ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    shape = dict(N=N, E=E, T=T)

    ms = timed(lambda: atom_graph.radius_graph(pos, batch, 5.0))
    emit("radius_graph (x2_radius_graph_count/fill + 1 host sync)", ms, 12 * N + 16 * E, **shape)
    ms = timed(lambda: edge_graph.vertex_to_edge_2(ei, N))
    emit("vertex_to_edge_2 (x2_triplets_count/fill + 1 host sync)", ms, 16 * E + 40 * T, **shape)

    def meta():
        graph_meta.clear_cache()
        graph_meta.build(tri, E)
    ms = timed(meta)
    emit("line-graph metadata (x2_meta_build + 1 host sync)", ms, 16 * T + 4 * (4 * T + 2 * E), **shape)

    sbf_layer = F_B_2D(7, 6, 5.0, 5)
    ms = timed(lambda: sbf_layer(d, ang, tri[0]))
    emit("F_B_2D.forward (x2_sbf_table + x2_sbf_fwd)", ms, 4 * T * (42 + 1) + 8 * T, **shape)
    rb = RadialBasis(6, 5.0).to(dev)
    with torch.no_grad():
        ms = timed(lambda: rb(d))
    emit("RadialBasis.forward (x2_radial_fwd)", ms, 4 * E * (6 + 1), **shape)

    # ---- OCELOT-sized molecules (60-146 atoms, mean ~84): inference throughput, batches of 12
    rng = np.random.default_rng(0)
    sizes = rng.integers(60, 147, size=12)
    mols = [synth.synth_mol(int(n), np.random.default_rng(100 + i)) for i, n in enumerate(sizes)]
    ob = synth.collate(mols, seed=0)
    data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in ob.items()}
    hp = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
    torch.manual_seed(0)
    net = XGNNPoly(**hp).to(dev).eval()
    with torch.no_grad():
        net(data)
        tri_o = edge_graph.vertex_to_edge_2(data["edge_index"], data["x"].size(0))[0]
        ms = timed(lambda: net(data), iters=10)
    print(json.dumps({"config": "OCELOT-sized inference (12 molecules of 60-146 atoms, full model forward, 4 conv layers)",
                      "atoms": int(data["x"].size(0)), "E": int(data["edge_index"].size(1)), "T": int(tri_o.size(1)),
                      "ms_per_batch": round(ms, 3), "molecules_per_sec": round(12 / (ms * 1e-3), 1),
                      "edge_messages_per_sec_4_layers": round(4 * tri_o.size(1) / (ms * 1e-3), 1)}), flush=True)


if __name__ == "__main__":
    main()
