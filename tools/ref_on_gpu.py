"""The "reference-on-GPU" bar of SURVEY.md 8(d): the reference's composite PyTorch path (the oracle port:
gather / scatter softmax / elementwise / index_add, ~35 launches forward and ~70 backward per layer) timed on
the B200 itself for the bench workload, next to this package's kernels on the same tensors.

    python tools/ref_on_gpu.py [steps]        -> one JSON line

Checker code (oracle/) on the device is a baseline here, never a product path."""
import json
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
import bench                                    # noqa: E402
from oracle import conv as oconv                # noqa: E402
from x2gnn_b200.sbftransformer_conv import SBFTransformerConv      # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
w = bench.host_workload()
D, H, S, R, A = (bench.DIMS[k] for k in "DHSRA")
torch.manual_seed(0)
ref = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A)
mine = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A)
mine.load_state_dict(ref.state_dict())
ref, mine = ref.to(dev), mine.to(dev)
t = {k: torch.from_numpy(w[k]).to(dev) for k in ("x", "rbf", "edge_attr", "sbf", "edge_index")}
for k in ("x", "rbf", "edge_attr"):
    t[k].requires_grad_(True)
gout = torch.randn(w["E"], D, generator=torch.Generator().manual_seed(1)).to(dev)


def make(layer):
    params = list(layer.parameters())

    def step():
        out = layer(t["sbf"], t["rbf"], x=t["x"], edge_index=t["edge_index"], edge_attr=t["edge_attr"])
        return out, torch.autograd.grad(out, [t["x"], t["rbf"], t["edge_attr"]] + params, gout)
    return step


def time_it(step):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


s_ref, s_mine = make(ref), make(mine)
o_ref, g_ref = s_ref()
o_mine, g_mine = s_mine()
err = lambda a, b: float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max())
ms_ref, ms_mine = time_it(s_ref), time_it(s_mine)
torch.cuda.reset_peak_memory_stats()
s_ref()
mem_ref = torch.cuda.max_memory_allocated()
print(json.dumps({
    "workload": "qm9_b128_sbfconv_layer_fwd_bwd", "E": w["E"], "T": w["T"], "steps": steps,
    "composite_pytorch_on_b200_ms": round(ms_ref, 3), "composite_edge_messages_per_sec": w["T"] / (ms_ref * 1e-3),
    "x2gnn_b200_ms": round(ms_mine, 3), "x2gnn_b200_edge_messages_per_sec": w["T"] / (ms_mine * 1e-3),
    "speedup": round(ms_ref / ms_mine, 2), "composite_peak_bytes": mem_ref,
    "fp32_vs_fp32_max_rel_diff": {"out": err(o_mine, o_ref), "dx": err(g_mine[0], g_ref[0]),
                                  "dedge_attr": err(g_mine[2], g_ref[2])}}))
