"""A/B probe of the block-centric kernels on the bench workload (QM9 batch 128): layer forward + backward, CUDA
events + the library's phase events, for
  dense        generic two-kernel backward, dense sbf
  blocks       one-kernel block backward, dense sbf
  fact         factorised lin_sbf (sbf = F_B_2D output with its factors), edge_attr [T, A]
  fact_table   factorised lin_sbf + segment-constant edge_attr table
  table        segment-constant edge_attr table, dense sbf (blocks on)
usage: python tools/blk_probe.py [nmol] [iters] [variants,comma,separated]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x2gnn_b200
from x2gnn_b200 import synth, _lib
import x2gnn_b200.sbftransformer_conv as sc
from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
from x2gnn_b200.angular_basis_layer import F_B_2D
from x2gnn_b200.edge_graph import vertex_to_edge_2


def main():
    nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    which = sys.argv[3].split(",") if len(sys.argv) > 3 else ["dense", "blocks", "fact", "fact_table", "table"]
    D, H, S, R, A = 128, 16, 42, 6, 128
    b = synth.qm9_batch(nmol, seed=0)
    ei = torch.from_numpy(b["edge_index"]).cuda()
    N = len(b["x"])
    tri, aj, ai, ak = vertex_to_edge_2(ei, N)
    pos = torch.from_numpy(b["atom_pos"]).cuda()
    d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
    ji, jk = pos[ai] - pos[aj], pos[ak] - pos[aj]
    ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    sbf_f = F_B_2D(7, 6, 5.0)(d, ang, tri[0])
    E, T = ei.size(1), tri.size(1)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(E, D, device="cuda", generator=g)
    rbf = torch.rand(E, R, device="cuda", generator=g) * 2 - 1
    ea = torch.randn(T, A, device="cuda", generator=g)
    sbf_d = torch.randn(T, S, device="cuda", generator=g)
    tab = torch.randn(N, A, device="cuda", generator=g)
    idx = torch.zeros(E, dtype=torch.int64, device="cuda")
    idx[tri[1]] = aj
    gout = torch.randn(E, D, device="cuda", generator=g)
    torch.manual_seed(0)
    conv = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).cuda()
    conv.precision = 1

    def step(kind):
        x_ = x.requires_grad_(True)
        if kind in ("dense", "blocks"):
            sc.USE_BLOCKS = kind == "blocks"
            out = conv(sbf_d, rbf, x=x_, edge_index=tri, edge_attr=ea)
        elif kind == "fact":
            out = conv(sbf_f, rbf, x=x_, edge_index=tri, edge_attr=ea)
        elif kind == "fact_table":
            out = conv(sbf_f, rbf, x=x_, edge_index=tri, edge_attr=tab, edge_attr_index=idx)
        else:
            out = conv(sbf_d, rbf, x=x_, edge_index=tri, edge_attr=tab, edge_attr_index=idx)
        out.backward(gout)
        sc.USE_BLOCKS = True

    for kind in which:
        for _ in range(5):
            step(kind)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(iters):
            step(kind)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / iters
        _lib.timing_read()
        _lib.timing_enable(True)
        for _ in range(iters):
            step(kind)
        torch.cuda.synchronize()
        _lib.timing_enable(False)
        ph = {k: round(v[0] / iters, 4) for k, v in _lib.timing_read().items()}
        print(json.dumps({"kind": kind, "E": E, "T": T, "ms": round(ms, 4), "phases": ph,
                          "plan": dict(sc.PLAN_COUNTS)}), flush=True)


main()
