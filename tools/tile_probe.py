"""Forward-only timing of one SBFTransformerConv layer on the bench workload (QM9 batch 128), CUDA events.
Development probe for the fused tile kernel: run with X2GNN_TA_DBG=<bits> for the ablations listed in
csrc/tile_attn.cuh, X2GNN_FUSED=0 for the unfused pair."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import x2gnn_b200
from x2gnn_b200 import synth
from x2gnn_b200.sbftransformer_conv import SBFTransformerConv

def main():
    nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    D, H, S, R, A = 128, 16, 42, 6, 128
    b = synth.qm9_batch(nmol, seed=0)
    tri = synth.triplets_host(b["edge_index"], len(b["x"]))[0]
    E = b["edge_index"].shape[1]
    ci = synth.conv_inputs(E, tri, D, S, R, A, seed=0)
    t = {k: torch.from_numpy(v).cuda() for k, v in ci.items()}
    torch.manual_seed(0)
    conv = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).cuda()
    conv.precision = 1 if os.environ.get("X2GNN_FUSED") == "0" else 2     # X2_MODE_TF32X3 / X2_MODE_TF32X3_FUSED
    def step():
        with torch.enable_grad():
            x = t["x"].requires_grad_(True)
            return conv(t["sbf"], t["rbf"], x=x, edge_index=t["edge_index"], edge_attr=t["edge_attr"])
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        step()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    if os.environ.get("X2GNN_TA_TRACE"):
        import ctypes
        from x2gnn_b200 import _lib
        h = ctypes.CDLL(_lib.LIB_PATH)
        buf = torch.zeros(16 * 256, dtype=torch.int64, device="cuda")
        h.x2_debug_tile_trace(ctypes.c_void_p(buf.data_ptr()))
        step(); torch.cuda.synchronize()
        h.x2_debug_tile_trace(ctypes.c_void_p(0))
        tr = buf.cpu().view(16, 256)
        t0 = int(tr[tr > 0].min())
        names = ["mma_wait", "mma_full", "tr_accfull", "tr_done", "prod_issue", "prod_landed", "prod_full",
                 "c0_wait", "c0_go", "c0_p1done", "c0_tiledone", "c14_wait", "c14_go", "c14_p1done", "c14_tiledone", "c14_pass1"]
        out = {n: [int(v) - t0 if v > 0 else -1 for v in tr[i, :40].tolist()] for i, n in enumerate(names)}
        with open(os.environ["X2GNN_TA_TRACE"], "w") as f:
            json.dump(out, f)
    print(json.dumps({"dbg": os.environ.get("X2GNN_TA_DBG", "0"), "fused": os.environ.get("X2GNN_FUSED", "1"),
                      "E": E, "T": int(t["edge_index"].shape[1]), "fwd_ms": ms}))

main()
