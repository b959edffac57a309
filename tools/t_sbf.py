"""Timing of the two F_B_2D kernels (x2_sbf_table, x2_sbf_fwd) on the bench batch and on a 3.8 M-triplet
graph where GPU time dominates the host overhead of the call.  Development tool: python tools/t_sbf.py"""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from x2gnn_b200 import atom_graph, edge_graph, synth
from x2gnn_b200.angular_basis_layer import F_B_2D
dev="cuda"
b = synth.qm9_batch(128, seed=0)
pos = torch.from_numpy(b["atom_pos"]).to(dev); batch = torch.from_numpy(b["batch"]).to(dev)
ei,_ = atom_graph.radius_graph(pos, batch, 5.0)
tri, aj, ai, ak = edge_graph.vertex_to_edge_2(ei, pos.size(0))
d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
ji, jk = pos[ai]-pos[aj], pos[ak]-pos[aj]
ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji*jk).sum(1))
L = F_B_2D(7,6,5.0,5)
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
print("table ms", timed(lambda: L.radial_table(d)))
print("full  ms", timed(lambda: L(d, ang, tri[0])))
idx=tri[0].contiguous()
print("idx contiguous?", tri[0].is_contiguous())
# larger graph: GPU time dominates the host overhead of the call
b = synth.ball_batch(5, n_atoms=500, seed=0)
pos = torch.from_numpy(b["atom_pos"]).to(dev); batch = torch.from_numpy(b["batch"]).to(dev)
ei,_ = atom_graph.radius_graph(pos, batch, 5.0)
tri, aj, ai, ak = edge_graph.vertex_to_edge_2(ei, pos.size(0))
d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
ji, jk = pos[ai]-pos[aj], pos[ak]-pos[aj]
ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji*jk).sum(1))
T=tri.size(1)
t_tab=timed(lambda: L.radial_table(d)); t_full=timed(lambda: L(d, ang, tri[0]))
print("ball5: T", T, "table ms", t_tab, "full ms", t_full, "sbf_fwd GB/s", (4*T*43+8*T)/((t_full-t_tab)*1e-3)/1e9)
