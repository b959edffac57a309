"""Kernel-time table of the graphed training step (x2gnn_b200.train_graph) at the bench shape:
python tools/prof_train_graph.py  -> per-kernel GPU time of 3 replays, sorted (torch profiler, CUDA activity)."""
import os, sys, collections, torch
sys.path.insert(0, os.getcwd())
from x2gnn_b200 import synth
from x2gnn_b200.train_graph import GraphedTrainStep
from x2gnn_b200.xgnn_model import XGNNPoly
dev = torch.device("cuda")
torch.manual_seed(0)
hp = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
model = XGNNPoly(**hp).to(dev)
b = synth.qm9_batch(128, seed=0)
data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in b.items()}
gs = GraphedTrainStep(model, data, torch.zeros(128, device=dev))
for _ in range(3):
    gs.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        gs.replay()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:110]
        tot[k][0] += e.device_time
        tot[k][1] += 1
rows = sorted(tot.items(), key=lambda kv: -kv[1][0])
total = sum(v[0] for _, v in rows)
print(f"GPU kernel time per replay: {total / N / 1e3:.3f} ms, {sum(v[1] for _, v in rows) // N} kernels")
for k, (us, n) in rows[:45]:
    print(f"{us / N:9.1f} us/step {n // N:5d} x {us / n:7.1f} us  {k}")
