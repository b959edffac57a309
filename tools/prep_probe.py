import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from x2gnn_b200 import synth
from x2gnn_b200.train_graph import GraphedTrainStep
from x2gnn_b200.xgnn_model import XGNNPoly
dev = torch.device("cuda")
hp = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
b = synth.qm9_batch(128, seed=0)
data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in b.items()}
torch.manual_seed(0)
model = XGNNPoly(**hp).to(dev)
for _ in range(3): model.prepare(data)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): model.prepare(data)
torch.cuda.synchronize()
print("prepare alone ms", (time.perf_counter() - t0) / 20 * 1e3)
gs = GraphedTrainStep(model, data, torch.zeros(128, device=dev))
ps = torch.cuda.Stream(priority=-1)
for _ in range(3): gs.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(ps):
        model.prepare(data)
    gs.replay()
torch.cuda.synchronize()
print("prepare + replay ms", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(20):
    gs.replay()
torch.cuda.synchronize()
print("replay alone ms", (time.perf_counter() - t0) / 20 * 1e3)
# where are the syncs: profile prepare on an idle device
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): model.prepare(data)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
