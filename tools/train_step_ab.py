"""A/B of the graphed training step (bench batch, harness model): wall time per replay (CUDA events, 30 replays)
with the TCLinear weight gradients deferred and batched (default) and launched per layer.
python tools/train_step_ab.py  -> one JSON line."""
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from x2gnn_b200 import synth
from x2gnn_b200.train_graph import GraphedTrainStep
from x2gnn_b200.xgnn_model import XGNNPoly
dev = torch.device("cuda")
hp = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
b = synth.qm9_batch(128, seed=0)
data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in b.items()}
out = {}
for name, kw in (("deferred_wgrads", dict(defer_wgrads=True)), ("per_layer_wgrads", dict(defer_wgrads=False))):
    torch.manual_seed(0)
    model = XGNNPoly(**hp).to(dev)
    gs = GraphedTrainStep(model, data, torch.zeros(128, device=dev), **kw)
    for _ in range(5):
        gs.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        gs.replay()
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"ms_per_step": e0.elapsed_time(e1) / 30, "loss": float(gs.loss), "deferred_tensors": getattr(gs, "deferred", 0)}
    del gs, model
print(json.dumps(out))
