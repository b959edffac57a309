import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from x2gnn_b200 import synth
from x2gnn_b200.xgnn_model import XGNNPoly
dev = torch.device("cuda")
torch.manual_seed(0)
hp = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
model = XGNNPoly(**hp).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
b = synth.qm9_batch(128, seed=0)
data = {k: (torch.from_numpy(v).to(dev) if hasattr(v, "shape") else v) for k, v in b.items()}
y = torch.zeros(128, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.smooth_l1_loss(model(data), y)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 100.0)
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
import time
t=time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); print("wall ms/step", (time.perf_counter()-t)/5*1e3)
