"""Fixed and per-tile cost of one x2_tc_gemm launch (K = N = 128): M = 128 x tiles-per-CTA x CTAs."""
import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from x2gnn_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda")
K = N = 128
w = torch.randn(N, K, device=dev) * 0.1; b = torch.randn(N, device=dev)
_lib.require_cuda(w, what="probe")
ws = _lib.workspace(L.x2_tc_gemm_workspace_bytes(128, 128), dev)
st = torch.cuda.current_stream().cuda_stream
out = {}
for M in (148 * 256, 148 * 256 + 128, 148 * 256 + 40, 148 * 256 + 41 * 128, 43048 - 40, 43048, 43136, 148 * 384):
    x = torch.randn(M, K, device=dev); c = torch.empty(M, N, device=dev)
    f = lambda: L.x2_tc_gemm(x.data_ptr(), K, M, K, w.data_ptr(), 1, K, N, b.data_ptr(), c.data_ptr(), N, 0, ws.data_ptr(), ws.numel(), st)
    for _ in range(5): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): f()
    e1.record(); torch.cuda.synchronize()
    out[M] = round(e0.elapsed_time(e1) / 100 * 1e3, 2)
# an empty-ish kernel for the launch floor: a 1-element fill
z = torch.zeros(1, device=dev)
torch.cuda.synchronize(); e0.record()
for _ in range(100): z.fill_(1.0)
e1.record(); torch.cuda.synchronize()
out["fill_1elem"] = round(e0.elapsed_time(e1) / 100 * 1e3, 2)
print(json.dumps(out))
