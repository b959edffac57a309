"""Per-kernel timing of the T-scale tensor-core GEMM building blocks on the bench workload's row count
(CUDA events, L2-exceeding operands): GB/s = HBM bytes the call must move / time.
usage: python tools/kbench.py [rows] [which ...]   which in {gemm_e, gemm_s, dgrad_e, wgrad_e, wgrad_s}"""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from x2gnn_b200 import _lib  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 811834
which = [a for a in sys.argv[1:] if not a.isdigit()] or ["gemm_e", "gemm_s", "dgrad_e", "wgrad_e", "wgrad_s"]
L = _lib.lib()
dev = "cuda"
g = torch.Generator(dev).manual_seed(0)
X128 = torch.randn(rows, 128, device=dev, generator=g)
X42 = torch.randn(rows, 42, device=dev, generator=g)
Y = torch.randn(rows, 128, device=dev, generator=g)
W128 = torch.randn(128, 128, device=dev, generator=g)
W42 = torch.randn(128, 42, device=dev, generator=g)
b = torch.randn(128, device=dev, generator=g)
C = torch.empty(rows, 128, device=dev)
dW = torch.empty(128, 128, device=dev)
dW42 = torch.empty(128, 42, device=dev)
db = torch.empty(128, device=dev)
wsg = _lib.workspace(L.x2_tc_gemm_workspace_bytes(128, 128), dev)
wsw = _lib.workspace(L.x2_tc_wgrad_workspace_bytes(rows, 128), dev)
st = _lib.stream()


def gemm(A, W, K, bias):
    _lib.check(L.x2_tc_gemm(_lib.ptr(A), A.stride(0), rows, K, _lib.ptr(W), 1, K, 128, _lib.ptr(bias), _lib.ptr(C),
                            128, 0, _lib.ptr(wsg), wsg.numel(), st))


def dgrad():
    _lib.check(L.x2_tc_gemm(_lib.ptr(Y), 128, rows, 128, _lib.ptr(W128), 128, 1, 128, None, _lib.ptr(C), 128, 0,
                            _lib.ptr(wsg), wsg.numel(), st))


def wgrad(X, N, out, bias):
    _lib.check(L.x2_tc_wgrad(_lib.ptr(Y), 128, _lib.ptr(X), X.stride(0), rows, N, _lib.ptr(out), N, _lib.ptr(bias),
                             _lib.ptr(wsw), wsw.numel(), st))


cases = {
    "gemm_e": (lambda: gemm(X128, W128, 128, None), 1024),
    "gemm_s": (lambda: gemm(X42, W42, 42, b), 168 + 512),
    "dgrad_e": (dgrad, 1024),
    "wgrad_e": (lambda: wgrad(X128, 128, dW, None), 1024),
    "wgrad_s": (lambda: wgrad(X42, 42, dW42, db), 512 + 168),
}
for name in which:
    fn, bpr = cases[name]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"kernel": name, "rows": rows, "ms": round(ms, 4), "GBs": round(rows * bpr / ms / 1e6, 1)}))
