"""GPU parity of the tcgen05 3xTF32 GEMM building blocks (x2_tc_gemm / x2_tc_wgrad) against fp64
matmul.  The bar is fp32-level accuracy: max error <= 1e-5 of the output scale (north_star), and in
practice ~1e-6 (two dropped 2^-22-order terms)."""
import pytest
import torch

from util import relerr

pytestmark = pytest.mark.gpu


def _tc_gemm(A, W, sbk, sbn, K, N, bias=None, C=None, beta=0):
    from x2gnn_b200 import _lib
    L = _lib.lib()
    M = A.size(0)
    if C is None:
        C = torch.full((M, N), float("nan"), device=A.device)
    ws = _lib.workspace(L.x2_tc_gemm_workspace_bytes(K, N), A.device)
    _lib.check(L.x2_tc_gemm(_lib.ptr(A), A.stride(0), M, K, _lib.ptr(W), sbk, sbn, N, _lib.ptr(bias),
                            _lib.ptr(C), C.stride(0), beta, _lib.ptr(ws), ws.numel(), _lib.stream()), "x2_tc_gemm")
    return C


@pytest.mark.parametrize("M,K,N", [(128, 32, 128), (128, 128, 128), (1000, 128, 128), (40000, 128, 128),
                                   (777, 42, 128), (513, 6, 128), (300, 128, 42), (300, 128, 6),
                                   (1, 128, 128), (129, 64, 16), (260, 100, 100)])
def test_forward_layout(M, K, N):
    """y = x W^T + b   (W [N,K] as torch.nn.Linear stores it)."""
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    C = _tc_gemm(A, W, 1, K, K, N, bias=b)
    ref = A.double() @ W.double().T + b.double()
    assert relerr(C, ref) < 2e-6, (M, K, N)


@pytest.mark.parametrize("M,K,N", [(1000, 128, 128), (555, 128, 42), (555, 128, 6), (200, 256 // 2, 128)])
def test_dgrad_layout_and_accumulate(M, K, N):
    """dx = dy W   (W [K,N] row-major => B(k,n) = W[k*N + n]); beta=1 accumulates into C."""
    g = torch.Generator().manual_seed(7 * M + N)
    dy = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(K, N, generator=g).cuda()
    C0 = torch.randn(M, N, generator=g).cuda()
    C = _tc_gemm(dy, W, N, 1, K, N)
    assert relerr(C, dy.double() @ W.double()) < 2e-6
    C = _tc_gemm(dy, W, N, 1, K, N, C=C0.clone(), beta=1)
    assert relerr(C, C0.double() + dy.double() @ W.double()) < 2e-6


def test_strided_operands():
    """A and C as column blocks of wider buffers (how the conv path passes dQ|dK|dV and Q|K|V|skip)."""
    g = torch.Generator().manual_seed(3)
    buf = torch.randn(900, 384, generator=g).cuda()
    A = buf[:, 128:256]
    W = torch.randn(128, 128, generator=g).cuda()
    out = torch.zeros(900, 512, device="cuda")
    _tc_gemm(A, W, 1, 128, 128, 128, C=out[:, 256:384])
    assert relerr(out[:, 256:384], A.double() @ W.double().T) < 2e-6
    assert float(out[:, :256].abs().max()) == 0 and float(out[:, 384:].abs().max()) == 0


@pytest.mark.parametrize("rows,N", [(32, 128), (1000, 128), (50000, 128), (4097, 42), (999, 6), (5, 128), (70000, 64),
                                    (811834, 128), (811834, 42), (300001, 20)])    # bench-batch row counts: see below
def test_wgrad(rows, N):
    from x2gnn_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(rows + N)
    Ybuf = torch.randn(rows, 384, generator=g).cuda()
    Y = Ybuf[:, 128:256]
    X = torch.randn(rows, N, generator=g).cuda()
    dW = torch.full((128, N), float("nan"), device="cuda")
    db = torch.full((128,), float("nan"), device="cuda")
    ws = _lib.workspace(L.x2_tc_wgrad_workspace_bytes(rows, N), "cuda")
    for _ in range(2):
        _lib.check(L.x2_tc_wgrad(_lib.ptr(Y), Y.stride(0), _lib.ptr(X), X.stride(0), rows, N, _lib.ptr(dW),
                                 dW.stride(0), _lib.ptr(db), _lib.ptr(ws), ws.numel(), _lib.stream()), "x2_tc_wgrad")
    ref = Y.double().T @ X.double()
    # (the tensor core accumulates with truncation: before the accumulation was cut into 256-row periods summed in
    # fp32 registers, the result shrank by 6e-9 per row a CTA accumulated -- 4.2e-5 at 811 834 rows, above the bar)
    # the tensor core's fp32 accumulation over thousands of rows costs a little more than one
    # rounding per add; still well inside the 1e-5 bar
    assert relerr(dW, ref) < 6e-6, (rows, N)
    assert relerr(db, Y.double().sum(0)) < 2e-6
    dW2 = dW.clone()
    _lib.check(L.x2_tc_wgrad(_lib.ptr(Y), Y.stride(0), _lib.ptr(X), X.stride(0), rows, N, _lib.ptr(dW),
                             dW.stride(0), None, _lib.ptr(ws), ws.numel(), _lib.stream()), "x2_tc_wgrad")
    assert torch.equal(dW, dW2)          # deterministic split-K


def test_precision_is_fp32_level_not_tf32():
    """Single-pass TF32 would give ~5e-4; the 3-term split must be ~1e-6."""
    g = torch.Generator().manual_seed(11)
    A = torch.randn(4096, 128, generator=g).cuda()
    W = torch.randn(128, 128, generator=g).cuda()
    C = _tc_gemm(A, W, 1, 128, 128, 128)
    ref = A.double() @ W.double().T
    err = relerr(C, ref)
    fp32 = relerr(A @ W.T, ref) if not torch.backends.cuda.matmul.allow_tf32 else None
    assert err < 2e-6, (err, fp32)


@pytest.mark.parametrize("M,K,N", [(1000, 128, 128), (777, 338, 256), (512, 256, 128), (300, 128, 1), (64, 6, 128)])
def test_tc_linear_module_fwd_bwd(M, K, N):
    """TCLinear == nn.Linear (fp64) for output, input grad, weight grad, bias grad; state_dict-compatible."""
    from x2gnn_b200.tc_linear import TCLinear
    torch.manual_seed(M + N)
    ref = torch.nn.Linear(K, N).double()
    lin = TCLinear(K, N)
    lin.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    lin = lin.cuda()
    x = torch.randn(M, K)
    g = torch.randn(M, N)
    xr = x.double().requires_grad_(True)
    ref(xr).backward(g.double())
    xc = x.cuda().requires_grad_(True)
    y = lin(xc)
    y.backward(g.cuda())
    assert relerr(y, ref(xr)) < 2e-6
    assert relerr(xc.grad, xr.grad) < 2e-6
    assert relerr(lin.weight.grad, ref.weight.grad) < 6e-6
    assert relerr(lin.bias.grad, ref.bias.grad) < 6e-6
