"""GPU end-to-end parity: the harness model (hot path on the sm_100a kernels) against
(1) predictions of the reference's own xgnn.py/model.py (golden, small dims) and
(2) the CPU oracle model at config.json dims -- "numerically matching the reference on U0"."""
import pytest
import torch

from oracle import model as omodel
from util import relerr, to_t

pytestmark = pytest.mark.gpu


# segment_edge_attr: True = conv layers get the per-atom edgenn table (opt-in fast path),
# False = the [T, A] per-triplet gather of the reference (xgnn.py:57-58).  Same function either way.
@pytest.mark.parametrize("segment_edge_attr", [True, False])
def test_golden_small_model(golden, segment_edge_attr):
    from x2gnn_b200.xgnn_model import XGNNPoly
    rec = golden("model")["small"]
    net = XGNNPoly(**rec["hparams"])
    net.segment_edge_attr = segment_edge_attr
    assert list(net.state_dict().keys()) == list(rec["state_dict"].keys())
    net.load_state_dict(rec["state_dict"])        # reference checkpoint loads unchanged
    net = net.cuda().eval()
    data = to_t(rec["batch"], device="cuda")
    with torch.no_grad():
        pred = net(data)
    assert relerr(pred, rec["pred_f64"]) < 2e-5


@pytest.mark.parametrize("segment_edge_attr", [True, False])
def test_config_dims_keys_and_u0_prediction(golden, segment_edge_attr):
    from x2gnn_b200 import synth
    from x2gnn_b200.xgnn_model import XGNNPoly
    hp = dict(conv_layers=4, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
    torch.manual_seed(0)
    ref = omodel.XGNNPoly(**hp)
    net = XGNNPoly(**hp)
    net.segment_edge_attr = segment_edge_attr
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == golden("model")["cfg_keys"]
    net.load_state_dict(ref.state_dict())
    net = net.cuda()
    b = synth.qm9_batch(6, seed=4)
    ref = ref.double().eval()
    pred_ref = ref(to_t(b, dtype=torch.float64))
    net.train()
    pred = net(to_t(b, device="cuda"))
    assert relerr(pred, pred_ref) < 5e-5
    # a full training-style backward: gradients of every parameter group against the oracle
    y = torch.linspace(-1, 1, pred.numel())
    torch.nn.functional.smooth_l1_loss(pred_ref, y.double()).backward()
    torch.nn.functional.smooth_l1_loss(pred, y.cuda()).backward()
    pr = dict(ref.named_parameters())
    checked = 0
    for k, p in net.named_parameters():
        g_ref = pr[k].grad
        if g_ref is None or p.grad is None:
            assert g_ref is None and p.grad is None, k
            continue
        if float(g_ref.abs().max()) < 1e-12:
            continue
        assert relerr(p.grad, g_ref) < 5e-4, k
        checked += 1
    assert checked > 100


def test_graphed_train_step_matches_eager():
    """x2gnn_b200.train_graph: the dense part of a training step captured into a CUDA graph (all library
    launches go through the C ABI on the capturing stream) and replayed against the same model stepped
    eagerly -- loss trajectory, parameters and EMA after 4 replays."""
    from x2gnn_b200 import synth
    from x2gnn_b200.train_graph import GraphedTrainStep, dense_step
    from x2gnn_b200.xgnn_model import XGNNPoly
    hp = dict(conv_layers=2, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
    data = to_t(synth.qm9_batch(5, seed=2), device="cuda")
    y = torch.linspace(-1, 1, 5, device="cuda")

    def make():
        torch.manual_seed(0)
        return XGNNPoly(**hp).cuda()

    net = make()
    gs = GraphedTrainStep(net, data, y, lr=1e-3, warmup=2)
    g_losses = [float(gs.replay()) for _ in range(4)]

    ref = make()
    ps = [p for p in ref.parameters() if p.requires_grad]
    opt = torch.optim.Adam(ps, lr=1e-3, fused=True)
    ema = [p.detach().clone() for p in ps]
    prep = ref.prepare(data)
    e_losses = [float(dense_step(ref, opt, ps, data, prep, y, ema)) for _ in range(2 + 4)]
    assert g_losses == pytest.approx(e_losses[2:], rel=2e-4)
    # Adam divides by sqrt(v): elements whose gradient is at rounding-noise level may step either way, so the
    # parameter check is a norm over the whole vector, not element-wise
    def flat(ts):
        return torch.cat([t.detach().double().reshape(-1) for t in ts])
    init = flat(make().parameters())
    pa, pb = flat(net.parameters()), flat(ref.parameters())
    assert float((pb - init).norm()) > 0                   # the replays did update the parameters
    assert float((pa - pb).norm() / (pb - init).norm()) < 0.1     # e.g. lin_key.bias: zero gradient analytically
    ea, eb = flat(gs.ema_params), flat(ema)
    assert float((ea - eb).norm() / (eb - init).norm()) < 0.1
    # a second batch's values written into the captured buffers are what the next replay consumes
    with torch.no_grad():
        before = float(gs.loss)
        gs.y.add_(1.0)
        after = float(gs.replay())
    assert abs(after - before) > 1e-3


@pytest.mark.parametrize("D,counts", [(128, [72, 0, 650, 1, 240]), (8, [3, 5]), (256, [1000]), (128, [])])
def test_graph_layernorm_kernel_matches_oracle(D, counts):
    """x2_graph_layernorm_fwd / _bwd (model.py:24,46: PyG LayerNorm with a batch vector, affine=False)
    against the fp64 oracle: ragged molecules, an empty one, a single row, one large graph, no graphs."""
    from x2gnn_b200.graph_norm import graph_layer_norm_rows, rowptr_from_counts
    from x2gnn_b200.xgnn_model import graph_layer_norm
    torch.manual_seed(3)
    cnt = torch.tensor(counts, dtype=torch.int64)
    rows, B = int(cnt.sum()), len(counts)
    x = (torch.randn(rows, D) * 2.0 + 0.7).requires_grad_()
    gy = torch.randn(rows, D)
    batch = torch.repeat_interleave(torch.arange(B), cnt)
    xr = x.detach().double().requires_grad_()
    yr = omodel.graph_layer_norm(xr, batch, B) if B else xr * 1.0
    yr.backward(gy.double())
    xc = x.detach().cuda().requires_grad_()
    rp = rowptr_from_counts(cnt.cuda())
    assert rp.dtype == torch.int32 and rp.tolist() == [0] + torch.cumsum(cnt, 0).tolist()
    y = graph_layer_norm_rows(xc, rp)
    y.backward(gy.cuda())
    if rows == 0:
        assert y.shape == (0, D) and xc.grad.shape == (0, D)
        return
    assert relerr(y, yr) < 1e-5
    assert relerr(xc.grad, xr.grad) < 1e-5
    # deterministic, and the composite path of the harness (any batch order) agrees
    y2 = graph_layer_norm_rows(xc.detach(), rp)
    assert torch.equal(y2, y.detach())
    y3 = graph_layer_norm(xc.detach(), batch.cuda(), B, counts=cnt.cuda())
    assert relerr(y3, yr) < 1e-5


@pytest.mark.parametrize("D,R,bias,deg", [(128, 6, True, [3, 0, 28, 1, 9]), (256, 16, True, [5, 7, 0]),
                                          (128, 9, False, [40] * 70), (128, 6, True, [0, 0]), (256, 3, True, [])])
def test_rbf_readout_kernel_matches_composite(D, R, bias, deg):
    """x2_rbf_readout_fwd / _bwd (readout.py:34-43: scatter-sum of lin_rbf(rbf) * x over the bonds of an atom)
    against the same composite in fp64 on the CPU: ragged atoms, atoms without bonds, both channel widths,
    both register variants of the backward (R <= 8, R <= 16), no bias, many blocks, nothing to do."""
    from x2gnn_b200.graph_norm import rowptr_from_counts
    from x2gnn_b200.readout_sum import rbf_readout
    torch.manual_seed(5)
    cnt = torch.tensor(deg, dtype=torch.int64)
    E, N = int(cnt.sum()), len(deg)
    idx = torch.repeat_interleave(torch.arange(N), cnt)
    leaves = dict(x=torch.randn(E, D), rbf=torch.rand(E, R) * 2 - 1, w=torch.randn(D, R) * 0.5)
    if bias:
        leaves["b"] = torch.randn(D) * 0.3
    g = torch.randn(N, D)
    ref = {k: v.double().requires_grad_() for k, v in leaves.items()}
    F = ref["rbf"] @ ref["w"].t() + (ref["b"] if bias else 0.0)
    out_ref = torch.zeros(N, D, dtype=torch.float64).index_add(0, idx, F * ref["x"])
    out_ref.backward(g.double())
    dev = {k: v.cuda().requires_grad_() for k, v in leaves.items()}
    out = rbf_readout(dev["x"], dev["rbf"], dev["w"], dev.get("b"), rowptr_from_counts(cnt.cuda()))
    out.backward(g.cuda())
    assert out.shape == (N, D)
    if N == 0:
        return
    if E == 0:
        assert float(out.abs().max()) == 0.0 and float(dev["w"].grad.abs().max()) == 0.0
        return
    assert relerr(out, out_ref) < 1e-5
    for k in leaves:
        assert relerr(dev[k].grad, ref[k].grad) < 1e-5, k
    out2 = rbf_readout(dev["x"].detach(), dev["rbf"].detach(), dev["w"].detach(),
                       dev["b"].detach() if bias else None, rowptr_from_counts(cnt.cuda()))
    assert torch.equal(out2, out.detach())                      # deterministic


def test_fused_optimizer_tail_matches_torch():
    """x2_optim_tail (clip_grad_norm_ + Adam + EMA over flat buffers, two launches) against torch's own
    clip_grad_norm_ / Adam / lerp on the same tensors, 6 updates, with and without clipping active."""
    from x2gnn_b200.optim_tail import FusedTail
    for max_norm in (100.0, 0.05):
        torch.manual_seed(3)
        shapes = [(128, 128), (128,), (42, 7), (1,), (338, 256)]
        pa = [torch.nn.Parameter(torch.randn(*s, device="cuda") * 0.1) for s in shapes]
        pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
        tail = FusedTail(pa, lr=1e-3, max_norm=max_norm, ema_decay=0.95)
        opt = torch.optim.Adam(pb, lr=1e-3)
        ema = [p.detach().clone() for p in pb]
        g = torch.Generator(device="cuda").manual_seed(1)
        for it in range(6):
            grads = [torch.randn(*s, device="cuda", generator=g) * (1.0 + it) for s in shapes]
            tail.zero_grad()
            for p, gr in zip(pa, grads):
                p.grad = gr.clone()                   # what autograd leaves
            tail.step()
            opt.zero_grad(set_to_none=True)
            for p, gr in zip(pb, grads):
                p.grad = gr.clone()
            norm = torch.nn.utils.clip_grad_norm_(pb, max_norm=max_norm)
            opt.step()
            torch._foreach_lerp_(ema, [p.detach() for p in pb], 1.0 - 0.95)
            assert float(tail.grad_norm) == pytest.approx(float(norm), rel=1e-6)
        assert float(tail.step_count) == 6.0
        for a, b in zip(pa, pb):
            assert relerr(a, b) < 1e-6
        for a, b in zip(tail.ema_views, ema):
            assert relerr(a, b) < 1e-6
        # the module's tensors are views of the flat buffers
        assert pa[0].data_ptr() == tail.flat_p.data_ptr() and all(p.data_ptr() % 256 == 0 for p in pa)


def test_graphed_train_step_torch_tail_still_available():
    from x2gnn_b200 import synth
    from x2gnn_b200.train_graph import GraphedTrainStep
    from x2gnn_b200.xgnn_model import XGNNPoly
    hp = dict(conv_layers=1, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128)
    data = to_t(synth.qm9_batch(3, seed=2), device="cuda")
    y = torch.linspace(-1, 1, 3, device="cuda")
    losses = {}
    for fused in (True, False):
        torch.manual_seed(0)
        net = XGNNPoly(**hp).cuda()
        gs = GraphedTrainStep(net, data, y, lr=1e-3, warmup=2, fused_tail=fused)
        losses[fused] = [float(gs.replay()) for _ in range(3)]
    assert losses[True] == pytest.approx(losses[False], rel=2e-4)


def test_deferred_weight_gradients_match_the_immediate_ones():
    """tc_linear.DeferredWgrads / x2_tc_wgrad_batch: the TCLinear weight gradients of one backward computed together at
    the end, straight into the optimizer's flat buffer, against the per-layer x2_tc_wgrad launches through autograd
    (and both against fp64 autograd of the same stack): every gradient, 1e-5 relative."""
    from x2gnn_b200.optim_tail import FusedTail
    from x2gnn_b200.tc_linear import TCLinear
    torch.manual_seed(0)
    rows = 30011                     # > 16 problems x 768 rows: the batched launch runs the accumulation periods

    class Stack(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.inp = TCLinear(338, 256)                 # K blocks 128 | 128 | 82, rows of 1 352 bytes (not 16-aligned)
            self.mid = torch.nn.ModuleList([TCLinear(256 if i == 0 else 128, 128, bias=(i % 3 != 2)) for i in range(19)])
            self.small = TCLinear(128, 64)                # out_features the kernels do not take: ordinary path
            self.out = torch.nn.Linear(64, 1)

        def forward(self, x):
            h = torch.nn.functional.silu(self.inp(x))
            for m in self.mid:
                h = torch.nn.functional.silu(m(h)) + (h if h.size(1) == 128 else 0)
            return self.out(self.small(h))

    net = Stack().cuda()
    x = torch.randn(rows, 338, device="cuda")
    grads = {}
    for defer in (False, True):
        tail = FusedTail(net.parameters(), lr=1e-3, max_norm=0.0)
        n = tail.defer_wgrads(net, defer)
        assert n == (2 * 20 - 6 if defer else 0)         # 20 deferring layers, 6 of them without bias... see below
        tail.zero_grad()
        tail.backward(net(x).square().mean())
        if defer:
            assert net.inp.weight.grad is None and net.small.weight.grad is not None
            assert len(tail.queue.pending) == 20
        tail._pack()
        assert not tail.queue.pending
        grads[defer] = [v.clone() for v in tail.grad_views]
        tail.defer_wgrads(net, False)
    ref = Stack().double().cuda()
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    ref(x.double()).square().mean().backward()
    for (k, p), a, b in zip(ref.named_parameters(), grads[False], grads[True]):
        assert relerr(b, p.grad) < 1e-5, k
        assert relerr(a, p.grad) < 1e-5, k
    # a weight used twice in one backward cannot defer
    tail = FusedTail(net.parameters(), lr=1e-3, max_norm=0.0)
    tail.defer_wgrads(net, True)
    h = torch.randn(64, 128, device="cuda")
    with pytest.raises(RuntimeError):
        tail.backward((net.mid[3](h) + net.mid[3](h * 2)).sum())
    tail.queue.flush(h.device)
    tail.defer_wgrads(net, False)
