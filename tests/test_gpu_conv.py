"""GPU parity: SBFTransformerConv forward + backward through the C-ABI against
(1) golden vectors produced by the reference's own source in fp64 and
(2) the CPU oracle on seeded inputs, within the north_star tolerance (1e-5 relative, fp32)."""
import pytest
import torch

from oracle import conv as oconv
from util import FP32_TOL, relerr

pytestmark = pytest.mark.gpu

INPUTS = ("x", "rbf", "sbf", "edge_attr")


MODES = {"fp32": 0, "tf32x3": 1, "fused": 2, "tf32": 3}   # X2_MODE_FP32 / _TF32X3 / _TF32X3_FUSED (tile kernel) / _TF32


def _mine(dims, state, mode=None, **kw):
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    D, H, S, R, A = dims
    if mode in ("tf32x3", "fused", "tf32") and D % 128:
        pytest.skip("tensor-core mode needs D % 128 == 0")
    c = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, dropout=kw.pop("dropout", 0),
                           edge_dim=A, **kw)
    c.precision = None if mode is None else MODES[mode]
    if state is not None:
        c.load_state_dict(state)          # reference state_dict loads unchanged
    return c.cuda()


def _run(conv, rec, dev, dtype, want_alpha=False, sbf_grad=True):
    conv.zero_grad(set_to_none=True)
    xs = {k: rec[k].to(device=dev, dtype=dtype).requires_grad_(k != "sbf" or sbf_grad) for k in INPUTS}
    ei = rec["edge_index"].to(dev)
    r = conv(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=ei, edge_attr=xs["edge_attr"],
             return_attention_weights=True if want_alpha else None)
    out, alpha = (r[0], r[1][1]) if want_alpha else (r, None)
    out.backward(rec["grad_out"][:, :out.size(1)].to(device=dev, dtype=dtype))
    return out, alpha, xs


def test_reduced_precision_tf32_mode(golden):
    """X2_MODE_TF32 (one tf32 pass per product in the Linear layers) against the reference's fp64 results:
    inside the north_star's reduced-precision tolerance (2e-2), and measurably NOT fp32-accurate, i.e.
    the mode really runs a different arithmetic."""
    rec = golden("conv")["cfg"]
    conv = _mine(rec["dims"], rec["state_dict"], mode="tf32")
    out, _, xs = _run(conv, rec, "cuda", torch.float32)
    errs = {"out": relerr(out, rec["out_f64"])}
    for k in INPUTS:
        errs[k] = relerr(xs[k].grad, rec[f"grad_{k}_f64"])
    for k, p in conv.named_parameters():
        ref = rec[f"gradp_{k}_f64"]
        if float(ref.abs().max()) > 1e-12:
            errs[k] = relerr(p.grad, ref)
    assert max(errs.values()) < 2e-2, errs
    assert max(errs.values()) > FP32_TOL, errs


@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "fused"])
@pytest.mark.parametrize("tag", ["cfg", "small", "c16"])
def test_golden_fwd_bwd(golden, tag, mode):
    rec = golden("conv")[tag]
    conv = _mine(rec["dims"], rec["state_dict"], mode=mode)
    assert list(conv.state_dict().keys()) == list(rec["state_dict"].keys())
    out, alpha, xs = _run(conv, rec, "cuda", torch.float32, want_alpha=True)
    assert relerr(out, rec["out_f64"]) < FP32_TOL
    if mode == "fused":       # the alpha request takes the unfused path; run the fused tile kernel as well
        out2, _, xs2 = _run(conv, rec, "cuda", torch.float32, want_alpha=False)
        assert relerr(out2, rec["out_f64"]) < FP32_TOL
        for k in INPUTS:
            assert relerr(xs2[k].grad, rec[f"grad_{k}_f64"]) < FP32_TOL, k
    if "alpha_f64" in rec:
        assert relerr(alpha, rec["alpha_f64"]) < FP32_TOL
    for k in INPUTS:
        assert relerr(xs[k].grad, rec[f"grad_{k}_f64"]) < FP32_TOL, k
    for k, p in conv.named_parameters():
        ref = rec[f"gradp_{k}_f64"]
        if k == "lin_key.bias":           # App. A identity: exactly zero in exact arithmetic
            assert float(p.grad.abs().max()) < 1e-5 * float(rec["gradp_lin_value.bias_f64"].abs().max())
            continue
        assert relerr(p.grad, ref) < FP32_TOL, k


def _oracle_pair(dims, seed=0, mode=None, **kw):
    D, H, S, R, A = dims
    torch.manual_seed(seed)
    ref = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A, **kw)
    with torch.no_grad():                   # non-trivial biases everywhere
        for p in ref.parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    mine = _mine(dims, ref.state_dict(), mode=mode, **kw)
    return ref.double(), mine


def _graph_inputs(nmol, dims, seed):
    from x2gnn_b200 import synth
    D, H, S, R, A = dims
    b = synth.qm9_batch(nmol, seed=seed)
    tri = synth.triplets_host(b["edge_index"], len(b["x"]))[0]
    E = b["edge_index"].shape[1]
    ci = synth.conv_inputs(E, tri, D, S, R, max(A or 1, 1), seed=seed)
    rec = {k: torch.from_numpy(v) for k, v in ci.items()}
    rec["grad_out"] = torch.randn(E, D, generator=torch.Generator().manual_seed(seed + 1))
    return rec


def _compare(ref, mine, rec, tol=FP32_TOL, check_alpha=True):
    o_ref, a_ref, x_ref = _run(ref, rec, "cpu", torch.float64, want_alpha=True)
    if mine.precision == MODES["fused"]:   # forward through the fused tile kernel (no alpha request)
        o_f, _, _ = _run(mine, rec, "cuda", torch.float32, want_alpha=False)
        assert relerr(o_f, o_ref) < tol
    o, a, x = _run(mine, rec, "cuda", torch.float32, want_alpha=True)
    assert relerr(o, o_ref) < tol
    if check_alpha:
        assert relerr(a, a_ref) < tol
    for k in INPUTS:
        if x_ref[k].grad is None:
            continue
        assert relerr(x[k].grad, x_ref[k].grad) < tol, k
    pm = dict(mine.named_parameters())
    for k, p in ref.named_parameters():
        if k == "lin_key.bias":
            continue
        if p.grad is None:                  # parameter unused by this variant
            assert pm[k].grad is None, k
            continue
        assert relerr(pm[k].grad, p.grad) < tol, k


@pytest.mark.parametrize("dims", [(128, 16, 42, 6, 128), (256, 16, 112, 16, 128), (64, 8, 10, 3, 20),
                                  (32, 1, 5, 2, 7), (128, 4, 42, 6, 128)])
@pytest.mark.parametrize("mode", ["fp32", "tf32x3", "fused"])
def test_vs_oracle_qm9_batch(dims, mode):
    """config.json dims, class-default dims (xgnn.py:16) and odd shapes on a 6-molecule batch."""
    ref, mine = _oracle_pair(dims, mode=mode)
    _compare(ref, mine, _graph_inputs(6, dims, seed=2))


@pytest.mark.parametrize("kw", [dict(concat=False), dict(beta=True), dict(root_weight=False),
                                dict(bias=False), dict(concat=False, beta=True)])
def test_variants(kw):
    dims = (64, 8, 10, 3, 20)
    ref, mine = _oracle_pair(dims, seed=3, **kw)
    _compare(ref, mine, _graph_inputs(3, dims, seed=4))



@pytest.mark.parametrize("shape", [(48, 2, 16), (16, 1, 16), (24, 3, 8), (200, 16, 8), (64, 4, 8), ((40, 40), 4, 8)])
@pytest.mark.parametrize("kw", [dict(), dict(concat=False), dict(root_weight=False)])
def test_shapes_outside_the_kernel_widths(shape, kw):
    """Reference ctor generality (sbftransformer_conv.py:19,47-48): in_channels != heads*out_channels, widths the
    kernels are not instantiated for, and a (equal) tuple in_channels run zero-padded at the next kernel width
    (sbftransformer_conv.padded_width) and must match the UNPADDED fp64 oracle on every output and gradient."""
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    in_ch, H, Cc = shape
    S, R, A = 10, 3, 20
    width = in_ch if isinstance(in_ch, int) else in_ch[0]
    torch.manual_seed(7)
    ref = oconv.OracleSBFTransformerConv(width, Cc, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A, **kw)
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    mine = SBFTransformerConv(in_ch, Cc, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A, **kw)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    rec = _graph_inputs(3, (width, H, S, R, A), seed=8)
    n_out = H * Cc if kw.get("concat", True) else Cc
    rec["grad_out"] = torch.randn(rec["x"].size(0), n_out, generator=torch.Generator().manual_seed(9))
    _compare(ref.double(), mine, rec)
    for k, p in mine.named_parameters():          # gradients have the PARAMETERS' shapes, not the padded ones
        assert p.grad is None or p.grad.shape == p.shape, k


def test_shape_outside_kernel_widths_golden(golden):
    """The same against the REFERENCE's own layer (fp64 golden `odd48`: in 48, heads 2, out 16): the kernels at the
    padded width 64 (4 heads, two of them zero) against the unpadded reference, every output and gradient."""
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv, padded_width
    rec = golden("conv")["odd48"]
    in_ch, H, C, S, R, A = rec["shape"]
    assert padded_width(in_ch, H, C) == 64
    mine = SBFTransformerConv(in_ch, C, heads=H, sbf_dim=S, rbf_dim=R, dropout=0, edge_dim=A)
    mine.load_state_dict(rec["state_dict"])
    mine = mine.cuda()
    out, alpha, xs = _run(mine, rec, "cuda", torch.float32, want_alpha=True)
    assert relerr(out, rec["out_f64"]) < FP32_TOL and relerr(alpha, rec["alpha_f64"]) < FP32_TOL
    for k in INPUTS:
        assert relerr(xs[k].grad, rec[f"grad_{k}_f64"]) < FP32_TOL, k
    for k, p in mine.named_parameters():
        if k == "lin_key.bias":
            continue
        assert relerr(p.grad, rec[f"gradp_{k}_f64"]) < FP32_TOL, k


def test_unsupported_shape_raises():
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    c = SBFTransformerConv(24, 24, heads=1, sbf_dim=4, rbf_dim=2).cuda()          # out_channels not a power of two
    x = torch.randn(4, 24, device="cuda")
    ei = torch.tensor([[0, 1, 2], [1, 2, 3]], device="cuda")
    with pytest.raises(NotImplementedError):
        c(torch.randn(3, 4, device="cuda"), torch.randn(4, 2, device="cuda"), x=x, edge_index=ei)
    with pytest.raises(ValueError):                                                  # x of the wrong width
        SBFTransformerConv(32, 8, heads=4, sbf_dim=4, rbf_dim=2).cuda()(
            torch.randn(3, 4, device="cuda"), torch.randn(4, 2, device="cuda"), x=x, edge_index=ei)


@pytest.mark.parametrize("heads", [8, 16])       # D = 64 and D = 128 (the latter: staged forward kernel)
def test_no_edge_dim(heads):
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    torch.manual_seed(0)
    ref = oconv.OracleSBFTransformerConv(8 * heads, 8, heads=heads, sbf_dim=10, rbf_dim=3, edge_dim=None)
    mine = SBFTransformerConv(8 * heads, 8, heads=heads, sbf_dim=10, rbf_dim=3, edge_dim=None)
    assert "lin_edge.weight" not in mine.state_dict()
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    rec = _graph_inputs(3, (8 * heads, heads, 10, 3, 1), seed=5)
    x = {k: rec[k] for k in ("x", "rbf", "sbf")}
    o_ref = ref.double()(x["sbf"].double(), x["rbf"].double(), x=x["x"].double(), edge_index=rec["edge_index"])
    o = mine(x["sbf"].cuda(), x["rbf"].cuda(), x=x["x"].cuda(), edge_index=rec["edge_index"].cuda())
    assert relerr(o, o_ref) < FP32_TOL


def test_empty_segments_and_no_triplets():
    """Diatomics: every segment is empty => out = lin_skip(x) (App. A); T = 0 must work."""
    dims = (64, 8, 10, 3, 20)
    ref, mine = _oracle_pair(dims, seed=6)
    E = 6
    rec = dict(x=torch.randn(E, 64), rbf=torch.rand(E, 3), sbf=torch.zeros(0, 10), edge_attr=torch.zeros(0, 20),
               edge_index=torch.zeros(2, 0, dtype=torch.int64), grad_out=torch.randn(E, 64))
    o, _, xs = _run(mine, rec, "cuda", torch.float32)
    want = torch.nn.functional.linear(rec["x"].double(), ref.lin_skip.weight, ref.lin_skip.bias)
    assert relerr(o, want) < FP32_TOL
    assert torch.isfinite(xs["x"].grad).all()
    # a mix: some targets have no incoming triplet
    rec = _graph_inputs(2, dims, seed=7)
    keep = rec["edge_index"][1] % 3 != 0
    for k in ("sbf", "edge_attr"):
        rec[k] = rec[k][keep]
    rec["edge_index"] = rec["edge_index"][:, keep]
    _compare(ref, mine, rec)


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_unsorted_edge_index_and_permutation_invariance(mode):
    dims = (128, 16, 42, 6, 128)
    ref, mine = _oracle_pair(dims, seed=8, mode=mode)
    rec = _graph_inputs(3, dims, seed=9)
    o_sorted, _, _ = _run(mine, rec, "cuda", torch.float32)
    T = rec["edge_index"].size(1)
    perm = torch.randperm(T, generator=torch.Generator().manual_seed(0))
    rec_p = dict(rec)
    rec_p["edge_index"] = rec["edge_index"][:, perm].contiguous()
    rec_p["sbf"], rec_p["edge_attr"] = rec["sbf"][perm], rec["edge_attr"][perm]
    _compare(ref, mine, rec_p)
    o_perm, _, _ = _run(mine, rec_p, "cuda", torch.float32)
    assert relerr(o_perm, o_sorted) < 3e-6     # only the summation / tile order differs


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_deterministic_bitwise(mode):
    dims = (128, 16, 42, 6, 128)
    _, mine = _oracle_pair(dims, seed=10, mode=mode)
    rec = _graph_inputs(8, dims, seed=11)
    runs = []
    for _ in range(2):
        mine.zero_grad()
        o, _, xs = _run(mine, rec, "cuda", torch.float32)
        runs.append([o.detach().clone()] + [xs[k].grad.clone() for k in INPUTS] +
                    [p.grad.clone() for p in mine.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)            # atomic-free => run-to-run identical


def test_batching_independence():
    """Block-diagonal graphs: a molecule's rows do not depend on what else is in the batch."""
    from x2gnn_b200 import synth
    dims = (128, 16, 42, 6, 128)
    _, mine = _oracle_pair(dims, seed=12)
    rec = _graph_inputs(4, dims, seed=13)
    with torch.no_grad():
        full = mine(rec["sbf"].cuda(), rec["rbf"].cuda(), x=rec["x"].cuda(),
                    edge_index=rec["edge_index"].cuda(), edge_attr=rec["edge_attr"].cuda())
    b = synth.qm9_batch(4, seed=13)
    E0 = int(b["edge_num"][0])
    m = rec["edge_index"][1] < E0
    with torch.no_grad():
        first = mine(rec["sbf"][m].cuda(), rec["rbf"][:E0].cuda(), x=rec["x"][:E0].cuda(),
                     edge_index=rec["edge_index"][:, m].cuda(), edge_attr=rec["edge_attr"][m].cuda())
    assert torch.equal(first, full[:E0])


def test_sbf_without_grad_and_frozen_inputs():
    dims = (64, 8, 10, 3, 20)
    ref, mine = _oracle_pair(dims, seed=14)
    rec = _graph_inputs(2, dims, seed=15)
    o, _, xs = _run(mine, rec, "cuda", torch.float32, sbf_grad=False)   # the reference graph: sbf has no grad
    assert xs["sbf"].grad is None
    o_ref, _, x_ref = _run(ref, rec, "cpu", torch.float64)
    assert relerr(xs["x"].grad, x_ref["x"].grad) < FP32_TOL


@pytest.mark.parametrize("dims,mode", [((128, 16, 42, 6, 128), "tf32x3"), ((128, 16, 42, 6, 128), "fp32"),
                                       ((64, 8, 10, 3, 20), "fp32"), ((256, 16, 112, 16, 128), "tf32x3")])
@pytest.mark.parametrize("rows", ["atoms", "few"])
def test_segment_constant_edge_attr_table(dims, mode, rows):
    """Opt-in fast path: edge_attr as a table [M, A] + edge_attr_index [E] must equal
    the reference layer fed the expanded edge_attr[edge_attr_index[edge_index[1]]] ([T, A]), including the
    gradient w.r.t. the table (= index_add of the per-triplet gradient) and every parameter gradient."""
    from x2gnn_b200 import synth
    D, H, S, R, A = dims
    ref, mine = _oracle_pair(dims, seed=20, mode=mode)
    b = synth.qm9_batch(5, seed=21)
    rec = _graph_inputs(5, dims, seed=21)
    E = rec["x"].size(0)
    g = torch.Generator().manual_seed(22)
    if rows == "atoms":      # xgnn.py:57-58: the row is a function of the bond's second (central) atom
        index = torch.from_numpy(b["edge_index"][1]).long()
        M = len(b["x"])
    else:                    # a handful of distinct rows (one per element), some rows unused
        M = 7
        index = torch.randint(0, 5, (E,), generator=g)
    table = torch.randn(M, A, generator=g)
    tgt = rec["edge_index"][1]

    tab_ref = table.double().requires_grad_(True)
    xs_ref = {k: rec[k].double().requires_grad_(True) for k in ("x", "rbf", "sbf")}
    o_ref = ref(xs_ref["sbf"], xs_ref["rbf"], x=xs_ref["x"], edge_index=rec["edge_index"],
                edge_attr=tab_ref[index[tgt]])
    o_ref.backward(rec["grad_out"].double())

    mine.zero_grad(set_to_none=True)
    tab = table.cuda().requires_grad_(True)
    xs = {k: rec[k].cuda().requires_grad_(True) for k in ("x", "rbf", "sbf")}
    o, (_, alpha) = mine(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"].cuda(), edge_attr=tab,
                         edge_attr_index=index.cuda(), return_attention_weights=True)
    o.backward(rec["grad_out"].cuda())
    assert relerr(o, o_ref) < FP32_TOL
    assert tab.grad.shape == table.shape
    assert relerr(tab.grad, tab_ref.grad) < FP32_TOL
    for k in xs:
        assert relerr(xs[k].grad, xs_ref[k].grad) < FP32_TOL, k
    pm = dict(mine.named_parameters())
    for k, p in ref.named_parameters():
        if k == "lin_key.bias":
            continue
        assert relerr(pm[k].grad, p.grad) < FP32_TOL, k
    # and it is the same function as the [T, A] form of this implementation
    with torch.no_grad():
        o_full = mine(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"].cuda(),
                      edge_attr=tab[index.cuda()[tgt.cuda()]])
    assert relerr(o, o_full) < 2e-6
    # without the alpha request the forward takes the plain (for D=128, C=8: bulk-copy staged) kernel
    with torch.no_grad():
        o_plain = mine(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"].cuda(), edge_attr=tab,
                       edge_attr_index=index.cuda())
    assert relerr(o_plain, o_ref) < FP32_TOL
    assert relerr(o_plain, o) < 2e-6
    with pytest.raises(IndexError):
        mine(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"].cuda(), edge_attr=tab,
             edge_attr_index=(index + M).cuda())
    with pytest.raises(ValueError):
        mine(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"].cuda(), edge_attr=tab,
             edge_attr_index=index[:-1].cuda())


def test_dropout_training_directional_derivative():
    dims = (64, 8, 10, 3, 20)
    _, mine = _oracle_pair(dims, seed=16, dropout=0.3)
    rec = _graph_inputs(2, dims, seed=17)
    dev = "cuda"
    args = {k: rec[k].to(dev) for k in INPUTS}
    ei = rec["edge_index"].to(dev)
    mine.eval()
    with torch.no_grad():
        o_eval = mine(args["sbf"], args["rbf"], x=args["x"], edge_index=ei, edge_attr=args["edge_attr"])
    mine.train()
    torch.manual_seed(0)
    x = args["x"].clone().requires_grad_(True)
    o1 = mine(args["sbf"], args["rbf"], x=x, edge_index=ei, edge_attr=args["edge_attr"])
    assert relerr(o1, o_eval) > 1e-3            # the mask is applied in training ...
    g = rec["grad_out"].to(dev)
    o1.backward(g)
    v = torch.randn_like(x)
    eps = 1e-2
    outs = []
    for sgn in (+1, -1):
        torch.manual_seed(0)                      # same dropout seed => same mask
        with torch.no_grad():
            outs.append(mine(args["sbf"], args["rbf"], x=args["x"] + sgn * eps * v, edge_index=ei,
                             edge_attr=args["edge_attr"]))
    fd = float(((outs[0] - outs[1]).double() * g.double()).sum() / (2 * eps))
    an = float((x.grad.double() * v.double()).sum())
    assert abs(fd - an) < 2e-2 * max(abs(an), 1.0)
    mine.eval()
    with torch.no_grad():                          # ... and never in eval
        assert torch.equal(mine(args["sbf"], args["rbf"], x=args["x"], edge_index=ei,
                                edge_attr=args["edge_attr"]), o_eval)


def test_loud_failures():
    from x2gnn_b200 import _lib
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    c = SBFTransformerConv(64, 8, heads=8, sbf_dim=10, rbf_dim=3, edge_dim=20)
    rec = _graph_inputs(1, (64, 8, 10, 3, 20), seed=1)
    with pytest.raises(_lib.X2Error):       # CPU tensors: no fallback
        c(rec["sbf"], rec["rbf"], x=rec["x"], edge_index=rec["edge_index"], edge_attr=rec["edge_attr"])
    c = c.cuda()
    with pytest.raises(TypeError):          # fp64 is not silently down-cast
        c(rec["sbf"].cuda().double(), rec["rbf"].cuda().double(), x=rec["x"].cuda().double(),
          edge_index=rec["edge_index"].cuda(), edge_attr=rec["edge_attr"].cuda().double())
    bad = SBFTransformerConv(96, 12, heads=8, sbf_dim=10, rbf_dim=3, edge_dim=20).cuda()   # C = 12: not a power of two
    with pytest.raises(NotImplementedError):
        bad(rec["sbf"].cuda(), rec["rbf"].cuda(), x=torch.randn(rec["x"].size(0), 96, device="cuda"),
            edge_index=rec["edge_index"].cuda(), edge_attr=rec["edge_attr"].cuda())


def test_inplace_modification_between_forward_and_backward_is_an_error():
    """The layer's inputs / buffers are saved through save_for_backward: changing an input in place after the forward
    must raise in the backward instead of producing a gradient of something else."""
    dims = (64, 8, 10, 3, 20)
    _, mine = _oracle_pair(dims, seed=11)
    rec = _graph_inputs(2, dims, seed=12)
    x = rec["x"].cuda().requires_grad_(True)
    xin = x * 1.0
    out = mine(rec["sbf"].cuda(), rec["rbf"].cuda(), x=xin, edge_index=rec["edge_index"].cuda(),
               edge_attr=rec["edge_attr"].cuda())
    xin.add_(1.0)
    with pytest.raises(RuntimeError, match="inplace|in-place"):
        out.sum().backward()
