"""Pin the oracle (CPU restatement) against golden vectors produced by the reference's
own source (tests/golden/make_golden.py).  CPU only."""
import math
import os

import pytest
import torch

from oracle import bases, conv as oconv, graph, model as omodel


# ------------------------------------------------------------------ graph indices
@pytest.mark.parametrize("case", ["kat5", "mol9", "mol17", "mol29", "ball30"])
def test_radius_graph_and_triplets(golden, case):
    g = golden("graph")[case]
    D = graph.calculate_Dij(g["pos"])
    if "Dij" in g:
        assert torch.equal(torch.nan_to_num(D, nan=-1.0), torch.nan_to_num(g["Dij"], nan=-1.0))
    ei = graph.gen_bonds_mini(D, g["cutoff"])
    assert torch.equal(ei, g["edge_index"])
    tri, ej, ei_, ek = graph.vertex_to_edge_2(ei, g["pos"].size(0))
    assert tri.dtype == torch.int64
    assert torch.equal(tri, g["triplets_index"].long())
    assert torch.equal(ej, g["edge_j"].long())
    assert torch.equal(ei_, g["edge_i"].long())
    assert torch.equal(ek, g["edge_k"].long())


@pytest.mark.parametrize("case", ["batch3", "directed_unsorted"])
def test_triplets_batched_and_directed(golden, case):
    g = golden("graph")[case]
    for fn in (graph.vertex_to_edge_2, graph.vertex_to_edge_2_bruteforce):
        tri, ej, ei_, ek = fn(g["edge_index"], g["num_nodes"])
        assert torch.equal(tri, g["triplets_index"].long())
        assert torch.equal(ej, g["edge_j"].long())
        assert torch.equal(ei_, g["edge_i"].long())
        assert torch.equal(ek, g["edge_k"].long())


def test_kat5_literal():
    """SURVEY.md App. E known-answer test."""
    pos = torch.tensor([[0, 0, 0], [1.2, 0, 0], [2, 1, 0], [9, 9, 9], [2.2, -1, 0.3]])
    ei = graph.gen_bonds_mini(graph.calculate_Dij(pos), 2.0)
    assert ei.tolist() == [[0, 1, 1, 1, 2, 4], [1, 0, 2, 4, 1, 1]]
    tri, ej, ei_, ek = graph.vertex_to_edge_2(ei, 5)
    assert tri.tolist() == [[2, 3, 1, 3, 1, 2], [0, 0, 4, 4, 5, 5]]
    assert ej.tolist() == [1] * 6
    assert ei_.tolist() == [0, 0, 2, 2, 4, 4]
    assert ek.tolist() == [2, 4, 0, 4, 0, 2]


# ------------------------------------------------------------------ bases
def test_envelope_and_radial(golden):
    b = golden("bases")
    d = b["d"]
    assert torch.allclose(bases.poly_envelop(d), b["env_f32"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(bases.poly_envelop(d.double()), b["env_f64"], rtol=1e-13, atol=1e-13)
    f = bases.radial_frequencies(6)
    assert torch.allclose(bases.radial_basis(d, f), b["rbf_f32"], rtol=0, atol=1e-6)
    assert torch.allclose(bases.radial_basis(d.double(), f.double()), b["rbf_f64"], atol=1e-13)
    # App. B KATs
    assert bases.envelope_coeffs(5) == (6, -28.0, 48, -21.0)
    assert float(bases.poly_envelop(torch.tensor([2.5]))) == 1.7109375


@pytest.mark.parametrize("LR", [(7, 6), (3, 4)])
def test_f_b_2d_vs_reference_fp64(golden, LR):
    b = golden("bases")
    L, R = LR
    z, _ = bases.bessel_tables(L, R)
    assert torch.equal(torch.from_numpy(z), b[f"zeros_{L}_{R}"])
    ref = b[f"sbf_{L}_{R}_f64"]
    got = bases.f_b_2d(b["d"].double(), b["angles"].double(), b["src"], L, R)
    # sympy prints its float32-derived constants with 15 significant digits
    assert torch.allclose(got, ref, rtol=1e-9, atol=1e-9 * float(ref.abs().max()))
    # the reference's own fp32 evaluation is only ~1e-3-accurate for l >= 5 (App. B)
    err32 = (b[f"sbf_{L}_{R}_f32"].double() - ref).abs().max() / ref.abs().max()
    assert err32 < 5e-4



@pytest.mark.parametrize("LR", [(7, 6), (3, 4)])
def test_basis_gradients_vs_reference_autograd(golden, LR):
    """d / d(distance) and d / d(angle) of the oracle's expansions against autograd through the REFERENCE's own
    lambdified expressions in fp64 (tests/golden/make_golden.py): the oracle's j_l' comes from scipy."""
    _rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
    b = golden("bases")
    L, R = LR
    d = b["d"].double().requires_grad_(True)
    a = b["angles"].double().requires_grad_(True)
    gd, ga = torch.autograd.grad(bases.f_b_2d(d, a, b["src"], L, R), (d, a), b[f"sbf_{L}_{R}_go"])
    assert _rel(gd, b[f"sbf_{L}_{R}_gd_f64"]) < 1e-7      # (float32-rounded zeros / normalisers in both)
    assert _rel(ga, b[f"sbf_{L}_{R}_gang_f64"]) < 1e-7
    d = b["d"].double().requires_grad_(True)
    assert _rel(torch.autograd.grad(bases.poly_envelop(d, 5.0, 5), d, b["env_go"])[0], b["env_gd_f64"]) < 1e-12
    a = b["angles"].double().requires_grad_(True)
    assert _rel(torch.autograd.grad(bases.angular_basis(a, 7), a, b["cbf_7_go"])[0], b["cbf_7_gang_f64"]) < 1e-12


def test_angular_basis(golden):
    b = golden("bases")
    assert torch.allclose(bases.angular_basis(b["angles"].double(), 7), b["cbf_7_f64"], atol=1e-12)


# ------------------------------------------------------------------ conv layer
def _oracle_conv(rec, dtype):
    D, H, S, R, A = rec["dims"]
    c = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, dropout=0,
                                       edge_dim=A).to(dtype)
    assert list(c.state_dict().keys()) == list(rec["state_dict"].keys())
    c.load_state_dict({k: v.to(dtype) for k, v in rec["state_dict"].items()})
    return c


@pytest.mark.parametrize("tag", ["cfg", "small", "c16"])
def test_conv_fp64_fwd_bwd(golden, tag):
    rec = golden("conv")[tag]
    c = _oracle_conv(rec, torch.float64)
    xs = {k: rec[k].double().requires_grad_(True) for k in ("x", "rbf", "sbf", "edge_attr")}
    out, (_, alpha) = c(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"],
                        edge_attr=xs["edge_attr"], return_attention_weights=True)
    out.backward(rec["grad_out"].double())
    tol = dict(rtol=1e-10, atol=1e-10)
    assert torch.allclose(out, rec["out_f64"], **tol)
    if "alpha_f64" in rec:
        assert torch.allclose(alpha, rec["alpha_f64"], **tol)
    for k in xs:
        assert torch.allclose(xs[k].grad, rec[f"grad_{k}_f64"], **tol), k
    for k, p in c.named_parameters():
        assert torch.allclose(p.grad, rec[f"gradp_{k}_f64"], **tol), k
    # App. A identity: grad(lin_key.bias) == 0
    assert c.lin_key.bias.grad.abs().max() < 1e-10


def test_conv_shape_outside_kernel_widths_vs_reference_fp64(golden):
    """in_channels != heads * out_channels (sbftransformer_conv.py:19,47-48): the oracle against the reference's own
    layer at (in 48, heads 2, out 16) in fp64 -- the case the kernels run zero-padded (padded_width)."""
    rec = golden("conv")["odd48"]
    in_ch, H, C, S, R, A = rec["shape"]
    c = oconv.OracleSBFTransformerConv(in_ch, C, heads=H, sbf_dim=S, rbf_dim=R, dropout=0, edge_dim=A).double()
    assert list(c.state_dict().keys()) == list(rec["state_dict"].keys())
    c.load_state_dict({k: v.double() for k, v in rec["state_dict"].items()})
    xs = {k: rec[k].double().requires_grad_(True) for k in ("x", "rbf", "sbf", "edge_attr")}
    out, (_, alpha) = c(xs["sbf"], xs["rbf"], x=xs["x"], edge_index=rec["edge_index"],
                        edge_attr=xs["edge_attr"], return_attention_weights=True)
    out.backward(rec["grad_out"].double())
    tol = dict(rtol=1e-10, atol=1e-10)
    assert out.shape == (rec["x"].size(0), H * C) and torch.allclose(out, rec["out_f64"], **tol)
    assert torch.allclose(alpha, rec["alpha_f64"], **tol)
    for k in xs:
        assert torch.allclose(xs[k].grad, rec[f"grad_{k}_f64"], **tol), k
    for k, p in c.named_parameters():
        assert torch.allclose(p.grad, rec[f"gradp_{k}_f64"], **tol), k


def test_conv_fp32(golden):
    rec = golden("conv")["cfg"]
    c = _oracle_conv(rec, torch.float32)
    out = c(rec["sbf"], rec["rbf"], x=rec["x"], edge_index=rec["edge_index"], edge_attr=rec["edge_attr"])
    scale = float(rec["out_f32"].abs().max())
    assert torch.allclose(out, rec["out_f32"], rtol=1e-5, atol=1e-5 * scale)


# ------------------------------------------------------------------ full model
def test_model_keys_match_reference(golden):
    m = golden("model")
    torch.manual_seed(0)
    mine = omodel.XGNNPoly(4, 7, 6, 128, 16, 128)
    assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == m["cfg_keys"]
    assert sum(p.numel() for p in mine.parameters()) == 1158795


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_model_prediction(golden, dtype):
    rec = golden("model")["small"]
    net = omodel.XGNNPoly(**rec["hparams"]).to(dtype)
    net.load_state_dict({k: v.to(dtype) for k, v in rec["state_dict"].items()})
    net.eval()
    data = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v)
            for k, v in rec["batch"].items()}
    with torch.no_grad():
        pred = net(data)
    ref = rec["pred_f64"] if dtype == torch.float64 else rec["pred_f32"]
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    assert torch.allclose(pred, ref, rtol=tol, atol=tol * float(ref.abs().max()))


def _same(a, b, path=""):
    """Structural equality of two fixtures: index tensors bit-equal, floating tensors to 1e-12."""
    if isinstance(a, dict):
        assert set(a) == set(b), path
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif torch.is_tensor(a):
        assert a.shape == b.shape and a.dtype == b.dtype, path
        if a.dtype.is_floating_point:
            assert float((a.double() - b.double()).abs().max()) <= 1e-12 if a.numel() else True, path
        else:
            assert torch.equal(a, b), path
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (u, v) in enumerate(zip(a, b)):
            _same(u, v, f"{path}[{i}]")
    else:
        assert a == b, path


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference mount exists only in the build container")
def test_committed_fixtures_reproduce_from_the_reference(tmp_path):
    """tests/golden/*.pt are what tests/golden/make_golden.py produces TODAY from the unmodified reference
    source (a stale fixture is still a valid reference output, but nobody could regenerate it)."""
    import subprocess
    import sys
    gen = os.path.join(os.path.dirname(__file__), "golden", "make_golden.py")
    subprocess.run([sys.executable, gen, "--out", str(tmp_path)], check=True, capture_output=True, timeout=900)
    for name in ("graph", "bases", "conv", "model"):
        _same(torch.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.pt")),
              torch.load(tmp_path / f"{name}.pt"), name)
